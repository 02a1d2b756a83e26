/* ORACLE / TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Restated subset of xiph/opus `dnn/nnet.h` @940d4e5 — the declarations that
 * /root/reference/src/{rade_enc,rade_dec,rade_*_data,test_rade_*}.c need in
 * order to compile unchanged.  opus itself is absent from /root/reference
 * (fetched from the network by cmake/BuildOpus.cmake), so every item below is
 * derived from in-repo usage; see SURVEY.md Appendix C for the evidence:
 *   - WeightArray field order {name,type,size,data}: positional initialisers at
 *     the tail of src/rade_enc_data.c; list[i].name/.size in src/test_rade_enc.c:61-63
 *   - WEIGHT_TYPE_* codes: headers of bin/model19_check3.bin
 *   - WeightHead layout / WEIGHT_BLOCK_SIZE: src/write_rade_weights.c:54-72
 *   - the seven prototypes: src/opus-nnet.h.diff:6-30
 *   - linear_init argument order: src/rade_enc_data.c:227866
 */
#ifndef ORACLE_NNET_H
#define ORACLE_NNET_H

#include <stddef.h>
#include "opus_types.h"

#ifndef RADE_EXPORT
#define RADE_EXPORT __attribute__((visibility("default")))
#endif

#define ACTIVATION_LINEAR  0
#define ACTIVATION_SIGMOID 1
#define ACTIVATION_TANH    2
#define ACTIVATION_RELU    3
#define ACTIVATION_SOFTMAX 4
#define ACTIVATION_SWISH   5

#define WEIGHT_BLOB_VERSION 0
#define WEIGHT_BLOCK_SIZE   64

#define WEIGHT_TYPE_float   0
#define WEIGHT_TYPE_int     1
#define WEIGHT_TYPE_qweight 2
#define WEIGHT_TYPE_int8    3

typedef struct {
  const char *name;
  int type;
  int size;
  const void *data;
} WeightArray;

typedef struct {
  char head[4];
  int version;
  int type;
  int size;
  int block_size;
  char name[44];
} WeightHead;

/* y = W x + b.  Either float_weights ([in][out]) or int8 weights (8x4 blocks,
 * optional block index list) with a per-output scale. */
typedef struct {
  const float *bias;
  const float *subias;
  const opus_int8 *weights;
  const float *float_weights;
  const int *weights_idx;
  const float *diag;
  const float *scale;
  int nb_inputs;
  int nb_outputs;
} LinearLayer;

void RADE_EXPORT compute_generic_dense(const LinearLayer *layer, float *output, const float *input, int activation, int arch);
void RADE_EXPORT compute_generic_gru(const LinearLayer *input_weights, const LinearLayer *recurrent_weights, float *state, const float *in, int arch);
void RADE_EXPORT compute_generic_conv1d(const LinearLayer *layer, float *output, float *mem, const float *input, int input_size, int activation, int arch);
void RADE_EXPORT compute_generic_conv1d_dilation(const LinearLayer *layer, float *output, float *mem, const float *input, int input_size, int dilation, int activation, int arch);
void RADE_EXPORT compute_glu(const LinearLayer *layer, float *output, const float *input, int arch);

int RADE_EXPORT parse_weights(WeightArray **list, const void *data, int len);

int RADE_EXPORT linear_init(LinearLayer *layer, const WeightArray *arrays,
  const char *bias, const char *subias, const char *weights, const char *float_weights,
  const char *weights_idx, const char *diag, const char *scale, int nb_inputs, int nb_outputs);

/* oracle-only instrumentation: largest |integer accumulator| seen by the int8
 * GEMV since the last reset (float accumulation in the generic C path is exact
 * only while this stays below 2^24). */
double RADE_EXPORT oracle_nnet_max_abs_acc(int reset);

#endif
