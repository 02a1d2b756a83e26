/* ORACLE / TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * CPU restatement of the seven xiph/opus DNN entry points the reference's C
 * core codec calls (src/rade_enc.c:69-113, src/rade_dec.c:63-101,
 * src/rade_enc_data.c:227866-227882, src/test_rade_enc.c:60).
 *
 * Third-party dependency being restated: xiph/opus, commit
 * 940d4e5af64351ca8ba8390df3f555484c567fbb (cmake/BuildOpus.cmake:10), files
 * dnn/nnet.c, dnn/nnet_arch.h, dnn/vec.h (generic-C branch, i.e. what `arch=0`
 * names: src/rade_api.c:421, src/test_rade_enc.c:88), dnn/parse_lpcnet_weights.c.
 * opus is NOT present in /root/reference, so this file follows the *published
 * algorithm* of that version:
 *
 *   int8 linear : x_q[i] = (int)floor(.5 + 127*x[i]);  acc = sum_j w_q[i][j]*x_q[j]
 *                 (weights stored as 8(out) x 4(in) blocks, block order
 *                 (out/8, in/4) — weight-exchange/wexchange/c_export/common.py:59-67;
 *                 optional per-8-output index list [count, in_pos...] — :156-170);
 *                 out[i] = acc*scale[i] + bias[i]           (float ops, in this order)
 *   float linear: out[i] = sum_j W[j][i]*x[j] accumulated sequentially in j, + bias[i]
 *   tanh        : rational approximation x*(N0+N1 x^2+N2 x^4)/(D0+D1 x^2+D2 x^4), clamped to [-1,1]
 *   sigmoid     : .5 + .5*tanh_approx(.5 x)
 *   GRU         : gates ordered z,r,n (common.py:360-368);  zr = sigmoid(Wi x + bi + Wr h + br);
 *                 n = tanh(Wi_n x + bi_n + r*(Wr_n h + br_n));  h = z*h + (1-z)*n
 *   GLU         : out = in * sigmoid(W in + b)
 *   conv1d      : tmp = [mem, in]; linear; act; mem = tmp[in_size:]
 *   conv1d_dil  : taps `dilation` frames apart from a FIFO of dilation*(k-1) frames
 *
 * PARITY STATUS: "parity unpinned" at the opus boundary — the reference ships no
 * value-level golden vectors for this path and its own x86 build uses SIMD
 * variants with hardware-approximate reciprocals.  What IS pinned (see
 * tools/make_golden.py, tests/test_oracle_core.py): the float-weight variant of
 * this file driving the reference's own rade_enc.c/rade_dec.c reproduces the
 * reference PyTorch CoreEncoder/DecoderStatefull to ~1e-5 (layouts, gate order,
 * conv tap order, state handling), and the int8 variant sits inside the
 * reference's own C-vs-Python acceptance band (|delta loss| < 0.01,
 * CMakeLists.txt:521-556).
 *
 * Build with -ffp-contract=off: every float op below is a separately rounded
 * IEEE-754 binary32 operation, which is what the CUDA epilogues replicate.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "nnet.h"
#include "os_support.h"

#define ORACLE_MAX_DIM 2048          /* >= RADE_MAX_CONV_INPUTS (1536), src/rade_constants.h:16 */
#define SPARSE_BLOCK 32              /* 8 outputs x 4 inputs */

static double g_max_abs_acc = 0.0;

double oracle_nnet_max_abs_acc(int reset)
{
  double v = g_max_abs_acc;
  if (reset) g_max_abs_acc = 0.0;
  return v;
}

/* ---- activations (generic C branch of dnn/vec.h) ------------------------ */

static float tanh_rational(float x)
{
#ifdef ORACLE_EXACT_ACT   /* validation-only build: isolates layout errors from approximation error */
  return tanhf(x);
#endif
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f;
  const float D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
  float x2 = x*x;
  float num = (N2*x2 + N1)*x2 + N0;
  float den = (D2*x2 + D1)*x2 + D0;
  float y = num*x/den;
  if (y > 1.f) y = 1.f;
  if (y < -1.f) y = -1.f;
  return y;
}

static float sigmoid_rational(float x)
{
#ifdef ORACLE_EXACT_ACT
  return 1.f/(1.f + expf(-x));
#endif
  return .5f + .5f*tanh_rational(.5f*x);
}

static void apply_activation(float *y, const float *x, int n, int activation)
{
  int i;
  switch (activation) {
  case ACTIVATION_SIGMOID: for (i=0;i<n;i++) y[i] = sigmoid_rational(x[i]); break;
  case ACTIVATION_TANH:    for (i=0;i<n;i++) y[i] = tanh_rational(x[i]); break;
  case ACTIVATION_LINEAR:  if (y != x) for (i=0;i<n;i++) y[i] = x[i]; break;
  default:
    fprintf(stderr, "oracle nnet_shim: activation %d not used by RADE\n", activation);
    abort();
  }
}

/* ---- linear layers ------------------------------------------------------- */

static void quantise_input(opus_int8 *xq, const float *x, int n)
{
  int i;
  /* C semantics: 127*x[i] is a FLOAT product (int promoted to float, result rounded to binary32); the literal .5
   * then promotes the SUM to double; floor; truncate to int8 */
  for (i=0;i<n;i++) xq[i] = (opus_int8)(int)floor(.5 + 127*x[i]);
}

static void note_acc(const float *acc, int n)
{
  int i;
  for (i=0;i<n;i++) { double a = fabs((double)acc[i]); if (a > g_max_abs_acc) g_max_abs_acc = a; }
}

/* dense 8x4-blocked int8 matrix times quantised vector; float accumulators
 * holding integer values, block after block, exactly as the generic C gemv */
static void gemv_int8_blocked(float *out, const opus_int8 *w, const int *idx,
                              const float *scale, int rows, int cols, const float *x)
{
  opus_int8 xq[ORACLE_MAX_DIM];
  int i, j, k;
  for (i=0;i<rows;i++) out[i] = 0;
  quantise_input(xq, x, cols);
  for (i=0;i<rows;i+=8) {
    int nblocks = idx ? *idx++ : cols/4;
    for (j=0;j<nblocks;j++) {
      int pos = idx ? *idx++ : 4*j;
      float x0 = xq[pos], x1 = xq[pos+1], x2 = xq[pos+2], x3 = xq[pos+3];
      for (k=0;k<8;k++)
        out[i+k] += (w[4*k]*x0 + w[4*k+1]*x1 + w[4*k+2]*x2 + w[4*k+3]*x3);
      w += SPARSE_BLOCK;
    }
  }
  note_acc(out, rows);
  for (i=0;i<rows;i++) out[i] *= scale[i];
}

/* float weights stored [in][out]; each output accumulates its products in
 * input order (the generic sgemv adds one column of W at a time) */
static void gemv_float(float *out, const float *w, int rows, int cols, const float *x)
{
  int i, j;
  for (i=0;i<rows;i++) out[i] = 0;
  for (j=0;j<cols;j++) {
    const float *wj = &w[(size_t)j*rows];
    float xj = x[j];
    for (i=0;i<rows;i++) out[i] += wj[i]*xj;
  }
}

/* float weights in the block-indexed ("sparse") storage: per 8 outputs a
 * [count, pos...] list, blocks stored [4 in][8 out] (common.py:166) */
static void gemv_float_blocked(float *out, const float *w, const int *idx, int rows, const float *x)
{
  int i, j, k;
  for (i=0;i<rows;i++) out[i] = 0;
  for (i=0;i<rows;i+=8) {
    int nblocks = *idx++;
    for (j=0;j<nblocks;j++) {
      int pos = *idx++;
      float x0 = x[pos], x1 = x[pos+1], x2 = x[pos+2], x3 = x[pos+3];
      for (k=0;k<8;k++)
        out[i+k] += w[k]*x0 + w[8+k]*x1 + w[16+k]*x2 + w[24+k]*x3;
      w += SPARSE_BLOCK;
    }
  }
}

static void linear_forward(const LinearLayer *l, float *out, const float *in)
{
  int i;
  const int M = l->nb_inputs, N = l->nb_outputs;
  if (in == out) { fprintf(stderr, "oracle nnet_shim: in-place linear\n"); abort(); }
  if (M > ORACLE_MAX_DIM) { fprintf(stderr, "oracle nnet_shim: nb_inputs %d too large\n", M); abort(); }
  if (l->float_weights != NULL) {
    if (l->weights_idx != NULL) gemv_float_blocked(out, l->float_weights, l->weights_idx, N, in);
    else gemv_float(out, l->float_weights, N, M, in);
  } else if (l->weights != NULL) {
    gemv_int8_blocked(out, l->weights, l->weights_idx, l->scale, N, M, in);
  } else {
    for (i=0;i<N;i++) out[i] = 0;
  }
  if (l->bias != NULL) for (i=0;i<N;i++) out[i] += l->bias[i];
  if (l->diag != NULL) {
    /* only GRU recurrent matrices carry a diagonal; RADE exports none */
    for (i=0;i<M;i++) {
      out[i]     += l->diag[i]*in[i];
      out[i+M]   += l->diag[i+M]*in[i];
      out[i+2*M] += l->diag[i+2*M]*in[i];
    }
  }
}

/* ---- the five compute entry points -------------------------------------- */

void compute_generic_dense(const LinearLayer *layer, float *output, const float *input, int activation, int arch)
{
  (void)arch;
  linear_forward(layer, output, input);
  apply_activation(output, output, layer->nb_outputs, activation);
}

void compute_generic_gru(const LinearLayer *input_weights, const LinearLayer *recurrent_weights, float *state, const float *in, int arch)
{
  float gates[3*ORACLE_MAX_DIM/4];
  float rec[3*ORACLE_MAX_DIM/4];
  const int N = recurrent_weights->nb_inputs;
  float *z = gates, *r = gates + N, *h = gates + 2*N;
  int i;
  (void)arch;
  if (3*N != recurrent_weights->nb_outputs || input_weights->nb_outputs != 3*N || 3*N > 3*ORACLE_MAX_DIM/4) {
    fprintf(stderr, "oracle nnet_shim: bad GRU shape\n"); abort();
  }
  linear_forward(input_weights, gates, in);
  linear_forward(recurrent_weights, rec, state);
  for (i=0;i<2*N;i++) gates[i] += rec[i];
  apply_activation(gates, gates, 2*N, ACTIVATION_SIGMOID);
  for (i=0;i<N;i++) h[i] += rec[2*N+i]*r[i];
  apply_activation(h, h, N, ACTIVATION_TANH);
  for (i=0;i<N;i++) h[i] = z[i]*state[i] + (1-z[i])*h[i];
  for (i=0;i<N;i++) state[i] = h[i];
}

void compute_glu(const LinearLayer *layer, float *output, const float *input, int arch)
{
  float gate[ORACLE_MAX_DIM];
  int i;
  (void)arch;
  linear_forward(layer, gate, input);
  apply_activation(gate, gate, layer->nb_outputs, ACTIVATION_SIGMOID);
  for (i=0;i<layer->nb_outputs;i++) output[i] = input[i]*gate[i];
}

void compute_generic_conv1d(const LinearLayer *layer, float *output, float *mem, const float *input, int input_size, int activation, int arch)
{
  float tmp[ORACLE_MAX_DIM];
  const int hist = layer->nb_inputs - input_size;
  (void)arch;
  if (hist) OPUS_COPY(tmp, mem, hist);
  OPUS_COPY(&tmp[hist], input, input_size);
  linear_forward(layer, output, tmp);
  apply_activation(output, output, layer->nb_outputs, activation);
  if (hist) OPUS_COPY(mem, &tmp[input_size], hist);
}

void compute_generic_conv1d_dilation(const LinearLayer *layer, float *output, float *mem, const float *input, int input_size, int dilation, int activation, int arch)
{
  float tmp[ORACLE_MAX_DIM];
  const int ksize = layer->nb_inputs/input_size;
  const int hist = layer->nb_inputs - input_size;
  int i;
  (void)arch;
  if (dilation == 1) OPUS_COPY(tmp, mem, hist);
  else for (i=0;i<ksize-1;i++) OPUS_COPY(&tmp[i*input_size], &mem[i*input_size*dilation], input_size);
  OPUS_COPY(&tmp[hist], input, input_size);
  linear_forward(layer, output, tmp);
  apply_activation(output, output, layer->nb_outputs, activation);
  if (dilation == 1) OPUS_COPY(mem, &tmp[input_size], hist);
  else {
    const int fifo = input_size*dilation*(ksize-1);
    OPUS_MOVE(mem, &mem[input_size], fifo - input_size);
    OPUS_COPY(&mem[fifo - input_size], input, input_size);
  }
}

/* ---- weight tables -------------------------------------------------------- */

static const WeightArray *lookup(const WeightArray *arrays, const char *name)
{
  while (arrays->name != NULL && strcmp(arrays->name, name) != 0) arrays++;
  return arrays->name ? arrays : NULL;
}

static const void *need(const WeightArray *arrays, const char *name, size_t size)
{
  const WeightArray *a = lookup(arrays, name);
  return (a && (size_t)a->size == size) ? a->data : NULL;
}

/* optional array: absent is fine, present with the wrong size is an error */
static const void *maybe(const WeightArray *arrays, const char *name, size_t size, int *err)
{
  const WeightArray *a = lookup(arrays, name);
  *err = (a != NULL && (size_t)a->size != size);
  return (a && (size_t)a->size == size) ? a->data : NULL;
}

/* validate a block index list and count its blocks */
static const int *need_idx(const WeightArray *arrays, const char *name, int nb_in, int nb_out, int *total_blocks)
{
  const WeightArray *a = lookup(arrays, name);
  const int *p; int remain, out = nb_out, total = 0;
  if (a == NULL) return NULL;
  p = (const int*)a->data; remain = a->size/(int)sizeof(int);
  while (remain > 0) {
    int n = *p++, i; remain--;
    if (remain < n) return NULL;
    for (i=0;i<n;i++) { int pos = *p++; remain--; if (pos < 0 || pos+3 >= nb_in || (pos&3)) return NULL; }
    out -= 8; total += n;
  }
  if (out != 0) return NULL;
  *total_blocks = total;
  return (const int*)a->data;
}

int linear_init(LinearLayer *layer, const WeightArray *arrays,
  const char *bias, const char *subias, const char *weights, const char *float_weights,
  const char *weights_idx, const char *diag, const char *scale, int nb_inputs, int nb_outputs)
{
  int err = 0;
  memset(layer, 0, sizeof(*layer));
  if (bias   && !(layer->bias   = need(arrays, bias,   nb_outputs*sizeof(float)))) return 1;
  if (subias && !(layer->subias = need(arrays, subias, nb_outputs*sizeof(float)))) return 1;
  if (weights_idx) {
    int total_blocks = 0;
    if (!(layer->weights_idx = need_idx(arrays, weights_idx, nb_inputs, nb_outputs, &total_blocks))) return 1;
    if (weights && !(layer->weights = need(arrays, weights, (size_t)SPARSE_BLOCK*total_blocks))) return 1;
    if (float_weights) {
      layer->float_weights = maybe(arrays, float_weights, (size_t)SPARSE_BLOCK*total_blocks*sizeof(float), &err);
      if (err) return 1;
    }
  } else {
    if (weights && !(layer->weights = need(arrays, weights, (size_t)nb_inputs*nb_outputs))) return 1;
    if (float_weights) {
      layer->float_weights = maybe(arrays, float_weights, (size_t)nb_inputs*nb_outputs*sizeof(float), &err);
      if (err) return 1;
    }
  }
  if (diag && !(layer->diag = need(arrays, diag, nb_outputs*sizeof(float)))) return 1;
  if (weights && !(layer->scale = need(arrays, scale, nb_outputs*sizeof(float)))) return 1;
  layer->nb_inputs = nb_inputs;
  layer->nb_outputs = nb_outputs;
  return 0;
}

/* "DNNw" blob: sequence of 64-byte WeightHead records each followed by the
 * payload padded to a multiple of 64 bytes (src/write_rade_weights.c:51-74).
 * Returns the number of arrays; *list is malloc'd and NULL-terminated and
 * points INTO data (which must stay mapped, src/test_rade_enc.c:56). */
int parse_weights(WeightArray **list, const void *data, int len)
{
  const unsigned char *p = (const unsigned char*)data;
  int n = 0, cap = 64;
  *list = (WeightArray*)calloc(cap, sizeof(WeightArray));
  while (len > 0) {
    const WeightHead *h = (const WeightHead*)p;
    if (len < WEIGHT_BLOCK_SIZE || memcmp(h->head, "DNNw", 4) != 0 || h->version != WEIGHT_BLOB_VERSION
        || h->size <= 0 || h->block_size < h->size || h->block_size > len - WEIGHT_BLOCK_SIZE
        || h->name[sizeof(h->name)-1] != 0) {
      free(*list); *list = NULL; return -1;
    }
    if (n + 2 > cap) { cap *= 2; *list = (WeightArray*)realloc(*list, cap*sizeof(WeightArray)); }
    (*list)[n].name = h->name;
    (*list)[n].type = h->type;
    (*list)[n].size = h->size;
    (*list)[n].data = p + WEIGHT_BLOCK_SIZE;
    n++;
    p += WEIGHT_BLOCK_SIZE + h->block_size;
    len -= WEIGHT_BLOCK_SIZE + h->block_size;
  }
  (*list)[n].name = NULL; (*list)[n].type = 0; (*list)[n].size = 0; (*list)[n].data = NULL;
  return n;
}
