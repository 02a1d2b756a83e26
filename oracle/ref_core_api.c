/* ORACLE / TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Thin C-callable wrapper around the REFERENCE's own core codec sources
 * (/root/reference/src/rade_enc.c, rade_dec.c, rade_enc_data.c, rade_dec_data.c,
 * compiled where they lie by oracle/build.py) linked against oracle/nnet_shim.
 * Mirrors what src/rade_api.c does around the codec: init_radeenc(…, 84) /
 * init_radedec(…, 84) on the compiled-in arrays (src/rade_api.c:216, :309),
 * arch = 0, bottleneck = 3 (src/rade_api.c:421-422), one RADEEncState /
 * RADEDecState per stream.  Multi-stream entry points parallelise over streams
 * with OpenMP so bench.py can time the reference C path on all host cores.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "rade_core.h"
#include "rade_enc.h"
#include "rade_enc_data.h"
#include "rade_dec.h"
#include "rade_dec_data.h"

typedef struct {
  RADEEnc enc;
  RADEDec dec;
  WeightArray *blob_list;   /* non-NULL when loaded from a DNNw blob */
  void *blob;
} ref_core_model;

#define REF_API __attribute__((visibility("default")))

REF_API void *ref_core_open(const void *blob, int len, int input_dim, int output_dim)
{
  ref_core_model *m = (ref_core_model*)calloc(1, sizeof(*m));
  const WeightArray *enc_arrays = radeenc_arrays, *dec_arrays = radedec_arrays;
  if (blob != NULL) {
    m->blob = malloc(len);
    memcpy(m->blob, blob, len);
    if (parse_weights(&m->blob_list, m->blob, len) <= 0) { free(m->blob); free(m); return NULL; }
    enc_arrays = dec_arrays = m->blob_list;
  }
  if (init_radeenc(&m->enc, enc_arrays, input_dim) != 0) { fprintf(stderr, "ref_core_open: init_radeenc failed\n"); return NULL; }
  if (init_radedec(&m->dec, dec_arrays, output_dim) != 0) { fprintf(stderr, "ref_core_open: init_radedec failed\n"); return NULL; }
  return m;
}

REF_API void ref_core_close(void *h)
{
  ref_core_model *m = (ref_core_model*)h;
  if (!m) return;
  free(m->blob_list); free(m->blob); free(m);
}

REF_API int ref_core_uses_float_weights(void *h)
{
  ref_core_model *m = (ref_core_model*)h;
  return m->enc.enc_gru1_input.float_weights != NULL;
}

REF_API int ref_enc_state_size(void) { return (int)sizeof(RADEEncState); }
REF_API int ref_dec_state_size(void) { return (int)sizeof(RADEDecState); }

REF_API void ref_enc_state_init(void *states, int n)
{
  RADEEncState *s = (RADEEncState*)states; int i;
  for (i=0;i<n;i++) rade_init_encoder(&s[i]);
}
REF_API void ref_dec_state_init(void *states, int n)
{
  RADEDecState *s = (RADEDecState*)states; int i;
  for (i=0;i<n;i++) rade_init_decoder(&s[i]);
}

/* features: [n_streams][n_steps][in_dim]  ->  z: [n_streams][n_steps][80] */
REF_API void ref_core_encode(void *h, void *states, int n_streams, int n_steps,
                             const float *features, int in_dim, float *z, int bottleneck, int nthreads)
{
  ref_core_model *m = (ref_core_model*)h;
  RADEEncState *st = (RADEEncState*)states;
  int s;
  (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (s=0;s<n_streams;s++) {
    int t;
    for (t=0;t<n_steps;t++)
      rade_core_encoder(&st[s], &m->enc, &z[((size_t)s*n_steps+t)*RADE_LATENT_DIM],
                        &features[((size_t)s*n_steps+t)*in_dim], 0, bottleneck);
  }
}

/* z: [n_streams][n_steps][80]  ->  features: [n_streams][n_steps][out_dim] */
REF_API void ref_core_decode(void *h, void *states, int n_streams, int n_steps,
                             const float *z, float *features, int out_dim, int nthreads)
{
  ref_core_model *m = (ref_core_model*)h;
  RADEDecState *st = (RADEDecState*)states;
  int s;
  (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (s=0;s<n_streams;s++) {
    int t;
    for (t=0;t<n_steps;t++)
      rade_core_decoder(&st[s], &m->dec, &features[((size_t)s*n_steps+t)*out_dim],
                        &z[((size_t)s*n_steps+t)*RADE_LATENT_DIM], 0);
  }
}

REF_API double ref_core_max_abs_acc(int reset) { return oracle_nnet_max_abs_acc(reset); }
