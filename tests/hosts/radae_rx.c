/* Stand-in for the reference host src/radae_rx.c where /root/reference is absent (GPU box): same call sequence
 * against include/rade_api.h — IQ complex64 on stdin (nin samples per call), features.f32 on stdout, eoo_rx.f32. */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include "rade_api.h"

int main(void) {
  rade_initialize();
  struct rade *r = rade_open("dummy", RADE_USE_C_DECODER | RADE_VERBOSE_0);
  assert(r != NULL);
  int nf = rade_n_features_in_out(r), nmax = rade_nin_max(r), nb = rade_n_eoo_bits(r);
  float *features = malloc(sizeof(float) * nf), *eoo = malloc(sizeof(float) * nb);
  RADE_COMP *rx = malloc(sizeof(RADE_COMP) * nmax);
  FILE *feoo = fopen("eoo_rx.f32", "wb"); assert(feoo != NULL);
  int nin = rade_nin(r), has_eoo;
  while ((size_t)nin == fread(rx, sizeof(RADE_COMP), nin, stdin)) {
    int n = rade_rx(r, features, &has_eoo, eoo, rx);
    if (n) { fwrite(features, sizeof(float), nf, stdout); fflush(stdout); }
    if (has_eoo) fwrite(eoo, sizeof(float), nb, feoo);
    nin = rade_nin(r);
  }
  rade_close(r);
  rade_finalize();
  fclose(feoo);
  return 0;
}
