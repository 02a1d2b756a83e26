/* Stand-in for the reference host src/radae_tx.c where /root/reference is absent (GPU box): same call sequence
 * against include/rade_api.h — features.f32 on stdin, IQ complex64 on stdout, EOO frame at the end. */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include "rade_api.h"

int main(void) {
  rade_initialize();
  struct rade *r = rade_open("dummy", RADE_USE_C_ENCODER | RADE_VERBOSE_0);
  assert(r != NULL);
  int nf = rade_n_features_in_out(r), nt = rade_n_tx_out(r), ne = rade_n_tx_eoo_out(r);
  float *features = malloc(sizeof(float) * nf);
  RADE_COMP *tx = malloc(sizeof(RADE_COMP) * nt), *eoo = malloc(sizeof(RADE_COMP) * ne);
  while (fread(features, sizeof(float), nf, stdin) == (size_t)nf) {
    rade_tx(r, tx, features);
    fwrite(tx, sizeof(RADE_COMP), nt, stdout);
    fflush(stdout);
  }
  rade_tx_eoo(r, eoo);
  fwrite(eoo, sizeof(RADE_COMP), ne, stdout);
  rade_close(r);
  rade_finalize();
  return 0;
}
