/* C host for the multi-device context (include/rade_b200.h: rade_b200_open_multi): N streams split over every visible GPU
 * (or over `ndev` contexts on the listed devices), each frame transmitted with rade_b200_multi_tx and fed straight back into
 * rade_b200_multi_rx; the same frames also go through ONE single-device context and the two must agree bit for bit.
 * usage: multi_loopback n_streams n_frames [device device ...]   -> prints "valid <count> mismatches <count> devices <n>" */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rade_b200.h"

int main(int argc, char **argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 12, F = argc > 2 ? atoi(argv[2]) : 12;
  int devices[16], nd = 0;
  for (int i = 3; i < argc && nd < 16; i++) devices[nd++] = atoi(argv[i]);
  rade_multi *m = nd ? rade_b200_open_devices(S, devices, nd, RADE_USE_C_ENCODER | RADE_USE_C_DECODER, NULL, 0)
                     : rade_b200_open_multi(S, 0, RADE_USE_C_ENCODER | RADE_USE_C_DECODER, NULL, 0);
  rade_batch *b = rade_b200_open(S, 0, RADE_USE_C_ENCODER | RADE_USE_C_DECODER, NULL, 0);
  if (!m || !b) { fprintf(stderr, "open failed\n"); return 2; }
  float *feat = calloc((size_t)S * 432, sizeof(float)), *fo_m = calloc((size_t)S * 432, sizeof(float)), *fo_b = calloc((size_t)S * 432, sizeof(float));
  float *eoo_m = calloc((size_t)S * 180, sizeof(float)), *eoo_b = calloc((size_t)S * 180, sizeof(float));
  RADE_COMP *tx_m = calloc((size_t)S * 960, sizeof(RADE_COMP)), *tx_b = calloc((size_t)S * 960, sizeof(RADE_COMP));
  RADE_COMP *rx = calloc((size_t)S * 1120, sizeof(RADE_COMP));
  int *nin_m = calloc(S, sizeof(int)), *nin_b = calloc(S, sizeof(int)), *ret_m = calloc(S, sizeof(int)), *ret_b = calloc(S, sizeof(int));
  long valid = 0, mism = 0;
  unsigned lcg = 12345u;
  for (int k = 0; k < F; k++) {
    for (int i = 0; i < S * 432; i++) { lcg = lcg * 1664525u + 1013904223u; feat[i] = (i % 36 < 20) ? ((float)(lcg >> 8) / 8388608.0f - 1.0f) : 0.f; }
    if (rade_b200_multi_tx(m, tx_m, feat) < 0 || rade_b200_tx(b, tx_b, feat) < 0) return 3;
    mism += memcmp(tx_m, tx_b, (size_t)S * 960 * sizeof(RADE_COMP)) != 0;
    rade_b200_multi_nin(m, nin_m); rade_b200_nin(b, nin_b);
    mism += memcmp(nin_m, nin_b, S * sizeof(int)) != 0;
    for (int s = 0; s < S; s++)                                  /* 960 samples per frame; nin stays 960 on a clean loop-back */
      for (int i = 0; i < nin_b[s] && i < 1120; i++) rx[(size_t)s * 1120 + i] = tx_b[(size_t)s * 960 + i % 960];
    if (rade_b200_multi_rx(m, fo_m, ret_m, eoo_m, rx, NULL) < 0 || rade_b200_rx(b, fo_b, ret_b, eoo_b, rx, NULL) < 0) return 4;
    mism += memcmp(ret_m, ret_b, S * sizeof(int)) != 0;
    for (int s = 0; s < S; s++)
      if (ret_b[s] & 1) { valid++; mism += memcmp(fo_m + (size_t)s * 432, fo_b + (size_t)s * 432, 432 * sizeof(float)) != 0; }
  }
  printf("valid %ld mismatches %ld devices %d\n", valid, mism, rade_b200_multi_n_devices(m));
  rade_b200_close_multi(m); rade_b200_close(b);
  return mism ? 1 : 0;
}
