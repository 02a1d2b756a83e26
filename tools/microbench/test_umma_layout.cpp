// CPU check of umma_layout.h: bake real-shaped int8 layers into ring-stage chunks and read them back the way the planned kernel's
// descriptors would (overlapping M = 128 gate tiles, stale bytes behind the copied rows), against a plain integer GEMM.
//   g++ -O2 -std=c++17 -o test_umma_layout test_umma_layout.cpp && ./test_umma_layout
#include <cstdio>
#include <cstdlib>
#include "../../radae_b200/csrc/umma_layout.h"

static int check_layer(const char *name, int n_rows, int K, int tile_step, int n_tiles, int n_streams) {
  const int STAGE = 32768;
  const int span = umma_span_rows(n_rows, tile_step, n_tiles), nk_max = umma_kblocks_per_stage(span, STAGE);
  if (nk_max < 1 || span * nk_max * 32 > STAGE) { printf("%s: stage sizing broken\n", name); return 1; }
  std::vector<int8_t> W((size_t)n_rows * K), X((size_t)n_streams * K);
  for (auto &w : W) w = (int8_t)(rand() % 255 - 127);
  for (auto &x : X) x = (int8_t)(rand() % 255 - 127);
  std::vector<long> D((size_t)n_tiles * 128 * n_streams, 0);
  std::vector<uint8_t> stage(STAGE, 0x5A);                        // stale bytes from "earlier chunks"
  int chunks = 0;
  for (int kb0 = 0; kb0 < K / 32; kb0 += nk_max, chunks++) {
    const int nk = (K / 32 - kb0) < nk_max ? (K / 32 - kb0) : nk_max, kbytes = nk * 32;
    std::vector<uint8_t> c = umma_bake_chunk(W.data(), n_rows, K, kb0, nk);
    if (c.size() % 16 || c.size() > (size_t)STAGE) { printf("%s: bad chunk size\n", name); return 1; }
    memcpy(stage.data(), c.data(), c.size());                     // what cp.async.bulk does
    for (int t = 0; t < n_tiles; t++)
      for (int k = 0; k < nk; k++) {
        const int start = (t * tile_step / 8) * (kbytes * 8) + k * 256, sbo = kbytes * 8, lbo = 128;
        for (int r = 0; r < 128; r++) {
          if (start + (r / 8) * sbo + lbo + (r % 8) * 16 + 15 >= STAGE) { printf("%s: tile reads past the stage\n", name); return 1; }
          for (int s = 0; s < n_streams; s++) {
            long acc = 0;
            for (int b = 0; b < 32; b++)
              acc += (long)(int8_t)stage[start + (r / 8) * sbo + (b / 16) * lbo + (r % 8) * 16 + b % 16] * X[(size_t)s * K + (kb0 + k) * 32 + b];
            D[((size_t)t * 128 + r) * n_streams + s] += acc;
          }
        }
      }
  }
  // lanes [0, tile_step) (or all real rows for a single tile) of tile t must equal rows t * tile_step + lane of W X^T
  int bad = 0;
  for (int t = 0; t < n_tiles; t++) {
    const int lanes = n_tiles > 1 ? tile_step : n_rows;
    for (int r = 0; r < lanes; r++)
      for (int s = 0; s < n_streams; s++) {
        long ref = 0;
        for (int k = 0; k < K; k++) ref += (long)W[(size_t)(t * tile_step + r) * K + k] * X[(size_t)s * K + k];
        bad += ref != D[((size_t)t * 128 + r) * n_streams + s];
      }
  }
  printf("%-22s rows %3d K %4d: span %3d rows, %d k-blocks per stage, %2d chunks -> %s\n", name, n_rows, K, span, nk_max, chunks, bad ? "WRONG" : "ok");
  return bad != 0;
}

int main() {
  srand(1);
  int bad = 0;
  bad += check_layer("enc_gru5_input", 192, 704, 64, 3, 8);
  bad += check_layer("enc_gru_recurrent", 192, 64, 64, 3, 8);
  bad += check_layer("dec_gru5_input", 288, 608, 96, 3, 8);
  bad += check_layer("enc_conv5 (one tap)", 96, 768, 0, 1, 8);
  bad += check_layer("dec_conv5 (one tap)", 32, 704, 0, 1, 8);
  bad += check_layer("dec_glu", 96, 96, 0, 1, 8);
  bad += check_layer("enc_gru5, 16 streams", 192, 704, 64, 3, 16);
  printf("%s\n", bad ? "FAILED" : "all layouts ok");
  return bad;
}
