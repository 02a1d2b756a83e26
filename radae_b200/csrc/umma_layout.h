// Host / device helpers for the tcgen05 operand layout the round-2 codec plan uses: both operands K-major,
// no swizzle, 8-row x 16-byte core matrices:
//     byte offset of (row r, k byte b) inside an operand image whose rows hold `kbytes` bytes of K
//         = (r / 8) * (kbytes * 8) + (b / 16) * 128 + (r % 8) * 16 + b % 16          (SBO = kbytes * 8, LBO = 128)
// Header-only so that the weight pre-bake can move into radae_b200/csrc/weights.cpp unchanged once the kernels exist;
// test_umma_layout.cpp checks it on the CPU (pytest: tests/test_umma_addressing.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(__CUDACC__)
#define UMMA_HD __host__ __device__
#else
#define UMMA_HD
#endif

UMMA_HD constexpr int umma_canon(int r, int b, int kbytes) { return (r / 8) * (kbytes * 8) + (b / 16) * 128 + (r % 8) * 16 + b % 16; }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46),
// version 1 at [46,48), layout type SWIZZLE_NONE
UMMA_HD constexpr uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (uint64_t)(lbo_bytes >> 4) << 16 | (uint64_t)(sbo_bytes >> 4) << 32 | (uint64_t)1 << 46;
}
// instruction descriptors (cute::UMMA::InstrDescriptor), K-major A and B
UMMA_HD constexpr uint32_t umma_idesc_i8(int m, int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
UMMA_HD constexpr uint32_t umma_idesc_tf32(int m, int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

// rows an M = 128 tile may touch when the tiles of a layer start at rows 0, step, 2 step, ... (GRU: step = units, 3 tiles;
// conv / GLU: one tile): the ring stage has to span that many rows even though only `n_rows` are copied
constexpr int umma_span_rows(int n_rows, int tile_step, int n_tiles) { int s = tile_step * (n_tiles - 1) + 128; return s > n_rows ? s : n_rows; }
// k-blocks (32 bytes of K each) of a layer that fit one ring stage
constexpr int umma_kblocks_per_stage(int span_rows, int stage_bytes) { return stage_bytes / (span_rows * 32); }

// One weight chunk: k-blocks [kb0, kb0 + nk) of ALL rows of a row-major int8 [n_rows][K] matrix, canonical order with
// kbytes = nk * 32.  Returns the n_rows * nk * 32 bytes the producer copies into a stage (a multiple of 16).
inline std::vector<uint8_t> umma_bake_chunk(const int8_t *W, int n_rows, int K, int kb0, int nk) {
  const int kbytes = nk * 32;
  std::vector<uint8_t> out((size_t)n_rows * kbytes);
  for (int r = 0; r < n_rows; r++)
    for (int b = 0; b < kbytes; b += 16)
      memcpy(&out[umma_canon(r, b, kbytes)], W + (size_t)r * K + kb0 * 32 + b, 16);
  return out;
}
