// Prototype of ONE int8 GRU layer of the core codec on tcgen05 (round-2 plan), self-checking.
//
//   gates = Wi x_q * scale_i + bias_i,  rec = Wr h_q * scale_r + bias_r,   z = sig(g_z + rec_z), r = sig(g_r + rec_r),
//   n = tanh(g_n + rec_n * r),  h' = z h + (1 - z) n                        (oracle/nnet_shim.c compute_generic_gru,
//                                                                            /root/reference/src/rade_enc.c:72-73 per layer)
// for TS = 8 streams per CTA, UNITS = 64 hidden units, K_IN inputs.  Operands swapped: A = weight rows (TMEM lanes = hidden
// unit), B = quantised activations of the 8 streams (TMEM columns), no-swizzle K-major canonical layouts (see
// umma_i8_swapab.cu).  The three gates of a unit must meet in one thread, so three OVERLAPPING M = 128 tiles are issued, starting
// at weight rows 0, UNITS and 2*UNITS: lanes 0..UNITS-1 of column block g then hold gate g (the upper lanes hold whatever rows
// follow in memory and are never read) -- no padding rows, no extra weight bytes.  Input and recurrent matrices keep separate
// accumulators (different per-output scales): 6 blocks of 8 columns.  The epilogue is one thread per hidden unit: its scales
// and biases are scalars, the 8 streams sit in its registers, float ops in the oracle's order (bit-exact contract).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -ffp-contract=off -o umma_gru_layer umma_gru_layer.cu
//   timeout 60 ./umma_gru_layer
// Prints "GRU layer: exact" (h' and its int8 image identical to the CPU restatement for every stream / unit) and the
// cycles from first MMA issue to the last h' written.  Written without GPU time left in round 1: compiled, not yet run.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int TS = 8, UNITS = 64, K_IN = 224, K_REC = UNITS, ROWS = 3 * UNITS;
constexpr int ROWS_ALLOC = 2 * UNITS + 128;       // the third overlapping tile reads 128 rows from row 2*UNITS
constexpr int TMEM_COLS = 64;                     // 6 blocks x 8 columns -> next power of two

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ constexpr int canon(int r, int b, int k_total) { return (r / 8) * (k_total * 8) + (b / 16) * 128 + (r % 8) * 16 + b % 16; }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(lbo_bytes >> 4) << 16 | (uint64_t)(sbo_bytes >> 4) << 32 | (uint64_t)1 << 46;
}
__host__ __device__ constexpr uint32_t instr_desc_i8(int m, int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; !ok; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spin > (1ll << 24)) __trap();
  }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}

// scalar math, bit-exact w.r.t. oracle/nnet_shim.c (same helpers as radae_b200/csrc/core_codec.cu)
__host__ __device__ inline float tanh_r(float x) {
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f, D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
#ifdef __CUDA_ARCH__
  float x2 = __fmul_rn(x, x);
  float num = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(N2, x2), N1), x2), N0);
  float den = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(D2, x2), D1), x2), D0);
  float y = __fdiv_rn(__fmul_rn(num, x), den);
#else
  float x2 = x * x;
  float num = (N2 * x2 + N1) * x2 + N0;
  float den = (D2 * x2 + D1) * x2 + D0;
  float y = num * x / den;
#endif
  return y > 1.f ? 1.f : (y < -1.f ? -1.f : y);
}
__host__ __device__ inline float sigmoid_r(float x) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(.5f, __fmul_rn(.5f, tanh_r(__fmul_rn(.5f, x))));
#else
  return .5f + .5f * tanh_r(.5f * x);
#endif
}
__host__ __device__ inline int8_t quant8(float x) {
#ifdef __CUDA_ARCH__
  return (int8_t)__double2int_rd((double)__fmul_rn(127.f, x) + 0.5);
#else
  float p = 127.f * x;
  return (int8_t)(int)floor(.5 + (double)p);
#endif
}
__device__ __forceinline__ float lin(int acc, float scale, float bias) { return __fadd_rn(__fmul_rn((float)acc, scale), bias); }

struct LayerArgs {
  const int8_t *Wi, *Wr;                 // row-major [ROWS][K_IN], [ROWS][K_REC] (rows: z | r | n, src: wexchange GRU export order)
  const float *si, *bi, *sr, *br;        // [ROWS] each
  const float *x;                        // [TS][K_IN] layer input (already in [-1, 1])
  float *h;                              // [TS][UNITS] state, updated in place
  int8_t *hq;                            // [TS][UNITS] quantised new state (what the next layer's B operand is built from)
  long long *cyc;
};

__global__ void __launch_bounds__(128, 1) gru_layer_umma(LayerArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sWi = smem;                                   // ROWS_ALLOC x K_IN, canonical
  uint8_t *sWr = sWi + ROWS_ALLOC * K_IN;                // ROWS_ALLOC x K_REC
  uint8_t *sX = sWr + ROWS_ALLOC * K_REC;                // TS x K_IN
  uint8_t *sH = sX + TS * K_IN;                          // TS x K_REC
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;

  // stage operands (the real kernel streams pre-baked weight chunks with cp.async.bulk; rows >= ROWS are never read back)
  for (int i = tid; i < ROWS_ALLOC * K_IN; i += 128) { int r = i / K_IN, b = i % K_IN; sWi[canon(r, b, K_IN)] = r < ROWS ? (uint8_t)a.Wi[i] : 0; }
  for (int i = tid; i < ROWS_ALLOC * K_REC; i += 128) { int r = i / K_REC, b = i % K_REC; sWr[canon(r, b, K_REC)] = r < ROWS ? (uint8_t)a.Wr[i] : 0; }
  for (int i = tid; i < TS * K_IN; i += 128) { int s = i / K_IN, k = i % K_IN; sX[canon(s, k, K_IN)] = (uint8_t)quant8(a.x[i]); }
  for (int i = tid; i < TS * K_REC; i += 128) { int s = i / K_REC, k = i % K_REC; sH[canon(s, k, K_REC)] = (uint8_t)quant8(a.h[i]); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = instr_desc_i8(128, TS);

  long long t0 = clock64();
  if (tid == 0) {
    // column block g: input part of gate g; block 3 + g: recurrent part.  A tile g starts at weight row g * UNITS.
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const uint32_t ai = smem_u32(sWi) + (g * UNITS / 8) * (K_IN * 8), ar = smem_u32(sWr) + (g * UNITS / 8) * (K_REC * 8);
      for (int k = 0; k < K_IN / 32; k++)
        umma_i8(tmem + g * TS, smem_desc(ai + k * 256, 128, K_IN * 8), smem_desc(smem_u32(sX) + k * 256, 128, K_IN * 8), idesc, k != 0);
      for (int k = 0; k < K_REC / 32; k++)
        umma_i8(tmem + (3 + g) * TS, smem_desc(ar + k * 256, 128, K_REC * 8), smem_desc(smem_u32(sH) + k * 256, 128, K_REC * 8), idesc, k != 0);
    }
    umma_commit(&bar);
  }
  // epilogue: warps 0 and 1 own hidden units 0..63 (TMEM lanes 0..63); prefetch their scalars while the MMAs run
  const int u = tid;
  float si[3], bi[3], sr[3], br[3];
  if (u < UNITS)
#pragma unroll
    for (int g = 0; g < 3; g++) { si[g] = a.si[g * UNITS + u]; bi[g] = a.bi[g * UNITS + u]; sr[g] = a.sr[g * UNITS + u]; br[g] = a.br[g * UNITS + u]; }
  if (warp < UNITS / 32) {
    mbar_wait(&bar, 0);
    __syncwarp();                                        // tcgen05.ld is warp-collective (.sync.aligned)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int acc[6][8];
#pragma unroll
    for (int q = 0; q < 6; q++) tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + q * TS, acc[q]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int s = 0; s < TS; s++) {
      float z = sigmoid_r(__fadd_rn(lin(acc[0][s], si[0], bi[0]), lin(acc[3][s], sr[0], br[0])));
      float r = sigmoid_r(__fadd_rn(lin(acc[1][s], si[1], bi[1]), lin(acc[4][s], sr[1], br[1])));
      float n = tanh_r(__fadd_rn(lin(acc[2][s], si[2], bi[2]), __fmul_rn(lin(acc[5][s], sr[2], br[2]), r)));
      float hold = a.h[s * UNITS + u];
      float h = __fadd_rn(__fmul_rn(z, hold), __fmul_rn(__fsub_rn(1.f, z), n));
      a.h[s * UNITS + u] = h;
      a.hq[s * UNITS + u] = quant8(h);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) *a.cyc = clock64() - t0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  if (prop.major != 10) { printf("needs sm_100 (found sm_%d%d)\n", prop.major, prop.minor); return 1; }
  srand(7);
  auto frand = [](float lo, float hi) { return lo + (hi - lo) * (float)rand() / (float)RAND_MAX; };
  std::vector<int8_t> Wi(ROWS * K_IN), Wr(ROWS * K_REC);
  std::vector<float> si(ROWS), bi(ROWS), sr(ROWS), br(ROWS), x(TS * K_IN), h(TS * UNITS);
  for (auto &w : Wi) w = (int8_t)(rand() % 255 - 127);
  for (auto &w : Wr) w = (int8_t)(rand() % 255 - 127);
  for (int i = 0; i < ROWS; i++) { si[i] = frand(2e-5f, 9e-5f); sr[i] = frand(2e-5f, 9e-5f); bi[i] = frand(-.5f, .5f); br[i] = frand(-.5f, .5f); }
  for (auto &v : x) v = frand(-1.f, 1.f);
  for (auto &v : h) v = frand(-1.f, 1.f);

  // CPU restatement (oracle/nnet_shim.c order of operations; build with -ffp-contract=off)
  std::vector<float> h_ref(h); std::vector<int8_t> hq_ref(TS * UNITS);
  for (int s = 0; s < TS; s++) {
    std::vector<int8_t> xq(K_IN), hq(K_REC);
    for (int k = 0; k < K_IN; k++) xq[k] = quant8(x[s * K_IN + k]);
    for (int k = 0; k < K_REC; k++) hq[k] = quant8(h[s * UNITS + k]);
    std::vector<float> g(ROWS), rec(ROWS);
    for (int o = 0; o < ROWS; o++) {
      int ai = 0, ar = 0;
      for (int k = 0; k < K_IN; k++) ai += (int)Wi[o * K_IN + k] * xq[k];
      for (int k = 0; k < K_REC; k++) ar += (int)Wr[o * K_REC + k] * hq[k];
      float t = (float)ai * si[o]; g[o] = t + bi[o];
      float v = (float)ar * sr[o]; rec[o] = v + br[o];
    }
    for (int u = 0; u < UNITS; u++) {
      float zs = g[u] + rec[u], rs = g[UNITS + u] + rec[UNITS + u];
      float z = sigmoid_r(zs), r = sigmoid_r(rs);
      float m = rec[2 * UNITS + u] * r; float ns = g[2 * UNITS + u] + m;
      float n = tanh_r(ns);
      float p = z * h[s * UNITS + u], q = (1 - z) * n;
      h_ref[s * UNITS + u] = p + q;
      hq_ref[s * UNITS + u] = quant8(h_ref[s * UNITS + u]);
    }
  }

  LayerArgs a; int8_t *dWi, *dWr, *dhq; float *dsi, *dbi, *dsr, *dbr, *dx, *dh; long long *dc;
  CK(cudaMalloc(&dWi, Wi.size())); CK(cudaMalloc(&dWr, Wr.size())); CK(cudaMalloc(&dhq, TS * UNITS));
  CK(cudaMalloc(&dsi, ROWS * 4)); CK(cudaMalloc(&dbi, ROWS * 4)); CK(cudaMalloc(&dsr, ROWS * 4)); CK(cudaMalloc(&dbr, ROWS * 4));
  CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dh, h.size() * 4)); CK(cudaMalloc(&dc, 8));
  CK(cudaMemcpy(dWi, Wi.data(), Wi.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dWr, Wr.data(), Wr.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsi, si.data(), ROWS * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbi, bi.data(), ROWS * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsr, sr.data(), ROWS * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbr, br.data(), ROWS * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dh, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  a = {dWi, dWr, dsi, dbi, dsr, dbr, dx, dh, dhq, dc};
  const int smem_bytes = ROWS_ALLOC * (K_IN + K_REC) + TS * (K_IN + K_REC);
  CK(cudaFuncSetAttribute(gru_layer_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  gru_layer_umma<<<1, 128, smem_bytes>>>(a);
  CK(cudaDeviceSynchronize());
  std::vector<float> h_out(h.size()); std::vector<int8_t> hq_out(TS * UNITS); long long cyc;
  CK(cudaMemcpy(h_out.data(), dh, h.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hq_out.data(), dhq, TS * UNITS, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (size_t i = 0; i < h.size(); i++) bad += (h_out[i] != h_ref[i]) || (hq_out[i] != hq_ref[i]);
  printf("GRU layer (%d units, K_in %d, %d streams): %s (%d of %zu differ), %lld cycles issue -> h' stored (%d MMAs)\n", UNITS, K_IN, TS,
         bad ? "WRONG" : "exact", bad, h.size(), cyc, 3 * (K_IN / 32 + K_REC / 32));
  return bad != 0;
}
