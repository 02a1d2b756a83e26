// Dependent chain of int8 GRU layers on tcgen05 with cross-layer issue-ahead (round-2 plan, DESIGN.md §8.1 "Orchestration"):
// measures what ONE layer of the codec's chain costs when the MMA issuer runs ahead of the epilogue.
//
// L = 3 DenseNet-style GRU layers for TS = 8 streams: layer l reads the whole concat prefix [x0 | out_0 | .. | out_{l-1}]
// (K = 64 (l + 1)) plus its own recurrent state, and appends its 64 outputs.  Roles: warp 3 lane 0 issues every MMA; warps 0-1
// (TMEM lanes 0..63 = hidden units) run the float epilogue and write the quantised outputs straight into the concat buffer,
// which IS the B operand of the following layers.  Only the last two k-blocks of a layer depend on the layer before it, so the
// issuer pushes the recurrent product and all older k-blocks of layer l while the epilogue of layer l - 1 is still running,
// waits on act_ready[l - 1], issues the two fresh k-blocks and commits acc_full[l].  Checked bit for bit against a CPU
// restatement (oracle/nnet_shim.c GRU math); prints cycles per layer.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -ffp-contract=off -o umma_gru_chain umma_gru_chain.cu
//   timeout 60 ./umma_gru_chain
// Written without GPU time left in round 1: compiled, not yet run.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int TS = 8, UNITS = 64, L = 3, K_REC = UNITS, ROWS = 3 * UNITS;
constexpr int ROWS_ALLOC = 2 * UNITS + 128;       // the third overlapping tile reads 128 rows from row 2*UNITS
constexpr int KCAT = UNITS * (L + 1);             // concat buffer: x0 (64) + one 64-wide output per layer
constexpr int TMEM_COLS = 256;                    // 48 columns per layer
__host__ __device__ constexpr int k_in(int l) { return UNITS * (l + 1); }
__host__ __device__ constexpr int wi_off(int l) { return ROWS_ALLOC * UNITS * (l * (l + 1) / 2); }   // bytes before layer l's input matrix

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

struct ChainArgs {
  const int8_t *Wi[L], *Wr[L];            // row-major [ROWS][k_in(l)], [ROWS][K_REC]
  const float *si[L], *bi[L], *sr[L], *br[L];
  const float *x0;                        // [TS][UNITS] input features of the step (already in [-1, 1])
  float *h;                               // [L][TS][UNITS] states, updated in place
  int8_t *cat;                            // [TS][KCAT] quantised concat buffer after the step (row-major copy for the check)
  long long *cyc;
};

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int cat_off(int s, int k) { return (k / 16) * 128 + s * 16 + k % 16; }      // B operand layout for 8 streams

__global__ void __launch_bounds__(128, 1) gru_chain_umma(ChainArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sWi = smem;                                           // layer l at wi_off(l): ROWS_ALLOC x k_in(l), canonical
  uint8_t *sWr = sWi + wi_off(L);                                // L x (ROWS_ALLOC x K_REC)
  uint8_t *sC = sWr + L * ROWS_ALLOC * K_REC;                    // TS x KCAT concat buffer (B operand)
  uint8_t *sH = sC + TS * KCAT;                                  // L x (TS x K_REC) quantised states of the previous step
  __shared__ __align__(8) uint64_t acc_full[L], act_ready[L];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

  for (int l = 0; l < L; l++) {
    const int K = k_in(l);
    for (int i = tid; i < ROWS_ALLOC * K; i += 128) { int r = i / K, b = i % K; sWi[wi_off(l) + canon(r, b, K)] = r < ROWS ? (uint8_t)a.Wi[l][i] : 0; }
    for (int i = tid; i < ROWS_ALLOC * K_REC; i += 128) { int r = i / K_REC, b = i % K_REC; sWr[l * ROWS_ALLOC * K_REC + canon(r, b, K_REC)] = r < ROWS ? (uint8_t)a.Wr[l][i] : 0; }
    for (int i = tid; i < TS * K_REC; i += 128) { int s = i / K_REC, k = i % K_REC; sH[l * TS * K_REC + cat_off(s, k)] = (uint8_t)quant8(a.h[(l * TS + s) * UNITS + k]); }
  }
  for (int i = tid; i < TS * KCAT; i += 128) sC[i] = 0;
  __syncthreads();
  for (int i = tid; i < TS * UNITS; i += 128) { int s = i / UNITS, k = i % UNITS; sC[cat_off(s, k)] = (uint8_t)quant8(a.x0[i]); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    for (int l = 0; l < L; l++) { mbar_init(&acc_full[l], 1); mbar_init(&act_ready[l], UNITS); }
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
  const long long t0 = clock64();

  if (warp == 3) {
    // ===== MMA issuer
    if (lane == 0) {
      for (int l = 0; l < L; l++) {
        const int K = k_in(l), nkb = K / 32, fresh = l == 0 ? 0 : 2;     // the last two k-blocks are out_{l-1}
        const uint32_t wi = smem_u32(sWi) + wi_off(l), wr = smem_u32(sWr) + l * ROWS_ALLOC * K_REC;
        const uint32_t col = tmem + l * 6 * TS;
        for (int g = 0; g < 3; g++) {                                    // runs ahead of the epilogue of layer l - 1
          const uint32_t ai = wi + (g * UNITS / 8) * (K * 8), ar = wr + (g * UNITS / 8) * (K_REC * 8);
          for (int k = 0; k < K_REC / 32; k++)
            umma_i8(col + (3 + g) * TS, smem_desc(ar + k * 256, 128, K_REC * 8), smem_desc(smem_u32(sH) + l * TS * K_REC + k * 256, 128, K_REC * 8), idesc, k != 0);
          for (int k = 0; k < nkb - fresh; k++)
            umma_i8(col + g * TS, smem_desc(ai + k * 256, 128, K * 8), smem_desc(smem_u32(sC) + k * 256, 128, KCAT * 8), idesc, k != 0);
        }
        if (fresh) {
          mbar_wait(&act_ready[l - 1], 0);                               // out_{l-1} is in the concat buffer (and fenced)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int g = 0; g < 3; g++) {
            const uint32_t ai = wi + (g * UNITS / 8) * (K * 8);
            for (int k = nkb - fresh; k < nkb; k++)
              umma_i8(col + g * TS, smem_desc(ai + k * 256, 128, K * 8), smem_desc(smem_u32(sC) + k * 256, 128, KCAT * 8), idesc, 1);
          }
        }
        umma_commit(&acc_full[l]);
      }
    }
  } else if (warp < UNITS / 32) {
    // ===== epilogue: one thread per hidden unit
    const int u = tid;
    for (int l = 0; l < L; l++) {
      float si[3], bi[3], sr[3], br[3];
#pragma unroll
      for (int g = 0; g < 3; g++) { si[g] = a.si[l][g * UNITS + u]; bi[g] = a.bi[l][g * UNITS + u]; sr[g] = a.sr[l][g * UNITS + u]; br[g] = a.br[l][g * UNITS + u]; }
      float hold[TS];
#pragma unroll
      for (int s = 0; s < TS; s++) hold[s] = a.h[(l * TS + s) * UNITS + u];
      mbar_wait(&acc_full[l], 0);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int acc[6][8];
#pragma unroll
      for (int q = 0; q < 6; q++) tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + l * 6 * TS + q * TS, acc[q]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int kout = UNITS * (l + 1) + u;
#pragma unroll
      for (int s = 0; s < TS; s++) {
        float z = sigmoid_r(__fadd_rn(lin(acc[0][s], si[0], bi[0]), lin(acc[3][s], sr[0], br[0])));
        float r = sigmoid_r(__fadd_rn(lin(acc[1][s], si[1], bi[1]), lin(acc[4][s], sr[1], br[1])));
        float n = tanh_r(__fadd_rn(lin(acc[2][s], si[2], bi[2]), __fmul_rn(lin(acc[5][s], sr[2], br[2]), r)));
        float h = __fadd_rn(__fmul_rn(z, hold[s]), __fmul_rn(__fsub_rn(1.f, z), n));
        a.h[(l * TS + s) * UNITS + u] = h;
        sC[cat_off(s, kout)] = (uint8_t)quant8(h);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // my stores -> visible to the tensor core's proxy
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&act_ready[l]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) *a.cyc = clock64() - t0;
  for (int i = tid; i < TS * KCAT; i += 128) { int s = i / KCAT, k = i % KCAT; a.cat[i] = (int8_t)sC[cat_off(s, k)]; }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
template <typename T> static T *to_dev(const std::vector<T> &v) { T *d = nullptr; cudaMalloc(&d, v.size() * sizeof(T)); cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice); return d; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  if (prop.major != 10) { printf("needs sm_100 (found sm_%d%d)\n", prop.major, prop.minor); return 1; }
  srand(5);
  auto frand = [](float lo, float hi) { return lo + (hi - lo) * (float)rand() / (float)RAND_MAX; };
  std::vector<int8_t> Wi[L], Wr[L]; std::vector<float> si[L], bi[L], sr[L], br[L];
  for (int l = 0; l < L; l++) {
    Wi[l].resize(ROWS * k_in(l)); Wr[l].resize(ROWS * K_REC); si[l].resize(ROWS); bi[l].resize(ROWS); sr[l].resize(ROWS); br[l].resize(ROWS);
    for (auto &w : Wi[l]) w = (int8_t)(rand() % 255 - 127);
    for (auto &w : Wr[l]) w = (int8_t)(rand() % 255 - 127);
    for (int i = 0; i < ROWS; i++) { si[l][i] = frand(2e-5f, 9e-5f); sr[l][i] = frand(2e-5f, 9e-5f); bi[l][i] = frand(-.5f, .5f); br[l][i] = frand(-.5f, .5f); }
  }
  std::vector<float> x0(TS * UNITS), h(L * TS * UNITS);
  for (auto &v : x0) v = frand(-1.f, 1.f);
  for (auto &v : h) v = frand(-1.f, 1.f);

  // CPU restatement: layer after layer over the quantised concat buffer
  std::vector<float> h_ref(h); std::vector<int8_t> cat_ref(TS * KCAT, 0);
  for (int s = 0; s < TS; s++) {
    int8_t *cat = &cat_ref[s * KCAT];
    for (int k = 0; k < UNITS; k++) cat[k] = quant8(x0[s * UNITS + k]);
    for (int l = 0; l < L; l++) {
      const int K = k_in(l);
      int8_t hq[K_REC];
      for (int k = 0; k < K_REC; k++) hq[k] = quant8(h[(l * TS + s) * UNITS + k]);
      float g[ROWS], rec[ROWS];
      for (int o = 0; o < ROWS; o++) {
        int ai = 0, ar = 0;
        for (int k = 0; k < K; k++) ai += (int)Wi[l][o * K + k] * cat[k];
        for (int k = 0; k < K_REC; k++) ar += (int)Wr[l][o * K_REC + k] * hq[k];
        float t = (float)ai * si[l][o]; g[o] = t + bi[l][o];
        float v = (float)ar * sr[l][o]; rec[o] = v + br[l][o];
      }
      for (int u = 0; u < UNITS; u++) {
        float zs = g[u] + rec[u], rs = g[UNITS + u] + rec[UNITS + u];
        float z = sigmoid_r(zs), r = sigmoid_r(rs);
        float m = rec[2 * UNITS + u] * r; float ns = g[2 * UNITS + u] + m;
        float n = tanh_r(ns);
        float p = z * h[(l * TS + s) * UNITS + u], q = (1 - z) * n;
        float hn = p + q;
        h_ref[(l * TS + s) * UNITS + u] = hn;
        cat[UNITS * (l + 1) + u] = quant8(hn);
      }
    }
  }

  ChainArgs a;
  for (int l = 0; l < L; l++) { a.Wi[l] = to_dev(Wi[l]); a.Wr[l] = to_dev(Wr[l]); a.si[l] = to_dev(si[l]); a.bi[l] = to_dev(bi[l]); a.sr[l] = to_dev(sr[l]); a.br[l] = to_dev(br[l]); }
  a.x0 = to_dev(x0); a.h = to_dev(h);
  CK(cudaMalloc(&a.cat, TS * KCAT)); CK(cudaMalloc(&a.cyc, 8));
  const int smem_bytes = wi_off(L) + L * ROWS_ALLOC * K_REC + TS * KCAT + L * TS * K_REC;
  CK(cudaFuncSetAttribute(gru_chain_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  gru_chain_umma<<<1, 128, smem_bytes>>>(a);
  CK(cudaDeviceSynchronize());
  std::vector<float> h_out(h.size()); std::vector<int8_t> cat_out(TS * KCAT); long long cyc;
  CK(cudaMemcpy(h_out.data(), a.h, h.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cat_out.data(), a.cat, TS * KCAT, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, a.cyc, 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (size_t i = 0; i < h.size(); i++) bad += h_out[i] != h_ref[i];
  for (size_t i = 0; i < cat_out.size(); i++) bad += cat_out[i] != cat_ref[i];
  printf("GRU chain (%d layers, %d units, %d streams, issue-ahead): %s (%d mismatches), %lld cycles total = %.0f per layer (%.2f us at 1.965 GHz)\n",
         L, UNITS, TS, bad ? "WRONG" : "exact", bad, cyc, (double)cyc / L, (double)cyc / L / 1965.0);
  return bad != 0;
}
