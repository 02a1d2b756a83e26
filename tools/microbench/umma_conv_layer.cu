// Prototype of ONE int8 conv1d (k = 2, causal, dilated) layer of the core codec on tcgen05 (round-2 plan),
// self-checking; companion of umma_gru_layer.cu.
//
//   y = tanh( (W [x_old ; x_cur]_q) * scale + bias )       (oracle/nnet_shim.c compute_generic_conv1d_dilation,
//                                                            /root/reference/src/rade_enc.c:76-104: tap 0 = the frame `dilation`
//                                                            steps back, tap 1 = the current frame)
// for TS = 8 streams, OUT = 96 outputs, KC inputs per tap.  A = the 96 weight rows (TMEM lanes 0..95 of one M = 128 tile; the
// upper 32 lanes read whatever follows in memory and are ignored), B = the quantised activations: the two taps live in two
// DIFFERENT concat buffers (previous and current frame), so the K loop simply switches the B descriptor half way while the
// accumulator chain continues -- no copy of the old frame.  Epilogue: one thread per output, tanh in the oracle's float order.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -ffp-contract=off -o umma_conv_layer umma_conv_layer.cu
//   timeout 60 ./umma_conv_layer
// Written without GPU time left in round 1: compiled, not yet run.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int TS = 8, OUT = 96, KC = 288, K = 2 * KC;     // enc_conv2: 2 taps x 288 inputs -> 96 outputs
constexpr int ROWS_ALLOC = 128;                   // one M = 128 tile; rows 96..127 are never read back
constexpr int TMEM_COLS = 32;

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
  const int8_t *W;                        // row-major [OUT][K]: columns [0, KC) multiply the OLD frame, [KC, 2 KC) the current one
  const float *scale, *bias;              // [OUT]
  const float *x_old, *x_cur;             // [TS][KC] each (already in [-1, 1])
  float *y;                               // [TS][OUT]
  int8_t *yq;                             // [TS][OUT] quantised outputs (appended to the concat buffer in the real kernel)
  long long *cyc;
};

__global__ void __launch_bounds__(128, 1) conv_layer_umma(LayerArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sW = smem;                                    // ROWS_ALLOC x K, canonical
  uint8_t *sOld = sW + ROWS_ALLOC * K;                   // TS x KC   (stands for the previous frame's concat buffer)
  uint8_t *sCur = sOld + TS * KC;                        // TS x KC   (the current frame's)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;

  for (int i = tid; i < ROWS_ALLOC * K; i += 128) { int r = i / K, b = i % K; sW[canon(r, b, K)] = r < OUT ? (uint8_t)a.W[i] : 0; }
  for (int i = tid; i < TS * KC; i += 128) {
    int s = i / KC, k = i % KC;
    sOld[canon(s, k, KC)] = (uint8_t)quant8(a.x_old[i]); sCur[canon(s, k, KC)] = (uint8_t)quant8(a.x_cur[i]);
  }
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
    for (int k = 0; k < K / 32; k++) {
      const bool old = k < KC / 32;                       // first half of K: the old frame's buffer
      const uint32_t b = (old ? smem_u32(sOld) : smem_u32(sCur)) + (old ? k : k - KC / 32) * 256;
      umma_i8(tmem, smem_desc(smem_u32(sW) + k * 256, 128, K * 8), smem_desc(b, 128, KC * 8), idesc, k != 0);
    }
    umma_commit(&bar);
  }
  const int o = tid;
  float sc = 0.f, bi = 0.f;
  if (o < OUT) { sc = a.scale[o]; bi = a.bias[o]; }
  if (warp < OUT / 32) {
    mbar_wait(&bar, 0);
    __syncwarp();                                        // tcgen05.ld is warp-collective (.sync.aligned)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int acc[8];
    tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16), acc);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int s = 0; s < TS; s++) {
      const float y = tanh_r(lin(acc[s], sc, bi));
      a.y[s * OUT + o] = y;
      a.yq[s * OUT + o] = quant8(y);
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
  srand(11);
  auto frand = [](float lo, float hi) { return lo + (hi - lo) * (float)rand() / (float)RAND_MAX; };
  std::vector<int8_t> W(OUT * K);
  std::vector<float> sc(OUT), bi(OUT), xo(TS * KC), xc(TS * KC);
  for (auto &w : W) w = (int8_t)(rand() % 255 - 127);
  for (int i = 0; i < OUT; i++) { sc[i] = frand(1e-5f, 5e-5f); bi[i] = frand(-.5f, .5f); }
  for (auto &v : xo) v = frand(-1.f, 1.f);
  for (auto &v : xc) v = frand(-1.f, 1.f);
  std::vector<float> y_ref(TS * OUT); std::vector<int8_t> yq_ref(TS * OUT);
  for (int s = 0; s < TS; s++) for (int o = 0; o < OUT; o++) {
    int acc = 0;
    for (int k = 0; k < KC; k++) acc += (int)W[o * K + k] * quant8(xo[s * KC + k]) + (int)W[o * K + KC + k] * quant8(xc[s * KC + k]);
    float t = (float)acc * sc[o]; float v = t + bi[o];
    y_ref[s * OUT + o] = tanh_r(v); yq_ref[s * OUT + o] = quant8(y_ref[s * OUT + o]);
  }
  LayerArgs a; int8_t *dW, *dyq; float *dsc, *dbi, *dxo, *dxc, *dy; long long *dc;
  CK(cudaMalloc(&dW, W.size())); CK(cudaMalloc(&dyq, TS * OUT)); CK(cudaMalloc(&dsc, OUT * 4)); CK(cudaMalloc(&dbi, OUT * 4));
  CK(cudaMalloc(&dxo, xo.size() * 4)); CK(cudaMalloc(&dxc, xc.size() * 4)); CK(cudaMalloc(&dy, TS * OUT * 4)); CK(cudaMalloc(&dc, 8));
  CK(cudaMemcpy(dW, W.data(), W.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsc, sc.data(), OUT * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbi, bi.data(), OUT * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dxo, xo.data(), xo.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dxc, xc.data(), xc.size() * 4, cudaMemcpyHostToDevice));
  a = {dW, dsc, dbi, dxo, dxc, dy, dyq, dc};
  const int smem_bytes = ROWS_ALLOC * K + 2 * TS * KC;
  CK(cudaFuncSetAttribute(conv_layer_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  conv_layer_umma<<<1, 128, smem_bytes>>>(a);
  CK(cudaDeviceSynchronize());
  std::vector<float> y(TS * OUT); std::vector<int8_t> yq(TS * OUT); long long cyc;
  CK(cudaMemcpy(y.data(), dy, y.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(yq.data(), dyq, yq.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (size_t i = 0; i < y.size(); i++) bad += (y[i] != y_ref[i]) || (yq[i] != yq_ref[i]);
  printf("conv layer (%d outputs, 2 x %d inputs, %d streams): %s (%d of %zu differ), %lld cycles issue -> outputs stored (%d MMAs)\n", OUT, KC, TS,
         bad ? "WRONG" : "exact", bad, y.size(), cyc, K / 32);
  return bad != 0;
}
