// Layout / numerics probe for the round-2 rx_track plan: tcgen05.mma kind::tf32, fp32 data in shared memory
// (the tensor core ignores or rounds the low 13 mantissa bits -- this program tells which), fp32 accumulators in TMEM.
// Same no-swizzle K-major canonical layout as umma_i8_swapab.cu (8-row x 16-byte core matrices; 4 floats per row chunk,
// K = 8 floats = 32 bytes per MMA).  Prints, per N, the largest deviation of D from three CPU models: exact fp32 operands,
// operands truncated to tf32, operands rounded to nearest tf32 -- and the cycles per MMA issued back to back.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o umma_tf32_probe umma_tf32_probe.cu
//   timeout 60 ./umma_tf32_probe
// Written at the end of round 1 without GPU time left: compiled (SASS shows UTCHMMA), not yet run.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

constexpr int M = 128;            // output features per MMA (TMEM lanes)
constexpr int K_TOTAL = 256;      // bytes of K staged in shared memory = 8 MMAs of K = 8 floats
constexpr int K_MMA = 32;
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46), version 1 at [46,48),
  // base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE (0) at [61,64)
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
  // cute::UMMA::InstrDescriptor: c_format F32 (1) at [4,6), a_format / b_format TF32 (2) at [7,10) / [10,13),
  // a_major = b_major = K (0), n >> 3 at [17,23), m >> 4 at [24,29)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  for (long long spin = 0; !ok; spin++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spin > (1ll << 24)) __trap();                // never hang the box
  }
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// A_rm [M][K_TOTAL / 4], B_rm [N][K_TOTAL / 4] row-major fp32 in global memory (addressed as bytes); D [M][N] fp32 bit
// patterns; cyc[0] = back-to-back cycles for
// `iters` passes of 8 MMAs, cyc[1] = cycles of `chain` dependent single-MMA round trips
__global__ void __launch_bounds__(128, 1)
umma_probe(const uint8_t *__restrict__ A_rm, const uint8_t *__restrict__ B_rm, int N, int iters, int chain,
           int *__restrict__ D, long long *__restrict__ cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sA = smem;                              // M * K_TOTAL
  uint8_t *sB = smem + M * K_TOTAL;                // N * K_TOTAL
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  constexpr uint32_t LBO = 128, SBO = K_TOTAL * 8;

  for (int i = tid; i < M * K_TOTAL; i += 128) { int r = i / K_TOTAL, b = i % K_TOTAL; sA[(r / 8) * SBO + (b / 16) * LBO + (r % 8) * 16 + b % 16] = A_rm[i]; }
  for (int i = tid; i < N * K_TOTAL; i += 128) { int r = i / K_TOTAL, b = i % K_TOTAL; sB[(r / 8) * SBO + (b / 16) * LBO + (r % 8) * 16 + b % 16] = B_rm[i]; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core's async proxy
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = instr_desc_tf32(M, N);
  const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
  uint32_t phase = 0;

  // ---- (1)+(2): `iters` passes over the 8 K-chunks, one commit at the end
  if (warp == 0) {
    long long t0 = clock64();
    if (lane == 0) {
      for (int it = 0; it < iters; it++)
#pragma unroll
        for (int k = 0; k < K_TOTAL / K_MMA; k++)
          umma_tf32(tmem, smem_desc(a0 + k * 2 * LBO, LBO, SBO), smem_desc(b0 + k * 2 * LBO, LBO, SBO), idesc, (it | k) != 0);
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, phase);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  }
  phase ^= 1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // every warp reads its 32 lanes (output features 32*warp .. +31), 8 streams at a time
  if (blockIdx.x == 0)
    for (int c = 0; c < N; c += 8) {
      int v[8];
      tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + c, v);
      for (int j = 0; j < 8; j++) D[(32 * warp + lane) * N + c + j] = v[j];
    }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- (3): dependent chain, one K = 32 MMA per link: issue -> commit -> wait -> ld (first 8 streams) -> next
  if (warp == 0) {
    int sink = 0;
    long long t0 = clock64();
    for (int i = 0; i < chain; i++) {
      if (lane == 0) { umma_tf32(tmem, smem_desc(a0, LBO, SBO), smem_desc(b0, LBO, SBO), idesc, 0); umma_commit(&bar); }
      __syncwarp();
      mbar_wait(&bar, phase); phase ^= 1;
      __syncwarp();                                      // tcgen05.ld is warp-collective (.sync.aligned)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int v[8];
      tmem_ld8(tmem, v);
      sink += v[0] + v[7];
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cyc[1] = t1 - t0;
    if (sink == 0x7fffffff) D[0] = sink;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  if (prop.major != 10) { printf("needs sm_100 (found sm_%d%d)\n", prop.major, prop.minor); return 1; }
  const int Ns[] = {16, 48, 96, 256};     // M = 128 wants N % 16 == 0 for the f16 / tf32 kinds
  const int KF = K_TOTAL / 4;                                      // floats of K per row
  std::vector<float> hA(M * KF), hB(256 * KF);
  srand(1);
  for (auto &x : hA) x = (float)rand() / (float)RAND_MAX * 2.f - 1.f;     // full 24-bit mantissas
  for (auto &x : hB) x = (float)rand() / (float)RAND_MAX * 2.f - 1.f;
  auto trunc_tf32 = [](float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; };
  auto rne_tf32 = [](float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x00000FFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; };
  uint8_t *dA, *dB; int *dD; long long *dC;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, M * 256 * 4)); CK(cudaMalloc(&dC, 16));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (M + 256) * K_TOTAL));
  for (int N : Ns) {
    const int iters = 512, chain = 256;
    std::vector<float> hD(M * N);
    long long hC[2];
    umma_probe<<<1, 128, (M + N) * K_TOTAL>>>(dA, dB, N, 1, 1, dD, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double e_exact = 0, e_trunc = 0, e_rne = 0;
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
      double r0 = 0, r1 = 0, r2 = 0;
      for (int k = 0; k < KF; k++) {
        const float a = hA[m * KF + k], b = hB[n * KF + k];
        r0 += (double)a * b; r1 += (double)trunc_tf32(a) * trunc_tf32(b); r2 += (double)rne_tf32(a) * rne_tf32(b);
      }
      const double d = hD[m * N + n];
      e_exact = fmax(e_exact, fabs(d - r0)); e_trunc = fmax(e_trunc, fabs(d - r1)); e_rne = fmax(e_rne, fabs(d - r2));
    }
    umma_probe<<<prop.multiProcessorCount, 128, (M + N) * K_TOTAL>>>(dA, dB, N, iters, chain, dD, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hC, dC, 16, cudaMemcpyDeviceToHost));
    printf("tf32 M=128 N=%3d K=%d: max |D - model|: fp32 operands %.2e, truncated %.2e, rounded %.2e   %.2f cyc/MMA back to back (floor %d)  %.0f cyc per dependent link\n",
           N, KF, e_exact, e_trunc, e_rne, (double)hC[0] / (iters * (K_TOTAL / K_MMA)), 128 * N / 256, (double)hC[1] / chain);
  }
  return 0;
}
