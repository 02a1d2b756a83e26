// Micro-benchmark + layout check for the round-2 codec plan: int8 layers on the 5th-gen tensor cores with the
// operands swapped -- tcgen05.mma kind::i8, A = a 128-row slice of an int8 weight matrix, B = the quantised activations of
// N = 8..256 streams, D = int32 accumulators in TMEM (lane = output feature, column = stream).  Both operands K-major in the
// no-swizzle canonical layout: 8-row x 16-byte core matrices, 128 B each,
//     offset(row r, k byte b) = (r / 8) * SBO + (b / 16) * LBO + (r % 8) * 16 + b % 16,      LBO = 128, SBO = K_TOTAL * 8
// i.e. what a producer of quantised activations can write with plain byte stores and what the host can pre-bake for weights
// (1-D cp.async.bulk streaming, no tensor map).
//
// Prints per N: (1) exactness of D against an int32 CPU GEMM (validates the descriptors / layouts), (2) cycles per MMA when
// issued back to back (the guide's floor is max(M,128) * N / 256), (3) the round trip one layer of a dependent chain pays:
// issue -> tcgen05.commit -> mbarrier wait -> tcgen05.ld -> wait::ld.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o umma_i8_swapab umma_i8_swapab.cu
//   timeout 60 ./umma_i8_swapab
// NOT part of the library; written at the end of round 1 without GPU time left, so it has been compiled (SASS shows
// UTCIMMA / UTCBAR / LDTM) but not yet run.  Every wait is bounded and traps, run it under `timeout`.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128;            // output features per MMA (TMEM lanes)
constexpr int K_TOTAL = 256;      // bytes of K staged in shared memory = 8 MMAs of K = 32
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
__host__ __device__ constexpr uint32_t instr_desc_i8(int m, int n) {
  // cute::UMMA::InstrDescriptor: c_format S32 (2) at [4,6), a_format / b_format signed 8 bit (1) at [7,10) / [10,13),
  // a_major = b_major = K (0), n >> 3 at [17,23), m >> 4 at [24,29)
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
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

// A_rm [M][K_TOTAL], B_rm [N][K_TOTAL] row-major int8 in global memory; D [M][N] int32; cyc[0] = back-to-back cycles for
// `iters` passes of 8 MMAs, cyc[1] = cycles of `chain` dependent single-MMA round trips
__global__ void __launch_bounds__(128, 1)
umma_probe(const int8_t *__restrict__ A_rm, const int8_t *__restrict__ B_rm, int N, int iters, int chain,
           int *__restrict__ D, long long *__restrict__ cyc, int swap_lbo_sbo) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *sA = smem;                              // M * K_TOTAL
  uint8_t *sB = smem + M * K_TOTAL;                // N * K_TOTAL
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  constexpr uint32_t LBO = 128, SBO = K_TOTAL * 8;
  // diagnostic: should the descriptor fields be the other way round on this toolchain / hardware, the host tries both
  const uint32_t DL = swap_lbo_sbo ? SBO : LBO, DS = swap_lbo_sbo ? LBO : SBO;

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
  const uint32_t idesc = instr_desc_i8(M, N);
  const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
  uint32_t phase = 0;

  // ---- (1)+(2): `iters` passes over the 8 K-chunks, one commit at the end
  if (warp == 0) {
    long long t0 = clock64();
    if (lane == 0) {
      for (int it = 0; it < iters; it++)
#pragma unroll
        for (int k = 0; k < K_TOTAL / K_MMA; k++)
          umma_i8(tmem, smem_desc(a0 + k * 2 * LBO, DL, DS), smem_desc(b0 + k * 2 * LBO, DL, DS), idesc, (it | k) != 0);
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
      if (lane == 0) { umma_i8(tmem, smem_desc(a0, DL, DS), smem_desc(b0, DL, DS), idesc, 0); umma_commit(&bar); }
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
  const int Ns[] = {8, 16, 32, 64, 128, 256};
  std::vector<int8_t> hA(M * K_TOTAL), hB(256 * K_TOTAL);
  srand(1);
  for (auto &x : hA) x = (int8_t)(rand() % 15 - 7);
  for (auto &x : hB) x = (int8_t)(rand() % 15 - 7);
  int8_t *dA, *dB; int *dD; long long *dC;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, M * 256 * 4)); CK(cudaMalloc(&dC, 16));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (M + 256) * K_TOTAL));
  for (int N : Ns) {
    for (int grid : {1, prop.multiProcessorCount}) {
      const int iters = 512, chain = 256;
      std::vector<int> hD(M * N);
      long long hC[2];
      // pass 1: iters = 1 for the exactness check (int32 would not overflow either way), pass 2: timing
      long bad = 0, bad_swapped = 0;
      for (int swap = 0; swap >= 0; swap--) {             // the documented field order (the exchanged one reads outside shared memory: round-2 run)
        umma_probe<<<grid, 128, (M + N) * K_TOTAL>>>(dA, dB, N, 1, 1, dD, dC, swap);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        long nb = 0;
        for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
          int ref = 0; for (int k = 0; k < K_TOTAL; k++) ref += (int)hA[m * K_TOTAL + k] * (int)hB[n * K_TOTAL + k];
          nb += ref != hD[m * N + n];
        }
        (swap ? bad_swapped : bad) = nb;
      }
      if (bad && !bad_swapped) printf("  NOTE: exact only with LBO and SBO exchanged in the descriptors\n");
      umma_probe<<<grid, 128, (M + N) * K_TOTAL>>>(dA, dB, N, iters, chain, dD, dC, 0);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hC, dC, 16, cudaMemcpyDeviceToHost));
      printf("M=128 N=%3d grid=%3d: %s (%ld mismatches)  %.2f cyc/MMA back to back (floor %d)  %.0f cyc per dependent link\n", N, grid,
             bad ? "WRONG" : "exact", bad, (double)hC[0] / (iters * (K_TOTAL / K_MMA)), 128 * N / 256, (double)hC[1] / chain);
    }
  }
  return 0;
}
