// Which issue-loop shape lets one thread feed tcgen05.mma kind::i8 M = 128 N = 8 at the tensor core's own pace (47 cycles)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue_loop umma_issue_loop.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t accumulate) {
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((8u >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t da = (uint64_t)a_hi << 32 | a_lo, db = (uint64_t)b_hi << 32 | b_lo;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }

__global__ void __launch_bounds__(128, 1) k(int variant, int nk, int nchunks, long long *cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t done_bar, chunk_bar[4];
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
  if (tid == 0) { mbar_init(&done_bar, 1); for (int i = 0; i < 4; i++) mbar_init(&chunk_bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_s)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (warp == 3) {
    const uint32_t a_hi = (1024u >> 4) | (1u << 14), b_hi = (8192u >> 4) | (1u << 14);
    const uint32_t a_base = ((smem_u32(smem) & 0x3FFFF) >> 4) | (8u << 16), b_base = ((smem_u32(smem + 32768) & 0x3FFFF) >> 4) | (8u << 16);
    long long t0 = clock64();
    if (variant == 0) {                        // everything inside if (lane == 0), 8 MMAs unrolled per chunk
      if (lane == 0)
        for (int c = 0; c < nchunks; c++) {
#pragma unroll
          for (int kk = 0; kk < 8; kk++) umma(tmem, a_base + kk * 16, a_hi, b_base + kk * 16, b_hi, (c | kk) != 0);
          umma_commit(&chunk_bar[c & 3]);
        }
    } else if (variant == 1) {                 // inside if (lane == 0), runtime inner loop, incremental descriptors
      if (lane == 0)
        for (int c = 0; c < nchunks; c++) {
          uint32_t a_lo = a_base, b_lo = b_base + (c & 3) * 128;
          for (int kk = 0; kk < nk; kk++) { umma(tmem, a_lo, a_hi, b_lo, b_hi, (c | kk) != 0); a_lo += 16; b_lo += 16; }
          umma_commit(&chunk_bar[c & 3]);
        }
    } else if (variant == 2) {                 // warp-uniform loop, if (leader) around each MMA (the codec's issuer)
      const bool leader = lane == 0;
      for (int c = 0; c < nchunks; c++) {
        uint32_t a_lo = a_base, b_lo = b_base + (c & 3) * 128;
        for (int kk = 0; kk < nk; kk++) { if (leader) umma(tmem, a_lo, a_hi, b_lo, b_hi, (c | kk) != 0); a_lo += 16; b_lo += 16; }
        if (leader) umma_commit(&chunk_bar[c & 3]);
        __syncwarp();
      }
    } else if (variant == 3) {                 // warp-uniform loop, leader branch around the whole inner loop
      const bool leader = lane == 0;
      for (int c = 0; c < nchunks; c++) {
        const uint32_t a_lo = a_base, b_lo = b_base + (c & 3) * 128;
        if (leader) {
          for (int kk = 0; kk < nk; kk++) umma(tmem, a_lo + kk * 16, a_hi, b_lo + kk * 16, b_hi, (c | kk) != 0);
          umma_commit(&chunk_bar[c & 3]);
        }
        __syncwarp();
      }
    } else if (variant == 4) {                 // as 2, two tiles per k-block (two accumulators)
      const bool leader = lane == 0;
      for (int c = 0; c < nchunks; c++) {
        uint32_t a_lo = a_base, b_lo = b_base + (c & 3) * 128;
        for (int kk = 0; kk < nk; kk++) {
          if (leader) { umma(tmem, a_lo, a_hi, b_lo, b_hi, (c | kk) != 0); umma(tmem + 8, a_lo + 1024, a_hi, b_lo, b_hi, (c | kk) != 0); }
          a_lo += 16; b_lo += 16;
        }
        if (leader) umma_commit(&chunk_bar[c & 3]);
        __syncwarp();
      }
    } else if (variant == 5) {                 // as 0 but the accumulate flag is a compile-time constant (1)
      if (lane == 0)
        for (int c = 0; c < nchunks; c++) {
#pragma unroll
          for (int kk = 0; kk < 8; kk++) umma(tmem, a_base + kk * 16, a_hi, b_base + kk * 16, b_hi, 1);
          umma_commit(&chunk_bar[c & 3]);
        }
    }
    __syncwarp();
    if (lane == 0) umma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0);
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128) : "memory");
}
int main() {
  long long *cyc; cudaMalloc(&cyc, sizeof(long long) * 256);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char *names[] = {"lane0 branch, 8 unrolled", "lane0 branch, runtime loop", "uniform loop, leader per MMA", "uniform loop, leader per chunk",
                         "uniform loop, 2 tiles per k-block", "lane0 branch, unrolled, const accumulate"};
  for (int v = 0; v < 6; v++) {
    const int nk = 8, nchunks = 64;
    k<<<148, 128, 64 * 1024>>>(v, nk, nchunks, cyc);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
    const int n = nk * nchunks * (v == 4 ? 2 : 1);
    printf("%-44s %8.1f cycles per MMA (%d MMAs, commit every %d)\n", names[v], (double)mx / n, n, nk * (v == 4 ? 2 : 1));
  }
  return 0;
}
