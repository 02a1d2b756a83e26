// How fast does cp.async.bulk (1-D TMA) fill shared memory from L2 when every CTA streams the SAME bytes (the codec's weight
// stream) — as a function of copy size and copies in flight?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_rate bulk_copy_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// ring of NST stages of `stage` bytes; each stage is filled by `split` copies of stage/split bytes; one thread does everything
__global__ void k(const unsigned char *src, size_t src_bytes, int stage, int nst, int split, int iters, int distinct, long long *cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) { for (int i = 0; i < nst; i++) mbar_init(&full[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned char *base = src + (distinct ? (size_t)blockIdx.x * (src_bytes / gridDim.x) : 0);
  const size_t span = distinct ? src_bytes / gridDim.x : src_bytes;
  size_t off = 0;
  auto issue = [&](int s) {
    mbar_expect_tx(&full[s], stage);
    const int piece = stage / split;
    for (int p = 0; p < split; p++) bulk_g2s(smem + (size_t)s * stage + p * piece, base + off + p * piece, piece, &full[s]);
    off += stage; if (off + stage > span) off = 0;
  };
  long long t0 = clock64();
  for (int s = 0; s < nst; s++) issue(s);
  uint32_t phase = 0; int s = 0;
  for (int i = 0; i < iters; i++) {
    mbar_wait(&full[s], phase);
    if (i + nst < iters) issue(s);
    if (++s == nst) { s = 0; phase ^= 1; }
  }
  cycles[blockIdx.x] = clock64() - t0;
}
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  const size_t src_bytes = 64u << 20;
  unsigned char *src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  long long *cyc; cudaMalloc(&cyc, sizeof(long long) * nsm);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h[256];
  printf("%-9s %-6s %-4s %-6s %-9s %10s %10s\n", "source", "stage", "nst", "split", "in flight", "B/clk/SM", "TB/s chip");
  for (int distinct = 0; distinct < 2; distinct++)
    for (int stage : {4096, 16384, 32768})
      for (int nst : {1, 2, 4, 6})
        for (int split : {1, 4, 16}) {
          if ((size_t)stage * nst > 196 * 1024 || stage / split < 512) continue;
          const int iters = (8 << 20) / stage;               // 8 MB per CTA
          // same-source runs cycle through a 1 MB window (the size of one codec's weight set), L2-resident after the first pass
          const size_t window = distinct ? src_bytes : (1u << 20);
          k<<<nsm, 32, (size_t)stage * nst>>>(src, window, stage, nst, split, iters, distinct, cyc);   // warm L2
          k<<<nsm, 32, (size_t)stage * nst>>>(src, window, stage, nst, split, iters, distinct, cyc);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < nsm; i++) mx = h[i] > mx ? h[i] : mx;
          const double bpc = (double)iters * stage / mx;
          printf("%-9s %-6d %-4d %-6d %-9d %10.1f %10.2f\n", distinct ? "per-CTA" : "shared", stage, nst, split, nst * split, bpc, bpc * nsm * 1.965e9 / 1e12);
        }
  return 0;
}
