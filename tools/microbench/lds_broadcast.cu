// Shared-memory cost of broadcast loads: cycles per warp-wide LDS.{32,64,128} when all lanes read the same address, 4 distinct
// addresses (8 lanes each), or 32 distinct consecutive addresses; 16 warps per SM issuing back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_broadcast lds_broadcast.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int W> __device__ __forceinline__ float lds(unsigned addr) {
  float a, b2, c, d;
  if (W == 4) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(addr)); return a; }
  if (W == 8) { asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b2) : "r"(addr)); return a + b2; }
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b2), "=f"(c), "=f"(d) : "r"(addr));
  return a + b2 + c + d;
}
template <int W> __global__ void k(int pattern, int iters, long long *cyc, float *sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 1.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int idx = pattern == 0 ? 0 : pattern == 1 ? (lane >> 3) : lane;       // same / 4 distinct / 32 distinct
  const unsigned base = (unsigned)__cvta_generic_to_shared(smem) + idx * W;
  float acc = 0.f;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) acc += lds<W>(base + u * 1024);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 1234.5f) sink[threadIdx.x] = acc;
}
template <int W> void run(const char *name) {
  long long *cyc; cudaMalloc(&cyc, 8 * 148); float *sink; cudaMalloc(&sink, 4096);
  const char *pn[] = {"all lanes same address", "4 addresses x 8 lanes", "32 consecutive"};
  for (int p = 0; p < 3; p++) {
    const int iters = 2000, warps = 16;
    k<W><<<148, warps * 32, 32768>>>(p, iters, cyc, sink);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-8s %-24s %6.2f cycles per warp instruction (SM-wide)\n", name, pn[p], (double)mx / ((double)iters * 16 * warps));
  }
}
int main() { run<4>("LDS.32"); run<8>("LDS.64"); run<16>("LDS.128"); return 0; }
