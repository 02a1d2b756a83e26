// Micro-benchmark: legacy mma.sync m16n8k32 s8 (IMMA.16832) on sm_100a — cycles per instruction for 1, 3 and 6 independent
// accumulator chains per warp, 1 / 2 / 3 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_rate imma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH> __global__ void k(int *out, int iters, long long *cyc) {
  int c[NCH][4]; unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int i = 0; i < NCH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  int s = 0; for (int i = 0; i < NCH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int NCH> void run(int wps, int *out, long long *cyc) {
  long long h; const int iters = 2048;
  k<NCH><<<1, wps * 128>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("chains %d, warps/SMSP %d: %.1f cycles per IMMA per warp, %.1f per SMSP-instr\n", NCH, wps, (double)h / (iters * NCH), (double)h / (iters * NCH) / wps);
}
int main() {
  int *out; long long *cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  for (int wps = 1; wps <= 3; wps++) { run<1>(wps, out, cyc); run<3>(wps, out, cyc); run<6>(wps, out, cyc); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
