// Micro-benchmark: FP64 issue rate on sm_100a — DFMA (vector) and DMMA m8n8k4 (tensor) per SM sub-partition, for
// 1..8 warps per SM sub-partition with 8 independent accumulator chains per thread.  Prints cycles per warp-instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_k(double *out, int iters, long long *cyc) {
  double a[8], x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
  for (int i = 0; i < 8; i++) a[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], x, y);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dmma_k(double *out, int iters, long long *cyc) {
  double c[8][2], a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int wps = 1; wps <= 8; wps *= 2) {
    int threads = wps * 4 * 32;
    dfma_k<<<1, threads>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 8.0);
    printf("DFMA  warps/SMSP %d: %.2f cycles per warp-instr per warp, %.2f per SMSP-instr\n", wps, per, per / wps);
    dmma_k<<<1, threads>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    per = (double)h / (iters * 8.0);
    printf("DMMA  warps/SMSP %d: %.2f cycles per warp-instr per warp, %.2f per SMSP-instr (256 FMA each)\n", wps, per, per / wps);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
