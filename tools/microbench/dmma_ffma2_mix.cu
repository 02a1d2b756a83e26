// Does DMMA (FP64 tensor) slow down when FFMA2-heavy warps share the SM sub-partition?  16 warps: 0-11 run packed fp32
// FMAs, 12-15 run DMMA m8n8k4; each group is timed with clock64 alone and together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_ffma2_mix dmma_ffma2_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                     rc = *reinterpret_cast<unsigned long long *>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2 *>(&rd);
}
__global__ void k(double *out, int iters, int mode, long long *cyc) {
  const int w = threadIdx.x >> 5;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (w < 12) {
    if (mode & 1) {
      float2 a[12], x = make_float2(1.0000001f, 0.9999999f), y = make_float2(1e-9f, -1e-9f);
      for (int i = 0; i < 12; i++) a[i] = make_float2(i, -i);
      t0 = clock64();
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 12; i++) a[i] = ffma2(a[i], x, y);
      }
      t1 = clock64();
      float s = 0; for (int i = 0; i < 12; i++) s += a[i].x + a[i].y;
      out[threadIdx.x] = s;
      if (threadIdx.x == 0) cyc[0] = t1 - t0;
    }
  } else {
    if (mode & 2) {
      double c[6][2], a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
      for (int i = 0; i < 6; i++) c[i][0] = c[i][1] = i;
      t0 = clock64();
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 6; i++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      }
      t1 = clock64();
      double s = 0; for (int i = 0; i < 6; i++) s += c[i][0] + c[i][1];
      out[threadIdx.x] = s;
      if (threadIdx.x == 384) cyc[1] = t1 - t0;
    }
  }
}
int main() {
  double *out; long long *cyc, h[2];
  cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 16);
  const int iters = 4096;
  for (int mode = 1; mode <= 3; mode++) {
    cudaMemset(cyc, 0, 16);
    k<<<1, 512>>>(out, iters, mode, cyc); cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    printf("mode %d: FFMA2 %.2f cycles/warp-instr/SMSP (3 warps per SMSP), DMMA %.2f cycles/instr (6 independent chains)\n", mode,
           h[0] / (iters * 12.0 * 3.0), h[1] / (iters * 6.0));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
