// Scalar arithmetic of the core codec epilogues, shared by the mma.sync kernels (core_codec.cu) and the tcgen05 kernels
// (core_codec_umma.cu).  Every float operation is separately rounded, in the order of the reference's generic C path
// (restated in oracle/nnet_shim), so that results are bit-identical to it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float tanh_r(float x) {
  const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f;
  const float D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
  float x2 = __fmul_rn(x, x);
  float num = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(N2, x2), N1), x2), N0);
  float den = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(D2, x2), D1), x2), D0);
  float y = __fdiv_rn(__fmul_rn(num, x), den);
  return fmaxf(-1.f, fminf(1.f, y));
}
__device__ __forceinline__ float sigmoid_r(float x) {
  return __fadd_rn(.5f, __fmul_rn(.5f, tanh_r(__fmul_rn(.5f, x))));
}
// C semantics of `(int)floor(.5+127*x)`: 127*x is a float product (rounded to binary32), the sum with .5 is double
__device__ __forceinline__ int8_t quant8(float x) {
  return (int8_t)__double2int_rd((double)__fmul_rn(127.f, x) + 0.5);
}
__device__ __forceinline__ float lin(int acc, float scale, float bias) {
  return __fadd_rn(__fmul_rn((float)acc, scale), bias);
}


// Float layers accumulate sequentially over inputs, product and sum separately rounded (the generic sgemv of the reference):
// (round(w.x * x), round(w.y * x)) with one packed multiply (SASS FMUL2); the accumulation stays scalar so that ptxas cannot
// contract product and sum into an FFMA2 (it does that to mul.rn.f32x2 + add.rn.f32x2, even under -fmad=false)
__device__ __forceinline__ float2 prod2_rn(float wx, float wy, float x) {
  float2 w = make_float2(wx, wy), xx = make_float2(x, x);
  unsigned long long ra = *reinterpret_cast<unsigned long long *>(&w), rb = *reinterpret_cast<unsigned long long *>(&xx), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2 *>(&rd);
}
template <int OPT>
__device__ __forceinline__ void mac_row(float (&acc)[OPT], const float4 *wrow, float x) {
#pragma unroll
  for (int v = 0; v < OPT / 4; v++) {
    const float4 w = wrow[v];
    const float2 p0 = prod2_rn(w.x, w.y, x), p1 = prod2_rn(w.z, w.w, x);
    acc[4 * v + 0] = __fadd_rn(acc[4 * v + 0], p0.x); acc[4 * v + 1] = __fadd_rn(acc[4 * v + 1], p0.y);
    acc[4 * v + 2] = __fadd_rn(acc[4 * v + 2], p1.x); acc[4 * v + 3] = __fadd_rn(acc[4 * v + 3], p1.y);
  }
}
