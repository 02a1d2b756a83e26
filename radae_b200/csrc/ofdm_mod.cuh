// OFDM modulator of one 120 ms modem frame as a device function (used by ofdm_mod_kernel and, fused, by the channel kernel).
// transmitter_one.transmitter_one (radae/dsp.py:340-378): z[3][80] -> 120 QPSK-like symbols row-major over [Ns=4][Nc=30],
// pilot row = pilot_gain*P, pruned 30->160 IDFT against the reference's Winv table, cyclic prefix = tail copy, PA model.
#pragma once
#include "rade_common.h"

__device__ __forceinline__ float2 mod_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// tanh(|x|) * exp(j angle(x)) == x * tanh(|x|)/|x|
__device__ __forceinline__ float2 pa_limit(float2 x) {
  float mag = hypotf(x.x, x.y);
  if (mag == 0.f) return make_float2(0.f, 0.f);
  float s = tanhf(mag) / mag;
  return make_float2(x.x * s, x.y * s);
}
// Called by EVERY thread of the CTA (contains a __syncthreads); thread n < 160 produces sample n of the five OFDM symbols.
// zs: the stream's 240 latents; out: 960 samples (shared or global memory); sym: shared scratch [Ns+1][Nc].
__device__ __forceinline__ void ofdm_mod_frame(const DspTables &T, const float *__restrict__ zs, float2 *out, float2 (*sym)[RADE_NC], int n) {
  if (n < RADE_NC) sym[0][n] = make_float2(T.pilot_gain * T.P[n].x, T.pilot_gain * T.P[n].y);
  if (n < RADE_NS * RADE_NC) sym[1 + n / RADE_NC][n % RADE_NC] = make_float2(zs[2 * n], zs[2 * n + 1]);
  __syncthreads();
  if (n >= RADE_M) return;
  float2 acc[RADE_NS + 1];
#pragma unroll
  for (int r = 0; r <= RADE_NS; r++) acc[r] = make_float2(0.f, 0.f);
  for (int c = 0; c < RADE_NC; c++) {
    const float2 w = T.Winv[c * RADE_M + n];
#pragma unroll
    for (int r = 0; r <= RADE_NS; r++) {
      float2 v = mod_cmul(sym[r][c], w);
      acc[r].x += v.x; acc[r].y += v.y;
    }
  }
#pragma unroll
  for (int r = 0; r <= RADE_NS; r++) {
    float2 y = pa_limit(acc[r]);
    out[r * RADE_SYM + RADE_NCP + n] = y;
    if (n >= RADE_M - RADE_NCP) out[r * RADE_SYM + n - (RADE_M - RADE_NCP)] = y;
  }
}
