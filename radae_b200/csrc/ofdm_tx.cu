// OFDM modulator for one 120 ms modem frame per stream.
//
// Replaces transmitter_one.transmitter_one (radae/dsp.py:340-378; batch twin RADAE.forward radae/radae.py:482-527)
// and the EOO frame of RADAE.__init__ / set_eoo_bits (radae/radae.py:208-219, :441-455) as used by
// radae_tx.do_radae_tx / do_eoo (radae_txe.py:108-144).
//
// One CTA per stream, one thread per time sample of an OFDM symbol: z[3][80] -> 120 QPSK-like symbols laid out
// row-major over [Ns=4][Nc=30] (symbol k -> row 1+k/30, carrier k%30), pilot row = pilot_gain*P, pruned 30->160
// IDFT against the reference's own Winv table (L1/L2 resident, read coalesced), cyclic prefix = tail copy,
// PA model tanh(|x|)*x/|x|.  HBM traffic: 960 B in, 7680 B out per stream-frame, nothing else.
#include "rade_common.h"
#include "ofdm_mod.cuh"

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return mod_cmul(a, b); }

__global__ void __launch_bounds__(RADE_M)
ofdm_mod_kernel(DspTables T, const float *__restrict__ z, float2 *__restrict__ tx, int S) {
  __shared__ float2 sym[RADE_NS + 1][RADE_NC];
  const int s = blockIdx.x;
  ofdm_mod_frame(T, z + (size_t)s * RADE_NZMF * RADE_LATENT, tx + (size_t)s * RADE_NMF, sym, threadIdx.x);
}

// EOO frame: skeleton (P E . . . E) + optionally 3 data symbols from 180 +-1 bits
__global__ void __launch_bounds__(RADE_M)
eoo_kernel(DspTables T, const float *__restrict__ bits, const int *__restrict__ has_bits, float2 *__restrict__ tx, int S) {
  __shared__ float2 sym[RADE_NS - 1][RADE_NC];
  const int s = blockIdx.x, n = threadIdx.x;
  float2 *out = tx + (size_t)s * RADE_NEOO;
  const bool data = has_bits[s] != 0;
  for (int i = n; i < RADE_NEOO; i += RADE_M)
    if (!data || i < 2 * RADE_SYM || i >= RADE_NMF) out[i] = T.eoo_base[i];     // data rows 2..4 are written below
  if (!data) return;
  if (n < (RADE_NS - 1) * RADE_NC) {
    const float *b = bits + (size_t)s * RADE_NEOO_BITS;
    sym[n / RADE_NC][n % RADE_NC] = make_float2(b[2 * n], b[2 * n + 1]);
  }
  __syncthreads();
  float2 acc[RADE_NS - 1];
#pragma unroll
  for (int r = 0; r < RADE_NS - 1; r++) acc[r] = make_float2(0.f, 0.f);
  for (int c = 0; c < RADE_NC; c++) {
    const float2 w = T.Winv[c * RADE_M + n];
#pragma unroll
    for (int r = 0; r < RADE_NS - 1; r++) {
      float2 v = cmul(sym[r][c], w);
      acc[r].x += v.x; acc[r].y += v.y;
    }
  }
#pragma unroll
  for (int r = 0; r < RADE_NS - 1; r++) {
    float2 y = pa_limit(make_float2(acc[r].x * T.pilot_gain, acc[r].y * T.pilot_gain));
    out[(2 + r) * RADE_SYM + RADE_NCP + n] = y;
    if (n >= RADE_M - RADE_NCP) out[(2 + r) * RADE_SYM + n - (RADE_M - RADE_NCP)] = y;
  }
}

// Optional TX band-pass filter + unit-magnitude clip, radae_tx(txbpf_en=True): radae_txe.py:74-81 (same complex_bpf
// object and band as the receive filter), :130-132 (modem frame) and :141-143 (EOO frame).  Applied in place to the
// n = 960 or 1152 samples the modulator has just written; one CTA per stream.  complex_bpf.bpf (radae/dsp.py:63-102):
// mix down with the running phase, 101 real taps over [memory, frame], mix up; the memory keeps Ntap+1 = 102 samples
// (dsp.py:96), so every call after the first is delayed by two samples -- restated exactly as in rx_bpf_kernel.
// clip(|y|, 0, 1) * exp(j angle(y)) == y * min(|y|, 1) / |y|.
constexpr int TXBPF_THREADS = 288;
__global__ void __launch_bounds__(TXBPF_THREADS)
tx_bpf_clip_kernel(DspTables T, float2 *__restrict__ tx, size_t stride, int n, TxBpfState *__restrict__ st) {
  __shared__ float2 X[RADE_BPF_MEM + RADE_NEOO + 2];
  __shared__ float h[RADE_BPF_NTAP];
  const int s = blockIdx.x, tid = threadIdx.x;
  TxBpfState &c = st[s];
  float2 *x = tx + (size_t)s * stride;
  const bool fresh = c.started == 0;                                 // zero-filled state == a new complex_bpf object
  const float2 ph = fresh ? make_float2(1.f, 0.f) : c.phase;
  const int off = fresh ? 2 : 0;                                     // 100 samples of memory on the first call, 102 afterwards
  for (int i = tid; i < RADE_BPF_MEM; i += TXBPF_THREADS) X[i] = c.mem[i];
  for (int i = tid; i < RADE_BPF_NTAP; i += TXBPF_THREADS) h[i] = T.bpf_h[i];
  for (int i = tid; i < n; i += TXBPF_THREADS) X[RADE_BPF_MEM + i] = cmul(x[i], cmul(ph, T.bpf_exp[i]));     // mix down
  for (int i = RADE_BPF_MEM + n + tid; i < RADE_BPF_MEM + RADE_NEOO + 2; i += TXBPF_THREADS) X[i] = make_float2(0.f, 0.f);
  __syncthreads();
  for (int i = tid; i < n; i += TXBPF_THREADS) {
    float2 acc = make_float2(0.f, 0.f);
    const float2 *w = X + i + off;
    for (int k = 0; k < RADE_BPF_NTAP; k++) { acc.x = fmaf(h[k], w[k].x, acc.x); acc.y = fmaf(h[k], w[k].y, acc.y); }
    const float2 u = cmul(ph, T.bpf_exp[i]);
    const float2 y = cmul(acc, make_float2(u.x, -u.y));              // mix up
    const float mag = hypotf(y.x, y.y);
    const float g = mag > 0.f ? fminf(mag, 1.f) / mag : 0.f;
    x[i] = make_float2(y.x * g, y.y * g);
  }
  for (int i = tid; i < RADE_BPF_MEM; i += TXBPF_THREADS) c.mem[i] = X[n + i];
  if (tid == 0) { c.phase = cmul(ph, T.bpf_exp[n - 1]); c.started = 1; }
}

}  // namespace

int tx_bpf_clip_launch(const DspTables &T, float2 *tx, size_t stride, int n, TxBpfState *st, int S, cudaStream_t stream) {
  if (n < 1 || n > RADE_NEOO) return -1;
  tx_bpf_clip_kernel<<<S, TXBPF_THREADS, 0, stream>>>(T, tx, stride, n, st);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}

int ofdm_mod_launch(const DspTables &T, const float *z, float2 *tx, int S, cudaStream_t stream) {
  ofdm_mod_kernel<<<S, RADE_M, 0, stream>>>(T, z, tx, S);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}

int eoo_launch(const DspTables &T, const float *bits, const int *has_bits, float2 *tx, int S, cudaStream_t stream) {
  eoo_kernel<<<S, RADE_M, 0, stream>>>(T, bits, has_bits, tx, S);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}
