// HF channel simulator + loop-back sample link.
//
// Replaces the rate-Fs channel branch of RADAE.forward (radae/radae.py:529-599): two-path multipath
// tx*G1 + delay_d(tx*G2) (:530-534), frequency/phase offset exp(j*cumsum(omega)) (:542-553), AWGN with
// sigma = sqrt(Fs/(EbNo*Rb)) (:570-578) and gain (:589).  Deviation, stated in DESIGN.md: the reference
// normalises multipath power over the WHOLE batch tensor (:536-539); here normalisation is per stream and by
// expectation (E|G1|^2 + E|G2|^2 = 1, which is what multipath_samples.m:27-31's hf_gain achieves on its files).
// The Watterson gains (Gaussian Doppler spectrum, doppler_spread.m:11-41) are synthesised in-kernel as a sum of
// sinusoids with Gaussian-distributed Doppler shifts, evaluated at frame edges and interpolated linearly, the
// way the reference interpolates its 10 Hz gain samples up to Fs.  Noise: Philox4x32-10 counter RNG keyed by
// (seed, stream), counter = absolute sample index, Box-Muller.
#include "rade_common.h"
#include "rade_host.h"
#include "ofdm_mod.cuh"

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// unit-variance circular complex normal for (seed, stream, sample index, salt)
__device__ __forceinline__ float2 cnormal(unsigned long long seed, uint32_t stream, unsigned long long idx, uint32_t salt) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), stream, salt};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float r = sqrtf(-logf(u01(c[0])));          // |n|^2 ~ Exp(1): variance 1 in total (torch.randn on a complex tensor)
  float sn, cs;
  sincospif(2.f * u01(c[1]), &sn, &cs);
  return make_float2(r * cs, r * sn);
}

// explicit form used by the parity tests: every random quantity supplied by the caller
__global__ void channel_apply_kernel(float2 *__restrict__ rx, const float2 *__restrict__ tx, const float2 *__restrict__ G1,
                                     const float2 *__restrict__ G2, const float2 *__restrict__ noise, int S, int n, int d,
                                     float mp_gain, float freq, float df_dt, float phase0, float sigma, float gain) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)S * n) return;
  const int k = (int)(i % n);
  float2 mp = cmul(tx[i], G1[i]);
  if (k >= d) { float2 e = cmul(tx[i - d], G2[i - d]); mp.x += e.x; mp.y += e.y; }
  double sn, cs;
  // cumsum of omega[i] = 2 pi (f + df_dt i / Fs) / Fs over i = 0..k (radae.py:546-550): f (k + 1) + df_dt k (k + 1) / (2 Fs)
  const double kk = (double)(k + 1);
  sincos((double)phase0 + 2.0 * M_PI / RADE_FS * ((double)freq * kk + (double)df_dt * (double)k * kk / (2.0 * RADE_FS)), &sn, &cs);
  float2 v = cmul(make_float2(mp_gain * mp.x, mp_gain * mp.y), make_float2((float)cs, (float)sn));
  rx[i] = make_float2(gain * (v.x + sigma * noise[i].x), gain * (v.y + sigma * noise[i].y));
}

constexpr int NSIN = 16;
constexpr int LINK_CAP = 4096;

// one sinusoid of a path gain at absolute time t (samples): exp(j(2*pi*f_i*t/Fs + phi_i)), f_i ~ N(0, (spread/2)^2);
// the gain is (1/sqrt(2*NSIN)) * the sum over i = 0..NSIN-1
__device__ float2 path_gain_term(unsigned long long seed, uint32_t stream, uint32_t path, int i, double t, float spread) {
  uint32_t c[4] = {(uint32_t)i, path, stream, 0x5EEDu};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float rr = sqrtf(-2.f * logf(u01(c[0])));
  float s1, c1; sincospif(2.f * u01(c[1]), &s1, &c1);
  const double f = 0.5 * (double)spread * (double)(rr * c1);           // Gaussian Doppler: sigma_f = spread/2
  double sn, cs;
  sincos(2.0 * M_PI * (f * t / RADE_FS + (double)u01(c[2])), &sn, &cs);
  return make_float2((float)cs, (float)sn);
}

// streaming generator form: one CTA per stream, one modem frame (960 samples) per call.  With z_mod non-null the modem frame is
// modulated right here from the stream's 240 latents (fused OFDM modulator: the tx samples never go through HBM).
__global__ void __launch_bounds__(256)
channel_stream_kernel(DspTables T, const float *__restrict__ z_mod, float2 *__restrict__ rx, const float2 *__restrict__ tx,
                      ChanState *__restrict__ st, int S,
                      float sigma, float freq0, float freq_spread, float doppler, int d, float gain, unsigned long long seed,
                      float2 *__restrict__ link_ring, long long *__restrict__ link_wr, const long long *__restrict__ link_rd,
                      int *__restrict__ link_overflow) {
  __shared__ float2 stx[64 + RADE_NMF];
  __shared__ float2 g[4];                       // G1(t0), G1(t0+960), G2(t0), G2(t0+960)
  __shared__ float2 sym[RADE_NS + 1][RADE_NC];
  const int s = blockIdx.x, tid = threadIdx.x;
  ChanState &cs_ = st[s];
  const long long t0 = cs_.t;
  const double ph0 = cs_.phase;
  const float2 *txs = tx + (size_t)s * RADE_NMF;
  for (int i = tid; i < 64; i += blockDim.x) stx[i] = cs_.delay[i];
  if (z_mod) ofdm_mod_frame(T, z_mod + (size_t)s * RADE_NZMF * RADE_LATENT, stx + 64, sym, tid);
  else for (int i = tid; i < RADE_NMF; i += blockDim.x) stx[64 + i] = txs[i];
  if (tid < 4 * NSIN) {                         // 4 gains x 16 sinusoids, one per thread, summed with shuffles
    const int gi = tid / NSIN, i = tid % NSIN;
    float2 v = make_float2(0.f, 0.f);
    if (doppler > 0.f) v = path_gain_term(seed, s, gi >> 1, i, (double)(t0 + (gi & 1) * RADE_NMF), doppler);
#pragma unroll
    for (int o = NSIN / 2; o > 0; o >>= 1) { v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o); }
    if (i == 0) {
      const float a = rsqrtf(2.f * NSIN);
      g[gi] = (doppler > 0.f) ? make_float2(a * v.x, a * v.y) : ((gi < 2) ? make_float2(1.f, 0.f) : make_float2(0.f, 0.f));
    }
  }
  // per-stream frequency offset: freq0 + U(-1,1)*freq_spread, fixed for the life of the stream
  uint32_t c[4] = {0u, 0u, (uint32_t)s, 0xF0FFu};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double f = (double)freq0 + (double)freq_spread * (2.0 * (double)u01(c[0]) - 1.0);
  const double dphi = 2.0 * M_PI * f / RADE_FS;
  // carrier exp(j(ph0 + dphi (i+1))): one complex128 sincos per thread, then rotations by exp(j dphi 256)
  double sn, cs, sb, cb;
  sincos(ph0 + dphi * (double)(tid + 1), &sn, &cs);
  sincos(dphi * 256.0, &sb, &cb);
  __syncthreads();
  // output: the caller's [S][960] array, or (loop-back runs) straight into the stream's link FIFO
  float2 *out = link_ring ? link_ring + (size_t)s * LINK_CAP : rx + (size_t)s * RADE_NMF;
  const long long lw = link_ring ? link_wr[s] : 0;
  const int omask = link_ring ? LINK_CAP - 1 : 0x7fffffff;
  // a full FIFO (the receiver is not keeping up) drops the frame instead of overwriting unread samples, and says so
  const bool full = link_ring && link_rd && lw - link_rd[s] + RADE_NMF > LINK_CAP;
  for (int i = tid; i < RADE_NMF; i += blockDim.x) {
    const float a = (float)i * (1.f / RADE_NMF);
    const float2 g1 = make_float2(g[0].x + a * (g[1].x - g[0].x), g[0].y + a * (g[1].y - g[0].y));
    const float2 g2 = make_float2(g[2].x + a * (g[3].x - g[2].x), g[2].y + a * (g[3].y - g[2].y));
    float2 mp = cmul(stx[64 + i], g1);
    const float2 e = cmul(stx[64 + i - d], g2);   // g2 of the current instant: gains vary by <1e-3 over the 2 ms delay
    mp.x += e.x; mp.y += e.y;
    const float2 v = cmul(mp, make_float2((float)cs, (float)sn));
    const float2 nz = cnormal(seed, s, (unsigned long long)(t0 + i), 0xA11CEu);
    if (!full) out[(int)((lw + i) & omask)] = make_float2(gain * (v.x + sigma * nz.x), gain * (v.y + sigma * nz.y));
    const double c2 = cs * cb - sn * sb, s2 = sn * cb + cs * sb;
    cs = c2; sn = s2;
  }
  __syncthreads();
  for (int i = tid; i < 64; i += blockDim.x) cs_.delay[i] = stx[RADE_NMF + i];
  if (link_ring) __threadfence();               // samples before the write pointer: the receiver may be running concurrently
  __syncthreads();
  if (tid == 0) { cs_.t = t0 + RADE_NMF; cs_.phase = fmod(ph0 + dphi * RADE_NMF, 2.0 * M_PI); if (link_ring && !full) link_wr[s] = lw + RADE_NMF; if (full && link_overflow) atomicAdd(link_overflow, 1); }
}

// ---------------------------------------------------------------- loop-back link (per-stream FIFO)

__global__ void link_push_kernel(float2 *__restrict__ ring, long long *__restrict__ wr, const long long *__restrict__ rd, int *__restrict__ overflow,
                                 const float2 *__restrict__ in, int S) {
  const int s = blockIdx.x;
  const long long w = wr[s];
  if (rd && w - rd[s] + RADE_NMF > LINK_CAP) { if (threadIdx.x == 0 && overflow) atomicAdd(overflow, 1); return; }   // full: drop, flag
  for (int i = threadIdx.x; i < RADE_NMF; i += blockDim.x)
    ring[(size_t)s * LINK_CAP + ((w + i) & (LINK_CAP - 1))] = in[(size_t)s * RADE_NMF + i];
  __syncthreads();
  if (threadIdx.x == 0) wr[s] = w + RADE_NMF;
}

__global__ void link_pop_kernel(const float2 *__restrict__ ring, const long long *__restrict__ wr, long long *__restrict__ rd,
                                const RxCtl *__restrict__ ctl, float2 *__restrict__ out, unsigned char *__restrict__ active, int S) {
  const int s = blockIdx.x;
  const int nin = ctl[s].nin;
  const long long r = rd[s];
  const bool ok = wr[s] - r >= nin;
  if (ok)
    for (int i = threadIdx.x; i < nin; i += blockDim.x)
      out[(size_t)s * RADE_NIN_MAX + i] = ring[(size_t)s * LINK_CAP + ((r + i) & (LINK_CAP - 1))];
  __syncthreads();
  if (threadIdx.x == 0) { active[s] = ok ? 1 : 0; if (ok) rd[s] = r + nin; }
}

}  // namespace

int channel_apply_launch(float2 *rx, const float2 *tx, const float2 *G1, const float2 *G2, const float2 *noise, int S, int n,
                         int d, float mp_gain, float freq, float df_dt, float phase0, float sigma, float gain, cudaStream_t stream) {
  const size_t total = (size_t)S * n;
  channel_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(rx, tx, G1, G2, noise, S, n, d, mp_gain, freq, df_dt, phase0, sigma, gain);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}

int channel_stream_launch(const DspTables &T, const float *z_mod, float2 *rx, const float2 *tx, ChanState *st, int S, float sigma, float freq0, float freq_spread,
                          float doppler, int d, float gain, unsigned long long seed, float2 *link_ring, long long *link_wr,
                          const long long *link_rd, int *link_overflow, cudaStream_t stream) {
  if (d < 0 || d > 64) return -1;
  channel_stream_kernel<<<S, 256, 0, stream>>>(T, z_mod, rx, tx, st, S, sigma, freq0, freq_spread, doppler, d, gain, seed, link_ring, link_wr, link_rd,
                                               link_overflow);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}

int link_push_launch(float2 *ring, long long *wr, const long long *rd, int *overflow, const float2 *in, int S, cudaStream_t stream) {
  link_push_kernel<<<S, 256, 0, stream>>>(ring, wr, rd, overflow, in, S);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}

int link_pop_launch(const float2 *ring, const long long *wr, long long *rd, const RxCtl *ctl, float2 *out,
                    unsigned char *active, int S, cudaStream_t stream) {
  link_pop_kernel<<<S, 256, 0, stream>>>(ring, wr, rd, ctl, out, active, S);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}
