// Streaming receiver DSP for S independent streams: band-pass filter, pilot acquisition (coarse grid search,
// fine refinement, sync check), frequency correction, OFDM demodulation, least-squares pilot equalisation,
// SNR estimate, end-of-over detection and the search/candidate/sync state machine.
//
// Replaces, per stream and per call of radae_rx.do_radae_rx (radae_rxe.py:171-330):
//   complex_bpf.bpf                 radae/dsp.py:63-102   (incl. the Ntap+1 memory quirk, dsp.py:96)
//   acquisition.detect_pilots       radae/dsp.py:178-231
//   acquisition.refine              radae/dsp.py:233-270  (complex128 correlations -> csingle, f outer / t inner, strict >)
//   acquisition.check_pilots        radae/dsp.py:273-320  (deterministic row schedule instead of np.random)
//   receiver_one.receiver_one       radae/dsp.py:487-526  (+ est_pilots :418-435, update_snr_est :438-456, do_pilot_eq_one :459-484)
//   the state machine               radae_rxe.py:248-297
//
// Data layout per stream in HBM: rx_buf as a 2112-sample RING (no 17 KB shift per call, only the nin new samples
// are written), 102-sample BPF history, two 960-entry row-sum vectors sum_f|Dt1|, sum_f|Dt2| (all the reference
// ever reads back from its two 960x40 complex grids = 614 KB/stream), a 128-byte control block.
// Kernels (one launch each per rade_rx call; every stream branches on its own state):
//   rx_bpf (all streams; builds the search list and the track list)
//     -> rx_refresh (streams in sync: 48-row refresh of the |Dt| row sums, fp32)   | two streams,
//     -> rx_track   (streams in sync: refine in complex128 + spot correlations)    | concurrently
//     -> rx_demod   (sync-state machine at its head, then frequency correction, DFT, pilot EQ)
//     -> rx_detect (persistent over the search list: coarse grid search) -> rx_finish (search / candidate state machine)
//   the search branch runs on a side stream concurrently with the sync branch; both join before the core decoder.
// The two search kernels evaluate the coarse grid in a low-rank basis (proj_tap / expand_*), the refine as moments of the
// window (refine_moments); rx_finish uses the same moments form with 18 terms for the +-10 Hz first fix (refine_first_fix).
#include "rade_common.h"
#include "rade_host.h"
#include "tma.cuh"

namespace {

enum { ST_SEARCH = 0, ST_CANDIDATE = 1, ST_SYNC = 2 };
// sqrt(-ln(Pacq_error / 5)) of radae/dsp.py:229, :300 for Pacq_error = 1e-4, 1e-5, as float64 evaluates them
constexpr double K_ACQ_1E4 = 3.2893431387452243, K_ACQ_1E5 = 3.622480279781289;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ double2 dcmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// Coarse-grid correlations D(+-f_k) = sum_n y[n] exp(+-j w_k n), y[n] = conj(x[n]) p[n], f_k = 2.5 k Hz, k = 0..20
// (acquisition.detect_pilots / check_pilots, radae/dsp.py:204-205, :291-295; the reference multiplies by p_w = exp(j w n) p).
// Only |D| is ever used, so the window may be centred: with n = 80 + m and n = 79 - m folded onto m = 0..79,
//   A_k = sum_m (y[80+m] + y[79-m]) cos(w_k (m + 1/2)),  B_k = sum_m (y[80+m] - y[79-m]) sin(w_k (m + 1/2)),  |D(+-f_k)| = |A_k +- j B_k|.
// The 21 cosine rows and the 20 sine rows are each spanned by RADE_SRANK = 6 basis vectors to 1e-9 (AcqTables, tables.cpp): a
// window is first projected on the 12 basis vectors (12 complex accumulators instead of 41), then expanded to the 40 grid points.
// Per window and timing offset: 80 x (6 + 12) packed FMAs + 21 x 12 for the expansion, against 160 x 56 for the direct sums —
// at the same distance from the exact (float64) correlations as the direct float32 sums (tests/test_oracle_dsp.py).
// Blackwell packed fp32: one FFMA2 does two independent IEEE FMAs on a register pair (SASS FFMA2 .F32x2)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                     rc = *reinterpret_cast<unsigned long long *>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
// y = conj(x) p with two packed FMAs; pp = (p.x, p.y, p.y, -p.x)
__device__ __forceinline__ float2 conj_mul_p(float2 x, float4 pp) {
  return ffma2(splat(x.x), make_float2(pp.x, pp.y), ffma2(splat(x.y), make_float2(pp.z, pp.w), make_float2(0.f, 0.f)));
}
struct Proj { float2 c[RADE_SRANK], s[RADE_SRANK]; };         // projections on the cosine / sine basis
__device__ __forceinline__ void proj_zero(Proj &P) {
#pragma unroll
  for (int r = 0; r < RADE_SRANK; r++) P.c[r] = P.s[r] = make_float2(0.f, 0.f);
}
// one folded tap pair of one window: ya = y[80 + m], yb = y[79 - m]
template <bool COS, bool SIN>
__device__ __forceinline__ void proj_tap(Proj &P, float2 ya, float2 yb, const float4 *bm /* basis[m] */) {
  if (COS) {
    const float2 e = ffma2(yb, splat(1.f), ya);                 // ya + yb as one packed op
    const float4 q0 = bm[0], q1 = bm[1];
    P.c[0] = ffma2(e, splat(q0.x), P.c[0]); P.c[1] = ffma2(e, splat(q0.y), P.c[1]); P.c[2] = ffma2(e, splat(q0.z), P.c[2]);
    P.c[3] = ffma2(e, splat(q0.w), P.c[3]); P.c[4] = ffma2(e, splat(q1.x), P.c[4]); P.c[5] = ffma2(e, splat(q1.y), P.c[5]);
  }
  if (SIN) {
    const float2 o = ffma2(yb, splat(-1.f), ya);                // ya - yb
    const float4 q2 = bm[2], q3 = bm[3];
    P.s[0] = ffma2(o, splat(q2.x), P.s[0]); P.s[1] = ffma2(o, splat(q2.y), P.s[1]); P.s[2] = ffma2(o, splat(q2.z), P.s[2]);
    P.s[3] = ffma2(o, splat(q2.w), P.s[3]); P.s[4] = ffma2(o, splat(q3.x), P.s[4]); P.s[5] = ffma2(o, splat(q3.y), P.s[5]);
  }
}
// A_k / B_k from the projections; ek = expand[k]
__device__ __forceinline__ float2 expand_cos(const Proj &P, const float4 *ek) {
  const float4 e0 = ek[0], e1 = ek[1];
  float2 a = make_float2(P.c[0].x * e0.x, P.c[0].y * e0.x);
  a = ffma2(P.c[1], splat(e0.y), a); a = ffma2(P.c[2], splat(e0.z), a); a = ffma2(P.c[3], splat(e0.w), a);
  a = ffma2(P.c[4], splat(e1.x), a); a = ffma2(P.c[5], splat(e1.y), a);
  return a;
}
__device__ __forceinline__ float2 expand_sin(const Proj &P, const float4 *ek) {
  const float4 e1 = ek[1], e2 = ek[2];
  float2 b = make_float2(P.s[0].x * e1.z, P.s[0].y * e1.z);
  b = ffma2(P.s[1], splat(e1.w), b); b = ffma2(P.s[2], splat(e2.x), b); b = ffma2(P.s[3], splat(e2.y), b);
  b = ffma2(P.s[4], splat(e2.z), b); b = ffma2(P.s[5], splat(e2.w), b);
  return b;
}
__device__ __forceinline__ float mag2(float x, float y) { return sqrtf(fmaf(x, x, y * y)); }
// |D(+f_k)|, |D(-f_k)| from A, B
__device__ __forceinline__ void mags_pm(float2 A, float2 B, float &mp, float &mm) {
  mp = mag2(A.x - B.y, A.y + B.x);
  mm = mag2(A.x + B.y, A.y - B.x);
}
__device__ __forceinline__ int ring_idx(int head, int i) { int k = head + i; return k >= RADE_RXBUF ? k - RADE_RXBUF : k; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The tables every lane reads at the same index live in constant memory: they reach the FP32 pipe through the uniform datapath
// (ULDC / constant operands) and leave the shared-memory return path (128 B/clk per SM, which a broadcast LDS.128 occupies for
// four cycles like any other) to the per-lane sample reads.  Measured on rx_track: the row refresh was bound by exactly those
// broadcast loads (14.3k cycles per stream with the tables in shared memory).
struct AcqConst {
  float4 ps4[RADE_M];              // = AcqTables::ps4
  float4 basis[RADE_M / 2][4];     // = AcqTables::basis
  float4 expand[21][3];            // = AcqTables::expand
};
__constant__ AcqConst c_acq;


// three block-wide sums at once (one pass of barriers), results valid in every thread; scratch: >= 96 floats
__device__ void block_sum3(float &a, float &b, float &c, float *scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  __syncthreads();
  if (lane == 0) { scratch[w] = a; scratch[32 + w] = b; scratch[64 + w] = c; }
  __syncthreads();
  if (w < 3) {
    float t = (lane < nw) ? scratch[32 * w + lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) scratch[32 * w] = t;
  }
  __syncthreads();
  a = scratch[0]; b = scratch[32]; c = scratch[64];
}

// ================================================================= band-pass filter + ring append
// 101-tap real FIR on complex samples.  Each thread produces four consecutive outputs from a sliding register window
// (one LDS.128 = two new samples per two taps) with the taps as constant-bank operands of the FMAs.
__constant__ float c_bpf_h[RADE_BPF_NTAP + 3];
constexpr int BPF_THREADS = 288;                  // 4 outputs per thread, nin <= 1120
__global__ void __launch_bounds__(BPF_THREADS)
rx_bpf_kernel(DspTables T, RxCtl *__restrict__ ctl, float2 *__restrict__ ring, float2 *__restrict__ bpf_mem,
              const float2 *__restrict__ rx_in, const unsigned char *__restrict__ active, int bpf_en,
              int *__restrict__ search_list, int *__restrict__ track_list, int *__restrict__ counters, LinkSrc link) {
  __shared__ __align__(16) float2 X[RADE_BPF_MEM + RADE_NIN_MAX + 42];
  constexpr int LINK_CAP = 4096;
  const int s = blockIdx.x, tid = threadIdx.x;
  RxCtl &c = ctl[s];
  const int nin = c.nin, head = c.ring_head;
  // input: row s of the caller's [S][1120] array, or nin samples popped from the stream's link FIFO
  const float2 *xin = rx_in + (size_t)s * RADE_NIN_MAX;
  int xoff = 0, xmask = 0x7fffffff;
  if (link.ring) {
    const long long r = link.rd[s];
    const bool ok = *reinterpret_cast<const volatile long long *>(&link.wr[s]) - r >= nin;
    __threadfence();                                    // write pointer before samples (the channel may be running concurrently)
    __syncthreads();                                    // everybody has read rd before it moves
    if (tid == 0) { link.active_out[s] = ok ? 1 : 0; if (ok) link.rd[s] = r + nin; }
    if (!ok) return;
    xin = link.ring + (size_t)s * LINK_CAP; xoff = (int)(r & (LINK_CAP - 1)); xmask = LINK_CAP - 1;
  } else if (active && !active[s]) return;
  float2 *rg = ring + (size_t)s * RADE_RXBUF;
  if (bpf_en) {
    const float2 ph = c.bpf_phase;
    const int off = c.bpf_first ? 2 : 0;
    float2 *mem = bpf_mem + (size_t)s * RADE_BPF_MEM;
    for (int i = tid; i < RADE_BPF_MEM; i += BPF_THREADS) X[i] = mem[i];
    for (int i = tid; i < nin; i += BPF_THREADS) X[RADE_BPF_MEM + i] = cmul(xin[(xoff + i) & xmask], cmul(ph, T.bpf_exp[i]));   // mix down
    for (int i = RADE_BPF_MEM + nin + tid; i < RADE_BPF_MEM + RADE_NIN_MAX + 42; i += BPF_THREADS) X[i] = make_float2(0.f, 0.f);
    __syncthreads();
    const int i0 = 4 * tid;
    if (i0 < nin) {
      // out[i0 + q] = sum_k h[k] X[i0 + q + k + off], k ascending (same summation order as a plain tap loop)
      const float4 *Xv = reinterpret_cast<const float4 *>(X + i0 + off);
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      float4 w0 = Xv[0], w1 = Xv[1];              // samples k..k+1, k+2..k+3
      // real tap x complex sample = one packed FMA on the (re, im) pair (SASS FFMA2): half the FP32-pipe time of two scalar
      // FMAs, the same two IEEE operations
#pragma unroll
      for (int k = 0; k < RADE_BPF_NTAP + 1; k += 2) {
        const float4 w2 = Xv[k / 2 + 2];          // samples k+4, k+5
        const float2 h0 = splat(c_bpf_h[k]), h1 = splat(c_bpf_h[k + 1]);        // h[101] = 0 pads the odd tap count
        acc[0] = ffma2(h0, make_float2(w0.x, w0.y), acc[0]);
        acc[1] = ffma2(h0, make_float2(w0.z, w0.w), acc[1]);
        acc[2] = ffma2(h0, make_float2(w1.x, w1.y), acc[2]);
        acc[3] = ffma2(h0, make_float2(w1.z, w1.w), acc[3]);
        if (k + 1 < RADE_BPF_NTAP) {
          acc[0] = ffma2(h1, make_float2(w0.z, w0.w), acc[0]);
          acc[1] = ffma2(h1, make_float2(w1.x, w1.y), acc[1]);
          acc[2] = ffma2(h1, make_float2(w1.z, w1.w), acc[2]);
          acc[3] = ffma2(h1, make_float2(w2.x, w2.y), acc[3]);
        }
        w0 = w1; w1 = w2;
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (i0 + q < nin) rg[ring_idx(head, i0 + q)] = cmul(acc[q], cconj(cmul(ph, T.bpf_exp[i0 + q])));      // mix up
    }
    __syncthreads();
    for (int i = tid; i < RADE_BPF_MEM; i += BPF_THREADS) mem[i] = X[nin + i];
    if (tid == 0) { c.bpf_phase = cmul(ph, T.bpf_exp[nin - 1]); c.bpf_first = 0; }
  } else {
    for (int i = tid; i < nin; i += BPF_THREADS) rg[ring_idx(head, i)] = xin[(xoff + i) & xmask];
  }
  if (tid == 0) {
    int nh = head + nin; if (nh >= RADE_RXBUF) nh -= RADE_RXBUF;
    c.ring_head = nh;                      // logical sample 0 of rx_buf now lives at ring[nh]
    c.detect_key = 0ull;
    c.candidate = 0; c.endofover = 0; c.valid_output = 0; c.uw_fail = 0; c.ran_sync = 0; c.ret = 0;
    c.tracking = c.state == ST_SYNC;
    if (!c.tracking) search_list[atomicAdd(&counters[0], 1)] = s;                  // work list of rx_detect / rx_finish
    else track_list[atomicAdd(&counters[2], 1)] = s;                              // work list of rx_refresh / rx_track
  }
}

// ================================================================= coarse pilot search (search / candidate streams)
// work item = 64 timing offsets x 40 frequency offsets x 2 pilot positions.  One thread per timing offset: both windows are
// projected on the 12 basis vectors in one pass over the 80 folded taps (48 packed accumulators), then expanded to the grid.
// The tables are in constant memory; 64-thread CTAs with 3.6 KB of shared memory: many are resident per SM, next to rx_track.
constexpr int DET_TB = 64, DET_THREADS = DET_TB;
struct DetectSmem {
  float2 r1[DET_TB + RADE_M];
  float2 r2[DET_TB + RADE_M];
  unsigned long long best[DET_THREADS / 32];
};

__global__ void __launch_bounds__(DET_THREADS)
rx_detect_kernel(DspTables T, RxCtl *__restrict__ ctl, const float2 *__restrict__ ring, float *__restrict__ rowsum,
                 const int *__restrict__ search_list, int *__restrict__ counters) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DetectSmem &sm = *reinterpret_cast<DetectSmem *>(smem_raw);
  __shared__ int work;
  const int tid = threadIdx.x;
  const int n_items = counters[0] * (RADE_NMF / DET_TB);
  if ((int)blockIdx.x >= n_items) return;         // steady state: (almost) nobody is searching
  // persistent CTAs pull (stream, 64-offset block) items off a device-side counter
  for (;;) {
    __syncthreads();
    if (tid == 0) work = atomicAdd(&counters[1], 1);
    __syncthreads();
    const int w = work;
    if (w >= n_items) break;
    const int s = search_list[w / (RADE_NMF / DET_TB)], t0 = (w % (RADE_NMF / DET_TB)) * DET_TB;
    RxCtl &c = ctl[s];
    const int head = c.ring_head;
    const float2 *rg = ring + (size_t)s * RADE_RXBUF;
    for (int i = tid; i < DET_TB + RADE_M; i += blockDim.x) {
      sm.r1[i] = rg[ring_idx(head, t0 + i)];
      sm.r2[i] = rg[ring_idx(head, t0 + RADE_NMF + i)];
    }
    __syncthreads();
    Proj P1, P2;
    proj_zero(P1); proj_zero(P2);
    const float2 *x1 = &sm.r1[tid], *x2 = &sm.r2[tid];
#pragma unroll 2
    for (int m = 0; m < RADE_M / 2; m++) {
      const float4 pa = c_acq.ps4[RADE_M / 2 + m], pb = c_acq.ps4[RADE_M / 2 - 1 - m];
      proj_tap<true, true>(P1, conj_mul_p(x1[RADE_M / 2 + m], pa), conj_mul_p(x1[RADE_M / 2 - 1 - m], pb), c_acq.basis[m]);
      proj_tap<true, true>(P2, conj_mul_p(x2[RADE_M / 2 + m], pa), conj_mul_p(x2[RADE_M / 2 - 1 - m], pb), c_acq.basis[m]);
    }
    float s1 = 0.f, s2 = 0.f, best = -1.f; int bestf = RADE_NFCOARSE;
#pragma unroll 3
    for (int k = 0; k <= 20; k++) {
      const float4 *ek = c_acq.expand[k];
      float p1, m1, p2, m2;
      mags_pm(expand_cos(P1, ek), expand_sin(P1, ek), p1, m1);
      mags_pm(expand_cos(P2, ek), expand_sin(P2, ek), p2, m2);
      if (k > 0) {                                    // -2.5k Hz -> grid index 20 - k
        s1 += m1; s2 += m2;
        const float d = m1 + m2; const int fi = 20 - k;
        if (d > best || (d == best && fi < bestf)) { best = d; bestf = fi; }
      }
      if (k < 20) {                                   // +2.5k Hz -> grid index 20 + k
        s1 += p1; s2 += p2;
        const float d = p1 + p2; const int fi = 20 + k;
        if (d > best || (d == best && fi < bestf)) { best = d; bestf = fi; }
      }
    }
    rowsum[((size_t)s * 2 + 0) * RADE_NMF + t0 + tid] = s1;
    rowsum[((size_t)s * 2 + 1) * RADE_NMF + t0 + tid] = s2;
    // arg-max with "first (t, f) wins": larger key = larger value, then smaller flat index
    unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) | (0xFFFFFFFFu - (unsigned)((t0 + tid) * RADE_NFCOARSE + bestf));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o); key = k2 > key ? k2 : key; }
    if ((tid & 31) == 0) sm.best[tid >> 5] = key;
    __syncthreads();
    if (tid == 0) {
      unsigned long long k = sm.best[0];
      for (int i = 1; i < DET_THREADS / 32; i++) k = sm.best[i] > k ? sm.best[i] : k;
      atomicMax(&c.detect_key, k);
    }
  }
}

// ================================================================= fine timing / frequency refinement (shared by track & finish)
// acquisition.refine (radae/dsp.py:233-270): argmax over (f outer, t inner) of |Dt1 + Dt2|, strict >, with
//   Dt1[t,f] = sum_n rx[t+n] conj(p[n]) exp(-j w_f n),  Dt2 = same at t + Nmf times exp(-j w_f Nmf),
// computed in complex128 and rounded to csingle like the reference.  Here: moments of the window in complex128 (refine_moments
// for the tracking search, refine_first_fix for the first fix after acquisition) instead of one steering vector per frequency.
constexpr int REF_NT = 16;                        // max timing offsets
constexpr int REF_NFP = 24;                       // room for the 20 or 21 frequencies of arange(f0 - 1, f0 + 1, 0.1)
constexpr int REF_RLEN = REF_NT + RADE_M + 8;     // widened window
constexpr int REF_THREADS = 128;                  // four warps
constexpr int REF_NK = 9;                         // Taylor terms of exp(-j dw n'), |dw n'| <= 0.0625: truncation 6e-15
struct MomentsSmem {                              // refine_moments (tracking)
  double2 q[RADE_M];                              // conj(p[n]) exp(-j w0 (n - 79.5)), w0 = the tracked frequency
  double2 part[4][32][REF_NK];                    // moments of (window, quarter of the taps)
  double2 ph0[2];                                 // exp(-j w0 (79.5 + 960 pos))
  double2 ra[REF_RLEN];                           // rx[t_lo ...] widened once (np.dot up-casts csingle to complex128)
  double2 pad_;
  double2 rb[REF_RLEN];                           // rx[t_lo + Nmf ...]
  float2 d1[REF_NFP][REF_NT];
  float2 d2[REF_NFP][REF_NT];
  float red_mag[4]; int red_ord[4];
  float best_mag; int best_t; int best_found; double best_f;
};
// values of np.arange(start, stop, step) for float64: v[i] = start + i*((start+step)-start), len = ceil((stop-start)/step)
__device__ __forceinline__ int arange_len(double start, double stop, double step) { return (int)ceil((stop - start) / step); }

// acquisition.refine for the tracking case — 16 timing offsets, f in arange(f0 - 1, f0 + 1, 0.1) — without a steering vector per
// frequency.  With n' = n - 79.5 and f = f0 + d:  exp(-j w_f n) = exp(-j w_f 79.5) exp(-j w0 n') exp(-j dw n'), and |dw n'| <= 0.0625 rad,
// so exp(-j dw n') = sum_k (-j 80 dw)^k (n'/80)^k / k! to 6e-15 with k <= 8.  Per window: nine complex128 moments
//   M_k = sum_n rx[t + n] q[n] (n'/80)^k / k!,   q[n] = conj(p[n]) exp(-j w0 n'),
// then every frequency is a Horner evaluation in u = -j 80 dw and one phase rotation.  FP64 work per stream: 32 windows x 160 taps
// x 22 FMA = 113k against 491k for the matrix form (20 steering vectors), and the 24 x 5 double sincos chains of the steering
// table are gone (two sincos per thread, in parallel).  Results agree with the complex128 sums to ~1e-14 relative before the
// rounding to csingle (radae/dsp.py:255-257 is reproduced from there on: d1, d2 rounded separately, then |csingle(d1 + d2)|).
template <typename Load>
__device__ void refine_moments(MomentsSmem &sm, const double (*bk)[10], const double2 (*phd)[24], const double2 *pcd, Load load, int t_lo, int nt, double f0,
                               int g, int bar) {
  const double f_start = f0 - 1, f_stop = f0 + 1, f_step = 0.1;
  const int nf_all = min(arange_len(f_start, f_stop, f_step), REF_NFP);
  const double delta = (f_start + f_step) - f_start;
  const double w0 = 2.0 * M_PI * f0 / RADE_FS;
  if (g == 0) { sm.best_mag = 0.f; sm.best_found = 0; sm.best_t = 0; sm.best_f = 0.0; }
  for (int i2 = g; i2 < REF_RLEN; i2 += REF_THREADS) {
    const bool in = i2 < nt + RADE_M;
    const float2 a = in ? load(t_lo + i2) : make_float2(0.f, 0.f), c = in ? load(t_lo + RADE_NMF + i2) : make_float2(0.f, 0.f);
    sm.ra[i2] = make_double2((double)a.x, (double)a.y);
    sm.rb[i2] = make_double2((double)c.x, (double)c.y);
  }
  {
    // q[n] for n = g (and g + 128), the two stream-dependent phases by threads 32 and 33: at most two sincos per thread
    double sn, cs;
    sincos(w0 * ((double)g - 79.5), &sn, &cs);
    sm.q[g] = dcmul(make_double2(cs, -sn), pcd[g]);
    if (g < RADE_M - REF_THREADS + 2) {
      const double arg = g < RADE_M - REF_THREADS ? (double)(g + REF_THREADS) - 79.5 : (g == RADE_M - REF_THREADS ? 79.5 : 79.5 + RADE_NMF);
      sincos(w0 * arg, &sn, &cs);
      if (g < RADE_M - REF_THREADS) sm.q[g + REF_THREADS] = dcmul(make_double2(cs, -sn), pcd[g + REF_THREADS]);
      else sm.ph0[g - (RADE_M - REF_THREADS)] = make_double2(cs, -sn);
    }
  }
  group_sync(bar, REF_THREADS);
  const int wdx = g & 31, qt = g >> 5;               // window = (pilot position, timing offset), quarter of the taps
  const int pos = wdx >> 4, ti = wdx & 15;
  {
    double2 M[REF_NK];
#pragma unroll
    for (int k = 0; k < REF_NK; k++) M[k] = make_double2(0.0, 0.0);
    const double2 *xr = (pos ? sm.rb : sm.ra) + ti;
#pragma unroll 2
    for (int n = qt * (RADE_M / 4); n < (qt + 1) * (RADE_M / 4); n++) {
      const double2 xv = xr[n], qv = sm.q[n];
      const double2 z = dcmul(xv, qv);
      const double *b = bk[n];
#pragma unroll
      for (int k = 0; k < REF_NK; k++) { M[k].x = fma(z.x, b[k], M[k].x); M[k].y = fma(z.y, b[k], M[k].y); }
    }
#pragma unroll
    for (int k = 0; k < REF_NK; k++) sm.part[qt][wdx][k] = M[k];
  }
  group_sync(bar, REF_THREADS);
  {
    // thread = (window, frequencies qt, qt + 4, ...): add the four partial moments, Horner per frequency, phase, round to csingle
    double2 M[REF_NK];
#pragma unroll
    for (int k = 0; k < REF_NK; k++) {
      const double2 a = sm.part[0][wdx][k], b = sm.part[1][wdx][k], c = sm.part[2][wdx][k], d = sm.part[3][wdx][k];
      M[k] = make_double2((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y));
    }
    const double2 ph0 = sm.ph0[pos];
    if (ti < nt)
      for (int fi = qt; fi < nf_all; fi += 4) {
        const double f = f_start + (double)fi * delta;
        const double al = 2.0 * M_PI * (f - f0) / RADE_FS * 80.0;          // u = -j al
        double2 D = M[REF_NK - 1];
#pragma unroll
        for (int k = REF_NK - 2; k >= 0; k--) D = make_double2(fma(al, D.y, M[k].x), fma(-al, D.x, M[k].y));   // D (-j al) + M_k
        const double2 e = dcmul(dcmul(D, ph0), phd[pos][fi]);
        (pos ? sm.d2 : sm.d1)[fi][ti] = make_float2((float)e.x, (float)e.y);
      }
  }
  group_sync(bar, REF_THREADS);
  float bm = -1.f; int bo = 0x7fffffff;
  for (int q = g; q < nf_all * nt; q += REF_THREADS) {
    const int fi = q / nt, t2 = q - fi * nt;         // ord = q: f outer loop, t inner loop
    const float2 a = sm.d1[fi][t2], c = sm.d2[fi][t2];
    const float m = hypotf(a.x + c.x, a.y + c.y);
    if (m > bm || (m == bm && q < bo)) { bm = m; bo = q; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, bm, o); const int o2 = __shfl_xor_sync(0xffffffffu, bo, o);
    if (m2 > bm || (m2 == bm && o2 < bo)) { bm = m2; bo = o2; }
  }
  if ((g & 31) == 0) { sm.red_mag[g >> 5] = bm; sm.red_ord[g >> 5] = bo; }
  group_sync(bar, REF_THREADS);
  if (g == 0) {
    for (int q = 1; q < REF_THREADS / 32; q++)
      if (sm.red_mag[q] > bm || (sm.red_mag[q] == bm && sm.red_ord[q] < bo)) { bm = sm.red_mag[q]; bo = sm.red_ord[q]; }
    if (bm > 0.f) {                 // strict > against the initial Dtmax = 0 (radae/dsp.py:262)
      sm.best_mag = bm; sm.best_found = 1;
      sm.best_t = t_lo + bo % nt;
      sm.best_f = f_start + (double)(bo / nt) * delta;
    }
  }
  group_sync(bar, REF_THREADS);
}

// acquisition.refine for the first fix after acquisition (radae_rxe.py:267-273): t in [max(0, tmax - 1), tmax + 2),
// f in arange(f0 - 10, f0 + 10, 0.25) — the same moments form with |80 dw| <= 0.63, i.e. 18 Taylor terms (truncation 4e-20).
// Only six windows, so the layout is by task instead of by window: (1) q[n], the power table (n'/80)^k / k! and the two
// stream-dependent phases, (2) z[w][n] = rx q, (3) one thread per (window, k) moment, (4) Horner + phase per (frequency, window),
// (5) arg-max.  About 8 k cycles per stream instead of four passes of steering table + FP64 tensor-core GEMM (~45 k): the search
// branch (rx_detect -> rx_finish) used to end after rx_demod and held the core decoder back (tools/step_timeline.py).
constexpr int FF_NK = 18, FF_NF = 80, FF_NT = 3, FF_NW = 2 * FF_NT;
struct FirstFixSmem {
  double bk[RADE_M][FF_NK];
  double2 z[FF_NW][RADE_M];
  double2 q[RADE_M];
  double2 M[FF_NW][FF_NK];
  double2 ph0[2];
  float2 d1[FF_NF][FF_NT], d2[FF_NF][FF_NT];
  float red_mag[4]; int red_ord[4];
  float best_mag; int best_t; int best_found; double best_f;
};
template <typename Load>
__device__ void refine_first_fix(FirstFixSmem &sm, const AcqTables &tab, Load load, int t_lo, int nt, double f0, int g, int bar) {
  const double f_start = f0 - 10, f_stop = f0 + 10, f_step = 0.25;
  const int nf_all = min(arange_len(f_start, f_stop, f_step), FF_NF);
  const double delta = (f_start + f_step) - f_start;
  const double w0 = 2.0 * M_PI * f0 / RADE_FS;
  const int nw = 2 * nt;
  if (g == 0) { sm.best_mag = 0.f; sm.best_found = 0; sm.best_t = 0; sm.best_f = 0.0; }
  {
    double sn, cs;
    sincos(w0 * ((double)g - 79.5), &sn, &cs);
    sm.q[g] = dcmul(make_double2(cs, -sn), make_double2(__ldg(&tab.pcd[g].x), __ldg(&tab.pcd[g].y)));
    if (g < RADE_M - REF_THREADS + 2) {
      const double arg = g < RADE_M - REF_THREADS ? (double)(g + REF_THREADS) - 79.5 : (g == RADE_M - REF_THREADS ? 79.5 : 79.5 + RADE_NMF);
      sincos(w0 * arg, &sn, &cs);
      if (g < RADE_M - REF_THREADS) sm.q[g + REF_THREADS] = dcmul(make_double2(cs, -sn), make_double2(__ldg(&tab.pcd[g + REF_THREADS].x), __ldg(&tab.pcd[g + REF_THREADS].y)));
      else sm.ph0[g - (RADE_M - REF_THREADS)] = make_double2(cs, -sn);
    }
    for (int n = g; n < RADE_M; n += REF_THREADS) {
      const double sc = ((double)n - 79.5) / 80.0;
      double v = 1.0;
#pragma unroll
      for (int k = 0; k < FF_NK; k++) { sm.bk[n][k] = v; v = v * sc / (double)(k + 1); }
    }
  }
  group_sync(bar, REF_THREADS);
  for (int i = g; i < nw * RADE_M; i += REF_THREADS) {
    const int w = i / RADE_M, n = i - w * RADE_M, pos = w / nt, ti = w - pos * nt;
    const float2 x = load(t_lo + ti + pos * RADE_NMF + n);
    sm.z[w][n] = dcmul(make_double2((double)x.x, (double)x.y), sm.q[n]);
  }
  group_sync(bar, REF_THREADS);
  if (g < nw * FF_NK) {
    const int w = g / FF_NK, k = g - w * FF_NK;
    double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
#pragma unroll 4
    for (int n = 0; n < RADE_M; n += 2) {
      const double2 z0 = sm.z[w][n], z1 = sm.z[w][n + 1];
      const double b0 = sm.bk[n][k], b1 = sm.bk[n + 1][k];
      ax = fma(z0.x, b0, ax); ay = fma(z0.y, b0, ay); bx = fma(z1.x, b1, bx); by = fma(z1.y, b1, by);
    }
    sm.M[w][k] = make_double2(ax + bx, ay + by);
  }
  group_sync(bar, REF_THREADS);
  for (int i = g; i < nf_all * nw; i += REF_THREADS) {
    const int fi = i / nw, w = i - fi * nw, pos = w / nt, ti = w - pos * nt;
    const double f = f_start + (double)fi * delta;
    const double al = 2.0 * M_PI * (f - f0) / RADE_FS * 80.0;              // u = -j al
    double2 D = sm.M[w][FF_NK - 1];
#pragma unroll
    for (int k = FF_NK - 2; k >= 0; k--) D = make_double2(fma(al, D.y, sm.M[w][k].x), fma(-al, D.x, sm.M[w][k].y));
    const double2 pd = make_double2(__ldg(&tab.phd10[pos][fi].x), __ldg(&tab.phd10[pos][fi].y));
    const double2 e = dcmul(dcmul(D, sm.ph0[pos]), pd);
    (pos ? sm.d2 : sm.d1)[fi][ti] = make_float2((float)e.x, (float)e.y);
  }
  group_sync(bar, REF_THREADS);
  float bm = -1.f; int bo = 0x7fffffff;
  for (int q = g; q < nf_all * nt; q += REF_THREADS) {
    const int fi = q / nt, t2 = q - fi * nt;         // ord = q: f outer loop, t inner loop
    const float2 a = sm.d1[fi][t2], c = sm.d2[fi][t2];
    const float m = hypotf(a.x + c.x, a.y + c.y);
    if (m > bm || (m == bm && q < bo)) { bm = m; bo = q; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, bm, o); const int o2 = __shfl_xor_sync(0xffffffffu, bo, o);
    if (m2 > bm || (m2 == bm && o2 < bo)) { bm = m2; bo = o2; }
  }
  if ((g & 31) == 0) { sm.red_mag[g >> 5] = bm; sm.red_ord[g >> 5] = bo; }
  group_sync(bar, REF_THREADS);
  if (g == 0) {
    for (int q = 1; q < REF_THREADS / 32; q++)
      if (sm.red_mag[q] > bm || (sm.red_mag[q] == bm && sm.red_ord[q] < bo)) { bm = sm.red_mag[q]; bo = sm.red_ord[q]; }
    if (bm > 0.f) {                 // strict > against the initial Dtmax = 0 (radae/dsp.py:262)
      sm.best_mag = bm; sm.best_found = 1;
      sm.best_t = t_lo + bo % nt;
      sm.best_f = f_start + (double)(bo / nt) * delta;
    }
  }
  group_sync(bar, REF_THREADS);
}

// sigma_r = (mean|Dt1| + mean|Dt2|) / (2*sqrt(pi/2)) from the row sums, float32 like the reference's np.mean
__device__ float sigma_r_from_rowsums(const float *rs /* [2][960] */, float *scratch /* >= 96 floats */) {
  float a = 0.f, b = 0.f, z = 0.f;
  for (int i = threadIdx.x; i < RADE_NMF; i += blockDim.x) { a += __ldg(rs + i); b += __ldg(rs + RADE_NMF + i); }
  block_sum3(a, b, z, scratch);
  const float k = 1.2533141373155001f;              // (pi/2)**0.5 as float32
  const float s1 = (a / (float)(RADE_NMF * RADE_NFCOARSE)) / k, s2 = (b / (float)(RADE_NMF * RADE_NFCOARSE)) / k;
  return (s1 + s2) / 2.0f;
}

// ================================================================= sync-state tracking: check_pilots row refresh | refine + spot correlations
// Streams in sync (the list rx_bpf builds).  Two kernels with many small CTAs per SM instead of one persistent CTA per SM: every
// phase of the per-stream work is a short dependent chain (FP64 latency, constant-bank latency), and with one warp per scheduler
// nothing hid it — the persistent kernel spent 25k cycles per stream with every pipe under 30 % (profiles/r02_track_phases.txt).
//   rx_refresh  CTA = one stream, 96 threads = (row, pilot position): 48 rows t_i = 20 i + rot of both |Dt| row-sum vectors
//   rx_track    CTA = one stream, 128 threads: refine in complex128 (moments form) + the four spot correlations -> TrackTmp
// The sync-state part of the state machine needs both results; it runs at the head of rx_demod (track_tail).
struct TrackTmp { double best_f; double spot[4]; int best_found, best_t, pad[2]; };

constexpr int RFR_STREAMS = 2, RFR_THREADS = RFR_STREAMS * RADE_NUPDATE;
__global__ void __launch_bounds__(RFR_THREADS)
rx_refresh_kernel(RxCtl *__restrict__ ctl, const float2 *__restrict__ ring, float *__restrict__ rowsum,
                  const int *__restrict__ track_list, const int *__restrict__ counters) {
  __shared__ __align__(16) float2 rx[RFR_STREAMS][RADE_RXBUF];
  const int tid = threadIdx.x, n_items = counters[2];
  for (int it0 = blockIdx.x * RFR_STREAMS; it0 < n_items; it0 += gridDim.x * RFR_STREAMS) {
    __syncthreads();
    for (int q = 0; q < RFR_STREAMS; q++) {
      if (it0 + q >= n_items) break;
      const int s = track_list[it0 + q], head = ctl[s].ring_head;
      const float2 *rg = ring + (size_t)s * RADE_RXBUF;
      for (int i = tid; i < RADE_RXBUF; i += RFR_THREADS) rx[q][i] = rg[ring_idx(head, i)];
    }
    __syncthreads();
    // check_pilots row refresh (radae/dsp.py:288-295, deterministic schedule): rows t_i = 20 i + rot, both pilot positions,
    // 40 grid frequencies.  thread = (stream of the pair, row i): both windows are projected on the 12 basis vectors in one pass
    // (the table values are loaded once for the two), then expanded to the grid.
    const int q = tid / RADE_NUPDATE, i = tid % RADE_NUPDATE;
    if (it0 + q < n_items) {
      const int s = track_list[it0 + q], rot = ctl[s].n_check % 20;
      const float2 *x1 = rx[q] + rot + 20 * i, *x2 = x1 + RADE_NMF;
      Proj P1, P2; proj_zero(P1); proj_zero(P2);
#pragma unroll 2
      for (int m = 0; m < RADE_M / 2; m++) {
        const float4 pa = c_acq.ps4[RADE_M / 2 + m], pb = c_acq.ps4[RADE_M / 2 - 1 - m];
        proj_tap<true, true>(P1, conj_mul_p(x1[RADE_M / 2 + m], pa), conj_mul_p(x1[RADE_M / 2 - 1 - m], pb), c_acq.basis[m]);
        proj_tap<true, true>(P2, conj_mul_p(x2[RADE_M / 2 + m], pa), conj_mul_p(x2[RADE_M / 2 - 1 - m], pb), c_acq.basis[m]);
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 3
      for (int k = 0; k <= 20; k++) {
        const float4 *ek = c_acq.expand[k];
        float p1, m1, p2, m2;
        mags_pm(expand_cos(P1, ek), expand_sin(P1, ek), p1, m1);
        mags_pm(expand_cos(P2, ek), expand_sin(P2, ek), p2, m2);
        if (k > 0) { s1 += m1; s2 += m2; }              // grid index 20 - k
        if (k < 20) { s1 += p1; s2 += p2; }             // grid index 20 + k
      }
      float *rs = rowsum + (size_t)s * 2 * RADE_NMF + 20 * i + rot;
      rs[0] = s1; rs[RADE_NMF] = s2;
    }
  }
}

struct TrackSmem {
  MomentsSmem ref;
  double bk[RADE_M][10];                          // AcqTables::bk, staged once per CTA
  double2 phd[2][24];
  double2 pcd[RADE_M];
  float2 pp[RADE_M], pend[RADE_M];                // pilot / end-of-over pilot (spot correlations)
  double2 spot_e[32], spot_step;                  // exp(-j w lane), exp(-j 32 w): the same for the four spot correlations
};
__global__ void __launch_bounds__(REF_THREADS)
rx_track_kernel(DspTables T, const RxCtl *__restrict__ ctl, const float2 *__restrict__ ring, const int *__restrict__ track_list,
                const int *__restrict__ counters, TrackTmp *__restrict__ tmp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TrackSmem &sm = *reinterpret_cast<TrackSmem *>(smem_raw);
  const int tid = threadIdx.x, n_items = counters[2];
  if ((int)blockIdx.x >= n_items) return;
  {
    const AcqTables &tab = *reinterpret_cast<const AcqTables *>(T.acq_tab);
    for (int i = tid; i < RADE_M * 10; i += REF_THREADS) (&sm.bk[0][0])[i] = __ldg(&tab.bk[0][0] + i);
    for (int i = tid; i < 2 * 24 * 2; i += REF_THREADS) reinterpret_cast<double *>(sm.phd)[i] = __ldg(reinterpret_cast<const double *>(tab.phd) + i);
    for (int i = tid; i < RADE_M; i += REF_THREADS) {
      sm.pcd[i] = make_double2(__ldg(&tab.pcd[i].x), __ldg(&tab.pcd[i].y));
      const float4 p4 = __ldg(&tab.ps4[i]);
      sm.pp[i] = make_float2(p4.x, p4.y); sm.pend[i] = __ldg(&tab.pend[i]);
    }
  }
  __syncthreads();
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int s = track_list[it];
    const int head = ctl[s].ring_head, tmax0 = ctl[s].tmax; const double fmax0 = ctl[s].fmax;
    const float2 *rg = ring + (size_t)s * RADE_RXBUF;
    auto load = [rg, head](int i) { return __ldg(rg + ring_idx(head, i)); };
    // ---- refine (radae_rxe.py:202-205, radae/dsp.py:233-270): t in [max(0,tmax-8), tmax+8), f in arange(fmax-1, fmax+1, .1)
    const int t_lo = max(0, tmax0 - 8);
    refine_moments(sm.ref, sm.bk, sm.phd, sm.pcd, load, t_lo, tmax0 + 8 - t_lo, fmax0, tid, 1);
    // ---- spot correlations in complex128 (radae/dsp.py:305-314): warp k -> |sum conj(w_vec*rx[o_k + n]) * q_k[n]| at the
    // refined timing / smoothed frequency
    {
      const int wp = tid >> 5, lane = tid & 31;
      const int tm = sm.ref.best_found ? sm.ref.best_t : tmax0;
      const double fm = 0.9 * fmax0 + 0.1 * (sm.ref.best_found ? sm.ref.best_f : fmax0);
      const int o = tm + (wp == 0 ? 0 : wp == 2 ? RADE_M + RADE_NCP : RADE_NMF);
      const double w = 2.0 * M_PI * fm / RADE_FS;
      float2 xv[5];
#pragma unroll
      for (int j = 0; j < 5; j++) xv[j] = load(o + lane + 32 * j);
      if (tid <= 32) {                               // 33 sincos instead of 256
        double sn, cs; sincos(w * (double)tid, &sn, &cs);
        if (tid < 32) sm.spot_e[tid] = make_double2(cs, -sn); else sm.spot_step = make_double2(cs, -sn);
      }
      group_sync(1, REF_THREADS);
      double2 e = sm.spot_e[lane]; const double2 step = sm.spot_step;
      double ax = 0.0, ay = 0.0;
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const int n = lane + 32 * j;
        const float2 qf = (wp < 2) ? sm.pp[n] : sm.pend[n];
        const double2 v = dcmul(e, make_double2((double)xv[j].x, (double)xv[j].y));                 // w_vec * rx
        const double2 r2 = dcmul(make_double2(v.x, -v.y), make_double2((double)qf.x, (double)qf.y));
        ax += r2.x; ay += r2.y;
        e = dcmul(e, step);
      }
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) { ax += __shfl_xor_sync(0xffffffffu, ax, q); ay += __shfl_xor_sync(0xffffffffu, ay, q); }
      if (lane == 0) tmp[s].spot[wp] = hypot(ax, ay);
      if (tid == 0) { tmp[s].best_found = sm.ref.best_found; tmp[s].best_t = sm.ref.best_t; tmp[s].best_f = sm.ref.best_f; }
    }
    group_sync(1, REF_THREADS);                     // best_* and spot_e are rewritten during the next item
  }
}

// Sync-state part of the state machine (radae_rxe.py:208-218, :276-296) and check_pilots' decisions (radae/dsp.py:297-320) for a
// stream that was tracked this call: called by every thread of the stream's rx_demod CTA before the demodulation proper.
__device__ void track_tail(RxCtl &c, const RxCtl &ci, const TrackTmp &tt, const float *__restrict__ rs, int *__restrict__ uw_errors, int s,
                           int *__restrict__ ret_out, unsigned char *__restrict__ dec_active, int *__restrict__ nin_out, float *scratch) {
  // sigma_r = (mean|Dt1| + mean|Dt2|) / (2 sqrt(pi/2)) from the refreshed row sums (its barriers also order the staging of ci / tt)
  const float sigma_r = sigma_r_from_rowsums(rs, scratch);
  if (threadIdx.x == 0) {
    const int tmax0 = ci.tmax; const double fmax0 = ci.fmax;
    int tmax = tt.best_found ? tt.best_t : tmax0;
    const double fhat = tt.best_found ? tt.best_f : fmax0;
    const double fmax = 0.9 * fmax0 + 0.1 * fhat;
    const double Dthresh = (double)(2.f * sigma_r) * K_ACQ_1E4;
    const double Dthresh_eoo = (double)(2.f * sigma_r) * K_ACQ_1E5;
    const double D = tt.spot[0] + tt.spot[1], De = tt.spot[2] + tt.spot[3];
    const int valid = D > Dthresh, endofover = De > Dthresh_eoo;
    c.Dthresh = (float)Dthresh; c.Dtmax12 = (float)D; c.Dtmax12_eoo = (float)De;
    // timing slips (radae_rxe.py:208-218): the adjusted tmax is used for this call's extraction too
    int nin = RADE_NMF;
    if (tmax >= RADE_NMF - RADE_M) { nin = RADE_NMF + RADE_M; tmax -= RADE_M; }
    if (tmax < RADE_M) { nin = RADE_NMF - RADE_M; tmax += RADE_M; }
    c.tmax = tmax; c.fmax = fmax; c.n_check = ci.n_check + 1;
    const int synced_count = ci.synced_count + 1;
    c.synced_count = synced_count;
    int uw_fail = 0;
    if (synced_count % RADE_SYNCED_ONE_SEC == 0) {
      if (uw_errors[s] > RADE_UW_THRESH) uw_fail = 1;
      uw_errors[s] = 0;
    }
    const int valid_output = !endofover;
    c.uw_fail = uw_fail; c.candidate = valid; c.endofover = endofover; c.valid_output = valid_output; c.ran_sync = 1;
    // sync-state branch of the state machine (radae_rxe.py:276-296); search / candidate streams: rx_finish_kernel
    int next = ST_SYNC, vc = ci.valid_count;
    if (valid) vc = RADE_NMF_UNSYNC;
    else { vc -= 1; if (vc == 0) next = ST_SEARCH; }
    if (endofover || uw_fail) next = ST_SEARCH;
    if (next == ST_SEARCH) nin = RADE_NMF;
    c.valid_count = vc; c.state = next; c.nin = nin;
    const int ret = valid_output | (endofover << 1);
    c.ret = ret; ret_out[s] = ret; dec_active[s] = (unsigned char)valid_output; nin_out[s] = nin;
  }
  __syncthreads();
}

// ================================================================= frequency correction + OFDM demod + pilot EQ
struct DemodSmem {
  float2 xs[RADE_NS + 2][RADE_M];
  float2 sym[RADE_NS + 2][RADE_NC];
  float2 dft_part[4][RADE_NS + 2][RADE_NC];
  float2 pil[2][RADE_NC];
  float2 rotc[RADE_NC];
  RxCtl ctl_in; TrackTmp tt;                        // the stream's control block / refine results as they were at kernel entry
  double2 e1[32], e32[6], estep[RADE_NS + 2];     // exp(-j w l), exp(-j w 32 h), exp(-j w 192 r): 45 double sincos per stream instead of 320
  float scratch[96];
  float mag;
};

__global__ void __launch_bounds__(192)
rx_demod_kernel(DspTables T, RxCtl *__restrict__ ctl, const float2 *__restrict__ ring, float *__restrict__ z_hat,
                float *__restrict__ eoo_out, const unsigned char *__restrict__ active, const TrackTmp *__restrict__ tmp,
                const float *__restrict__ rowsum, int *__restrict__ uw_errors, int *__restrict__ ret_out,
                unsigned char *__restrict__ dec_active, int *__restrict__ nin_out) {
  __shared__ DemodSmem sm;
  const int s = blockIdx.x, tid = threadIdx.x;
  if (active && !active[s]) return;
  RxCtl &c = ctl[s];
  if (!c.tracking) return;                        // in sync when this call began (flag written by rx_bpf with the track list)
  // one coalesced read of the control block and the refine results: the state machine is a single thread's chain of dependent
  // decisions and should not wait on global memory for every field
  if (tid < (int)(sizeof(RxCtl) / 4)) reinterpret_cast<uint32_t *>(&sm.ctl_in)[tid] = reinterpret_cast<const uint32_t *>(&c)[tid];
  else if (tid >= 64 && tid < 64 + (int)(sizeof(TrackTmp) / 4)) reinterpret_cast<uint32_t *>(&sm.tt)[tid - 64] = reinterpret_cast<const uint32_t *>(tmp + s)[tid - 64];
  track_tail(c, sm.ctl_in, sm.tt, rowsum + (size_t)s * 2 * RADE_NMF, uw_errors, s, ret_out, dec_active, nin_out, sm.scratch);
  const int head = c.ring_head, tmax = c.tmax, endofover = c.endofover;
  const float2 *rg = ring + (size_t)s * RADE_RXBUF;
  const double w = 2.0 * M_PI * c.fmax / RADE_FS;
  const double2 P0 = make_double2(c.rx_phase_re, c.rx_phase_im);
  // rx_phase_vec[n] = rx_phase * exp(-j w (n+1)) (closed form of the recursion radae_rxe.py:227-231), stored csingle;
  // keep only the M samples of each symbol after Ncp + time_offset = 16.  The phasors are products of three table entries:
  // n + 1 = 32 h + l + 192 r.  Double sincos is a long dependent chain on this part; 46 threads do one each instead of 160 doing two.
  double2 phase_next = make_double2(0.0, 0.0);
  if (tid < 32 + 6 + RADE_NS + 2 + 1) {
    const double arg = tid < 32 ? (double)tid : tid < 38 ? 32.0 * (tid - 32) : tid < 38 + RADE_NS + 2 ? (double)RADE_SYM * (tid - 38) : (double)RADE_NEOO;
    double sn, cs; sincos(w * arg, &sn, &cs);
    const double2 e = make_double2(cs, -sn);
    if (tid < 32) sm.e1[tid] = e; else if (tid < 38) sm.e32[tid - 32] = e; else if (tid < 38 + RADE_NS + 2) sm.estep[tid - 38] = e;
    else phase_next = dcmul(P0, e);
  }
  float2 xin[RADE_NS + 2];
  const int n0 = RADE_NCP + RADE_TIME_OFFSET + tid;
  if (tid < RADE_M) {
#pragma unroll
    for (int r = 0; r < RADE_NS + 2; r++) xin[r] = __ldg(rg + ring_idx(head, tmax - RADE_NCP + n0 + r * RADE_SYM));
  }
  __syncthreads();                                // everybody has read rx_phase
  if (tid == 32 + 6 + RADE_NS + 2) { c.rx_phase_re = phase_next.x; c.rx_phase_im = phase_next.y; }
  if (tid < RADE_M) {
    const int a = n0 + 1;
    const double2 v0 = dcmul(P0, dcmul(sm.e32[a >> 5], sm.e1[a & 31]));
#pragma unroll
    for (int r = 0; r < RADE_NS + 2; r++) {
      const double2 v = dcmul(v0, sm.estep[r]);
      sm.xs[r][tid] = cmul(xin[r], make_float2((float)v.x, (float)v.y));
    }
  }
  __syncthreads();
  // DFT: 6 symbols x 30 carriers, 160 taps each.  thread = (carrier, quarter of the taps) on warps 0-3: one twiddle load serves the
  // six symbols, every complex MAC is two packed FMAs, the four partial sums meet in shared memory (added in tap order).
  if (tid < 128 && (tid & 31) < RADE_NC) {
    const int cc = tid & 31, kq = tid >> 5;
    float2 acc[RADE_NS + 2];
#pragma unroll
    for (int r = 0; r < RADE_NS + 2; r++) acc[r] = make_float2(0.f, 0.f);
    const float2 *wf = T.Wfwd + cc;
#pragma unroll 1
    for (int k0 = kq * (RADE_M / 4); k0 < (kq + 1) * (RADE_M / 4); k0 += 8) {
      float2 w[8];
#pragma unroll
      for (int j = 0; j < 8; j++) w[j] = __ldg(wf + (k0 + j) * RADE_NC);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float2 ws = make_float2(-w[j].y, w[j].x);
#pragma unroll
        for (int r = 0; r < RADE_NS + 2; r++) {
          const float2 x = sm.xs[r][k0 + j];
          acc[r] = ffma2(splat(x.x), w[j], acc[r]);
          acc[r] = ffma2(splat(x.y), ws, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RADE_NS + 2; r++) sm.dft_part[kq][r][cc] = acc[r];
  }
  __syncthreads();
  if (tid < (RADE_NS + 2) * RADE_NC) {
    const int r = tid / RADE_NC, cc = tid % RADE_NC;
    const float2 a = sm.dft_part[0][r][cc], b = sm.dft_part[1][r][cc], c2 = sm.dft_part[2][r][cc], d = sm.dft_part[3][r][cc];
    sm.sym[r][cc] = make_float2(((a.x + b.x) + c2.x) + d.x, ((a.y + b.y) + c2.y) + d.y);
  }
  __syncthreads();
  if (!endofover) {
    // 3-pilot least-squares fit per carrier on the two pilot rows (dsp.py:418-435)
    if (tid < 2 * RADE_NC) {
      const int i = tid / RADE_NC, cc = tid % RADE_NC, row = i ? RADE_NS + 1 : 0;
      const int cm = min(max(cc, 1), RADE_NC - 2);
      float2 g0 = make_float2(0.f, 0.f), g1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float pk = T.P[cm - 1 + k].x;
        const float2 hk = make_float2(sm.sym[row][cm - 1 + k].x / pk, sm.sym[row][cm - 1 + k].y / pk);
        const float2 a = cmul(T.Pmat[cc * 6 + k], hk), b = cmul(T.Pmat[cc * 6 + 3 + k], hk);
        g0.x += a.x; g0.y += a.y; g1.x += b.x; g1.y += b.y;
      }
      const float2 e = cmul(g1, T.eq_rot[cc]);
      sm.pil[i][cc] = make_float2(g0.x + e.x, g0.y + e.y);
    }
    __syncthreads();
    // SNR estimate from pilot row 0 (dsp.py:438-456) and coarse magnitude (dsp.py:477-482)
    float s1 = 0.f, s2 = 0.f, pw = 0.f;
    if (tid < RADE_NC) {
      const float2 pc = sm.sym[0][tid], pl = sm.pil[0][tid];
      s1 = pc.x * pc.x + pc.y * pc.y;
      const float m = hypotf(pl.x, pl.y);
      const float im = (m > 0.f) ? (pc.y * pl.x - pc.x * pl.y) / m : pc.y;          // imag(Pcn * exp(-j angle(pilot)))
      s2 = im * im;
    }
    if (tid < 2 * RADE_NC) { const float2 pl = sm.pil[tid / RADE_NC][tid % RADE_NC]; pw = pl.x * pl.x + pl.y * pl.y; }
    block_sum3(s1, s2, pw, sm.scratch);
    const float S1 = s1, S2 = s2, PW = pw;
    if (tid == 0) {
      float mag = sqrtf(PW / (float)(2 * RADE_NC)) + 1e-6f;
      sm.mag = mag * T.p0_abs / T.pilot_gain;
    }
    __syncthreads();
    if (tid == 160) {                               // a thread with nothing to do in the equaliser: the float64 log10 is a long chain
      double snr = (double)S1 / (2.0 * ((double)S2 + 1e-12)) - 1.0;
      if (snr <= 0.0) snr = 0.1;
      const double snrdB = (10.0 * log10(snr) - 2.513) / 0.8070;
      // + 10 log10(Rs Nc / 3000) + 10 log10((M + Ncp) / M), Rs = Fs / M (dsp.py:452-455), as float64 evaluates them
      const double snr3k = snrdB + -3.010299956639812 + 0.7918124604762482;
      c.snr_est = 0.9 * c.snr_est + 0.1 * snr3k;
    }
    // phase-only EQ with the channel linearly interpolated between the two pilot rows, then demap (dsp.py:466-474, :507-512)
    if (tid < RADE_NS * RADE_NC) {
      const int r = 1 + tid / RADE_NC, cc = tid % RADE_NC;
      const float2 p0 = sm.pil[0][cc], p1 = sm.pil[1][cc];
      const float2 slope = make_float2((p1.x - p0.x) / (float)(RADE_NS + 1), (p1.y - p0.y) / (float)(RADE_NS + 1));
      const float2 ch = make_float2(slope.x * (float)r + p0.x, slope.y * (float)r + p0.y);
      const float m = hypotf(ch.x, ch.y);
      const float2 rotv = (m > 0.f) ? make_float2(ch.x / m, -ch.y / m) : make_float2(1.f, 0.f);
      const float2 v = cmul(sm.sym[r][cc], rotv);
      float *z = z_hat + (size_t)s * RADE_NZMF * RADE_LATENT;
      reinterpret_cast<float2 *>(z)[tid] = make_float2(v.x / sm.mag, v.y / sm.mag);
    }
  } else {
    // end of over: common phase from P, E, E (dsp.py:513-524); rows 2..4 carry 90 data symbols
    if (tid < RADE_NC) {
      const float2 a = sm.sym[0][tid], b = sm.sym[1][tid], d = sm.sym[RADE_NS + 1][tid];
      const float p = T.P[tid].x, pe = T.Pend[tid].x;
      const float2 acc = make_float2(a.x / p + b.x / pe + d.x / pe, a.y / p + b.y / pe + d.y / pe);
      const float m = hypotf(acc.x, acc.y);
      sm.rotc[tid] = (m > 0.f) ? make_float2(acc.x / m, -acc.y / m) : make_float2(1.f, 0.f);
    }
    __syncthreads();
    if (tid < (RADE_NS - 1) * RADE_NC) {
      const int r = 2 + tid / RADE_NC, cc = tid % RADE_NC;
      const float2 v = cmul(sm.sym[r][cc], sm.rotc[cc]);
      float *e = eoo_out + (size_t)s * RADE_NEOO_BITS;
      e[2 * tid] = v.x; e[2 * tid + 1] = v.y;
    }
  }
}

// ================================================================= coarse-search post-processing + state machine
// Streams that entered this call in search / candidate state (the sync-state branch lives at the end of rx_track_kernel).
// CTAs stride over the search list; inactive streams only get their outputs filled in.
struct FinishSmem {
  FirstFixSmem ref;
  float scratch[96];
  int do_refine;
};

__global__ void __launch_bounds__(256)
rx_finish_kernel(DspTables T, RxCtl *__restrict__ ctl, const float2 *__restrict__ ring, const float *__restrict__ rowsum,
                 int *__restrict__ uw_errors, DecStreamState *__restrict__ dec_state, int reset_dec_on_sync,
                 int *__restrict__ ret_out, unsigned char *__restrict__ dec_active, int *__restrict__ nin_out,
                 const unsigned char *__restrict__ active, const int *__restrict__ search_list,
                 const int *__restrict__ counters, int *__restrict__ counters_next, int S) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FinishSmem &sm = *reinterpret_cast<FinishSmem *>(smem_raw);
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid < 4) counters_next[tid] = 0;       // the set the NEXT call's kernels will count into
  if (active)
    for (int s = blockIdx.x * blockDim.x + tid; s < S; s += gridDim.x * blockDim.x)
      if (!active[s]) { ret_out[s] = 0; dec_active[s] = 0; nin_out[s] = ctl[s].nin; }
  const int n_items = counters[0];
  for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
    const int s = search_list[w];
    RxCtl &c = ctl[s];
    const int state = c.state;
    // detect_pilots epilogue (radae/dsp.py:217-231)
    const float sigma_r = sigma_r_from_rowsums(rowsum + (size_t)s * 2 * RADE_NMF, sm.scratch);
    if (tid == 0) {
      const unsigned long long key = c.detect_key;
      const float val = __uint_as_float((unsigned)(key >> 32));
      const unsigned idx = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
      int tmax = 0; double fmax = 0.0;
      if (val > 0.f) { tmax = (int)(idx / RADE_NFCOARSE); fmax = (double)T.fcoarse[idx % RADE_NFCOARSE]; }
      c.tmax = tmax; c.fmax = fmax;
      const double Dthresh = (double)(2.f * sigma_r) * K_ACQ_1E5;
      c.Dthresh = (float)Dthresh; c.Dtmax12 = val;
      const int candidate = ((double)val > Dthresh) ? 1 : 0;
      c.candidate = candidate;
      // state machine on the state held at entry (radae_rxe.py:248-275)
      int next = state, do_refine = 0;
      if (state == ST_SEARCH) {
        if (candidate) { next = ST_CANDIDATE; c.tmax_candidate = tmax; c.valid_count = 1; }
      } else {
        if (candidate && abs(tmax - c.tmax_candidate) < RADE_NCP) {
          const int vc = c.valid_count + 1;
          c.valid_count = vc;
          if (vc > 3) {
            next = ST_SYNC; do_refine = 1;
            c.synced_count = 0; c.uw_fail = 0; uw_errors[s] = 0; c.valid_count = RADE_NMF_UNSYNC;
          }
        } else next = ST_SEARCH;
      }
      c.state = next;
      sm.do_refine = do_refine;
    }
    __syncthreads();
    if (sm.do_refine) {
      // first fix after acquisition: t in [max(0,tmax-1), tmax+2), f in arange(fmax-10, fmax+10, 0.25)  (radae_rxe.py:267-273)
      const int tm = c.tmax; const double fm = c.fmax;
      const int t_lo = max(0, tm - 1);
      if (tid < REF_THREADS) {
        const float2 *rg = ring + (size_t)s * RADE_RXBUF; const int head = c.ring_head;
        refine_first_fix(sm.ref, *reinterpret_cast<const AcqTables *>(T.acq_tab), [rg, head](int i) { return __ldg(rg + ring_idx(head, i)); },
                         t_lo, tm + 2 - t_lo, fm, tid, 1);
      }
      __syncthreads();
      if (tid == 0) {
        if (sm.ref.best_found) { c.tmax = sm.ref.best_t; c.fmax = sm.ref.best_f; }
        c.fmax += c.foff_err; c.foff_err = 0.0;
      }
      if (reset_dec_on_sync) {        // model.core_decoder_statefull.module.reset() (radae_rxe.py:263)
        uint32_t *d = reinterpret_cast<uint32_t *>(dec_state + s);
        for (int i = tid; i < (int)(sizeof(DecStreamState) / 4); i += blockDim.x) d[i] = 0u;
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (c.state == ST_SEARCH) c.nin = RADE_NMF;      // radae_rxe.py:294-296
      c.ret = c.valid_output | (c.endofover << 1);
      ret_out[s] = c.ret; dec_active[s] = (unsigned char)c.valid_output; nin_out[s] = c.nin;
    }
    __syncthreads();
  }
}

__global__ void rx_init_kernel(RxCtl *ctl, int *uw_errors, int S, double foff_err) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  RxCtl c;
  memset(&c, 0, sizeof(c));
  c.nin = RADE_NMF; c.state = ST_SEARCH; c.rx_phase_re = 1.0; c.bpf_phase = make_float2(1.f, 0.f); c.bpf_first = 1;
  c.foff_err = foff_err;
  ctl[s] = c;
  uw_errors[s] = 0;
}

}  // namespace

int rx_init_launch(RxCtl *ctl, int *uw_errors, int S, double foff_err, cudaStream_t stream) {
  rx_init_kernel<<<(S + 127) / 128, 128, 0, stream>>>(ctl, uw_errors, S, foff_err);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}

int rx_dsp_init_device() {
  {
    DspTablesHost th; dsp_tables_host(th);
    float h[RADE_BPF_NTAP + 3] = {0.f};
    for (int i = 0; i < RADE_BPF_NTAP; i++) h[i] = th.bpf_h[i];
    CUDA_CHECK(cudaMemcpyToSymbol(c_bpf_h, h, sizeof(h)));
    static AcqConst ac;
    for (int n = 0; n < RADE_M; n++) { const float px = th.p[n].real(), py = th.p[n].imag(); ac.ps4[n] = make_float4(px, py, py, -px); }
    memcpy(ac.basis, th.srch_basis.data(), sizeof(ac.basis));
    memcpy(ac.expand, th.srch_expand.data(), sizeof(ac.expand));
    CUDA_CHECK(cudaMemcpyToSymbol(c_acq, &ac, sizeof(ac)));
  }
  CUDA_CHECK(cudaFuncSetAttribute(rx_detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DetectSmem)));
  CUDA_CHECK(cudaFuncSetAttribute(rx_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TrackSmem)));
  CUDA_CHECK(cudaFuncSetAttribute(rx_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FinishSmem)));
  return 0;
}

int rx_dsp_launch(const DspTables &T, RxBuffers &B, const float2 *rx_in, const unsigned char *active, const LinkSrc *link, int S,
                  int bpf_en, int reset_dec_on_sync, int *ret_out, cudaStream_t stream, Profiler *prof) {
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
  int *cnt = B.counters + 4 * B.parity, *cnt_next = B.counters + 4 * (B.parity ^ 1);
  B.parity ^= 1;
  // Two independent branches follow the band-pass kernel: streams in sync (rx_track -> rx_demod) and streams searching
  // (rx_detect -> rx_finish).  They touch disjoint streams, so the search branch runs on a side stream, concurrently
  // (fork / join with events); with the per-kernel profiler on everything is serialised on the main stream instead.
  const bool fork = !prof->on && B.side_stream;
  cudaStream_t ss = fork ? B.side_stream : stream;
  prof->begin(K_RX_BPF);
  LinkSrc ls = {nullptr, nullptr, nullptr, nullptr};
  if (link) { ls = *link; active = link->active_out; }      // downstream kernels read the flags the band-pass kernel writes
  rx_bpf_kernel<<<S, BPF_THREADS, 0, stream>>>(T, B.ctl, B.ring, B.bpf_mem, rx_in, active, bpf_en, B.search_list, B.track_list, cnt, ls);
  prof->end(K_RX_BPF);
  if (fork) { CUDA_CHECK(cudaEventRecord(B.ev_fork, stream)); CUDA_CHECK(cudaStreamWaitEvent(ss, B.ev_fork, 0)); }
  // sync branch: the row refresh (fp32) and the refine + spot correlations (fp64) are independent -> two streams, joined before
  // rx_demod, whose head holds the state machine that needs both
  const int trk_grid = S < n_sm * 8 ? S : n_sm * 8;
  const bool fork2 = fork && B.side2_stream;
  cudaStream_t s2 = fork2 ? B.side2_stream : stream;
  if (fork2) CUDA_CHECK(cudaStreamWaitEvent(s2, B.ev_fork, 0));
  prof->begin(K_RX_TRACK, s2);
  rx_track_kernel<<<S < n_sm * 4 ? S : n_sm * 4, REF_THREADS, sizeof(TrackSmem), s2>>>(T, B.ctl, B.ring, B.track_list, cnt, (TrackTmp *)B.track_tmp);
  prof->end(K_RX_TRACK, s2);
  if (fork2) CUDA_CHECK(cudaEventRecord(B.ev_join2, s2));
  prof->begin(K_RX_REFRESH);
  rx_refresh_kernel<<<(trk_grid + RFR_STREAMS - 1) / RFR_STREAMS, RFR_THREADS, 0, stream>>>(B.ctl, B.ring, B.rowsum, B.track_list, cnt);
  prof->end(K_RX_REFRESH);
  if (fork2) CUDA_CHECK(cudaStreamWaitEvent(stream, B.ev_join2, 0));
  prof->begin(K_RX_DETECT, ss);
  int det_grid = S * (RADE_NMF / DET_TB); if (det_grid > n_sm * 12) det_grid = n_sm * 12;      // 64-thread CTAs, 81 registers: 12 per SM
  rx_detect_kernel<<<det_grid, DET_THREADS, sizeof(DetectSmem), ss>>>(T, B.ctl, B.ring, B.rowsum, B.search_list, cnt);
  prof->end(K_RX_DETECT, ss); prof->begin(K_RX_DEMOD);
  rx_demod_kernel<<<S, 192, 0, stream>>>(T, B.ctl, B.ring, B.z_hat, B.eoo, active, (const TrackTmp *)B.track_tmp, B.rowsum, B.uw_errors,
                                         ret_out, B.dec_active, B.nin);
  prof->end(K_RX_DEMOD); prof->begin(K_RX_FINISH, ss);
  rx_finish_kernel<<<S < 2 * n_sm ? S : 2 * n_sm, 256, sizeof(FinishSmem), ss>>>(T, B.ctl, B.ring, B.rowsum, B.uw_errors, B.dec_state,
      reset_dec_on_sync, ret_out, B.dec_active, B.nin, active, B.search_list, cnt, cnt_next, S);
  prof->end(K_RX_FINISH, ss);
  if (fork) { CUDA_CHECK(cudaEventRecord(B.ev_join, ss)); CUDA_CHECK(cudaStreamWaitEvent(stream, B.ev_join, 0)); }
  CUDA_CHECK(cudaGetLastError());
  return 6;
}
