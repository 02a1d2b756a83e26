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
//     -> rx_track (persistent, streams in sync: refine on the FP64 tensor cores + row refresh + sync-state machine) -> rx_demod
//     -> rx_detect (persistent over the search list: coarse grid search) -> rx_finish (search / candidate state machine)
//   the second branch runs on a side stream concurrently with the first; both join before the core decoder.
#include "rade_common.h"
#include "rade_host.h"
#include "tma.cuh"

namespace {

enum { ST_SEARCH = 0, ST_CANDIDATE = 1, ST_SYNC = 2 };

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ double2 dcmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// acc += a*b with four FMAs (the compiler otherwise emits FMUL+FFMA+FADD per component pair)
__device__ __forceinline__ void cmac(float2 &acc, float2 a, float2 b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
// Coarse-grid correlations for two sample windows x0, x1 against the pilot p shifted to the grid
// frequencies +-2.5k Hz, k = 6*kg .. 6*kg+5:  D(+-f_k) = A_k +- j B_k with A_k = sum_n y[n] cos(w_k n), B_k = sum_n y[n] sin(w_k n),
// y[n] = conj(x[n]) p[n]  (acquisition.detect_pilots / check_pilots, radae/dsp.py:204-205, :291-295; p_w = exp(j w n) p there).
// One (cos, sin) pair serves the +f and -f grid points: 4 FMAs per tap per pair instead of 8.
// Blackwell packed fp32: one FFMA2 does two independent IEEE FMAs on a register pair (SASS FFMA2 .F32x2)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                     rc = *reinterpret_cast<unsigned long long *>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ void corr6(float2 (&A0)[6], float2 (&B0)[6], float2 (&A1)[6], float2 (&B1)[6], const float2 *x0,
                                      const float2 *x1, const float4 *ps4, const float2 (*cs)[RADE_CSK], int kg) {
#pragma unroll
  for (int j = 0; j < 6; j++) { A0[j] = B0[j] = A1[j] = B1[j] = make_float2(0.f, 0.f); }
#pragma unroll 2
  for (int n = 0; n < RADE_M; n++) {
    const float4 pp = ps4[n];                      // (p.x, p.y, p.y, -p.x): y = conj(x) p with two packed FMAs
    const float2 a = x0[n], c = x1[n];
    const float2 y0 = ffma2(splat(a.x), make_float2(pp.x, pp.y), ffma2(splat(a.y), make_float2(pp.z, pp.w), make_float2(0.f, 0.f)));
    const float2 y1 = ffma2(splat(c.x), make_float2(pp.x, pp.y), ffma2(splat(c.y), make_float2(pp.z, pp.w), make_float2(0.f, 0.f)));
    const float4 *t = reinterpret_cast<const float4 *>(&cs[n][kg * 6]);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const float4 q = t[j];                       // (cos_k, sin_k, cos_k+1, sin_k+1)
      A0[2 * j] = ffma2(y0, splat(q.x), A0[2 * j]);         B0[2 * j] = ffma2(y0, splat(q.y), B0[2 * j]);
      A0[2 * j + 1] = ffma2(y0, splat(q.z), A0[2 * j + 1]); B0[2 * j + 1] = ffma2(y0, splat(q.w), B0[2 * j + 1]);
      A1[2 * j] = ffma2(y1, splat(q.x), A1[2 * j]);         B1[2 * j] = ffma2(y1, splat(q.y), B1[2 * j]);
      A1[2 * j + 1] = ffma2(y1, splat(q.z), A1[2 * j + 1]); B1[2 * j + 1] = ffma2(y1, splat(q.w), B1[2 * j + 1]);
    }
  }
}
// |D(+f_k)|, |D(-f_k)| from A, B
__device__ __forceinline__ void mags_pm(float2 A, float2 B, float &mp, float &mm) {
  mp = hypotf(A.x - B.y, A.y + B.x);
  mm = hypotf(A.x + B.y, A.y - B.x);
}
__device__ __forceinline__ int ring_idx(int head, int i) { int k = head + i; return k >= RADE_RXBUF ? k - RADE_RXBUF : k; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in every thread; scratch: >= 32 floats of shared memory
__device__ float block_sum(float v, float *scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
  if (w == 0) { t = warp_sum(t); if (lane == 0) scratch[0] = t; }
  __syncthreads();
  return scratch[0];
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
#pragma unroll
      for (int k = 0; k < RADE_BPF_NTAP + 1; k += 2) {
        const float4 w2 = Xv[k / 2 + 2];          // samples k+4, k+5
        const float h0 = c_bpf_h[k], h1 = c_bpf_h[k + 1];        // h[101] = 0 pads the odd tap count
        acc[0].x = fmaf(h0, w0.x, acc[0].x); acc[0].y = fmaf(h0, w0.y, acc[0].y);
        acc[1].x = fmaf(h0, w0.z, acc[1].x); acc[1].y = fmaf(h0, w0.w, acc[1].y);
        acc[2].x = fmaf(h0, w1.x, acc[2].x); acc[2].y = fmaf(h0, w1.y, acc[2].y);
        acc[3].x = fmaf(h0, w1.z, acc[3].x); acc[3].y = fmaf(h0, w1.w, acc[3].y);
        if (k + 1 < RADE_BPF_NTAP) {
          acc[0].x = fmaf(h1, w0.z, acc[0].x); acc[0].y = fmaf(h1, w0.w, acc[0].y);
          acc[1].x = fmaf(h1, w1.x, acc[1].x); acc[1].y = fmaf(h1, w1.y, acc[1].y);
          acc[2].x = fmaf(h1, w1.z, acc[2].x); acc[2].y = fmaf(h1, w1.w, acc[2].y);
          acc[3].x = fmaf(h1, w2.x, acc[3].x); acc[3].y = fmaf(h1, w2.y, acc[3].y);
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
    if (c.state != ST_SYNC) search_list[atomicAdd(&counters[0], 1)] = s;         // work list of rx_detect / rx_finish
    else track_list[atomicAdd(&counters[2], 1)] = s;                              // work list of rx_track
  }
}

// ================================================================= coarse pilot search (search / candidate streams)
// work item = 32 timing offsets x 40 frequency offsets x 2 pilot positions, 160-tap complex correlations; 128-thread CTAs
// (80 registers) fit next to a resident rx_track CTA, so the search branch runs concurrently with the tracking branch
constexpr int DET_TB = 32, DET_THREADS = 4 * DET_TB;
struct DetectSmem {
  AcqTables tab;                   // one TMA bulk copy per CTA
  float2 r1[DET_TB + RADE_M];
  float2 r2[DET_TB + RADE_M];
  float part[2][4][DET_TB];
  unsigned long long best[DET_THREADS / 32];
  uint64_t tab_bar;
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
  if (tid == 0) {
    mbar_init(&sm.tab_bar, 1); mbar_fence_init();
    mbar_expect_tx(&sm.tab_bar, (uint32_t)sizeof(AcqTables));
    bulk_g2s(&sm.tab, T.acq_tab, (uint32_t)sizeof(AcqTables), &sm.tab_bar);
  }
  __syncthreads();
  mbar_wait(&sm.tab_bar, 0);
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
    const int tl = tid & (DET_TB - 1), fg = tid / DET_TB;  // fg = k group: k = 6 fg .. 6 fg + 5
    float2 A0[6], B0[6], A1[6], B1[6];
    corr6(A0, B0, A1, B1, &sm.r1[tl], &sm.r2[tl], sm.tab.ps4, sm.tab.cs, fg);
    float s1 = 0.f, s2 = 0.f, best = -1.f; int bestf = RADE_NFCOARSE;
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const int k = fg * 6 + j;
      if (k > 20) continue;
      float p1, m1, p2, m2;
      mags_pm(A0[j], B0[j], p1, m1); mags_pm(A1[j], B1[j], p2, m2);
      if (k < 20) {                                   // +2.5k Hz -> grid index 20 + k
        s1 += p1; s2 += p2;
        const float d = p1 + p2; const int fi = 20 + k;
        if (d > best || (d == best && fi < bestf)) { best = d; bestf = fi; }
      }
      if (k > 0) {                                    // -2.5k Hz -> grid index 20 - k
        s1 += m1; s2 += m2;
        const float d = m1 + m2; const int fi = 20 - k;
        if (d > best || (d == best && fi < bestf)) { best = d; bestf = fi; }
      }
    }
    sm.part[0][fg][tl] = s1; sm.part[1][fg][tl] = s2;
    // arg-max with "first (t, f) wins": larger key = larger value, then smaller flat index
    unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) | (0xFFFFFFFFu - (unsigned)((t0 + tl) * RADE_NFCOARSE + bestf));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o); key = k2 > key ? k2 : key; }
    if ((tid & 31) == 0) sm.best[tid >> 5] = key;
    __syncthreads();
    if (tid < 2 * DET_TB) {
      const int half = tid / DET_TB, t = tid & (DET_TB - 1);
      rowsum[((size_t)s * 2 + half) * RADE_NMF + t0 + t] = ((sm.part[half][0][t] + sm.part[half][1][t]) + sm.part[half][2][t]) + sm.part[half][3][t];
    }
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
// computed in complex128 and rounded to csingle like the reference.  Here: a real GEMM on the FP64 tensor cores
// (DMMA m8n8k4, 256 FMA per instruction): A = Toeplitz view of the widened samples [8 t][(n, re/im)], B = [(n, re/im)]
// [(f, re/im)] formed on the fly from the steering table vtab; 24 frequencies (6 n-tiles) per pass.
constexpr int REF_NT = 16;                        // max timing offsets
constexpr int REF_NFP = 24;                       // frequencies per pass = 6 DMMA n-tiles of 4 complex columns
constexpr int REF_VLD = 165;                      // vtab row: tap n lives at n + (n >> 5); 165 = 5 mod 8 spreads f over the banks
constexpr int REF_RLEN = REF_NT + RADE_M + 8;     // widened window (+ over-read of the last k-step)
constexpr int REF_THREADS = 128;                  // four warps
struct RefineSmem {
  double2 vtab[REF_NFP][REF_VLD];                 // conj(p[n]) exp(-j w_f n)
  double2 ramp[REF_NFP];                          // exp(-j w_f Nmf)
  double2 ra[REF_RLEN];                           // rx[t_lo ...] widened once (np.dot up-casts csingle to complex128)
  double2 pad_;                                   // shifts rb by one entry: ra/rb reads of one warp hit different banks
  double2 rb[REF_RLEN];                           // rx[t_lo + Nmf ...]
  float2 d1[REF_NFP][REF_NT];
  float2 d2[REF_NFP][REF_NT];
  float red_mag[4]; int red_ord[4];
  float best_mag; int best_t; int best_found; double best_f;
};

// values of np.arange(start, stop, step) for float64: v[i] = start + i*((start+step)-start), len = ceil((stop-start)/step)
__device__ __forceinline__ int arange_len(double start, double stop, double step) { return (int)ceil((stop - start) / step); }

// Searches t in [t_lo, t_lo+nt) x f in arange(f_start, f_stop, f_step); result in sm.best_* (best_found == 0: nothing beat 0).
// Executed by REF_THREADS threads (g = index within the group) that synchronise on named barrier `bar`.
// load(i) returns rx_buf[i] (logical index).  WIDE_T: nt > 8 -> warp = (pilot position, 8-row t tile) x 6 n-tiles;
// otherwise warp = (pilot position, half of the n-tiles) with a single t tile.
template <bool WIDE_T, typename Load>
__device__ void refine_dmma(RefineSmem &sm, const double2 *__restrict__ pcd, Load load, int t_lo, int nt,
                            double f_start, double f_stop, double f_step, int g, int bar) {
  constexpr int NQ = WIDE_T ? 6 : 3;              // n-tiles per warp
  const int nf_all = arange_len(f_start, f_stop, f_step);
  const double delta = (f_start + f_step) - f_start;
  if (g == 0) { sm.best_mag = 0.f; sm.best_found = 0; sm.best_t = 0; sm.best_f = 0.0; }
  for (int i2 = g; i2 < REF_RLEN; i2 += REF_THREADS) {
    const bool in = i2 < nt + RADE_M;
    const float2 a = in ? load(t_lo + i2) : make_float2(0.f, 0.f), c = in ? load(t_lo + RADE_NMF + i2) : make_float2(0.f, 0.f);
    sm.ra[i2] = make_double2((double)a.x, (double)a.y);
    sm.rb[i2] = make_double2((double)c.x, (double)c.y);
  }
  for (int c0 = 0; c0 < nf_all; c0 += REF_NFP) {
    const int nf = min(REF_NFP, nf_all - c0);
    group_sync(bar, REF_THREADS);                 // previous pass done with vtab / d1 / d2
    // steering vectors for this pass: thread = (f, 32-tap segment), one sincos pair + 31 rotations (|error| ~ 1e-15)
    for (int task = g; task < REF_NFP * 5; task += REF_THREADS) {
      const int fi = task / 5, seg = task % 5;
      double2 *row = sm.vtab[fi] + 33 * seg;
      if (fi < nf) {
        const double f = f_start + (double)(c0 + fi) * delta;
        const double w = 2.0 * M_PI * f / RADE_FS;
        double sn, cs, s1, c1; sincos(w * (double)(32 * seg), &sn, &cs); sincos(w, &s1, &c1);
        double2 e = make_double2(cs, -sn); const double2 step = make_double2(c1, -s1);
#pragma unroll 4
        for (int q = 0; q < 32; q++) { row[q] = dcmul(e, pcd[32 * seg + q]); e = dcmul(e, step); }
        if (seg == 0) {             // pilots of the NEXT frame: extra phase ramp exp(-j w Nmf)
          sincos(w * (double)RADE_NMF, &sn, &cs);
          sm.ramp[fi] = make_double2(cs, -sn);
        }
      } else {
        for (int q = 0; q < 32; q++) row[q] = make_double2(0.0, 0.0);
      }
    }
    group_sync(bar, REF_THREADS);
    {
      const int wr = g >> 5, lane = g & 31, gq = lane >> 2, c = lane & 3;
      const int half = wr >> 1, mt = WIDE_T ? (wr & 1) : 0, q0 = WIDE_T ? 0 : 3 * (wr & 1);
      const double *ap = reinterpret_cast<const double *>((half ? sm.rb : sm.ra) + mt * 8 + gq + (c >> 1)) + (c & 1);
      const int reim = gq & 1, comp = c & 1;
      const unsigned flip = (reim == 0 && comp == 1) ? 0x80000000u : 0u;      // B = [vr; -vi] for Re columns, [vi; vr] for Im
      const double *bp[NQ];
#pragma unroll
      for (int q = 0; q < NQ; q++) bp[q] = reinterpret_cast<const double *>(sm.vtab[(q0 + q) * 4 + (gq >> 1)] + (c >> 1)) + (comp ^ reim);
      double acc[NQ][2];
#pragma unroll
      for (int q = 0; q < NQ; q++) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll 1
      for (int blk = 0; blk < 5; blk++) {
#pragma unroll 4
        for (int kk = 0; kk < 16; kk++) {
          const double av = ap[4 * (16 * blk + kk)];
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            double bv = bp[q][2 * (33 * blk + 2 * kk)];
            bv = __hiloint2double(__double2hiint(bv) ^ flip, __double2loint(bv));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[q][0]), "+d"(acc[q][1]) : "d"(av), "d"(bv));
          }
        }
      }
      const int ti = mt * 8 + gq;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int fi = (q0 + q) * 4 + c;
        if (fi < nf && ti < nt) {
          if (half) { const double2 e = dcmul(make_double2(acc[q][0], acc[q][1]), sm.ramp[fi]); sm.d2[fi][ti] = make_float2((float)e.x, (float)e.y); }
          else sm.d1[fi][ti] = make_float2((float)acc[q][0], (float)acc[q][1]);
        }
      }
    }
    group_sync(bar, REF_THREADS);
    float bm = -1.f; int bo = 0x7fffffff;
    for (int q = g; q < nf * nt; q += REF_THREADS) {
      const int fi = q / nt, ti = q % nt;            // ord = q: f outer loop, t inner loop
      const float2 a = sm.d1[fi][ti], c = sm.d2[fi][ti];
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
      if (bm > sm.best_mag) {       // strict >: passes are visited in increasing f, the initial Dtmax is 0 (radae/dsp.py:262)
        sm.best_mag = bm; sm.best_found = 1;
        sm.best_t = t_lo + bo % nt;
        sm.best_f = f_start + (double)(c0 + bo / nt) * delta;
      }
    }
  }
  group_sync(bar, REF_THREADS);
}

// sigma_r = (mean|Dt1| + mean|Dt2|) / (2*sqrt(pi/2)) from the row sums, float32 like the reference's np.mean
__device__ float sigma_r_from_rowsums(const float *rs /* [2][960] */, float *scratch) {
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < RADE_NMF; i += blockDim.x) { a += rs[i]; b += rs[RADE_NMF + i]; }
  const float sa = block_sum(a, scratch), sb = block_sum(b, scratch);
  const float k = 1.2533141373155001f;              // (pi/2)**0.5 as float32
  const float s1 = (sa / (float)(RADE_NMF * RADE_NFCOARSE)) / k, s2 = (sb / (float)(RADE_NMF * RADE_NFCOARSE)) / k;
  return (s1 + s2) / 2.0f;
}

// ================================================================= sync-state tracking: refine + check_pilots + slips
// Persistent kernel: one CTA per SM walks the list of streams in sync (built by rx_bpf).  The constant tables (37 KB) are
// bulk-copied into shared memory once per CTA; a producer warp prefetches the NEXT stream's sample ring (in logical
// order), |Dt| row sums and control block with TMA bulk copies into the other half of a double buffer while the 16
// consumer warps work on the current one:
//   warps 0-11  check_pilots' refresh of 48 rows of the |Dt| row sums (fp32, packed FFMA2; even/odd tap split per lane pair)
//   warps 12-15 refine: 16 timing x 20(21) frequency x 2 pilot positions in complex128 as a real GEMM on the FP64
//               tensor cores (DMMA m8n8k4: 256 FMA per instruction instead of 32)
// then sigma_r, the four complex128 spot correlations, slips and the sync-state part of the state machine.
constexpr int TRK_REFRESH = 384, TRK_REFINE = REF_THREADS, TRK_CONSUMERS = TRK_REFRESH + TRK_REFINE, TRK_THREADS = TRK_CONSUMERS + 32;
struct TrackStage {
  alignas(128) float2 rx[RADE_RXBUF];             // rx_buf in LOGICAL order (two bulk copies around ring_head)
  float rs[2 * RADE_NMF];                         // row sums
  RxCtl ctl;
};
struct TrackSmem {
  AcqTables tab;
  TrackStage st[2];
  RefineSmem ref;
  float scratch[32];
  double spot[4];
  uint64_t full[2], empty[2], tab_bar;
  int item[2];                     // stream handled from stage b (-1: no more work)
};
static_assert(sizeof(RxCtl) % 16 == 0 && sizeof(TrackStage) % 128 == 0, "bulk-copy alignment");

__global__ void __launch_bounds__(TRK_THREADS, 1)
rx_track_kernel(DspTables T, RxCtl *__restrict__ ctl, const float2 *__restrict__ ring, float *__restrict__ rowsum,
                int *__restrict__ uw_errors, const int *__restrict__ track_list, int *__restrict__ counters,
                int *__restrict__ ret_out, unsigned char *__restrict__ dec_active, int *__restrict__ nin_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TrackSmem &sm = *reinterpret_cast<TrackSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const int n_items = counters[2];
  if ((int)blockIdx.x >= n_items) return;
  if (tid == 0) {
    mbar_init(&sm.full[0], 1); mbar_init(&sm.full[1], 1); mbar_init(&sm.empty[0], 1); mbar_init(&sm.empty[1], 1);
    mbar_init(&sm.tab_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid >= TRK_CONSUMERS) {                     // ---- producer warp: TMA prefetch, one stream ahead
    if (tid == TRK_CONSUMERS) {
      mbar_expect_tx(&sm.tab_bar, (uint32_t)sizeof(AcqTables));
      bulk_g2s(&sm.tab, T.acq_tab, (uint32_t)sizeof(AcqTables), &sm.tab_bar);
      // streams are handed out dynamically (atomic counter): a CTA that starts late — its SM was still busy with another
      // kernel of the frame pipeline — simply takes fewer of them
      for (int k = 0;; k++) {
        const int b = k & 1;
        if (k >= 2) mbar_wait(&sm.empty[b], ((k >> 1) - 1) & 1);
        const int it = atomicAdd(&counters[3], 1);
        if (it >= n_items) { sm.item[b] = -1; mbar_arrive(&sm.full[b]); break; }
        const int s = track_list[it];
        sm.item[b] = s;
        const int head = ctl[s].ring_head;        // even by construction (nin is 800 / 960 / 1120)
        const float2 *rg = ring + (size_t)s * RADE_RXBUF;
        TrackStage &st = sm.st[b];
        mbar_expect_tx(&sm.full[b], (uint32_t)(sizeof(float2) * RADE_RXBUF + sizeof(float) * 2 * RADE_NMF + sizeof(RxCtl)));
        bulk_g2s(st.rx, rg + head, (uint32_t)(sizeof(float2) * (RADE_RXBUF - head)), &sm.full[b]);
        if (head) bulk_g2s(st.rx + (RADE_RXBUF - head), rg, (uint32_t)(sizeof(float2) * head), &sm.full[b]);
        bulk_g2s(st.rs, rowsum + (size_t)s * 2 * RADE_NMF, (uint32_t)(sizeof(float) * 2 * RADE_NMF), &sm.full[b]);
        bulk_g2s(&st.ctl, ctl + s, (uint32_t)sizeof(RxCtl), &sm.full[b]);
      }
    }
    return;
  }
  mbar_wait(&sm.tab_bar, 0);
  for (int k = 0;; k++) {
    const int b = k & 1;
    mbar_wait(&sm.full[b], (k >> 1) & 1);
    const int s = sm.item[b];
    if (s < 0) break;
    TrackStage &st = sm.st[b];
    const int tmax0 = st.ctl.tmax; const double fmax0 = st.ctl.fmax;
    const int rot = st.ctl.n_check % 20;
    if (tid < TRK_REFRESH) {
      // ---- check_pilots row refresh (radae/dsp.py:288-295, deterministic schedule): rows t_i = 20 i + rot, both pilot
      // positions, 40 grid frequencies.  lane = (row i, k group kg, tap parity ks): conflict-free shared-memory reads
      const int ks = tid & 1, kg = (tid >> 1) & 3, i = tid >> 3;
      const float2 *x0 = st.rx + rot + 20 * i, *x1 = x0 + RADE_NMF;
      float2 A0[6], B0[6], A1[6], B1[6];
#pragma unroll
      for (int j = 0; j < 6; j++) { A0[j] = B0[j] = A1[j] = B1[j] = make_float2(0.f, 0.f); }
      // software pipeline: the six shared-memory loads of tap n+2 are issued before the 28 packed FMAs of tap n
      const float4 *csr = reinterpret_cast<const float4 *>(&sm.tab.cs[0][kg * 6]);       // row stride RADE_CSK/2 float4
      float4 pp = sm.tab.ps4[ks], q0 = csr[ks * (RADE_CSK / 2)], q1 = csr[ks * (RADE_CSK / 2) + 1], q2 = csr[ks * (RADE_CSK / 2) + 2];
      float2 a = x0[ks], c = x1[ks];
#pragma unroll 2
      for (int n = ks; n < RADE_M; n += 2) {
        const int nn = (n + 2 < RADE_M) ? n + 2 : n;
        const float4 ppn = sm.tab.ps4[nn], q0n = csr[nn * (RADE_CSK / 2)], q1n = csr[nn * (RADE_CSK / 2) + 1], q2n = csr[nn * (RADE_CSK / 2) + 2];
        const float2 an = x0[nn], cn = x1[nn];
        const float2 y0 = ffma2(splat(a.x), make_float2(pp.x, pp.y), ffma2(splat(a.y), make_float2(pp.z, pp.w), make_float2(0.f, 0.f)));
        const float2 y1 = ffma2(splat(c.x), make_float2(pp.x, pp.y), ffma2(splat(c.y), make_float2(pp.z, pp.w), make_float2(0.f, 0.f)));
        const float4 qq[3] = {q0, q1, q2};
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const float4 q = qq[j];
          A0[2 * j] = ffma2(y0, splat(q.x), A0[2 * j]);         B0[2 * j] = ffma2(y0, splat(q.y), B0[2 * j]);
          A0[2 * j + 1] = ffma2(y0, splat(q.z), A0[2 * j + 1]); B0[2 * j + 1] = ffma2(y0, splat(q.w), B0[2 * j + 1]);
          A1[2 * j] = ffma2(y1, splat(q.x), A1[2 * j]);         B1[2 * j] = ffma2(y1, splat(q.y), B1[2 * j]);
          A1[2 * j + 1] = ffma2(y1, splat(q.z), A1[2 * j + 1]); B1[2 * j + 1] = ffma2(y1, splat(q.w), B1[2 * j + 1]);
        }
        pp = ppn; q0 = q0n; q1 = q1n; q2 = q2n; a = an; c = cn;
      }
      // the even-tap lane finishes k = 6 kg + {0,1,2}, the odd-tap lane k = 6 kg + {3,4,5}: swap the other three partials
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float2 mine[4], send[4];
        mine[0] = ks ? A0[j + 3] : A0[j]; send[0] = ks ? A0[j] : A0[j + 3];
        mine[1] = ks ? B0[j + 3] : B0[j]; send[1] = ks ? B0[j] : B0[j + 3];
        mine[2] = ks ? A1[j + 3] : A1[j]; send[2] = ks ? A1[j] : A1[j + 3];
        mine[3] = ks ? B1[j + 3] : B1[j]; send[3] = ks ? B1[j] : B1[j + 3];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          mine[q].x += __shfl_xor_sync(0xffffffffu, send[q].x, 1);
          mine[q].y += __shfl_xor_sync(0xffffffffu, send[q].y, 1);
        }
        const int kk = kg * 6 + ks * 3 + j;
        if (kk <= 20) {
          float p0, m0, p1, m1;
          mags_pm(mine[0], mine[1], p0, m0); mags_pm(mine[2], mine[3], p1, m1);
          if (kk < 20) { s0 += p0; s1 += p1; }
          if (kk > 0) { s0 += m0; s1 += m1; }
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
      if ((tid & 7) == 0) {
        const int r = 20 * i + rot;
        st.rs[r] = s0; st.rs[RADE_NMF + r] = s1;
        float *rs = rowsum + (size_t)s * 2 * RADE_NMF;
        rs[r] = s0; rs[RADE_NMF + r] = s1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // st.rs is overwritten by a bulk copy two streams later
      }
    } else {
      // ---- refine (radae_rxe.py:202-205, radae/dsp.py:233-270): t in [max(0,tmax-8), tmax+8), f in arange(fmax-1, fmax+1, .1)
      const int t_lo = max(0, tmax0 - 8);
      const float2 *rxl = st.rx;
      refine_dmma<true>(sm.ref, sm.tab.pcd, [rxl](int i) { return rxl[i]; }, t_lo, tmax0 + 8 - t_lo, fmax0 - 1, fmax0 + 1, 0.1,
                        tid - TRK_REFRESH, 1);
    }
    if (tid >= TRK_REFRESH) {
      // spot correlations in complex128 (radae/dsp.py:305-314) by the refine warps, which finish before the row refresh does:
      // refine warp k -> |sum conj(w_vec*rx[o_k + n]) * q_k[n]| at the refined timing / smoothed frequency
      const int wp = (tid - TRK_REFRESH) >> 5, lane = tid & 31;
      const int tm = sm.ref.best_found ? sm.ref.best_t : tmax0;
      const double fm = 0.9 * fmax0 + 0.1 * (sm.ref.best_found ? sm.ref.best_f : fmax0);
      const int o = tm + (wp == 0 ? 0 : wp == 2 ? RADE_M + RADE_NCP : RADE_NMF);
      const double w = 2.0 * M_PI * fm / RADE_FS;
      double sn, cs, s32, c32; sincos(w * (double)lane, &sn, &cs); sincos(w * 32.0, &s32, &c32);
      double2 e = make_double2(cs, -sn); const double2 step = make_double2(c32, -s32);
      double ax = 0.0, ay = 0.0;
#pragma unroll
      for (int n = lane; n < RADE_M; n += 32) {
        const float2 xv = st.rx[o + n];
        const float2 qf = (wp < 2) ? make_float2(sm.tab.ps4[n].x, sm.tab.ps4[n].y) : sm.tab.pend[n];
        const double2 v = dcmul(e, make_double2((double)xv.x, (double)xv.y));                       // w_vec * rx
        const double2 r2 = dcmul(make_double2(v.x, -v.y), make_double2((double)qf.x, (double)qf.y));
        ax += r2.x; ay += r2.y;
        e = dcmul(e, step);
      }
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) { ax += __shfl_xor_sync(0xffffffffu, ax, q); ay += __shfl_xor_sync(0xffffffffu, ay, q); }
      if (lane == 0) sm.spot[wp] = hypot(ax, ay);
    }
    group_sync(3, TRK_CONSUMERS);
    int tmax = sm.ref.best_found ? sm.ref.best_t : tmax0;
    const double fhat = sm.ref.best_found ? sm.ref.best_f : fmax0;
    const double fmax = 0.9 * fmax0 + 0.1 * fhat;
    // sigma_r = (mean|Dt1| + mean|Dt2|) / (2 sqrt(pi/2)) from the refreshed row sums (radae/dsp.py:297-300): per-warp partial
    // sums now, thread 0 adds the 16 + 16 partials
    {
      float a = 0.f, c = 0.f;
      for (int q = tid; q < RADE_NMF; q += TRK_CONSUMERS) { a += st.rs[q]; c += st.rs[RADE_NMF + q]; }
      a = warp_sum(a); c = warp_sum(c);
      if ((tid & 31) == 0) { sm.scratch[tid >> 5] = a; sm.scratch[16 + (tid >> 5)] = c; }
    }
    group_sync(3, TRK_CONSUMERS);
    if (tid == 0) {
      float sa = 0.f, sb = 0.f;
      for (int q = 0; q < TRK_CONSUMERS / 32; q++) { sa += sm.scratch[q]; sb += sm.scratch[16 + q]; }
      const float kf = 1.2533141373155001f;
      const float sigma_r = ((sa / (float)(RADE_NMF * RADE_NFCOARSE)) / kf + (sb / (float)(RADE_NMF * RADE_NFCOARSE)) / kf) / 2.0f;
      RxCtl &c = ctl[s];
      const double Dthresh = (double)(2.f * sigma_r) * sqrt(-log(1e-4 / 5.0));
      const double Dthresh_eoo = (double)(2.f * sigma_r) * sqrt(-log(1e-5 / 5.0));
      const double D = sm.spot[0] + sm.spot[1], De = sm.spot[2] + sm.spot[3];
      const int valid = D > Dthresh, endofover = De > Dthresh_eoo;
      c.Dthresh = (float)Dthresh; c.Dtmax12 = (float)D; c.Dtmax12_eoo = (float)De;
      // timing slips (radae_rxe.py:208-218): the adjusted tmax is used for this call's extraction too
      int nin = RADE_NMF;
      if (tmax >= RADE_NMF - RADE_M) { nin = RADE_NMF + RADE_M; tmax -= RADE_M; }
      if (tmax < RADE_M) { nin = RADE_NMF - RADE_M; tmax += RADE_M; }
      c.tmax = tmax; c.fmax = fmax; c.n_check = st.ctl.n_check + 1;
      const int synced_count = st.ctl.synced_count + 1;
      c.synced_count = synced_count;
      int uw_fail = 0;
      if (synced_count % RADE_SYNCED_ONE_SEC == 0) {
        if (uw_errors[s] > RADE_UW_THRESH) uw_fail = 1;
        uw_errors[s] = 0;
      }
      const int valid_output = !endofover;
      c.uw_fail = uw_fail; c.candidate = valid; c.endofover = endofover; c.valid_output = valid_output; c.ran_sync = 1;
      // sync-state branch of the state machine (radae_rxe.py:276-296); search / candidate streams: rx_finish_kernel
      int next = ST_SYNC, vc = st.ctl.valid_count;
      if (valid) vc = RADE_NMF_UNSYNC;
      else { vc -= 1; if (vc == 0) next = ST_SEARCH; }
      if (endofover || uw_fail) next = ST_SEARCH;
      if (next == ST_SEARCH) nin = RADE_NMF;
      c.valid_count = vc; c.state = next; c.nin = nin;
      const int ret = valid_output | (endofover << 1);
      c.ret = ret; ret_out[s] = ret; dec_active[s] = (unsigned char)valid_output; nin_out[s] = nin;
    }
    group_sync(3, TRK_CONSUMERS);
    if (tid == 0) mbar_arrive(&sm.empty[b]);
  }
}

// ================================================================= frequency correction + OFDM demod + pilot EQ
struct DemodSmem {
  float2 xs[RADE_NS + 2][RADE_M];
  float2 sym[RADE_NS + 2][RADE_NC];
  float2 pil[2][RADE_NC];
  float2 rotc[RADE_NC];
  float scratch[32];
  float mag;
};

__global__ void __launch_bounds__(192)
rx_demod_kernel(DspTables T, RxCtl *__restrict__ ctl, const float2 *__restrict__ ring, float *__restrict__ z_hat,
                float *__restrict__ eoo_out, const unsigned char *__restrict__ active) {
  __shared__ DemodSmem sm;
  const int s = blockIdx.x, tid = threadIdx.x;
  if (active && !active[s]) return;
  RxCtl &c = ctl[s];
  if (!c.ran_sync) return;
  const int head = c.ring_head, tmax = c.tmax, endofover = c.endofover;
  const float2 *rg = ring + (size_t)s * RADE_RXBUF;
  const double w = 2.0 * M_PI * c.fmax / RADE_FS;
  const double2 P0 = make_double2(c.rx_phase_re, c.rx_phase_im);
  // rx_phase_vec[n] = rx_phase * exp(-j w (n+1)) (closed form of the recursion radae_rxe.py:227-231), stored csingle;
  // keep only the M samples of each symbol after Ncp + time_offset = 16.  Thread t owns sample k = t of every symbol
  // (t < 160): one complex128 sincos, then one rotation by exp(-j w 192) per symbol.
  if (tid < RADE_M) {
    const int n0 = RADE_NCP + RADE_TIME_OFFSET + tid;
    double sn, cs, ss, cc; sincos(w * (double)(n0 + 1), &sn, &cs); sincos(w * (double)RADE_SYM, &ss, &cc);
    double2 v = dcmul(P0, make_double2(cs, -sn)); const double2 step = make_double2(cc, -ss);
#pragma unroll
    for (int r = 0; r < RADE_NS + 2; r++) {
      sm.xs[r][tid] = cmul(rg[ring_idx(head, tmax - RADE_NCP + n0 + r * RADE_SYM)], make_float2((float)v.x, (float)v.y));
      v = dcmul(v, step);
    }
  }
  __syncthreads();
  if (tid == 0) {
    double sn, cs; sincos(w * (double)RADE_NEOO, &sn, &cs);
    const double2 v = dcmul(P0, make_double2(cs, -sn));
    c.rx_phase_re = v.x; c.rx_phase_im = v.y;
  }
  if (tid < (RADE_NS + 2) * RADE_NC) {              // 180 DFT outputs, 160-point each
    const int r = tid / RADE_NC, cc = tid % RADE_NC;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < RADE_M; k++) cmac(acc, sm.xs[r][k], T.Wfwd[k * RADE_NC + cc]);
    sm.sym[r][cc] = acc;
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
    const float S1 = block_sum(s1, sm.scratch), S2 = block_sum(s2, sm.scratch), PW = block_sum(pw, sm.scratch);
    if (tid == 0) {
      double snr = (double)S1 / (2.0 * ((double)S2 + 1e-12)) - 1.0;
      if (snr <= 0.0) snr = 0.1;
      const double snrdB = (10.0 * log10(snr) - 2.513) / 0.8070;
      const double Rs = (double)RADE_FS / RADE_M;
      const double snr3k = snrdB + 10.0 * log10(Rs * RADE_NC / 3000.0) + 10.0 * log10((double)(RADE_M + RADE_NCP) / RADE_M);
      c.snr_est = 0.9 * c.snr_est + 0.1 * snr3k;
      float mag = sqrtf(PW / (float)(2 * RADE_NC)) + 1e-6f;
      sm.mag = mag * T.p0_abs / T.pilot_gain;
    }
    __syncthreads();
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
      z[2 * tid] = v.x / sm.mag; z[2 * tid + 1] = v.y / sm.mag;
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
  RefineSmem ref;
  float scratch[32];
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
      const double Dthresh = (double)(2.f * sigma_r) * sqrt(-log(1e-5 / 5.0));
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
        refine_dmma<false>(sm.ref, reinterpret_cast<const AcqTables *>(T.acq_tab)->pcd, [rg, head](int i) { return rg[ring_idx(head, i)]; },
                           t_lo, tm + 2 - t_lo, fm - 10, fm + 10, 0.25, tid, 1);
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
  prof->begin(K_RX_TRACK);
  rx_track_kernel<<<S < n_sm ? S : n_sm, TRK_THREADS, sizeof(TrackSmem), stream>>>(T, B.ctl, B.ring, B.rowsum, B.uw_errors, B.track_list,
                                                                               cnt, ret_out, B.dec_active, B.nin);
  prof->end(K_RX_TRACK); prof->begin(K_RX_DETECT);
  int det_grid = S * (RADE_NMF / DET_TB); if (det_grid > n_sm * 4) det_grid = n_sm * 4;
  rx_detect_kernel<<<det_grid, DET_THREADS, sizeof(DetectSmem), ss>>>(T, B.ctl, B.ring, B.rowsum, B.search_list, cnt);
  prof->end(K_RX_DETECT); prof->begin(K_RX_DEMOD);
  rx_demod_kernel<<<S, 192, 0, stream>>>(T, B.ctl, B.ring, B.z_hat, B.eoo, active);
  prof->end(K_RX_DEMOD); prof->begin(K_RX_FINISH);
  rx_finish_kernel<<<S < 2 * n_sm ? S : 2 * n_sm, 256, sizeof(FinishSmem), ss>>>(T, B.ctl, B.ring, B.rowsum, B.uw_errors, B.dec_state,
      reset_dec_on_sync, ret_out, B.dec_active, B.nin, active, B.search_list, cnt, cnt_next, S);
  prof->end(K_RX_FINISH);
  if (fork) { CUDA_CHECK(cudaEventRecord(B.ev_join, ss)); CUDA_CHECK(cudaStreamWaitEvent(stream, B.ev_join, 0)); }
  CUDA_CHECK(cudaGetLastError());
  return 5;
}
