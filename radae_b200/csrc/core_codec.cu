// RADE core encoder / decoder on B200 — batched stateful streams, weights streamed by TMA.
//
// Replaces (per stream, per 40 ms step):
//   rade_core_encoder  /root/reference/src/rade_enc.c:55-114   (PyTorch twin radae/radae_base.py:260-286)
//   rade_core_decoder  /root/reference/src/rade_dec.c:50-102   (PyTorch twin radae/radae_base.py:400-416)
// and the opus DNN primitives they call (compute_generic_dense/gru/conv1d[_dilation], compute_glu).
//
// Design.  One CTA owns a tile of TS = 8 or 16 independent streams (template parameter) and walks
// the whole layer stack for n_steps consecutive steps with every activation on chip:
//   * a PRODUCER warp issues `cp.async.bulk` (TMA 1-D bulk copies, mbarrier complete_tx) that stream the step's weights —
//     pre-arranged on the host as one contiguous sequence of <= 32 KB chunks in consumption order, descriptors in constant
//     memory — from L2 into a 4/5-stage shared-memory ring; the consumer warps wait on the stage's "full" mbarrier, use the
//     chunk and arrive on its "empty" mbarrier.  No consumer ever waits on a global/L2 load for a weight.  The per-stream state
//     rows arrive the same way (three bulk copies per stream, once per launch).
//   * the consumers are split into I-WARPS (int8 GRU / conv / GLU layers on the tensor cores + their epilogues) and F-WARPS
//     (the float layers: dense1, and the wide zdense / output layer accumulated incrementally as each DenseNet segment is
//     produced).  The float layers never feed the int8 chain within a step, so they run concurrently with it; segments are
//     handed over through two shared-memory buffers guarded by named barriers (bar.arrive / bar.sync).  dense1 of step t+1 is
//     computed one step ahead and published when the I-warps finish step t.
//   * the DenseNet concat buffer lives in shared memory as int8 (exactly the quantised values floor(.5+127x) the
//     reference feeds its int8 GEMVs) — a ring of the current and the previous one/two steps, which is also the conv1d
//     tap memory and the GRU recurrent input, so conv "state" costs no extra storage;
//   * int8 layers: s8 x s8 -> s32 tensor-core MMA (m16n8k32: 16 streams x 8 outputs x 32 inputs per instruction), A fragments
//     by ldmatrix, B fragments read from the staged chunk with one conflict-free LDS.64 per lane;
//   * epilogues (scale, bias, rational tanh/sigmoid, GRU gating, GLU) straight from the accumulator registers, each
//     float operation separately rounded in the reference's order => bit-identical to the C oracle; conv layers split
//     their two taps over different warps and reduce the exact int32 partial sums through shared memory;
//   * the float layers accumulate sequentially over inputs (the generic sgemv order): packed FMUL2 products, scalar FADDs.
// Per-stream HBM state: GRU h (fp32) + int8 concat of step t-1 (and t-2 for the encoder's dilation-2 convs).
#include <string.h>
#include "rade_common.h"
#include "rade_host.h"
#include "tma.cuh"
#include "codec_math.cuh"

namespace {

constexpr int SEG_LD = 100;                // float rows are 16-byte aligned (float4 reads) with strides that spread banks
constexpr int FIN_LD = 92;
constexpr int ZIN_LD = 84;
constexpr int HQ_LD = 496;                 // row stride of the decoder's quantised hidden states (conflict-free A fragments)

template <int NST> struct PipeSmem {
  alignas(128) unsigned char ring[NST][CORE_STAGE_BYTES];
  alignas(8) uint64_t full[NST];
  alignas(8) uint64_t empty[NST];
  alignas(8) uint64_t state_bar;           // per-stream state rows, bulk-copied once at kernel start
};

// consumer-side cursor over the chunk stream; every consumer thread carries an identical copy
template <int NST> struct Cursor {
  PipeSmem<NST> *p; int stage; uint32_t phase;
  __device__ __forceinline__ const unsigned char *acquire() { mbar_wait(&p->full[stage], phase); return p->ring[stage]; }
  __device__ __forceinline__ void release() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&p->empty[stage]);
    if (++stage == NST) { stage = 0; phase ^= 1; }
  }
};

// named barriers: the consumer warps are split into I-warps (int8 layers) and F-warps (float layers)
enum { BAR_I = 1, BAR_F = 2, BAR_SEG_FULL = 3 /* +buf */, BAR_SEG_EMPTY = 5 /* +buf */, BAR_D1 = 7, BAR_ALL = 8 };
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
template <int NI> __device__ __forceinline__ void i_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NI * 32) : "memory"); }
template <int NF> __device__ __forceinline__ void f_sync() { asm volatile("bar.sync 2, %0;" ::"n"(NF * 32) : "memory"); }

// chunk descriptors of both weight streams in constant memory: the producer must not pay a global-memory round trip per chunk
// (the layout depends only on the layer dimensions, so one table per device serves every context)
__constant__ ChunkDesc c_chunks[2][CORE_MAX_CHUNKS];

// producer: lane 0 of the last warp streams every chunk of every step
template <int NST> __device__ void producer_loop(PipeSmem<NST> *p, const CodecStreamDev &ws, int which, int T) {
  int stage = 0; uint32_t phase = 0;
  for (int t = 0; t < T; t++)
    for (int c = (t == 0 ? 0 : ws.n_prologue); c < ws.n_chunks; c++) {
      const ChunkDesc d = c_chunks[which][c];
      mbar_wait(&p->empty[stage], phase ^ 1);
      mbar_expect_tx(&p->full[stage], d.bytes);
      bulk_g2s(p->ring[stage], ws.stream + d.offset, d.bytes, &p->full[stage]);
      if (++stage == NST) { stage = 0; phase ^= 1; }
    }
}

// ---------------------------------------------------------------- tensor-core int8 tile GEMM on staged weights
__device__ __forceinline__ void mma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Walk the KB k-blocks of one A operand; the weights arrive as chunks of core_kbc(NTL) k-blocks laid out [kb][nt][lane].
// Every consumer warp calls this (acquire/release are collective); only warps with `work` issue MMAs, for their NT n-tiles.
// A fragment of m16n8k32 (16 rows x 32 bytes) in one instruction: four 8x8 b16 matrices = rows 0-7 / 8-15 x bytes 0-15 / 16-31
__device__ __forceinline__ void ldsm_a(uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t &a3, const int8_t *lane_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(smem_u32(lane_ptr)));
}

template <int NT, int TS, typename CX>
__device__ __forceinline__ void gemm_stream(CX &cx, int (&acc)[NT][4], const int (&nt)[NT], const bool (&use)[NT], bool work,
                                            const int8_t *A, int lda, int KB, int NTL) {
  const int lane = threadIdx.x & 31;
  // ldmatrix row address of this lane: matrix (lane>>3): rows +8 for odd matrices, bytes +16 for matrices 2,3.  8-stream tiles
  // have no rows 8..15: those lanes re-read rows 0..7 (the MMA rows they feed are never used) instead of whatever lies behind
  const int8_t *arow = A + ((lane & 7) + (TS == 16 ? ((lane >> 3) & 1) * 8 : 0)) * lda + ((lane >> 4) & 1) * 16;
  const int kbc = core_kbc(NTL);
  for (int kb0 = 0; kb0 < KB; kb0 += kbc) {
    const int nk = min(kbc, KB - kb0);
    const uint2 *Wc = reinterpret_cast<const uint2 *>(cx.acquire());
    if (work) {
#pragma unroll 2
      for (int kb = 0; kb < nk; kb++) {
        uint32_t a0, a1, a2, a3;
        ldsm_a(a0, a1, a2, a3, arow + (kb0 + kb) * 32);
#pragma unroll
        for (int i = 0; i < NT; i++)
          if (use[i]) {                        // warp-uniform
            const uint2 b = Wc[(kb * NTL + nt[i]) * 32 + lane];
            mma_s8(acc[i], a0, a1, a2, a3, b.x, b.y);
          }
      }
    }
    cx.release();
  }
}

// chunks a warp group does not need are acquired and released untouched (keeps the ring's arrival counts uniform)
template <typename CX> __device__ __forceinline__ void skip_chunks(CX &cx, int n) {
  for (int i = 0; i < n; i++) { cx.acquire(); cx.release(); }
}
__device__ __forceinline__ int i8_chunks(int K, int N) { const int kbc = core_kbc(N / 8); return (K / 32 + kbc - 1) / kbc; }
__device__ __forceinline__ int f32_chunks(int K, int NOUTP) { const int rpc = core_f32_rpc(NOUTP); return (K + rpc - 1) / rpc; }

// ---------------------------------------------------------------- float layers: sequential-in-j accumulation
// Run by the F-warps concurrently with the int8 layers.  F-thread (s = ft % TS, grp = ft / TS) owns the OPT = TS outputs
// o = OPT*grp .. OPT*grp+OPT-1 of stream s; the W rows of the concat segment (zero-padded to NOUTP floats) come from the
// staged chunk(s) as broadcast LDS.128, the inputs as one LDS.128 per four.
// acc[i] = ((acc[i] + W[j0][o] x[j0]) + W[j0+1][o] x[j0+1]) + ...  — separately rounded, in input order
// (round(w.x * x), round(w.y * x)) with one packed multiply (SASS FMUL2); the accumulation stays scalar so that ptxas cannot
// contract product and sum into an FFMA2 (it does that to mul.rn.f32x2 + add.rn.f32x2, which would round once instead of twice)
template <int NOUTP, int OPT, typename CX>
__device__ __forceinline__ void dense_seg(CX &cx, float (&acc)[OPT], const float *xrow, int K, int grp) {
  constexpr int RPC = core_f32_rpc(NOUTP);
  const bool act = grp < NOUTP / OPT;
  for (int r0 = 0; r0 < K; r0 += RPC) {
    const int n = min(RPC, K - r0);
    const float4 *W4 = reinterpret_cast<const float4 *>(cx.acquire()) + grp * (OPT / 4);
    if (act) {
      const float4 *x4 = reinterpret_cast<const float4 *>(xrow + r0);
#pragma unroll 2
      for (int j = 0; j < n; j += 4) {
        const float4 x = x4[j >> 2];
        mac_row<OPT>(acc, W4 + (j + 0) * (NOUTP / 4), x.x); mac_row<OPT>(acc, W4 + (j + 1) * (NOUTP / 4), x.y);
        mac_row<OPT>(acc, W4 + (j + 2) * (NOUTP / 4), x.z); mac_row<OPT>(acc, W4 + (j + 3) * (NOUTP / 4), x.w);
      }
    }
    cx.release();
  }
}

// GRU layer (I-warps): warp w owns unit tile w (8 hidden units x 16 streams); gates z,r,n of a unit land in the same accumulator slot.
// The per-output scale / bias values are fetched BEFORE the GEMM so their L2 latency hides behind it.
template <int UNITS, int TS, typename CX, typename Emit>
__device__ __forceinline__ void gru_layer(CX &cx, const I8LayerDev &Li, const I8LayerDev &Lr, const int8_t *Xin, int ldx,
                                          const int8_t *Hq, int ldh, float *hs, int ldhs, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  constexpr int U = UNITS / 8;
  const bool work = warp < U;                // one unit tile per I-warp
  const int u = work ? warp : 0;
  const int j0 = u * 8 + 2 * tig;
  float2 si[3], bi[3], sr[3], br[3];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    si[q] = *reinterpret_cast<const float2 *>(Li.scale + q * UNITS + j0); bi[q] = *reinterpret_cast<const float2 *>(Li.bias + q * UNITS + j0);
    sr[q] = *reinterpret_cast<const float2 *>(Lr.scale + q * UNITS + j0); br[q] = *reinterpret_cast<const float2 *>(Lr.bias + q * UNITS + j0);
  }
  int ai[3][4] = {}, ar[3][4] = {};
  const int nt[3] = {u, U + u, 2 * U + u};
  const bool use[3] = {true, true, true};
  gemm_stream<3, TS>(cx, ai, nt, use, work, Xin, ldx, Li.K / 32, 3 * U);
  gemm_stream<3, TS>(cx, ar, nt, use, work, Hq, ldh, Lr.K / 32, 3 * U);
  if (!work) return;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int row = g + ((e & 2) ? 8 : 0);
    if (row >= TS) continue;                  // 8-stream tiles leave MMA rows 8..15 unused
    const int j = j0 + (e & 1);
#define PICK(v) ((e & 1) ? (v).y : (v).x)
    float z = sigmoid_r(__fadd_rn(lin(ai[0][e], PICK(si[0]), PICK(bi[0])), lin(ar[0][e], PICK(sr[0]), PICK(br[0]))));
    float r = sigmoid_r(__fadd_rn(lin(ai[1][e], PICK(si[1]), PICK(bi[1])), lin(ar[1][e], PICK(sr[1]), PICK(br[1]))));
    float n = tanh_r(__fadd_rn(lin(ai[2][e], PICK(si[2]), PICK(bi[2])), __fmul_rn(lin(ar[2][e], PICK(sr[2]), PICK(br[2])), r)));
#undef PICK
    float hold = hs[row * ldhs + j];
    float h = __fadd_rn(__fmul_rn(z, hold), __fmul_rn(__fsub_rn(1.f, z), n));
    hs[row * ldhs + j] = h;
    emit(row, j, h);
  }
}

// conv1d (k=2): work unit = (n-tile, tap); units are dealt round-robin to the consumer warps, each accumulates its tap's
// K range, the exact int32 partial sums meet in shared memory `red[2][TS][N]`, then all consumer threads run the epilogue
template <int N, int NCW, int TS, typename CX, typename Emit>
__device__ __forceinline__ void conv_layer(CX &cx, const I8LayerDev &L, const int8_t *Aold, const int8_t *Acur, int Ktap, int lda,
                                           int *red, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  constexpr int NTL = N / 8;
  constexpr int UNITS = 2 * NTL;
  constexpr int PER = (UNITS + NCW - 1) / NCW;
  constexpr int NEP = (TS * N + NCW * 32 - 1) / (NCW * 32);      // epilogue elements per thread
  float es[NEP], eb[NEP];
#pragma unroll
  for (int q = 0; q < NEP; q++) {
    const int e = threadIdx.x + q * NCW * 32;
    es[q] = (e < TS * N) ? L.scale[e % N] : 0.f; eb[q] = (e < TS * N) ? L.bias[e % N] : 0.f;
  }
#pragma unroll
  for (int tap = 0; tap < 2; tap++) {
    int acc[PER][4] = {};
    int nt[PER], ntc[PER]; bool use[PER]; bool any = false;
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const int u = warp + NCW * i;
      const bool mine = (u < UNITS) && (u / NTL == tap);
      nt[i] = mine ? (u % NTL) : -1;
      ntc[i] = mine ? (u % NTL) : 0;
      use[i] = mine;
      any |= mine;
    }
    gemm_stream<PER, TS>(cx, acc, ntc, use, any, tap ? Acur : Aold, lda, Ktap / 32, NTL);
#pragma unroll
    for (int i = 0; i < PER; i++)
      if (nt[i] >= 0) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int row = g + ((e & 2) ? 8 : 0);
          const int n = nt[i] * 8 + 2 * tig + (e & 1);
          if (row < TS) red[(tap * TS + row) * N + n] = acc[i][e];
        }
      }
  }
  i_sync<NCW>();
#pragma unroll
  for (int q = 0; q < NEP; q++) {
    const int e = threadIdx.x + q * NCW * 32;
    if (e < TS * N) {
      const int row = e / N, n = e % N;
      emit(row, n, lin(red[row * N + n] + red[(TS + row) * N + n], es[q], eb[q]));
    }
  }
}

// ================================================================= encoder
template <int TS, int NST> struct EncSmem {
  PipeSmem<NST> pipe;
  alignas(16) int8_t cb[3][TS][ENC_LDA];
  alignas(16) float hs[TS][5 * ENC_GRU];
  alignas(16) float seg[3][TS][SEG_LD];       // [0], [1]: segments handed from the I-warps to the F-warps; [2]: dense1 output
  alignas(16) float fin[TS][FIN_LD];
  alignas(16) int8_t d1q[TS][64];             // dense1 output of the next step, parked until its concat buffer is free
  int red[2 * TS * ENC_CONV];
  int any_active;
};

template <int TS, int NST>
__global__ void __launch_bounds__((ENC_NI + ENC_NF + 1) * 32, 1)
core_encoder_kernel(CoreWeightsDev W, EncStreamState *__restrict__ state, const float *__restrict__ in, int in_mode,
                    float *__restrict__ z_out, const uint8_t *__restrict__ active, int S, int T) {
  constexpr int NI = ENC_NI, NF = ENC_NF, NIT = NI * 32, NCT = (NI + NF) * 32, OPT = TS / 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef EncSmem<TS, NST> Smem;
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * TS;               // TS = 16 (full MMA tile) or 8 (more CTAs when the batch is small)

  if (tid == 0) sm.any_active = 0;
  __syncthreads();
  if (tid < TS && s0 + tid < S && (!active || active[s0 + tid])) sm.any_active = 1;
  if (tid == 0) {
    for (int i = 0; i < NST; i++) { mbar_init(&sm.pipe.full[i], 1); mbar_init(&sm.pipe.empty[i], NI + NF); }
    mbar_init(&sm.pipe.state_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (!sm.any_active) return;
  if (tid >= NCT) {                               // ---- producer warp
    if (tid == NCT) {
      // per-stream state rows first (one bulk copy per array, all in flight at once), then the weight stream
      const int rows = min(TS, S - s0);
      mbar_expect_tx(&sm.pipe.state_bar, (uint32_t)(rows * (sizeof(float) * 5 * ENC_GRU + 2 * ENC_LDA)));
      for (int r = 0; r < rows; r++) {
        const EncStreamState *st = state + (s0 + r);
        bulk_g2s(sm.hs[r], st->h, sizeof(float) * 5 * ENC_GRU, &sm.pipe.state_bar);
        bulk_g2s(sm.cb[2][r], st->cat1, ENC_LDA, &sm.pipe.state_bar);
        bulk_g2s(sm.cb[1][r], st->cat2, ENC_LDA, &sm.pipe.state_bar);
      }
      producer_loop<NST>(&sm.pipe, W.enc_stream, 0, T);
    }
    return;
  }
  Cursor<NST> cx{&sm.pipe, 0, 0u};

  if (tid >= NIT) {
    // =========================== F-warps: input staging, dense1, incremental zdense, z output
    const int ft = tid - NIT, sl = ft % TS, grp = ft / TS;
    const int sg = s0 + sl;
    bar_arrive(BAR_SEG_EMPTY + 0, NCT); bar_arrive(BAR_SEG_EMPTY + 1, NCT);       // both hand-over buffers start free
    bar_sync(BAR_ALL, NCT);                                                       // concat buffers initialised by the I-warps
    // input of step t -> fin (API layout: [S][4T][36]; 20 used features + aux = -1, src/rade_api.c:426-432)
    auto stage_input = [&](int t) {
      for (int i = ft; i < TS * ENC_IN; i += NF * 32) {
        const int r = i / ENC_IN, k = i % ENC_IN;
        float v = 0.f;
        if (s0 + r < S) {
          if (in_mode == 0) v = in[((size_t)(s0 + r) * T + t) * ENC_IN + k];
          else {
            const int fr = k / 21, f = k % 21;
            v = (f == 20) ? -1.f : in[((size_t)(s0 + r) * 4 * T + 4 * t + fr) * RADE_NB_TOTAL_FEATURES + f];
          }
        }
        sm.fin[r][k] = v;
      }
    };
    // dense1: tanh(W f + b), 84 -> 64, from fin into seg[2] (float, for zdense) and d1q (int8, copied into the concat buffer
    // of its step once the I-warps are done with the previous step)
    auto dense1 = [&]() {
      float a[OPT];
#pragma unroll
      for (int i = 0; i < OPT; i++) a[i] = 0.f;
      dense_seg<64, OPT>(cx, a, sm.fin[sl], ENC_IN, grp);
      if (grp < 64 / OPT) {
#pragma unroll
        for (int i = 0; i < OPT; i++) {
          const int o = OPT * grp + i;
          float y = tanh_r(__fadd_rn(a[i], W.enc_dense1.bias[o]));
          sm.seg[2][sl][o] = y;
          sm.d1q[sl][o] = quant8(y);
        }
      }
    };
    auto publish_d1 = [&](int t) {               // d1q -> cur(t)[:, 0:64), then the I-warps may start GRU 1 of step t
      for (int i = ft; i < TS * 16; i += NF * 32)
        reinterpret_cast<uint32_t *>(sm.cb[t % 3][i / 16])[i % 16] = reinterpret_cast<const uint32_t *>(sm.d1q[i / 16])[i % 16];
      f_sync<NF>();
      bar_arrive(BAR_D1, NCT);
    };
    stage_input(0);
    f_sync<NF>();
    dense1();                                    // prologue chunk
    f_sync<NF>();
    publish_d1(0);
    for (int t = 0; t < T; t++) {
      float zacc[OPT];
#pragma unroll
      for (int i = 0; i < OPT; i++) zacc[i] = 0.f;
      dense_seg<RADE_LATENT, OPT>(cx, zacc, sm.seg[2][sl], 64, grp);
      if (t + 1 < T) stage_input(t + 1);         // long before it is needed: the global-load latency is off the critical path
      int off = 64;
#pragma unroll 1
      for (int l = 0; l < 5; l++) {
        skip_chunks(cx, i8_chunks(off, 3 * ENC_GRU) + i8_chunks(ENC_GRU, 3 * ENC_GRU));
        bar_sync(BAR_SEG_FULL + 1, NCT);
        dense_seg<RADE_LATENT, OPT>(cx, zacc, sm.seg[1][sl], ENC_GRU, grp);
        bar_arrive(BAR_SEG_EMPTY + 1, NCT);
        off += ENC_GRU;
        if (l == 4) {                            // dense1 of the NEXT step while the I-warps run the last conv layer
          f_sync<NF>();
          if (t + 1 < T) dense1(); else skip_chunks(cx, f32_chunks(ENC_IN, 64));
          f_sync<NF>();
        }
        skip_chunks(cx, 2 * i8_chunks(off, ENC_CONV));
        bar_sync(BAR_SEG_FULL + 0, NCT);         // l == 4: the I-warps have finished this step
        if (l == 4 && t + 1 < T) publish_d1(t + 1);
        dense_seg<RADE_LATENT, OPT>(cx, zacc, sm.seg[0][sl], ENC_CONV, grp);
        bar_arrive(BAR_SEG_EMPTY + 0, NCT);
        off += ENC_CONV;
      }
      // ---- z = zdense(cat) + b   (bottleneck 3: linear; bottleneck 1: tanh -- src/rade_enc.c:107-113)
      if (sg < S && (!active || active[sg]) && grp < RADE_LATENT / OPT) {
#pragma unroll
        for (int i = 0; i < OPT; i++) {
          const float v = __fadd_rn(zacc[i], W.enc_zdense.bias[OPT * grp + i]);
          z_out[((size_t)sg * T + t) * RADE_LATENT + OPT * grp + i] = W.enc_z_tanh ? tanh_r(v) : v;
        }
      }
    }
    return;
  }

  // =========================== I-warps: GRU / conv layers on the tensor cores + their epilogues
  for (int r = 0; r < TS; r++) {
    const bool ok = s0 + r < S;
    for (int i = tid; i < ENC_LDA / 4; i += NIT) {
      reinterpret_cast<uint32_t *>(sm.cb[0][r])[i] = 0u;
      if (!ok) { reinterpret_cast<uint32_t *>(sm.cb[2][r])[i] = 0u; reinterpret_cast<uint32_t *>(sm.cb[1][r])[i] = 0u; }
    }
    if (!ok) for (int i = tid; i < 5 * ENC_GRU; i += NIT) sm.hs[r][i] = 0.f;
  }
  mbar_wait(&sm.pipe.state_bar, 0);
  bar_sync(BAR_ALL, NCT);

  constexpr int dil[5] = {1, 2, 2, 2, 2};
  for (int t = 0; t < T; t++) {
    int8_t(*cur)[ENC_LDA] = sm.cb[t % 3];
    int8_t(*prev1)[ENC_LDA] = sm.cb[(t + 2) % 3];
    int8_t(*prev2)[ENC_LDA] = sm.cb[(t + 1) % 3];
    skip_chunks(cx, (t == 0 ? f32_chunks(ENC_IN, 64) : 0) + f32_chunks(64, RADE_LATENT));
    bar_sync(BAR_D1, NCT);                       // dense1 output (int8) is in cur[:, 0:64)
    int off = 64;
#pragma unroll 1
    for (int l = 0; l < 5; l++) {
      // GRU l: input = cur[0:off), recurrent input = quantised h(t-1) = prev1[off : off+64)
      bar_sync(BAR_SEG_EMPTY + 1, NCT);
      gru_layer<ENC_GRU, TS>(cx, W.enc_gru_in[l], W.enc_gru_rec[l], &cur[0][0], ENC_LDA, &prev1[0][off], ENC_LDA,
                             &sm.hs[0][l * ENC_GRU], 5 * ENC_GRU,
                             [&](int row, int j, float h) { sm.seg[1][row][j] = h; cur[row][off + j] = quant8(h); });
      bar_arrive(BAR_SEG_FULL + 1, NCT);
      i_sync<NI>();
      skip_chunks(cx, f32_chunks(ENC_GRU, RADE_LATENT) + (l == 4 ? f32_chunks(ENC_IN, 64) : 0));
      off += ENC_GRU;
      // conv l (k = 2): tap 0 = concat prefix of step t-dilation, tap 1 = current prefix
      const int8_t *old = (dil[l] == 1) ? &prev1[0][0] : &prev2[0][0];
      bar_sync(BAR_SEG_EMPTY + 0, NCT);
      conv_layer<ENC_CONV, NI, TS>(cx, W.enc_conv[l], old, &cur[0][0], off, ENC_LDA, sm.red,
                                   [&](int row, int n, float v) { float y = tanh_r(v); sm.seg[0][row][n] = y; cur[row][off + n] = quant8(y); });
      bar_arrive(BAR_SEG_FULL + 0, NCT);
      i_sync<NI>();
      skip_chunks(cx, f32_chunks(ENC_CONV, RADE_LATENT));
      off += ENC_CONV;
    }
  }

  const int last = (T + 2) % 3, last2 = (T + 1) % 3;
  for (int r = 0; r < TS; r++) {
    if (s0 + r >= S || (active && !active[s0 + r])) continue;
    EncStreamState *st = state + (s0 + r);
    for (int i = tid; i < 5 * ENC_GRU; i += NIT) st->h[i] = sm.hs[r][i];
    for (int i = tid; i < ENC_LDA / 4; i += NIT) {
      reinterpret_cast<uint32_t *>(st->cat1)[i] = reinterpret_cast<const uint32_t *>(sm.cb[last][r])[i];
      reinterpret_cast<uint32_t *>(st->cat2)[i] = reinterpret_cast<const uint32_t *>(sm.cb[last2][r])[i];
    }
  }
}

// ================================================================= decoder
template <int TS, int NST> struct DecSmem {
  PipeSmem<NST> pipe;
  alignas(16) int8_t cb[2][TS][DEC_LDA];
  alignas(16) int8_t hq[2][TS][HQ_LD];
  alignas(16) float hs[TS][5 * DEC_GRU];
  alignas(16) float seg[3][TS][SEG_LD];
  alignas(16) float zin[TS][ZIN_LD];
  alignas(16) int8_t d1q[TS][96];
  int red[2 * TS * DEC_CONV];
  int any_active;
};

// out_mode 0: features [S][T][84];  out_mode 1: API layout [S][4T][36] (20 used, rest zero, src/rade_api.c:488-500)
// uw_count (optional): += number of steps whose first aux symbol (feature 20) is > 0 (src/rade_api.c:502-505)
template <int TS, int NST>
__global__ void __launch_bounds__((DEC_NI + DEC_NF + 1) * 32, 1)
core_decoder_kernel(CoreWeightsDev W, DecStreamState *__restrict__ state, const float *__restrict__ z_in,
                    float *__restrict__ out, int out_mode, int *__restrict__ uw_count,
                    const uint8_t *__restrict__ active, int S, int T) {
  constexpr int NI = DEC_NI, NF = DEC_NF, NIT = NI * 32, NCT = (NI + NF) * 32, OPT = TS;
  constexpr int NGRP = NF * 32 / TS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef DecSmem<TS, NST> Smem;
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * TS;               // TS = 16 (full MMA tile) or 8 (more CTAs when the batch is small)

  if (tid == 0) sm.any_active = 0;
  __syncthreads();
  if (tid < TS && s0 + tid < S && (!active || active[s0 + tid])) sm.any_active = 1;
  if (tid == 0) {
    for (int i = 0; i < NST; i++) { mbar_init(&sm.pipe.full[i], 1); mbar_init(&sm.pipe.empty[i], NI + NF); }
    mbar_init(&sm.pipe.state_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (!sm.any_active) return;
  if (tid >= NCT) {
    if (tid == NCT) {
      const int rows = min(TS, S - s0);
      mbar_expect_tx(&sm.pipe.state_bar, (uint32_t)(rows * (sizeof(float) * 5 * DEC_GRU + DEC_LDA)));
      for (int r = 0; r < rows; r++) {
        const DecStreamState *st = state + (s0 + r);
        bulk_g2s(sm.hs[r], st->h, sizeof(float) * 5 * DEC_GRU, &sm.pipe.state_bar);
        bulk_g2s(sm.cb[1][r], st->cat1, DEC_LDA, &sm.pipe.state_bar);
      }
      producer_loop<NST>(&sm.pipe, W.dec_stream, 1, T);
    }
    return;
  }
  Cursor<NST> cx{&sm.pipe, 0, 0u};

  if (tid >= NIT) {
    // =========================== F-warps: z staging, dense1, incremental output layer, feature output
    const int ft = tid - NIT, sl = ft % TS, grp = ft / TS;
    const int sg = s0 + sl;
    bar_arrive(BAR_SEG_EMPTY + 0, NCT); bar_arrive(BAR_SEG_EMPTY + 1, NCT);
    bar_sync(BAR_ALL, NCT);
    auto stage_input = [&](int t) {
      for (int i = ft; i < TS * DEC_IN; i += NF * 32) {
        const int r = i / DEC_IN, k = i % DEC_IN;
        sm.zin[r][k] = (s0 + r < S) ? z_in[((size_t)(s0 + r) * T + t) * DEC_IN + k] : 0.f;
      }
    };
    // dense1: tanh(W z + b), 80 -> 96, into seg[2] (float, for the output layer) and d1q (int8, published to the concat
    // buffer of its step once the I-warps are done with the previous step)
    auto dense1 = [&]() {
      float a[OPT];
#pragma unroll
      for (int i = 0; i < OPT; i++) a[i] = 0.f;
      dense_seg<96, OPT>(cx, a, sm.zin[sl], DEC_IN, grp);
#pragma unroll
      for (int i = 0; i < OPT; i++) {
        const int o = OPT * grp + i;
        float y = tanh_r(__fadd_rn(a[i], W.dec_dense1.bias[o]));
        sm.seg[2][sl][o] = y;
        sm.d1q[sl][o] = quant8(y);
      }
    };
    auto publish_d1 = [&](int t) {
      for (int i = ft; i < TS * 24; i += NF * 32)
        reinterpret_cast<uint32_t *>(sm.cb[t & 1][i / 24])[i % 24] = reinterpret_cast<const uint32_t *>(sm.d1q[i / 24])[i % 24];
      f_sync<NF>();
      bar_arrive(BAR_D1, NCT);
    };
    stage_input(0);
    f_sync<NF>();
    dense1();                                    // prologue chunk
    f_sync<NF>();
    publish_d1(0);
    for (int t = 0; t < T; t++) {
      float oacc[OPT];
#pragma unroll
      for (int i = 0; i < OPT; i++) oacc[i] = 0.f;
      dense_seg<DEC_OUTP, OPT>(cx, oacc, sm.seg[2][sl], 96, grp);
      if (t + 1 < T) stage_input(t + 1);
      int off = 96;
#pragma unroll 1
      for (int l = 0; l < 5; l++) {
        skip_chunks(cx, i8_chunks(off, 3 * DEC_GRU) + i8_chunks(DEC_GRU, 3 * DEC_GRU) + i8_chunks(DEC_GRU, DEC_GRU));
        bar_sync(BAR_SEG_FULL + 1, NCT);
        dense_seg<DEC_OUTP, OPT>(cx, oacc, sm.seg[1][sl], DEC_GRU, grp);
        bar_arrive(BAR_SEG_EMPTY + 1, NCT);
        off += DEC_GRU;
        if (l == 4) {                            // dense1 of the NEXT step while the I-warps run the last conv layer
          f_sync<NF>();
          if (t + 1 < T) dense1(); else skip_chunks(cx, f32_chunks(DEC_IN, 96));
          f_sync<NF>();
        }
        skip_chunks(cx, 2 * i8_chunks(off, DEC_CONV));
        bar_sync(BAR_SEG_FULL + 0, NCT);
        if (l == 4 && t + 1 < T) publish_d1(t + 1);
        dense_seg<DEC_OUTP, OPT>(cx, oacc, sm.seg[0][sl], DEC_CONV, grp);
        bar_arrive(BAR_SEG_EMPTY + 0, NCT);
        off += DEC_CONV;
      }
      if (sg < S && (!active || active[sg])) {
#pragma unroll
        for (int i = 0; i < OPT; i++) {
          const int o = OPT * grp + i;
          if (o >= DEC_OUT) continue;
          const float v = __fadd_rn(oacc[i], W.dec_output.bias[o]);
          if (out_mode == 0) out[((size_t)sg * T + t) * DEC_OUT + o] = v;
          else {
            const int fr = o / 21, f = o % 21;
            if (f < 20) out[((size_t)sg * 4 * T + 4 * t + fr) * RADE_NB_TOTAL_FEATURES + f] = v;
          }
          if (o == 20 && uw_count && v > 0.f) atomicAdd(&uw_count[sg], 1);
        }
        if (out_mode == 1) {            // zero the 16 unused slots of each 36-wide vector
          for (int k = grp; k < 4 * 16; k += NGRP)
            out[((size_t)sg * 4 * T + 4 * t + k / 16) * RADE_NB_TOTAL_FEATURES + 20 + (k % 16)] = 0.f;
        }
      }
    }
    return;
  }

  // =========================== I-warps
  for (int r = 0; r < TS; r++) {
    const bool ok = s0 + r < S;
    for (int i = tid; i < DEC_LDA / 4; i += NIT) {
      reinterpret_cast<uint32_t *>(sm.cb[0][r])[i] = 0u;
      if (!ok) reinterpret_cast<uint32_t *>(sm.cb[1][r])[i] = 0u;
    }
    if (!ok) for (int i = tid; i < 5 * DEC_GRU; i += NIT) sm.hs[r][i] = 0.f;
  }
  mbar_wait(&sm.pipe.state_bar, 0);
  i_sync<NI>();
  for (int i = tid; i < TS * 5 * DEC_GRU; i += NIT) {
    const int r = i / (5 * DEC_GRU), k = i % (5 * DEC_GRU);
    sm.hq[0][r][k] = quant8(sm.hs[r][k]);
  }
  bar_sync(BAR_ALL, NCT);

  for (int t = 0; t < T; t++) {
    int8_t(*cur)[DEC_LDA] = sm.cb[t & 1];
    int8_t(*prev1)[DEC_LDA] = sm.cb[(t + 1) & 1];
    int8_t(*hq_rd)[HQ_LD] = sm.hq[t & 1];
    int8_t(*hq_wr)[HQ_LD] = sm.hq[(t + 1) & 1];
    skip_chunks(cx, (t == 0 ? f32_chunks(DEC_IN, 96) : 0) + f32_chunks(96, DEC_OUTP));
    bar_sync(BAR_D1, NCT);
    int off = 96;
#pragma unroll 1
    for (int l = 0; l < 5; l++) {
      // GRU l on cur[0:off); its new state is kept un-gated (src/rade_dec.c:66-67)
      gru_layer<DEC_GRU, TS>(cx, W.dec_gru_in[l], W.dec_gru_rec[l], &cur[0][0], DEC_LDA, &hq_rd[0][l * DEC_GRU], HQ_LD,
                             &sm.hs[0][l * DEC_GRU], 5 * DEC_GRU,
                             [&](int row, int j, float h) { hq_wr[row][l * DEC_GRU + j] = quant8(h); });
      i_sync<NI>();
      // GLU l: out = h * sigmoid(Wg h + b)  -> concat;  12 n-tiles, one per warp, a single 9 KB chunk
      bar_sync(BAR_SEG_EMPTY + 1, NCT);
      {
        const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
        const I8LayerDev &L = W.dec_glu[l];
        const float2 gs = *reinterpret_cast<const float2 *>(L.scale + warp * 8 + 2 * tig);
        const float2 gb = *reinterpret_cast<const float2 *>(L.bias + warp * 8 + 2 * tig);
        int acc[1][4] = {};
        const int nt[1] = {warp};
        const bool use[1] = {true};
        gemm_stream<1, TS>(cx, acc, nt, use, true, &hq_wr[0][l * DEC_GRU], HQ_LD, DEC_GRU / 32, DEC_GRU / 8);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int row = g + ((e & 2) ? 8 : 0);
          if (row >= TS) continue;
          const int n = warp * 8 + 2 * tig + (e & 1);
          float y = __fmul_rn(sm.hs[row][l * DEC_GRU + n], sigmoid_r(lin(acc[0][e], (e & 1) ? gs.y : gs.x, (e & 1) ? gb.y : gb.x)));
          sm.seg[1][row][n] = y; cur[row][off + n] = quant8(y);
        }
      }
      bar_arrive(BAR_SEG_FULL + 1, NCT);
      i_sync<NI>();
      skip_chunks(cx, f32_chunks(DEC_GRU, DEC_OUTP) + (l == 4 ? f32_chunks(DEC_IN, 96) : 0));
      off += DEC_GRU;
      bar_sync(BAR_SEG_EMPTY + 0, NCT);
      conv_layer<DEC_CONV, NI, TS>(cx, W.dec_conv[l], &prev1[0][0], &cur[0][0], off, DEC_LDA, sm.red,
                                   [&](int row, int n, float v) { float y = tanh_r(v); sm.seg[0][row][n] = y; cur[row][off + n] = quant8(y); });
      bar_arrive(BAR_SEG_FULL + 0, NCT);
      i_sync<NI>();
      skip_chunks(cx, f32_chunks(DEC_CONV, DEC_OUTP));
      off += DEC_CONV;
    }
  }

  const int last = (T + 1) & 1;      // buffer that held the final step's concat
  for (int r = 0; r < TS; r++) {
    if (s0 + r >= S || (active && !active[s0 + r])) continue;
    DecStreamState *st = state + (s0 + r);
    for (int i = tid; i < 5 * DEC_GRU; i += NIT) st->h[i] = sm.hs[r][i];
    for (int i = tid; i < DEC_LDA / 4; i += NIT)
      reinterpret_cast<uint32_t *>(st->cat1)[i] = reinterpret_cast<const uint32_t *>(sm.cb[last][r])[i];
  }
}

}  // namespace

// ----------------------------------------------------------------- host launchers
// streams per CTA: a full 16-row MMA tile when that still fills the 148 SMs, otherwise 8 so twice as many SMs work
static int core_tile_streams(int S) {
  const char *e = getenv("RADE_B200_TILE_STREAMS");
  if (e && (atoi(e) == 8 || atoi(e) == 16)) return atoi(e);
  return ((S + 15) / 16 >= 148) ? 16 : 8;
}
// weight-ring depth: 8-stream tiles have the shared memory for a deeper ring
constexpr int NST16 = 4, NST8 = 5;
int core_codec_set_chunk_table(int which, const ChunkDesc *d, int n) {
  if (n > CORE_MAX_CHUNKS) { fprintf(stderr, "libradae_b200: weight stream has %d chunks (max %d)\n", n, CORE_MAX_CHUNKS); return -1; }
  CUDA_CHECK(cudaMemcpyToSymbol(c_chunks, d, sizeof(ChunkDesc) * n, sizeof(ChunkDesc) * CORE_MAX_CHUNKS * which));
  return 0;
}
// per-device kernel attributes (opt-in to > 48 KB dynamic shared memory); called by rade_b200_open on its device
int core_codec_init_device() {
  CUDA_CHECK(cudaFuncSetAttribute(core_encoder_kernel<16, NST16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem<16, NST16>)));
  CUDA_CHECK(cudaFuncSetAttribute(core_encoder_kernel<8, NST8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem<8, NST8>)));
  CUDA_CHECK(cudaFuncSetAttribute(core_decoder_kernel<16, NST16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem<16, NST16>)));
  CUDA_CHECK(cudaFuncSetAttribute(core_decoder_kernel<8, NST8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem<8, NST8>)));
  return core_codec_umma_init_device();
}

// kernel family: tcgen05 by default; RADE_B200_CODEC=mma selects the mma.sync kernels of this file (kept for A/B measurements)
int core_codec_use_umma() {
  static const int on = !(getenv("RADE_B200_CODEC") && strcmp(getenv("RADE_B200_CODEC"), "mma") == 0);
  return on;
}

int core_encoder_launch(const CoreWeightsDev &W, EncStreamState *state, const float *in, int in_mode, float *z,
                        const uint8_t *active, int S, int T, cudaStream_t stream) {
  if (core_codec_use_umma()) return core_encoder_umma_launch(W, state, in, in_mode, z, active, S, T, stream);
  const int ts = core_tile_streams(S);
  const int grid = (S + ts - 1) / ts;
  if (ts == 16)
    core_encoder_kernel<16, NST16><<<grid, (ENC_NI + ENC_NF + 1) * 32, sizeof(EncSmem<16, NST16>), stream>>>(W, state, in, in_mode, z, active, S, T);
  else
    core_encoder_kernel<8, NST8><<<grid, (ENC_NI + ENC_NF + 1) * 32, sizeof(EncSmem<8, NST8>), stream>>>(W, state, in, in_mode, z, active, S, T);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}

int core_decoder_launch(const CoreWeightsDev &W, DecStreamState *state, const float *z, float *out, int out_mode,
                        int *uw_count, const uint8_t *active, int S, int T, cudaStream_t stream) {
  if (core_codec_use_umma()) return core_decoder_umma_launch(W, state, z, out, out_mode, uw_count, active, S, T, stream);
  const int ts = core_tile_streams(S);
  const int grid = (S + ts - 1) / ts;
  if (ts == 16)
    core_decoder_kernel<16, NST16><<<grid, (DEC_NI + DEC_NF + 1) * 32, sizeof(DecSmem<16, NST16>), stream>>>(W, state, z, out, out_mode, uw_count, active, S, T);
  else
    core_decoder_kernel<8, NST8><<<grid, (DEC_NI + DEC_NF + 1) * 32, sizeof(DecSmem<8, NST8>), stream>>>(W, state, z, out, out_mode, uw_count, active, S, T);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}
