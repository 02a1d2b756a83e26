// RADE core encoder / decoder on the 5th-generation tensor cores (tcgen05, sm_100a) — the default codec path of libradae_b200.
//
// Replaces (per stream, per 40 ms step):
//   rade_core_encoder  /root/reference/src/rade_enc.c:55-114   (PyTorch twin radae/radae_base.py:260-286)
//   rade_core_decoder  /root/reference/src/rade_dec.c:50-102   (PyTorch twin radae/radae_base.py:400-416)
// and the opus DNN primitives they call (compute_generic_dense/gru/conv1d[_dilation], compute_glu).
// Same inputs, outputs, per-stream state and float arithmetic as the mma.sync kernels in core_codec.cu: bit-identical results.
//
// Formulation: every int8 layer is D[out feature][stream] = W[out][K] x X[stream][K]^T as tcgen05.mma kind::i8 with
// A = a 128-row slice of the weight matrix (M = 128 TMEM lanes), B = the quantised activations of the tile's NS streams (N = NS),
// int32 accumulators in TMEM.  Both operands K-major, no swizzle (8-row x 16-byte core matrices); the host pre-bakes the weights
// in that order (weights.cpp) and the DenseNet concat buffers ARE the B operand:
//     offset(stream n, feature k) = (n / 8) * SBO + (k / 16) * 128 + (n % 8) * 16 + k % 16.
// One CTA owns a tile of <= NS streams for all layers and all steps of the launch; tiles are uneven so that every SM gets one.
// Warp roles (all hand-overs are mbarriers; nobody but the float warps touches a float weight, nobody but the issuer an int8 one):
//   E  8 epilogue warps: warp w reads TMEM lanes 32 (w % 4) .. +31 (= output features), column half w / 4 (= streams); float
//      epilogue in the oracle's rounding order straight from tcgen05.ld registers, appends the quantised outputs to the concat
//      buffer (byte stores + fence.proxy.async) and the float outputs to a 4-deep segment ring for the float warps;
//   I  issuer: ONE elected thread runs the per-step MMA program — fixed at compile time (umma_program.h) and unrolled into
//      straight-line tcgen05.mma with immediate descriptor offsets — waits for a weight stage, issues its MMAs, releases the
//      stage with tcgen05.commit; only the last 1-3 k-blocks of a layer depend on the previous layer's
//      output, so all other k-blocks are issued ahead while the epilogue of the previous layer is still running;
//   F  float warps: dense1 (one step ahead, written straight into the next step's concat buffer) and the wide zdense / output
//      layer, accumulated segment by segment in concat order (= the reference's sequential summation order) from their own ring;
//   P  two producer threads: cp.async.bulk (TMA) weight streams -> the int8 ring (3 x 40 KB) and the float ring (2 x 22 KB); one
//      bulk copy per stage: the TMA unit of an SM retires about one bulk copy per 440 cycles whatever its size (measured,
//      tools/microbench/bulk_copy_rate.cu), so the streams are packed into few, large copies.
// Warp ids are assigned by criticality: the hardware arbiter prefers the highest warp id of a sub-partition (measured on B200: an
// issuer below the float warps needed ~470 cycles per MMA issue), so F = warps 0.., E above them, producers and the issuer on top.
// Encoder GRUs (64 units) use TWO M = 128 tiles — [z; r] and [n; -] — instead of one per gate: lanes 64..127 compute r and pass it
// through shared memory to lanes 0..63 (z, n, h); decoder GRUs (96 units) use three overlapping tiles (rows 0, 96, 192).
#include "rade_common.h"
#include "rade_host.h"
#include "tma.cuh"
#include "codec_math.cuh"
#include "umma_program.h"

namespace {

// debug timeline (rade_b200_debug_trace_*): CTA 0 stamps clock64() into a global buffer; slots: issuer (t*128+rec)*2+{dep wait begins, dep satisfied}, accumulator commits at 1024+t*16+layer,
// epilogue warp 0 at 2048+(t*16+layer)*2+{acc ready, done}, float warp 0 at 4096+(t*16+seg)*2, int8 producer at 6144+t*64+chunk
#define TR(slot) do { if (trace && blockIdx.x == 0) trace[(slot)] = clock64(); } while (0)

constexpr int NE = 8;                        // epilogue warps
constexpr int SEG_LD = 100;                  // float row strides (16-byte aligned, bank-spreading)
constexpr int FIN_LD = 92;
constexpr int ZIN_LD = 84;
constexpr int F32_NST = 2;                   // stages of the float weight ring
// Per tile width NS (streams per CTA): stages of the int8 weight ring and depth of the float segment ring E -> F.  The 16-stream
// tile has twice the activations, state and segments in shared memory and pays for them with shallower rings.
template <int NS> struct RingCfg { static constexpr int I8_NST = NS > 8 ? 2 : 3, NSEG = NS > 8 ? 2 : 4, RSF = NS > 8 ? 2 : 1; };

// named barriers (0 = __syncthreads)
enum { NB_F = 1, NB_MAIN = 2, NB_E = 3, NB_Z = 4, NB_START = 5, NB_R0 = 8 /* .. 11: z/r warp pairs */ };
__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// (shared-memory matrix descriptors — K-major, no swizzle, LBO = 128 B — are assembled by the issuer: desc_lo / desc_hi below)
template <int N> __device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  // instruction descriptor: D = s32, A = B = s8, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// NC consecutive accumulator columns of this thread's TMEM lane (32x32b shape: lane i of the warp <-> TMEM lane base + i)
template <int NC> __device__ __forceinline__ void tmem_ld(uint32_t taddr, int (&v)[NC]) {
  static_assert(NC == 4 || NC == 8, "column count");
  if constexpr (NC == 4)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
  else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
// exactly one lane of the (converged) warp gets `true`; ptxas knows a region guarded by elect.sync has a single active thread and
// issues the tcgen05 instructions in it directly instead of through a per-lane waterfall loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// B-operand (concat buffer) addressing: KB = K extent of the buffer in bytes per stream
template <int KB> __device__ __forceinline__ int b_off(int n, int k) { return (n >> 3) * (KB * 8) + (k >> 4) * 128 + (n & 7) * 16 + (k & 15); }

// ---------------------------------------------------------------- rings
template <int NST, int STAGE> struct RingSmem {
  alignas(128) unsigned char buf[NST][STAGE];
  alignas(8) uint64_t full[NST];
  alignas(8) uint64_t empty[NST];
};
template <int NST, int STAGE> struct RingCursor {
  RingSmem<NST, STAGE> *r; int stage; uint32_t phase;
  __device__ __forceinline__ unsigned char *acquire() { mbar_wait(&r->full[stage], phase); return r->buf[stage]; }
  __device__ __forceinline__ void advance() { if (++stage == NST) { stage = 0; phase ^= 1; } }
  __device__ __forceinline__ void release_warp() {           // consumer side of the float ring: one arrival per warp
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&r->empty[stage]);
    advance();
  }
};
// producer thread: chunks[0, n_pro) once, then chunks[n_pro, n) once per step
template <int NST, int STAGE>
__device__ void produce(RingSmem<NST, STAGE> *r, const unsigned char *stream, const ChunkDesc *chunks, int n_pro, int n, int T, long long *trace) {
  int stage = 0; uint32_t phase = 0;
  for (int t = 0; t < T; t++)
    for (int c = (t == 0 ? 0 : n_pro); c < n; c++) {
      const ChunkDesc d = chunks[c];
      mbar_wait(&r->empty[stage], phase ^ 1);
      TR(6144 + t * 64 + c);
      mbar_expect_tx(&r->full[stage], d.bytes);
      bulk_g2s(r->buf[stage], stream + d.offset, d.bytes, &r->full[stage]);
      if (++stage == NST) { stage = 0; phase ^= 1; }
    }
}

// ---------------------------------------------------------------- float layers (F-warps)
// The float stream is a sequence of rows (dense1: 64 or 96 floats wide, zdense: 80, output: 96) packed into ring stages; a stage
// always holds a multiple of 4 rows of a segment.  Every float warp walks it with an identical cursor.
template <int NST, int STAGE> struct FloatCursor {
  RingSmem<NST, STAGE> *r; const ChunkDesc *chunks; int n_pro, n;
  int stage; uint32_t phase; int ci; const unsigned char *p; int left;
  __device__ __forceinline__ void next_stage() {
    mbar_wait(&r->full[stage], phase);
    p = r->buf[stage]; left = (int)chunks[ci].bytes;
  }
  __device__ __forceinline__ void consumed(int bytes) {
    p += bytes; left -= bytes;
    if (left == 0) {                             // stage drained: hand it back (one arrival per warp)
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&r->empty[stage]);
      if (++stage == NST) { stage = 0; phase ^= 1; }
      if (++ci == n) ci = n_pro;
    }
  }
};
// Float thread = (4 consecutive outputs, RS streams): acc[r][i] += sum_j W[j][4 grp + i] x_r[j] over the next K rows (NOUTP floats
// each) of the float stream, sequentially in j.  One LDS.128 of weights serves RS streams: the float layers are bound by the
// shared-memory port (every weight used to be delivered once per stream), not by arithmetic.
// FMA = false: product and sum separately rounded (the generic C sgemv of the reference, bit for bit).  FMA = true: one fused
// multiply-add per MAC (what the reference computes when opus is built with its AVX2 / NEON sgemv): 1/3 of the FP32-pipe time.
__device__ __forceinline__ float2 fma2_rn(float wx, float wy, float x, float ax, float ay) {
  float2 w = make_float2(wx, wy), xx = make_float2(x, x), a = make_float2(ax, ay);
  unsigned long long rw = *reinterpret_cast<unsigned long long *>(&w), rx = *reinterpret_cast<unsigned long long *>(&xx),
                     ra = *reinterpret_cast<unsigned long long *>(&a), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rw), "l"(rx), "l"(ra));
  return *reinterpret_cast<float2 *>(&rd);
}
template <int RS, bool FMA> __device__ __forceinline__ void mac4(float (&acc)[RS][4], const float4 w, const float (&x)[RS], float one) {
#pragma unroll
  for (int r = 0; r < RS; r++) {
    if constexpr (FMA) {
      const float2 a0 = fma2_rn(w.x, w.y, x[r], acc[r][0], acc[r][1]), a1 = fma2_rn(w.z, w.w, x[r], acc[r][2], acc[r][3]);
      acc[r][0] = a0.x; acc[r][1] = a0.y; acc[r][2] = a1.x; acc[r][3] = a1.y;
    } else {
      // acc + round(w x): the rounded product goes through a packed fused multiply-add as p * one + acc — exact and packed.
      // `one` is 1.0f from a kernel parameter, so ptxas cannot contract it with the multiply (it fuses mul.rn.f32x2 + add.rn.f32x2
      // and even mul + fma(p, 1.0f, acc) into one FFMA2, which rounds once instead of twice)
      const float2 p0 = prod2_rn(w.x, w.y, x[r]), p1 = prod2_rn(w.z, w.w, x[r]);
      const float2 a0 = fma2_rn(p0.x, p0.y, one, acc[r][0], acc[r][1]), a1 = fma2_rn(p1.x, p1.y, one, acc[r][2], acc[r][3]);
      acc[r][0] = a0.x; acc[r][1] = a0.y; acc[r][2] = a1.x; acc[r][3] = a1.y;
    }
  }
}
template <int NOUTP, int RS, bool FMA = false, typename CX>
__device__ __forceinline__ void dense_seg(CX &cx, float (&acc)[RS][4], const float *x0, int ldx, int K, int grp, bool act, float one) {
  for (int r0 = 0; r0 < K;) {
    if (cx.left == 0) cx.next_stage();
    const int n = min(K - r0, cx.left / (NOUTP * 4));
    if (act) {
      const float4 *W4 = reinterpret_cast<const float4 *>(cx.p) + grp;
#pragma unroll 2
      for (int j = 0; j < n; j += 4) {
        float4 xv[RS];
#pragma unroll
        for (int r = 0; r < RS; r++) xv[r] = *reinterpret_cast<const float4 *>(x0 + r * ldx + r0 + j);
        float x[RS];
#pragma unroll
        for (int r = 0; r < RS; r++) x[r] = xv[r].x;
        mac4<RS, FMA>(acc, W4[(j + 0) * (NOUTP / 4)], x, one);
#pragma unroll
        for (int r = 0; r < RS; r++) x[r] = xv[r].y;
        mac4<RS, FMA>(acc, W4[(j + 1) * (NOUTP / 4)], x, one);
#pragma unroll
        for (int r = 0; r < RS; r++) x[r] = xv[r].z;
        mac4<RS, FMA>(acc, W4[(j + 2) * (NOUTP / 4)], x, one);
#pragma unroll
        for (int r = 0; r < RS; r++) x[r] = xv[r].w;
        mac4<RS, FMA>(acc, W4[(j + 3) * (NOUTP / 4)], x, one);
      }
    }
    cx.consumed(n * NOUTP * 4);
    r0 += n;
  }
}
template <int NOUTP, typename CX> __device__ __forceinline__ void skip_seg(CX &cx, int K) {
  for (int r0 = 0; r0 < K;) {
    if (cx.left == 0) cx.next_stage();
    const int n = min(K - r0, cx.left / (NOUTP * 4));
    cx.consumed(n * NOUTP * 4);
    r0 += n;
  }
}

// ---------------------------------------------------------------- issuer: the compile-time MMA program as straight-line code
// ONE elected thread issues everything.  The record list (umma_program.h) is unrolled by template recursion: per record a block
// of nk x n_tiles tcgen05.mma whose descriptors are `base register + immediate`, plus the waits and commits the record carries.
// History of this loop, measured on the B200 (tools/microbench/umma_issue_loop.cu, tools/codec_trace.py): descriptors computed
// inside `if (lane == 0)` -> R2UR waterfall code, 380 cycles per MMA; warp-uniform loop with `if (leader)` per MMA -> 100;
// elect.sync + unrolled chunks interpreted from a run-time record list -> 30 per MMA but ~480 per record; the tensor core itself
// needs 47-52 per MMA (M = 128, N = 8, kind::i8: bound by the 4 KB operand fetch, not by the 16 cycles of math).
struct IssuerBufs { uint32_t lo[5]; uint32_t hi_cat, hi_hq; };     // low descriptor words (address + LBO) of the five B buffers
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFF) >> 4) | ((128u >> 4) << 16); }       // start address, LBO = 128
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14); }        // SBO, descriptor version 1
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return (uint64_t)hi << 32 | lo; }
template <typename I8Ring> struct IssueCtx {
  I8Ring cx; IssuerBufs B; uint32_t tmem, stage_lo, par; uint64_t *act_ready, *acc_full; long long *trace; int t;
};
template <int NS, bool ENC, int I, typename CTX>
__device__ __forceinline__ void issue_rec(CTX &c) {
  constexpr UmmaRecC r = ENC ? kUmmaEncProg.r[I] : kUmmaDecProg.r[I];
  long long *const trace = c.trace;
  if constexpr ((r.flags & UR_STAGE_FIRST) != 0) { c.stage_lo = (smem_u32(c.cx.acquire()) & 0x3FFFF) >> 4; tc_fence_after(); }
  if constexpr (r.dep >= 0) {                                      // this image consumes the outputs of the layer before
    TR((c.t * 128 + I) * 2);
    mbar_wait(&c.act_ready[r.dep], c.par); tc_fence_after();
    TR((c.t * 128 + I) * 2 + 1);
  }
  constexpr uint32_t asbo = (uint32_t)r.nk * 256u;                 // 8-row group stride of the image: kbytes * 8
  constexpr uint32_t a_hi = desc_hi(asbo);
  constexpr uint32_t tile_off = ((uint32_t)(r.tile_step >> 3) * asbo) >> 4;      // descriptor units (16 B) between tiles
  const uint32_t a_lo = c.stage_lo + ((uint32_t)r.a_off16 | ((128u >> 4) << 16));
  const uint32_t b_lo = c.B.lo[r.b_buf] + (uint32_t)r.b_kb * 16u;
  const uint32_t b_hi = r.b_buf >= UB_HQ_RD ? c.B.hi_hq : c.B.hi_cat;
#pragma unroll
  for (int k = 0; k < r.nk; k++) {
#pragma unroll
    for (int g = 0; g < r.n_tiles; g++)
      umma_i8<NS>(c.tmem + (uint32_t)(r.d_blk + g * r.d_tile_stride) * NS, desc64(a_lo + k * 16 + g * tile_off, a_hi),
                  desc64(b_lo + k * 16, b_hi), (k == 0 && (r.flags & UR_ZERO_FIRST)) ? 0u : 1u);
  }
  if constexpr ((r.flags & UR_STAGE_LAST) != 0) { umma_commit(&c.cx.r->empty[c.cx.stage]); c.cx.advance(); }   // stage free once these MMAs have read it
  if constexpr (r.commit >= 0) { umma_commit(&c.acc_full[r.commit]); TR(1024 + c.t * 16 + r.commit); }
}
template <int NS, bool ENC, int I, int N> struct IssueAll {
  template <typename CTX> static __device__ __forceinline__ void run(CTX &c) { issue_rec<NS, ENC, I>(c); IssueAll<NS, ENC, I + 1, N>::run(c); }
};
template <int NS, bool ENC, int N> struct IssueAll<NS, ENC, N, N> {
  template <typename CTX> static __device__ __forceinline__ void run(CTX &) {}
};

// ================================================================= encoder
template <int NS> struct EncCfg {
  static constexpr int NC = NS / 2;                                // accumulator columns per epilogue thread
  static constexpr int RSF = RingCfg<NS>::RSF, I8_NST = RingCfg<NS>::I8_NST, NSEG = RingCfg<NS>::NSEG;   // RSF: streams per float thread (x 4 outputs)
  static constexpr int NFT = (NS / RSF) * (RADE_LATENT / 4);       // float threads: (stream pair, group of 4 outputs)
  static constexpr int NF = (NFT + 31) / 32;
  static constexpr int KB = ENC_CAT;                               // concat bytes per stream
  static constexpr int CB_BYTES = (NS / 8) * KB * 8;
  static constexpr int NCB = 4;                                    // concat buffers: t, t-1, t-2 and the one dense1(t+1) is written to
  static constexpr int THREADS = (NE + NF + 2) * 32;                // float + epilogue warps, one producer warp (two lanes), one issuer warp
  static constexpr int TMEM_COLS = (10 * NS <= 128) ? 128 : 256;   // GRU slots 2 x 4 blocks, conv slots 2 x 1 block
};
template <int NS> struct EncSmemU {
  RingSmem<RingCfg<NS>::I8_NST, UMMA_I8_STAGE_BYTES> i8;
  RingSmem<F32_NST, UMMA_F32_STAGE_BYTES> f32;
  alignas(128) uint8_t cb[EncCfg<NS>::NCB][EncCfg<NS>::CB_BYTES];
  alignas(16) float hs[NS][5 * ENC_GRU];
  alignas(16) float seg[RingCfg<NS>::NSEG][NS][SEG_LD];
  alignas(16) float d1f[NS][SEG_LD];                               // dense1 output (float) = concat segment 0
  alignas(16) float fin[NS][FIN_LD];
  alignas(16) float rx[NS][ENC_GRU];                               // reset gates handed from the r lanes to the z / n lanes
  alignas(8) uint64_t acc_full[10], act_ready[10], seg_full[RingCfg<NS>::NSEG], seg_empty[RingCfg<NS>::NSEG], d1_ready[2];   // d1_ready ping-pongs by step: the float warps run one step ahead
  ChunkDesc i8_chunks[UMMA_MAX_I8_CHUNKS], f32_chunks[UMMA_MAX_F32_CHUNKS];
  uint32_t tmem_base;
  int any_active;
};

template <int NS>
__global__ void __launch_bounds__(EncCfg<NS>::THREADS, 1)
core_encoder_umma_kernel(const __grid_constant__ CoreWeightsDev W, EncStreamState *__restrict__ state, const float *__restrict__ in, int in_mode,
                         float *__restrict__ z_out, const uint8_t *__restrict__ active, int S, int T) {
  typedef EncCfg<NS> C;
  constexpr int NSEG = C::NSEG, I8_NST = C::I8_NST;
  constexpr int NC = C::NC, RSF = C::RSF, NSP = NS / RSF, NF = C::NF, KB = C::KB, NCB = C::NCB;
  constexpr int N_MAIN = (NE + 1 + NF) * 32;                       // epilogue + issuer + float threads
  extern __shared__ __align__(128) unsigned char smem_raw[];
  EncSmemU<NS> &sm = *reinterpret_cast<EncSmemU<NS> *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // uneven tiles: CTA c owns streams [c S / G, (c + 1) S / G)
  const int s0 = (int)((long long)blockIdx.x * S / gridDim.x), s1 = (int)((long long)(blockIdx.x + 1) * S / gridDim.x);
  const int ns = s1 - s0;
  const UmmaCodecDev &U = W.enc_umma;

  if (tid == 0 && W.trace && blockIdx.x == 0) W.trace[8000] = clock64();
  if (tid == 0) sm.any_active = 0;
  __syncthreads();
  if (tid < ns && (!active || active[s0 + tid])) sm.any_active = 1;
  if (tid == 0) {
    for (int i = 0; i < I8_NST; i++) { mbar_init(&sm.i8.full[i], 1); mbar_init(&sm.i8.empty[i], 1); }
    for (int i = 0; i < F32_NST; i++) { mbar_init(&sm.f32.full[i], 1); mbar_init(&sm.f32.empty[i], NF); }
    for (int i = 0; i < 10; i++) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.act_ready[i], NE); }
    for (int i = 0; i < NSEG; i++) { mbar_init(&sm.seg_full[i], NE); mbar_init(&sm.seg_empty[i], NF); }
    mbar_init(&sm.d1_ready[0], 1); mbar_init(&sm.d1_ready[1], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < U.n_i8_chunks; i += blockDim.x) sm.i8_chunks[i] = U.i8_chunks[i];
  for (int i = tid; i < U.n_f32_chunks; i += blockDim.x) sm.f32_chunks[i] = U.f32_chunks[i];
  __syncthreads();
  if (!sm.any_active) return;
  if (warp == NF) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&sm.tmem_base)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  long long *const trace = W.trace;
  if (tid == 0) TR(8001);

  // ---------------------------------------------------------------- producers: lanes 0 and 1 of one warp, each walking its own ring
  // (independent thread scheduling keeps the two blocking loops apart; a thread block of 512 instead of 544 threads also lifts the
  // register cap from 96 to 128 per thread — the allocation unit is four warps — which removed the decoder's spills)
  if (warp == NF + NE) {
    if (lane == 0) produce(&sm.i8, U.i8_stream, sm.i8_chunks, 0, U.n_i8_chunks, T, W.trace);
    else if (lane == 1) produce(&sm.f32, U.f32_stream, sm.f32_chunks, U.n_f32_prologue, U.n_f32_chunks, T, nullptr);
    return;
  }

  // ---------------------------------------------------------------- float warps
  if (warp < NF) {
    const int ft = tid, sp = ft % NSP, grp = ft / NSP;    // streams RSF sp .. +RSF-1 of the tile, outputs 4 grp .. +3
    const int sl = sp * RSF;
    const bool f_on = grp < RADE_LATENT / 4;              // the last warp is only partly populated
    FloatCursor<F32_NST, UMMA_F32_STAGE_BYTES> cx{&sm.f32, sm.f32_chunks, U.n_f32_prologue, U.n_f32_chunks, 0, 0u, 0, nullptr, 0};
    auto stage_input = [&](int t) {             // API layout [S][4T][36]: 20 used features + aux = -1 (src/rade_api.c:426-432)
      for (int i = ft; i < NS * ENC_IN; i += NF * 32) {
        const int r = i / ENC_IN, k = i % ENC_IN;
        float v = 0.f;
        if (r < ns) {
          if (in_mode == 0) v = in[((size_t)(s0 + r) * T + t) * ENC_IN + k];
          else {
            const int fr = k / 21, f = k % 21;
            v = (f == 20) ? -1.f : in[((size_t)(s0 + r) * 4 * T + 4 * t + fr) * RADE_NB_TOTAL_FEATURES + f];
          }
        }
        sm.fin[r][k] = v;
      }
    };
    // dense1 of step t: tanh(W f + b), 84 -> 64: float copy = concat segment 0 (d1f), int8 copy straight into step t's concat buffer
    auto dense1 = [&](int t) {
      float a[RSF][4];
#pragma unroll
      for (int r = 0; r < RSF; r++) for (int i = 0; i < 4; i++) a[r][i] = 0.f;
      const bool act = grp < 64 / 4;
      dense_seg<64, RSF>(cx, a, sm.fin[sl], FIN_LD, ENC_IN, grp, act, W.one);
      if (act) {
        uint8_t *cb = sm.cb[t % NCB];
#pragma unroll
        for (int r = 0; r < RSF; r++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int o = 4 * grp + i;
            const float y = tanh_r(__fadd_rn(a[r][i], W.enc_dense1.bias[o]));
            sm.d1f[sl + r][o] = y;
            cb[b_off<KB>(sl + r, o)] = (uint8_t)quant8(y);
          }
      }
      fence_async_smem();
      nb_sync(NB_F, NF * 32);
      if (ft == 0) mbar_arrive(&sm.d1_ready[t & 1]);
    };
    nb_sync(NB_Z, NE * 32 + NF * 32);            // concat buffers zeroed; the state load of the epilogue warps runs beside dense1(0)
    stage_input(0);
    nb_sync(NB_F, NF * 32);
    dense1(0);
    int nseg = 0;
    for (int t = 0; t < T; t++) {
      float zacc[RSF][4];
#pragma unroll
      for (int r = 0; r < RSF; r++) for (int i = 0; i < 4; i++) zacc[r][i] = 0.f;
      if (W.float_fma) dense_seg<RADE_LATENT, RSF, true>(cx, zacc, sm.d1f[sl], SEG_LD, 64, grp, f_on, W.one);
      else dense_seg<RADE_LATENT, RSF>(cx, zacc, sm.d1f[sl], SEG_LD, 64, grp, f_on, W.one);
      nb_sync(NB_F, NF * 32);                    // everybody is done with d1f and fin of this step
      if (t + 1 < T) { stage_input(t + 1); nb_sync(NB_F, NF * 32); dense1(t + 1); }
      else skip_seg<64>(cx, ENC_IN);
#pragma unroll 1
      for (int j = 0; j < 10; j++, nseg++) {     // GRU 1, conv 1, GRU 2, ... in concat order
        const int slot = nseg % NSEG;
        mbar_wait(&sm.seg_full[slot], (nseg / NSEG) & 1);
        if (ft == 0) TR(4096 + (t * 16 + j) * 2);
        if (W.float_fma) dense_seg<RADE_LATENT, RSF, true>(cx, zacc, sm.seg[slot][sl], SEG_LD, (j & 1) ? ENC_CONV : ENC_GRU, grp, f_on, W.one);
        else dense_seg<RADE_LATENT, RSF>(cx, zacc, sm.seg[slot][sl], SEG_LD, (j & 1) ? ENC_CONV : ENC_GRU, grp, f_on, W.one);
        __syncwarp();
        if (ft == 0) TR(4096 + (t * 16 + j) * 2 + 1);
        if (lane == 0) mbar_arrive(&sm.seg_empty[slot]);
      }
      if (f_on) {                                // z = zdense(cat) + b (linear for bottleneck 3, tanh for 1: src/rade_enc.c:107-113)
#pragma unroll
        for (int r = 0; r < RSF; r++) {
          const int sg = s0 + sl + r;
          if (sl + r >= ns || (active && !active[sg])) continue;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float v = __fadd_rn(zacc[r][i], W.enc_zdense.bias[4 * grp + i]);
            z_out[((size_t)sg * T + t) * RADE_LATENT + 4 * grp + i] = W.enc_z_tanh ? tanh_r(v) : v;
          }
        }
      }
    }
    nb_sync(NB_MAIN, N_MAIN);
    return;
  }

  // ---------------------------------------------------------------- issuer warp
  if (warp == NF + NE + 1) {
    nb_sync(NB_START, (NE + 1) * 32);            // state loaded by the epilogue warps
    if (elect_one()) {                           // ONE thread runs the whole issue program (no reconvergence points inside)
      typedef RingCursor<I8_NST, UMMA_I8_STAGE_BYTES> Ring;
      IssueCtx<Ring> c;
      c.cx = Ring{&sm.i8, 0, 0u}; c.tmem = tmem; c.stage_lo = 0; c.act_ready = sm.act_ready; c.acc_full = sm.acc_full; c.trace = trace;
      c.B.lo[3] = c.B.lo[4] = 0; c.B.hi_cat = c.B.hi_hq = desc_hi(KB * 8);
      for (int t = 0; t < T; t++) {
        c.B.lo[0] = desc_lo(smem_u32(sm.cb[t % NCB])); c.B.lo[1] = desc_lo(smem_u32(sm.cb[(t + NCB - 1) % NCB]));
        c.B.lo[2] = desc_lo(smem_u32(sm.cb[(t + NCB - 2) % NCB]));
        c.par = t & 1; c.t = t;
        // accumulator columns are reused by layer parity; within a step every overwrite is ordered behind the epilogue that last
        // read the columns by the layer dependencies, EXCEPT the first conv of a step (no dependency of its own) against the last
        // conv epilogue of the step before: wait for that epilogue here (found as run-to-run differences in frame 1 of the CLI test)
        if (t > 0) { mbar_wait(&sm.act_ready[9], (t - 1) & 1); tc_fence_after(); }
        mbar_wait(&sm.d1_ready[t & 1], (t >> 1) & 1);        // dense1 output (int8) is in cur, features [0, 64)
        tc_fence_after();
        IssueAll<NS, true, 0, kUmmaEncProg.n>::run(c);
      }
    }
    __syncwarp();
    nb_sync(NB_MAIN, N_MAIN);
    return;
  }

  // ---------------------------------------------------------------- epilogue warps
  const int q = warp & 3, ch = (warp - NF) >> 2; // TMEM lane quadrant (fixed by the warp id), column (stream) half
  const int et = tid - NF * 32;                  // thread index among the epilogue warps
  const int c0 = ch * NC;                        // first stream of this thread
  const uint32_t tlane = (uint32_t)(32 * q) << 16;
  constexpr int NET = NE * 32;
  // concat buffers: zero, then steps t-1 (cb[NCB-1]) and t-2 (cb[NCB-2]) from the per-stream row-major state
  for (int i = et; i < NCB * C::CB_BYTES / 4; i += NET) reinterpret_cast<uint32_t *>(sm.cb[0])[i] = 0u;
  nb_sync(NB_E, NET);
  nb_arrive(NB_Z, NE * 32 + NF * 32);            // the float warps may start dense1(0)
  // all streams' state in ONE pass: the loads of a thread (one per stream and array slice) are issued back to back, so the tile
  // pays one global-memory latency instead of one per stream (the per-stream loop cost ~8 k cycles before the first MMA)
  {
    constexpr int HN = 5 * ENC_GRU, CN = KB / 4;
    float hv[(NS * HN + NET - 1) / NET];
#pragma unroll
    for (int q = 0; q < (NS * HN + NET - 1) / NET; q++) {
      const int idx = et + q * NET, r = idx / HN, i = idx - r * HN;
      hv[q] = (idx < NS * HN && r < ns) ? state[s0 + r].h[i] : 0.f;
    }
    uint32_t c1[(NS * CN + NET - 1) / NET], c2[(NS * CN + NET - 1) / NET];
#pragma unroll
    for (int q = 0; q < (NS * CN + NET - 1) / NET; q++) {
      const int idx = et + q * NET, r = idx / CN, i = idx - r * CN;
      const bool ok = idx < NS * CN && r < ns;
      c1[q] = ok ? reinterpret_cast<const uint32_t *>(state[s0 + r].cat1)[i] : 0u;
      c2[q] = ok ? reinterpret_cast<const uint32_t *>(state[s0 + r].cat2)[i] : 0u;
    }
#pragma unroll
    for (int q = 0; q < (NS * HN + NET - 1) / NET; q++) {
      const int idx = et + q * NET, r = idx / HN, i = idx - r * HN;
      if (idx < NS * HN) sm.hs[r][i] = hv[q];
    }
#pragma unroll
    for (int q = 0; q < (NS * CN + NET - 1) / NET; q++) {
      const int idx = et + q * NET, r = idx / CN, i = idx - r * CN;
      if (idx < NS * CN && r < ns) {
        *reinterpret_cast<uint32_t *>(sm.cb[NCB - 1] + b_off<KB>(r, 4 * i)) = c1[q];
        *reinterpret_cast<uint32_t *>(sm.cb[NCB - 2] + b_off<KB>(r, 4 * i)) = c2[q];
      }
    }
  }
  fence_async_smem();
  nb_sync(NB_START, (NE + 1) * 32);               // with the issuer: state and concat buffers are in place

  int nseg = 0;
  for (int t = 0; t < T; t++) {
    uint8_t *cur = sm.cb[t % NCB];
    const uint32_t par = t & 1;
    int off = 64;
#pragma unroll 1
    for (int l = 0; l < 5; l++) {
      // ---- GRU l.  Accumulator blocks (x NS columns) at gslot: [z; r] input, [z; r] recurrent, [n; -] input, [n; -] recurrent
      const uint32_t gcol = tmem + (uint32_t)((l & 1) * 4) * NS + c0;
      const int u = (q & 1) * 32 + lane;         // hidden unit
      const int gate = (q < 2) ? 0 : 1;          // lanes 0..63 hold z (and n in the second tile), lanes 64..127 hold r
      const I8LayerDev &Li = W.enc_gru_in[l], &Lr = W.enc_gru_rec[l];
      const float s_i = Li.scale[gate * ENC_GRU + u], b_i = Li.bias[gate * ENC_GRU + u];
      const float s_r = Lr.scale[gate * ENC_GRU + u], b_r = Lr.bias[gate * ENC_GRU + u];
      float sn_i = 0.f, bn_i = 0.f, sn_r = 0.f, bn_r = 0.f;
      if (q < 2) { sn_i = Li.scale[2 * ENC_GRU + u]; bn_i = Li.bias[2 * ENC_GRU + u]; sn_r = Lr.scale[2 * ENC_GRU + u]; bn_r = Lr.bias[2 * ENC_GRU + u]; }
      {
        const int slot = nseg % NSEG;
        mbar_wait(&sm.seg_empty[slot], ((nseg / NSEG) & 1) ^ 1);
        mbar_wait(&sm.acc_full[2 * l], par);
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 2 * l) * 2);
        tc_fence_after();
        int a_in[NC], a_rec[NC];
        tmem_ld<NC>(gcol + tlane, a_in); tmem_ld<NC>(gcol + tlane + NS, a_rec);
        if (q >= 2) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NC; j++)
            sm.rx[c0 + j][u] = sigmoid_r(__fadd_rn(lin(a_in[j], s_i, b_i), lin(a_rec[j], s_r, b_r)));
          nb_arrive(NB_R0 + (q & 1) * 2 + ch, 64);
        } else {
          int n_in[NC], n_rec[NC];
          tmem_ld<NC>(gcol + tlane + 2 * NS, n_in); tmem_ld<NC>(gcol + tlane + 3 * NS, n_rec);
          tmem_ld_wait();
          float z[NC], na[NC], nb[NC];
#pragma unroll
          for (int j = 0; j < NC; j++) {
            z[j] = sigmoid_r(__fadd_rn(lin(a_in[j], s_i, b_i), lin(a_rec[j], s_r, b_r)));
            na[j] = lin(n_in[j], sn_i, bn_i); nb[j] = lin(n_rec[j], sn_r, bn_r);
          }
          nb_sync(NB_R0 + q * 2 + ch, 64);       // the matching r warp has published its gates
#pragma unroll
          for (int j = 0; j < NC; j++) {
            const int s = c0 + j;
            const float r = sm.rx[s][u];
            const float n = tanh_r(__fadd_rn(na[j], __fmul_rn(nb[j], r)));
            const float hold = sm.hs[s][l * ENC_GRU + u];
            const float h = __fadd_rn(__fmul_rn(z[j], hold), __fmul_rn(__fsub_rn(1.f, z[j]), n));
            sm.hs[s][l * ENC_GRU + u] = h;
            sm.seg[slot][s][u] = h;
            cur[b_off<KB>(s, off + u)] = (uint8_t)quant8(h);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 2 * l) * 2 + 1);
        if (lane == 0) { mbar_arrive(&sm.act_ready[2 * l]); mbar_arrive(&sm.seg_full[slot]); }
        nseg++;
      }
      off += ENC_GRU;
      // ---- conv l: 96 outputs on lanes 0..95
      {
        const int o = q * 32 + lane;
        float es = 0.f, eb = 0.f;
        if (q < 3) { es = W.enc_conv[l].scale[o]; eb = W.enc_conv[l].bias[o]; }
        const int slot = nseg % NSEG;
        mbar_wait(&sm.seg_empty[slot], ((nseg / NSEG) & 1) ^ 1);
        mbar_wait(&sm.acc_full[2 * l + 1], par);
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 2 * l + 1) * 2);
        tc_fence_after();
        if (q < 3) {
          int acc[NC];
          tmem_ld<NC>(tmem + (uint32_t)(8 + (l & 1)) * NS + c0 + tlane, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NC; j++) {
            const int s = c0 + j;
            const float y = tanh_r(lin(acc[j], es, eb));
            sm.seg[slot][s][o] = y;
            cur[b_off<KB>(s, off + o)] = (uint8_t)quant8(y);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 2 * l + 1) * 2 + 1);
        if (lane == 0) { mbar_arrive(&sm.act_ready[2 * l + 1]); mbar_arrive(&sm.seg_full[slot]); }
        nseg++;
      }
      off += ENC_CONV;
    }
  }
  nb_sync(NB_MAIN, N_MAIN);                      // every MMA has been consumed, the float warps are done with the segments
  if (et == 0) TR(8002);
  if (warp == NF) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(C::TMEM_COLS) : "memory");
  const int last = (T + NCB - 1) % NCB, last2 = (T + NCB - 2) % NCB;
  for (int r = 0; r < ns; r++) {
    if (active && !active[s0 + r]) continue;
    EncStreamState *st = state + (s0 + r);
    for (int i = et; i < 5 * ENC_GRU; i += NET) st->h[i] = sm.hs[r][i];
    for (int i = et; i < KB / 4; i += NET) {
      reinterpret_cast<uint32_t *>(st->cat1)[i] = *reinterpret_cast<const uint32_t *>(sm.cb[last] + b_off<KB>(r, 4 * i));
      reinterpret_cast<uint32_t *>(st->cat2)[i] = *reinterpret_cast<const uint32_t *>(sm.cb[last2] + b_off<KB>(r, 4 * i));
    }
  }
  if (et == 0) TR(8003);
}

// ================================================================= decoder
// Three int8 products per DenseNet stage: GRU (input + recurrent), GLU gate on the new state, conv (two taps).  The quantised
// hidden states live in their own ping-pong buffers (hq), like the concat buffers in the B-operand layout.
template <int NS> struct DecCfg {
  static constexpr int NC = NS / 2;
  static constexpr int RSF = RingCfg<NS>::RSF, I8_NST = RingCfg<NS>::I8_NST, NSEG = RingCfg<NS>::NSEG;
  static constexpr int NFT = (NS / RSF) * (DEC_OUTP / 4);
  static constexpr int NF = (NFT + 31) / 32;
  static constexpr int KB = DEC_CAT, KH = 5 * DEC_GRU;
  static constexpr int CB_BYTES = (NS / 8) * KB * 8, HQ_BYTES = (NS / 8) * KH * 8;
  static constexpr int NCB = 3;                                    // t, t-1 and the one dense1(t+1) is written to
  static constexpr int THREADS = (NE + NF + 2) * 32;                // float + epilogue warps, one producer warp (two lanes), one issuer warp
  static constexpr int TMEM_COLS = (16 * NS <= 128) ? 128 : 256;   // GRU slots 2 x 6 blocks, GLU 2 x 1, conv 2 x 1
};
template <int NS> struct DecSmemU {
  RingSmem<RingCfg<NS>::I8_NST, UMMA_I8_STAGE_BYTES> i8;
  RingSmem<F32_NST, UMMA_F32_STAGE_BYTES> f32;
  alignas(128) uint8_t cb[DecCfg<NS>::NCB][DecCfg<NS>::CB_BYTES];
  alignas(128) uint8_t hq[2][DecCfg<NS>::HQ_BYTES];
  alignas(16) float hs[NS][5 * DEC_GRU];
  alignas(16) float seg[RingCfg<NS>::NSEG][NS][SEG_LD];
  alignas(16) float d1f[NS][SEG_LD];
  alignas(16) float zin[NS][ZIN_LD];
  alignas(8) uint64_t acc_full[15], act_ready[15], seg_full[RingCfg<NS>::NSEG], seg_empty[RingCfg<NS>::NSEG], d1_ready[2];   // d1_ready ping-pongs by step: the float warps run one step ahead
  ChunkDesc i8_chunks[UMMA_MAX_I8_CHUNKS], f32_chunks[UMMA_MAX_F32_CHUNKS];
  uint32_t tmem_base;
  int any_active;
};

// out_mode 0: features [S][T][84];  out_mode 1: API layout [S][4T][36] (20 used, rest zero, src/rade_api.c:488-500)
// uw_count (optional): += number of steps whose first aux symbol (feature 20) is > 0 (src/rade_api.c:502-505)
template <int NS>
__global__ void __launch_bounds__(DecCfg<NS>::THREADS, 1)
core_decoder_umma_kernel(const __grid_constant__ CoreWeightsDev W, DecStreamState *__restrict__ state, const float *__restrict__ z_in,
                         float *__restrict__ out, int out_mode, int *__restrict__ uw_count,
                         const uint8_t *__restrict__ active, int S, int T) {
  typedef DecCfg<NS> C;
  constexpr int NSEG = C::NSEG, I8_NST = C::I8_NST;
  constexpr int NC = C::NC, RSF = C::RSF, NSP = NS / RSF, NF = C::NF, KB = C::KB, KH = C::KH, NCB = C::NCB;
  constexpr int N_MAIN = (NE + 1 + NF) * 32, NGRP = DEC_OUTP / 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DecSmemU<NS> &sm = *reinterpret_cast<DecSmemU<NS> *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = (int)((long long)blockIdx.x * S / gridDim.x), s1 = (int)((long long)(blockIdx.x + 1) * S / gridDim.x);
  const int ns = s1 - s0;
  const UmmaCodecDev &U = W.dec_umma;

  if (tid == 0 && W.trace && blockIdx.x == 0) W.trace[8000] = clock64();
  if (tid == 0) sm.any_active = 0;
  __syncthreads();
  if (tid < ns && (!active || active[s0 + tid])) sm.any_active = 1;
  if (tid == 0) {
    for (int i = 0; i < I8_NST; i++) { mbar_init(&sm.i8.full[i], 1); mbar_init(&sm.i8.empty[i], 1); }
    for (int i = 0; i < F32_NST; i++) { mbar_init(&sm.f32.full[i], 1); mbar_init(&sm.f32.empty[i], NF); }
    for (int i = 0; i < 15; i++) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.act_ready[i], NE); }
    for (int i = 0; i < NSEG; i++) { mbar_init(&sm.seg_full[i], NE); mbar_init(&sm.seg_empty[i], NF); }
    mbar_init(&sm.d1_ready[0], 1); mbar_init(&sm.d1_ready[1], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < U.n_i8_chunks; i += blockDim.x) sm.i8_chunks[i] = U.i8_chunks[i];
  for (int i = tid; i < U.n_f32_chunks; i += blockDim.x) sm.f32_chunks[i] = U.f32_chunks[i];
  __syncthreads();
  if (!sm.any_active) return;
  if (warp == NF) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&sm.tmem_base)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  long long *const trace = W.trace;
  if (tid == 0) TR(8001);

  // ---------------------------------------------------------------- producers: lanes 0 and 1 of one warp, each walking its own ring
  // (independent thread scheduling keeps the two blocking loops apart; a thread block of 512 instead of 544 threads also lifts the
  // register cap from 96 to 128 per thread — the allocation unit is four warps — which removed the decoder's spills)
  if (warp == NF + NE) {
    if (lane == 0) produce(&sm.i8, U.i8_stream, sm.i8_chunks, 0, U.n_i8_chunks, T, W.trace);
    else if (lane == 1) produce(&sm.f32, U.f32_stream, sm.f32_chunks, U.n_f32_prologue, U.n_f32_chunks, T, nullptr);
    return;
  }

  if (warp < NF) {                               // ---- float warps: z staging, dense1, incremental output layer, feature output
    const int ft = tid, sp = ft % NSP, grp = ft / NSP;
    const int sl = sp * RSF;
    const bool f_on = grp < NGRP;
    FloatCursor<F32_NST, UMMA_F32_STAGE_BYTES> cx{&sm.f32, sm.f32_chunks, U.n_f32_prologue, U.n_f32_chunks, 0, 0u, 0, nullptr, 0};
    auto stage_input = [&](int t) {
      for (int i = ft; i < NS * DEC_IN; i += NF * 32) {
        const int r = i / DEC_IN, k = i % DEC_IN;
        sm.zin[r][k] = (r < ns) ? z_in[((size_t)(s0 + r) * T + t) * DEC_IN + k] : 0.f;
      }
    };
    auto dense1 = [&](int t) {                   // tanh(W z + b), 80 -> 96
      float a[RSF][4];
#pragma unroll
      for (int r = 0; r < RSF; r++) for (int i = 0; i < 4; i++) a[r][i] = 0.f;
      dense_seg<96, RSF>(cx, a, sm.zin[sl], ZIN_LD, DEC_IN, grp, f_on, W.one);
      if (f_on) {
        uint8_t *cb = sm.cb[t % NCB];
#pragma unroll
        for (int r = 0; r < RSF; r++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int o = 4 * grp + i;
            const float y = tanh_r(__fadd_rn(a[r][i], W.dec_dense1.bias[o]));
            sm.d1f[sl + r][o] = y;
            cb[b_off<KB>(sl + r, o)] = (uint8_t)quant8(y);
          }
      }
      fence_async_smem();
      nb_sync(NB_F, NF * 32);
      if (ft == 0) mbar_arrive(&sm.d1_ready[t & 1]);
    };
    nb_sync(NB_Z, NE * 32 + NF * 32);            // concat buffers zeroed (see the encoder)
    stage_input(0);
    nb_sync(NB_F, NF * 32);
    dense1(0);
    int nseg = 0;
    for (int t = 0; t < T; t++) {
      float oacc[RSF][4];
#pragma unroll
      for (int r = 0; r < RSF; r++) for (int i = 0; i < 4; i++) oacc[r][i] = 0.f;
      if (W.float_fma) dense_seg<DEC_OUTP, RSF, true>(cx, oacc, sm.d1f[sl], SEG_LD, 96, grp, f_on, W.one);
      else dense_seg<DEC_OUTP, RSF>(cx, oacc, sm.d1f[sl], SEG_LD, 96, grp, f_on, W.one);
      nb_sync(NB_F, NF * 32);
      if (t + 1 < T) { stage_input(t + 1); nb_sync(NB_F, NF * 32); dense1(t + 1); }
      else skip_seg<96>(cx, DEC_IN);
#pragma unroll 1
      for (int j = 0; j < 10; j++, nseg++) {     // GLU 1, conv 1, GLU 2, ...
        const int slot = nseg % NSEG;
        mbar_wait(&sm.seg_full[slot], (nseg / NSEG) & 1);
        if (ft == 0) TR(4096 + (t * 16 + j) * 2);
        if (W.float_fma) dense_seg<DEC_OUTP, RSF, true>(cx, oacc, sm.seg[slot][sl], SEG_LD, (j & 1) ? DEC_CONV : DEC_GRU, grp, f_on, W.one);
        else dense_seg<DEC_OUTP, RSF>(cx, oacc, sm.seg[slot][sl], SEG_LD, (j & 1) ? DEC_CONV : DEC_GRU, grp, f_on, W.one);
        __syncwarp();
        if (ft == 0) TR(4096 + (t * 16 + j) * 2 + 1);
        if (lane == 0) mbar_arrive(&sm.seg_empty[slot]);
      }
      if (f_on) {
#pragma unroll
        for (int r = 0; r < RSF; r++) {
          const int sg = s0 + sl + r;
          if (sl + r >= ns || (active && !active[sg])) continue;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int o = 4 * grp + i;
            if (o >= DEC_OUT) continue;
            const float v = __fadd_rn(oacc[r][i], W.dec_output.bias[o]);
            if (out_mode == 0) out[((size_t)sg * T + t) * DEC_OUT + o] = v;
            else {
              const int fr = o / 21, f = o % 21;
              if (f < 20) out[((size_t)sg * 4 * T + 4 * t + fr) * RADE_NB_TOTAL_FEATURES + f] = v;
            }
            if (o == 20 && uw_count && v > 0.f) atomicAdd(&uw_count[sg], 1);
          }
          if (out_mode == 1) {                   // zero the 16 unused slots of each 36-wide vector
            for (int k = grp; k < 4 * 16; k += NGRP)
              out[((size_t)sg * 4 * T + 4 * t + k / 16) * RADE_NB_TOTAL_FEATURES + 20 + (k % 16)] = 0.f;
          }
        }
      }
    }
    nb_sync(NB_MAIN, N_MAIN);
    return;
  }

  if (warp == NF + NE + 1) {                     // ---- issuer
    nb_sync(NB_START, (NE + 1) * 32);
    if (elect_one()) {
      typedef RingCursor<I8_NST, UMMA_I8_STAGE_BYTES> Ring;
      IssueCtx<Ring> c;
      c.cx = Ring{&sm.i8, 0, 0u}; c.tmem = tmem; c.stage_lo = 0; c.act_ready = sm.act_ready; c.acc_full = sm.acc_full; c.trace = trace;
      c.B.lo[2] = 0; c.B.hi_cat = desc_hi(KB * 8); c.B.hi_hq = desc_hi(KH * 8);
      for (int t = 0; t < T; t++) {
        c.B.lo[0] = desc_lo(smem_u32(sm.cb[t % NCB])); c.B.lo[1] = desc_lo(smem_u32(sm.cb[(t + NCB - 1) % NCB]));
        c.B.lo[3] = desc_lo(smem_u32(sm.hq[t & 1])); c.B.lo[4] = desc_lo(smem_u32(sm.hq[(t + 1) & 1]));
        c.par = t & 1; c.t = t;
        if (t > 0) { mbar_wait(&sm.act_ready[14], (t - 1) & 1); tc_fence_after(); }      // as in the encoder (here the GLU dependency already orders it)
        mbar_wait(&sm.d1_ready[t & 1], (t >> 1) & 1);
        tc_fence_after();
        IssueAll<NS, false, 0, kUmmaDecProg.n>::run(c);
      }
    }
    __syncwarp();
    nb_sync(NB_MAIN, N_MAIN);
    return;
  }

  // ---------------------------------------------------------------- epilogue warps
  const int q = warp & 3, ch = (warp - NF) >> 2;
  const int et = tid - NF * 32;
  const int c0 = ch * NC;
  const uint32_t tlane = (uint32_t)(32 * q) << 16;
  constexpr int NET = NE * 32;
  for (int i = et; i < (NCB * C::CB_BYTES + 2 * C::HQ_BYTES) / 4; i += NET) reinterpret_cast<uint32_t *>(sm.cb[0])[i] = 0u;
  nb_sync(NB_E, NET);
  nb_arrive(NB_Z, NE * 32 + NF * 32);
  {                                              // all streams' state in one pass (see the encoder)
    constexpr int HN = 5 * DEC_GRU, CN = KB / 4;
    float hv[(NS * HN + NET - 1) / NET];
#pragma unroll
    for (int q = 0; q < (NS * HN + NET - 1) / NET; q++) {
      const int idx = et + q * NET, r = idx / HN, i = idx - r * HN;
      hv[q] = (idx < NS * HN && r < ns) ? state[s0 + r].h[i] : 0.f;
    }
    uint32_t c1[(NS * CN + NET - 1) / NET];
#pragma unroll
    for (int q = 0; q < (NS * CN + NET - 1) / NET; q++) {
      const int idx = et + q * NET, r = idx / CN, i = idx - r * CN;
      c1[q] = (idx < NS * CN && r < ns) ? reinterpret_cast<const uint32_t *>(state[s0 + r].cat1)[i] : 0u;
    }
#pragma unroll
    for (int q = 0; q < (NS * HN + NET - 1) / NET; q++) {
      const int idx = et + q * NET, r = idx / HN, i = idx - r * HN;
      if (idx < NS * HN) { sm.hs[r][i] = hv[q]; sm.hq[0][b_off<KH>(r, i)] = (uint8_t)quant8(hv[q]); }
    }
#pragma unroll
    for (int q = 0; q < (NS * CN + NET - 1) / NET; q++) {
      const int idx = et + q * NET, r = idx / CN, i = idx - r * CN;
      if (idx < NS * CN && r < ns) *reinterpret_cast<uint32_t *>(sm.cb[NCB - 1] + b_off<KB>(r, 4 * i)) = c1[q];
    }
  }
  fence_async_smem();
  nb_sync(NB_START, (NE + 1) * 32);               // with the issuer: state and concat buffers are in place

  int nseg = 0;
  const int u = q * 32 + lane;                   // output feature / hidden unit of this thread (valid for u < 96)
  for (int t = 0; t < T; t++) {
    uint8_t *cur = sm.cb[t % NCB], *hq_wr = sm.hq[(t + 1) & 1];
    const uint32_t par = t & 1;
    int off = 96;
#pragma unroll 1
    for (int l = 0; l < 5; l++) {
      // ---- GRU l: new state, kept un-gated (src/rade_dec.c:66-67); quantised copy -> hq_wr.  Blocks: z in, z rec, r in, r rec, n in, n rec
      {
        float si[3], bi[3], sr[3], br[3];
        if (q < 3)
#pragma unroll
          for (int g = 0; g < 3; g++) {
            si[g] = W.dec_gru_in[l].scale[g * DEC_GRU + u]; bi[g] = W.dec_gru_in[l].bias[g * DEC_GRU + u];
            sr[g] = W.dec_gru_rec[l].scale[g * DEC_GRU + u]; br[g] = W.dec_gru_rec[l].bias[g * DEC_GRU + u];
          }
        mbar_wait(&sm.acc_full[3 * l], par);
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 3 * l) * 2);
        tc_fence_after();
        if (q < 3) {
          const uint32_t gcol = tmem + (uint32_t)((l & 1) * 6) * NS + c0 + tlane;
          int acc[6][NC];
#pragma unroll
          for (int g = 0; g < 6; g++) tmem_ld<NC>(gcol + g * NS, acc[g]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NC; j++) {
            const int s = c0 + j;
            const float z = sigmoid_r(__fadd_rn(lin(acc[0][j], si[0], bi[0]), lin(acc[1][j], sr[0], br[0])));
            const float r = sigmoid_r(__fadd_rn(lin(acc[2][j], si[1], bi[1]), lin(acc[3][j], sr[1], br[1])));
            const float n = tanh_r(__fadd_rn(lin(acc[4][j], si[2], bi[2]), __fmul_rn(lin(acc[5][j], sr[2], br[2]), r)));
            const float hold = sm.hs[s][l * DEC_GRU + u];
            const float h = __fadd_rn(__fmul_rn(z, hold), __fmul_rn(__fsub_rn(1.f, z), n));
            sm.hs[s][l * DEC_GRU + u] = h;
            hq_wr[b_off<KH>(s, l * DEC_GRU + u)] = (uint8_t)quant8(h);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 3 * l) * 2 + 1);
        if (lane == 0) mbar_arrive(&sm.act_ready[3 * l]);
      }
      // ---- GLU l: out = h * sigmoid(Wg h + b) -> concat
      {
        float gs = 0.f, gb = 0.f;
        if (q < 3) { gs = W.dec_glu[l].scale[u]; gb = W.dec_glu[l].bias[u]; }
        const int slot = nseg % NSEG;
        mbar_wait(&sm.seg_empty[slot], ((nseg / NSEG) & 1) ^ 1);
        mbar_wait(&sm.acc_full[3 * l + 1], par);
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 3 * l + 1) * 2);
        tc_fence_after();
        if (q < 3) {
          int acc[NC];
          tmem_ld<NC>(tmem + (uint32_t)(12 + (l & 1)) * NS + c0 + tlane, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NC; j++) {
            const int s = c0 + j;
            const float y = __fmul_rn(sm.hs[s][l * DEC_GRU + u], sigmoid_r(lin(acc[j], gs, gb)));
            sm.seg[slot][s][u] = y;
            cur[b_off<KB>(s, off + u)] = (uint8_t)quant8(y);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 3 * l + 1) * 2 + 1);
        if (lane == 0) { mbar_arrive(&sm.act_ready[3 * l + 1]); mbar_arrive(&sm.seg_full[slot]); }
        nseg++;
      }
      off += DEC_GRU;
      // ---- conv l: 32 outputs on lanes 0..31
      {
        float es = 0.f, eb = 0.f;
        if (q == 0) { es = W.dec_conv[l].scale[u]; eb = W.dec_conv[l].bias[u]; }
        const int slot = nseg % NSEG;
        mbar_wait(&sm.seg_empty[slot], ((nseg / NSEG) & 1) ^ 1);
        mbar_wait(&sm.acc_full[3 * l + 2], par);
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 3 * l + 2) * 2);
        tc_fence_after();
        if (q == 0) {
          int acc[NC];
          tmem_ld<NC>(tmem + (uint32_t)(14 + (l & 1)) * NS + c0 + tlane, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NC; j++) {
            const int s = c0 + j;
            const float y = tanh_r(lin(acc[j], es, eb));
            sm.seg[slot][s][u] = y;
            cur[b_off<KB>(s, off + u)] = (uint8_t)quant8(y);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (q == 0 && ch == 0 && lane == 0) TR(2048 + (t * 16 + 3 * l + 2) * 2 + 1);
        if (lane == 0) { mbar_arrive(&sm.act_ready[3 * l + 2]); mbar_arrive(&sm.seg_full[slot]); }
        nseg++;
      }
      off += DEC_CONV;
    }
  }
  nb_sync(NB_MAIN, N_MAIN);
  if (warp == NF) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(C::TMEM_COLS) : "memory");
  const int last = (T + NCB - 1) % NCB;
  for (int r = 0; r < ns; r++) {
    if (active && !active[s0 + r]) continue;
    DecStreamState *st = state + (s0 + r);
    for (int i = et; i < 5 * DEC_GRU; i += NET) st->h[i] = sm.hs[r][i];
    for (int i = et; i < KB / 4; i += NET)
      reinterpret_cast<uint32_t *>(st->cat1)[i] = *reinterpret_cast<const uint32_t *>(sm.cb[last] + b_off<KB>(r, 4 * i));
  }
}

}  // namespace

// ----------------------------------------------------------------- host side
// tiles: at most NS streams each, and a whole number of waves over the SMs when there are more tiles than SMs
// full_tiles: fewer tiles than SMs are NOT spread over all SMs — the free SMs go to the kernels that run beside the codec in the
// frame pipeline (a resident codec CTA owns its SM).  1024 streams: 128 CTAs + 20 free SMs, step 0.304 -> 0.297 ms; the codec kernel
// alone is 2 % slower that way (96.7 vs 94.4 us), so a stand-alone call keeps the uneven tiles.
static int umma_grid(int S, int NS, int n_sm, int full_tiles) {
  int g = (S + NS - 1) / NS;
  if (g < n_sm && full_tiles) return g < 1 ? 1 : g;
  if (g < n_sm) g = S < n_sm ? S : n_sm;                          // fewer, smaller tiles than SMs: one CTA per SM
  else g = ((g + n_sm - 1) / n_sm) * n_sm;
  return g < 1 ? 1 : g;
}
static int g_n_sm = 148;
// tile width: 16 streams per CTA once 8-stream tiles no longer fit one wave over the SMs (RADE_B200_CODEC_NS = 8 | 16 overrides)
static int umma_tile(int S, int n_sm) {
  static const int forced = getenv("RADE_B200_CODEC_NS") ? atoi(getenv("RADE_B200_CODEC_NS")) : 0;
  if (forced == 8 || forced == 16) return forced;
  return S > 8 * n_sm ? 16 : 8;
}
int core_codec_umma_init_device() {
  int dev = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_n_sm, cudaDevAttrMultiProcessorCount, dev);
  CUDA_CHECK(cudaFuncSetAttribute(core_encoder_umma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmemU<8>)));
  CUDA_CHECK(cudaFuncSetAttribute(core_decoder_umma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmemU<8>)));
  CUDA_CHECK(cudaFuncSetAttribute(core_encoder_umma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmemU<16>)));
  CUDA_CHECK(cudaFuncSetAttribute(core_decoder_umma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmemU<16>)));
  return 0;
}
int core_encoder_umma_launch(const CoreWeightsDev &W, EncStreamState *state, const float *in, int in_mode, float *z,
                             const uint8_t *active, int S, int T, cudaStream_t stream) {
  if (umma_tile(S, g_n_sm) == 16)
    core_encoder_umma_kernel<16><<<umma_grid(S, 16, g_n_sm, W.full_tiles), EncCfg<16>::THREADS, sizeof(EncSmemU<16>), stream>>>(W, state, in, in_mode, z, active, S, T);
  else
    core_encoder_umma_kernel<8><<<umma_grid(S, 8, g_n_sm, W.full_tiles), EncCfg<8>::THREADS, sizeof(EncSmemU<8>), stream>>>(W, state, in, in_mode, z, active, S, T);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}
int core_decoder_umma_launch(const CoreWeightsDev &W, DecStreamState *state, const float *z, float *out, int out_mode,
                             int *uw_count, const uint8_t *active, int S, int T, cudaStream_t stream) {
  if (umma_tile(S, g_n_sm) == 16)
    core_decoder_umma_kernel<16><<<umma_grid(S, 16, g_n_sm, W.full_tiles), DecCfg<16>::THREADS, sizeof(DecSmemU<16>), stream>>>(W, state, z, out, out_mode, uw_count, active, S, T);
  else
    core_decoder_umma_kernel<8><<<umma_grid(S, 8, g_n_sm, W.full_tiles), DecCfg<8>::THREADS, sizeof(DecSmemU<8>), stream>>>(W, state, z, out, out_mode, uw_count, active, S, T);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}
