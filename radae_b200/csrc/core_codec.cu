// RADE core encoder / decoder on B200 — batched stateful streams.
//
// Replaces (per stream, per 40 ms step):
//   rade_core_encoder  /root/reference/src/rade_enc.c:55-114   (PyTorch twin radae/radae_base.py:260-286)
//   rade_core_decoder  /root/reference/src/rade_dec.c:50-102   (PyTorch twin radae/radae_base.py:400-416)
// and the opus DNN primitives they call (compute_generic_dense/gru/conv1d[_dilation], compute_glu).
//
// Design: one CTA owns a tile of CORE_TS = 16 independent streams and walks the
// whole layer stack for n_steps consecutive steps with every activation on chip:
//   * the DenseNet concat buffer lives in shared memory as int8 (exactly the quantised values floor(.5+127x) the
//     reference feeds its int8 GEMVs) — a ring of the current and the previous one/two steps, which is also the
//     conv1d tap memory and the GRU recurrent input, so conv "state" costs no extra storage;
//   * int8 layers: s8 x s8 -> s32 tensor-core MMA (m16n8k32; 16 streams x 8 outputs x 32 inputs per instruction),
//     weights pre-tiled in fragment order on the host and streamed from L2 (1.7 MB total, L2 resident);
//   * epilogues (scale, bias, rational tanh/sigmoid, GRU gating, GLU) straight from the accumulator registers,
//     each float operation separately rounded in the reference's order => bit-identical to the C oracle;
//   * the four float layers accumulate sequentially over inputs (the generic sgemv order), the two wide ones
//     (enc_zdense 864->80, dec_output 736->84) incrementally as each concat segment is produced.
// Per-stream HBM state: GRU h (fp32) + int8 concat of step t-1 (and t-2 for the encoder's dilation-2 convs).
#include "rade_common.h"

namespace {

constexpr int NTHREADS = 128;
constexpr int SEG_LD = 97;
constexpr int FIN_LD = 85;
constexpr int ZIN_LD = 81;
constexpr int HQ_LD = 496;                 // row stride of the decoder's quantised hidden states (conflict-free A fragments)

// ---------------------------------------------------------------- scalar math, bit-exact w.r.t. oracle/nnet_shim.c
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

// ---------------------------------------------------------------- tensor-core int8 tile GEMM
__device__ __forceinline__ void mma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// acc[i] += A[16 x 32*KB] * W[n-tile nt[i]][k-blocks kb0 .. kb0+KB)
// A: shared memory, int8 row-major, row stride lda bytes (lda/4 odd multiple of 4 words => conflict-free)
// Wt: global, fragment order [(nt*KBtot + kb)*32 + lane] = {b0,b1}
template <int NT>
__device__ __forceinline__ void gemm_i8(int (&acc)[NT][4], const int8_t *A, int lda, int KB,
                                        const uint2 *__restrict__ Wt, const int (&nt)[NT], int KBtot, int kb0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const uint32_t *r0 = reinterpret_cast<const uint32_t *>(A + g * lda) + tig;
  const uint32_t *r1 = reinterpret_cast<const uint32_t *>(A + (g + 8) * lda) + tig;
  const uint2 *w[NT];
#pragma unroll
  for (int i = 0; i < NT; i++) w[i] = Wt + ((size_t)nt[i] * KBtot + kb0) * 32 + lane;
  constexpr int U = 4;                        // k-blocks of B fragments in flight per warp
  int kb = 0;
  for (; kb + U <= KB; kb += U) {
    uint2 b[U][NT];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int i = 0; i < NT; i++) b[u][i] = __ldg(w[i] + (kb + u) * 32);
#pragma unroll
    for (int u = 0; u < U; u++) {
      uint32_t a0 = r0[(kb + u) * 8], a1 = r1[(kb + u) * 8], a2 = r0[(kb + u) * 8 + 4], a3 = r1[(kb + u) * 8 + 4];
#pragma unroll
      for (int i = 0; i < NT; i++) mma_s8(acc[i], a0, a1, a2, a3, b[u][i].x, b[u][i].y);
    }
  }
  for (; kb < KB; kb++) {
    uint32_t a0 = r0[kb * 8], a1 = r1[kb * 8], a2 = r0[kb * 8 + 4], a3 = r1[kb * 8 + 4];
#pragma unroll
    for (int i = 0; i < NT; i++) {
      uint2 b = __ldg(w[i] + kb * 32);
      mma_s8(acc[i], a0, a1, a2, a3, b.x, b.y);
    }
  }
}

// ---------------------------------------------------------------- float layers: sequential-in-j accumulation
// thread (s = tid&15, grp = tid>>4) owns outputs o = grp + 8*i of stream s
template <int NOUT, int NACC>
__device__ __forceinline__ void dense_acc(float (&acc)[NACC], const float *xrow, int K, const float *__restrict__ wf, int grp) {
  constexpr int U = 8;                        // rows of W in flight per thread: hides the L2 latency of the weight fetch
  int j = 0;
  for (; j + U <= K; j += U) {
    float w[U][NACC];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int i = 0; i < NACC; i++)
        w[u][i] = (grp + 8 * i < NOUT) ? __ldg(wf + (size_t)(j + u) * NOUT + grp + 8 * i) : 0.f;
#pragma unroll
    for (int u = 0; u < U; u++) {
      const float x = xrow[j + u];
#pragma unroll
      for (int i = 0; i < NACC; i++)
        if (grp + 8 * i < NOUT) acc[i] = __fadd_rn(acc[i], __fmul_rn(w[u][i], x));
    }
  }
  for (; j < K; j++) {
    const float x = xrow[j];
    const float *wj = wf + (size_t)j * NOUT + grp;
#pragma unroll
    for (int i = 0; i < NACC; i++)
      if (grp + 8 * i < NOUT) acc[i] = __fadd_rn(acc[i], __fmul_rn(__ldg(wj + 8 * i), x));
  }
}

// GRU layer for one tile of 16 streams.  Xin: current concat (K = layer.K inputs); Hq: quantised previous hidden state.
// Each warp owns unit tiles u = warp, warp+4, ...; gates z,r,n of a unit land in the same accumulator slot.
template <int UNITS, typename Emit>
__device__ __forceinline__ void gru_layer(const I8LayerDev &Li, const I8LayerDev &Lr, const int8_t *Xin, int ldx,
                                          const int8_t *Hq, int ldh, float *hs /*[16][ldhs] slice for this layer*/, int ldhs,
                                          Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  constexpr int U = UNITS / 8;
  const int KBi = Li.K / 32, KBr = Lr.K / 32;
  for (int u = warp; u < U; u += 4) {
    int ai[3][4] = {}, ar[3][4] = {};
    const int nt[3] = {u, U + u, 2 * U + u};
    gemm_i8<3>(ai, Xin, ldx, KBi, Li.wt, nt, KBi, 0);
    gemm_i8<3>(ar, Hq, ldh, KBr, Lr.wt, nt, KBr, 0);
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int row = g + ((e & 2) ? 8 : 0);
      const int j = u * 8 + 2 * tig + (e & 1);
      float z = sigmoid_r(__fadd_rn(lin(ai[0][e], Li.scale[j], Li.bias[j]), lin(ar[0][e], Lr.scale[j], Lr.bias[j])));
      float r = sigmoid_r(__fadd_rn(lin(ai[1][e], Li.scale[UNITS + j], Li.bias[UNITS + j]),
                                    lin(ar[1][e], Lr.scale[UNITS + j], Lr.bias[UNITS + j])));
      float n = tanh_r(__fadd_rn(lin(ai[2][e], Li.scale[2 * UNITS + j], Li.bias[2 * UNITS + j]),
                                 __fmul_rn(lin(ar[2][e], Lr.scale[2 * UNITS + j], Lr.bias[2 * UNITS + j]), r)));
      float hold = hs[row * ldhs + j];
      float h = __fadd_rn(__fmul_rn(z, hold), __fmul_rn(__fsub_rn(1.f, z), n));
      hs[row * ldhs + j] = h;
      emit(row, j, h);
    }
  }
}

// plain int8 linear layer (conv taps / GLU gate): N outputs, up to two A operands accumulated into the same tile
template <int N, typename Emit>
__device__ __forceinline__ void i8_layer(const I8LayerDev &L, const int8_t *A0, int K0, const int8_t *A1, int K1, int lda, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  constexpr int NTILES = N / 8;
  constexpr int PER = (NTILES + 3) / 4;      // n-tiles per warp
  const int KBtot = L.K / 32;
  int acc[PER][4] = {};
  int nt[PER];
#pragma unroll
  for (int i = 0; i < PER; i++) nt[i] = warp + 4 * i;     // NTILES is a multiple of 4 for every RADE layer
  gemm_i8<PER>(acc, A0, lda, K0 / 32, L.wt, nt, KBtot, 0);
  if (A1) gemm_i8<PER>(acc, A1, lda, K1 / 32, L.wt, nt, KBtot, K0 / 32);
#pragma unroll
  for (int i = 0; i < PER; i++)
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int row = g + ((e & 2) ? 8 : 0);
      const int n = nt[i] * 8 + 2 * tig + (e & 1);
      emit(row, n, lin(acc[i][e], L.scale[n], L.bias[n]));
    }
}

// ================================================================= encoder
struct EncSmem {
  int8_t cb[3][CORE_TS][ENC_LDA];
  float hs[CORE_TS][5 * ENC_GRU];
  float seg[CORE_TS][SEG_LD];
  float fin[CORE_TS][FIN_LD];
};

__global__ void __launch_bounds__(NTHREADS)
core_encoder_kernel(CoreWeightsDev W, EncStreamState *__restrict__ state, const float *__restrict__ in, int in_mode,
                    float *__restrict__ z_out, const uint8_t *__restrict__ active, int S, int T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EncSmem &sm = *reinterpret_cast<EncSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * CORE_TS;
  const int sl = tid & 15, grp = tid >> 4;
  const int sg = s0 + sl;                         // this thread's stream for the float layers / IO

  // tile-level early exit when no stream of the tile is active
  __shared__ int any_active;
  if (tid == 0) any_active = 0;
  __syncthreads();
  if (tid < CORE_TS && s0 + tid < S && (!active || active[s0 + tid])) any_active = 1;
  __syncthreads();
  if (!any_active) return;

  // ---- load state: h -> hs, cat1 -> cb[2], cat2 -> cb[1]; zero cb[0]
  for (int r = 0; r < CORE_TS; r++) {
    const bool ok = s0 + r < S;
    const EncStreamState *st = state + (s0 + r);
    for (int i = tid; i < 5 * ENC_GRU; i += NTHREADS) sm.hs[r][i] = ok ? st->h[i] : 0.f;
    for (int i = tid; i < ENC_LDA / 4; i += NTHREADS) {
      reinterpret_cast<uint32_t *>(sm.cb[2][r])[i] = ok ? reinterpret_cast<const uint32_t *>(st->cat1)[i] : 0u;
      reinterpret_cast<uint32_t *>(sm.cb[1][r])[i] = ok ? reinterpret_cast<const uint32_t *>(st->cat2)[i] : 0u;
      reinterpret_cast<uint32_t *>(sm.cb[0][r])[i] = 0u;
    }
  }
  __syncthreads();

  constexpr int dil[5] = {1, 2, 2, 2, 2};
  for (int t = 0; t < T; t++) {
    int8_t(*cur)[ENC_LDA] = sm.cb[t % 3];
    int8_t(*prev1)[ENC_LDA] = sm.cb[(t + 2) % 3];
    int8_t(*prev2)[ENC_LDA] = sm.cb[(t + 1) % 3];

    // ---- input features -> fin[16][84]
    for (int i = tid; i < CORE_TS * ENC_IN; i += NTHREADS) {
      const int r = i / ENC_IN, k = i % ENC_IN;
      float v = 0.f;
      if (s0 + r < S) {
        if (in_mode == 0) v = in[((size_t)(s0 + r) * T + t) * ENC_IN + k];
        else {                                   // API layout: [S][4T][36]; 20 used features + aux = -1 (src/rade_api.c:426-432)
          const int fr = k / 21, f = k % 21;
          v = (f == 20) ? -1.f : in[((size_t)(s0 + r) * 4 * T + 4 * t + fr) * RADE_NB_TOTAL_FEATURES + f];
        }
      }
      sm.fin[r][k] = v;
    }
    __syncthreads();

    // ---- dense1: tanh(W f + b), 84 -> 64
    {
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = 0.f;
      dense_acc<64, 8>(a, sm.fin[sl], ENC_IN, W.enc_dense1.wf, grp);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int o = grp + 8 * i;
        float y = tanh_r(__fadd_rn(a[i], W.enc_dense1.bias[o]));
        sm.seg[sl][o] = y;
        cur[sl][o] = quant8(y);
      }
    }
    __syncthreads();
    float zacc[10];
#pragma unroll
    for (int i = 0; i < 10; i++) zacc[i] = 0.f;
    dense_acc<80, 10>(zacc, sm.seg[sl], 64, W.enc_zdense.wf, grp);
    __syncthreads();

    int off = 64;
#pragma unroll 1
    for (int l = 0; l < 5; l++) {
      // GRU l: input = cur[0:off), recurrent input = quantised h(t-1) = prev1[off : off+64)
      gru_layer<ENC_GRU>(W.enc_gru_in[l], W.enc_gru_rec[l], &cur[0][0], ENC_LDA, &prev1[0][off], ENC_LDA,
                         &sm.hs[0][l * ENC_GRU], 5 * ENC_GRU,
                         [&](int row, int j, float h) { sm.seg[row][j] = h; cur[row][off + j] = quant8(h); });
      __syncthreads();
      dense_acc<80, 10>(zacc, sm.seg[sl], ENC_GRU, W.enc_zdense.wf + (size_t)off * 80, grp);
      __syncthreads();
      off += ENC_GRU;
      // conv l (k = 2): tap 0 = concat prefix of step t-dilation, tap 1 = current prefix
      const int8_t *old = (dil[l] == 1) ? &prev1[0][0] : &prev2[0][0];
      i8_layer<ENC_CONV>(W.enc_conv[l], old, off, &cur[0][0], off, ENC_LDA,
                         [&](int row, int n, float v) { float y = tanh_r(v); sm.seg[row][n] = y; cur[row][off + n] = quant8(y); });
      __syncthreads();
      dense_acc<80, 10>(zacc, sm.seg[sl], ENC_CONV, W.enc_zdense.wf + (size_t)off * 80, grp);
      __syncthreads();
      off += ENC_CONV;
    }
    // ---- z = zdense(cat) + b   (bottleneck 3: linear, src/rade_enc.c:107-113)
    if (sg < S && (!active || active[sg])) {
#pragma unroll
      for (int i = 0; i < 10; i++) {
        const int o = grp + 8 * i;
        z_out[((size_t)sg * T + t) * RADE_LATENT + o] = __fadd_rn(zacc[i], W.enc_zdense.bias[o]);
      }
    }
  }

  // ---- store state
  const int last = (T + 2) % 3, last2 = (T + 1) % 3;
  for (int r = 0; r < CORE_TS; r++) {
    if (s0 + r >= S || (active && !active[s0 + r])) continue;
    EncStreamState *st = state + (s0 + r);
    for (int i = tid; i < 5 * ENC_GRU; i += NTHREADS) st->h[i] = sm.hs[r][i];
    for (int i = tid; i < ENC_LDA / 4; i += NTHREADS) {
      reinterpret_cast<uint32_t *>(st->cat1)[i] = reinterpret_cast<const uint32_t *>(sm.cb[last][r])[i];
      reinterpret_cast<uint32_t *>(st->cat2)[i] = reinterpret_cast<const uint32_t *>(sm.cb[last2][r])[i];
    }
  }
}

// ================================================================= decoder
struct DecSmem {
  int8_t cb[2][CORE_TS][DEC_LDA];
  int8_t hq[2][CORE_TS][HQ_LD];
  float hs[CORE_TS][5 * DEC_GRU];
  float seg[CORE_TS][SEG_LD];
  float zin[CORE_TS][ZIN_LD];
};

// out_mode 0: features [S][T][84];  out_mode 1: API layout [S][4T][36] (20 used, rest zero, src/rade_api.c:488-500)
// uw_count (optional): += number of steps whose first aux symbol (feature 20) is > 0 (src/rade_api.c:502-505)
__global__ void __launch_bounds__(NTHREADS)
core_decoder_kernel(CoreWeightsDev W, DecStreamState *__restrict__ state, const float *__restrict__ z_in,
                    float *__restrict__ out, int out_mode, int *__restrict__ uw_count,
                    const uint8_t *__restrict__ active, int S, int T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DecSmem &sm = *reinterpret_cast<DecSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * CORE_TS;
  const int sl = tid & 15, grp = tid >> 4;
  const int sg = s0 + sl;

  __shared__ int any_active;
  if (tid == 0) any_active = 0;
  __syncthreads();
  if (tid < CORE_TS && s0 + tid < S && (!active || active[s0 + tid])) any_active = 1;
  __syncthreads();
  if (!any_active) return;

  for (int r = 0; r < CORE_TS; r++) {
    const bool ok = s0 + r < S;
    const DecStreamState *st = state + (s0 + r);
    for (int i = tid; i < 5 * DEC_GRU; i += NTHREADS) {
      float h = ok ? st->h[i] : 0.f;
      sm.hs[r][i] = h;
      sm.hq[0][r][i] = quant8(h);
    }
    for (int i = tid; i < DEC_LDA / 4; i += NTHREADS) {
      reinterpret_cast<uint32_t *>(sm.cb[1][r])[i] = ok ? reinterpret_cast<const uint32_t *>(st->cat1)[i] : 0u;
      reinterpret_cast<uint32_t *>(sm.cb[0][r])[i] = 0u;
    }
  }
  __syncthreads();

  for (int t = 0; t < T; t++) {
    int8_t(*cur)[DEC_LDA] = sm.cb[t & 1];
    int8_t(*prev1)[DEC_LDA] = sm.cb[(t + 1) & 1];
    int8_t(*hq_rd)[HQ_LD] = sm.hq[t & 1];
    int8_t(*hq_wr)[HQ_LD] = sm.hq[(t + 1) & 1];

    for (int i = tid; i < CORE_TS * DEC_IN; i += NTHREADS) {
      const int r = i / DEC_IN, k = i % DEC_IN;
      sm.zin[r][k] = (s0 + r < S) ? z_in[((size_t)(s0 + r) * T + t) * DEC_IN + k] : 0.f;
    }
    __syncthreads();

    // ---- dense1: tanh(W z + b), 80 -> 96
    {
      float a[12];
#pragma unroll
      for (int i = 0; i < 12; i++) a[i] = 0.f;
      dense_acc<96, 12>(a, sm.zin[sl], DEC_IN, W.dec_dense1.wf, grp);
#pragma unroll
      for (int i = 0; i < 12; i++) {
        const int o = grp + 8 * i;
        float y = tanh_r(__fadd_rn(a[i], W.dec_dense1.bias[o]));
        sm.seg[sl][o] = y;
        cur[sl][o] = quant8(y);
      }
    }
    __syncthreads();
    float oacc[11];
#pragma unroll
    for (int i = 0; i < 11; i++) oacc[i] = 0.f;
    dense_acc<DEC_OUT, 11>(oacc, sm.seg[sl], 96, W.dec_output.wf, grp);
    __syncthreads();

    int off = 96;
#pragma unroll 1
    for (int l = 0; l < 5; l++) {
      // GRU l on cur[0:off); its new state is kept un-gated (src/rade_dec.c:66-67)
      gru_layer<DEC_GRU>(W.dec_gru_in[l], W.dec_gru_rec[l], &cur[0][0], DEC_LDA, &hq_rd[0][l * DEC_GRU], HQ_LD,
                         &sm.hs[0][l * DEC_GRU], 5 * DEC_GRU,
                         [&](int row, int j, float h) { hq_wr[row][l * DEC_GRU + j] = quant8(h); });
      __syncthreads();
      // GLU l: out = h * sigmoid(Wg h + b)  -> concat
      i8_layer<DEC_GRU>(W.dec_glu[l], &hq_wr[0][l * DEC_GRU], DEC_GRU, nullptr, 0, HQ_LD,
                        [&](int row, int n, float v) {
                          float y = __fmul_rn(sm.hs[row][l * DEC_GRU + n], sigmoid_r(v));
                          sm.seg[row][n] = y; cur[row][off + n] = quant8(y);
                        });
      __syncthreads();
      dense_acc<DEC_OUT, 11>(oacc, sm.seg[sl], DEC_GRU, W.dec_output.wf + (size_t)off * DEC_OUT, grp);
      __syncthreads();
      off += DEC_GRU;
      i8_layer<DEC_CONV>(W.dec_conv[l], &prev1[0][0], off, &cur[0][0], off, DEC_LDA,
                         [&](int row, int n, float v) { float y = tanh_r(v); sm.seg[row][n] = y; cur[row][off + n] = quant8(y); });
      __syncthreads();
      dense_acc<DEC_OUT, 11>(oacc, sm.seg[sl], DEC_CONV, W.dec_output.wf + (size_t)off * DEC_OUT, grp);
      __syncthreads();
      off += DEC_CONV;
    }

    if (sg < S && (!active || active[sg])) {
#pragma unroll
      for (int i = 0; i < 11; i++) {
        const int o = grp + 8 * i;
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
        for (int k = grp; k < 4 * 16; k += 8)
          out[((size_t)sg * 4 * T + 4 * t + k / 16) * RADE_NB_TOTAL_FEATURES + 20 + (k % 16)] = 0.f;
      }
    }
  }

  const int last = (T + 1) & 1;      // buffer that held the final step's concat
  for (int r = 0; r < CORE_TS; r++) {
    if (s0 + r >= S || (active && !active[s0 + r])) continue;
    DecStreamState *st = state + (s0 + r);
    for (int i = tid; i < 5 * DEC_GRU; i += NTHREADS) st->h[i] = sm.hs[r][i];
    for (int i = tid; i < DEC_LDA / 4; i += NTHREADS)
      reinterpret_cast<uint32_t *>(st->cat1)[i] = reinterpret_cast<const uint32_t *>(sm.cb[last][r])[i];
  }
}

}  // namespace

// ----------------------------------------------------------------- host launchers
int core_encoder_launch(const CoreWeightsDev &W, EncStreamState *state, const float *in, int in_mode, float *z,
                        const uint8_t *active, int S, int T, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(core_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EncSmem)));
    attr_set = true;
  }
  const int grid = (S + CORE_TS - 1) / CORE_TS;
  core_encoder_kernel<<<grid, NTHREADS, sizeof(EncSmem), stream>>>(W, state, in, in_mode, z, active, S, T);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}

int core_decoder_launch(const CoreWeightsDev &W, DecStreamState *state, const float *z, float *out, int out_mode,
                        int *uw_count, const uint8_t *active, int S, int T, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(core_decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem)));
    attr_set = true;
  }
  const int grid = (S + CORE_TS - 1) / CORE_TS;
  core_decoder_kernel<<<grid, NTHREADS, sizeof(DecSmem), stream>>>(W, state, z, out, out_mode, uw_count, active, S, T);
  CUDA_CHECK(cudaGetLastError());
  return 0;
}
