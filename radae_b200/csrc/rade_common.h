// Shared definitions for libradae_b200 (sm_100a only; there is no CPU fallback anywhere in this library).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

// ---- waveform / model constants for model19_check3 (reference: radae/radae.py:128-232) ----
#define RADE_FS 8000
#define RADE_M 160
#define RADE_NCP 32
#define RADE_NC 30
#define RADE_NS 4
#define RADE_NZMF 3
#define RADE_LATENT 80
#define RADE_SYM (RADE_M + RADE_NCP)          // 192
#define RADE_NMF 960
#define RADE_NEOO 1152
#define RADE_NIN_MAX 1120
#define RADE_RXBUF 2112
#define RADE_NFEAT 432                        // 12 x 36 floats per modem frame at the API
#define RADE_NB_TOTAL_FEATURES 36
#define RADE_NUM_USED_FEATURES 20
#define RADE_NEOO_BITS 180
#define RADE_NFCOARSE 40
#define RADE_CSK 24                            // padded k per row of cs_tab
#define RADE_NUPDATE 48
#define RADE_NMF_UNSYNC 25
#define RADE_SYNCED_ONE_SEC 8
#define RADE_UW_THRESH 7
#define RADE_BPF_NTAP 101
#define RADE_BPF_MEM 102                      // the reference keeps Ntap+1 samples (radae/dsp.py:96)
#define RADE_TIME_OFFSET (-16)

// ---- core codec dimensions (reference: src/rade_enc_data.h, src/rade_dec_data.h) ----
#define ENC_IN 84
#define ENC_CAT 864
#define ENC_GRU 64
#define ENC_CONV 96
#define DEC_IN 80
#define DEC_CAT 736
#define DEC_GRU 96
#define DEC_CONV 32
#define DEC_OUT 84

#define ENC_LDA 880                           // smem/global row stride of the int8 concat buffer (bank-conflict-free A fragments)
#define DEC_LDA 752

// per-stream persistent state in HBM
struct __align__(16) EncStreamState {
  float h[5 * ENC_GRU];                       // GRU hidden states, fp32 (src/rade_enc.h:37-41)
  int8_t cat1[ENC_LDA];                       // quantised concat buffer of step t-1: conv taps (dilation 1) + recurrent GEMM input
  int8_t cat2[ENC_LDA];                       // step t-2: dilation-2 conv taps (src/rade_enc.h:43-46 as int8)
};
struct __align__(16) DecStreamState {
  float h[5 * DEC_GRU];                       // src/rade_dec.h:36-40
  int8_t cat1[DEC_LDA];                       // step t-1 concat (all five convs have dilation 1, src/rade_dec.c:68-95)
};

// ---- weight streaming (core_codec.cu): all weights of one 40 ms step, laid out as a sequence of chunks in the exact
// order the layer walk consumes them; a producer warp TMA-bulk-copies chunk after chunk into a shared-memory ring.
#define CORE_STAGE_BYTES 32768
#define ENC_NI 8                              // int8-layer warps (encoder): one GRU unit tile each, three conv (n-tile, tap) units each
#define DEC_NI 12                             // int8-layer warps (decoder): 12 GRU unit tiles / 12 GLU n-tiles -> one per warp
#define ENC_NF 5                              // float-layer warps (dense1 + the incremental zdense / output accumulation)
#define DEC_NF 3
#define DEC_OUTP 96                           // dec_output rows are padded from 84 to 96 floats in the weight stream

struct I8LayerDev { const float *scale; const float *bias; int K; int N; };
struct F32LayerDev { const float *bias; int K; int N; };
struct ChunkDesc { unsigned int offset; unsigned int bytes; };
#define CORE_MAX_CHUNKS 96
// chunks[0 .. n_prologue) are streamed once per launch (dense1 of the first step), chunks[n_prologue .. n_chunks) once per step
struct CodecStreamDev { const unsigned char *stream; const ChunkDesc *chunks; int n_chunks; int n_prologue; };
// ---- tcgen05 formulation of the codec (core_codec_umma.cu): D[out feature][stream] = W[out][K] x X[stream][K]^T, kind::i8,
// A = weight rows (M = 128 TMEM lanes), B = the quantised activations of the tile's streams, int32 accumulators in TMEM.
// The host compiles, next to the weight streams, the per-step MMA PROGRAM the issuer thread walks: one UmmaOp per int8 matrix.
#define UMMA_I8_STAGE_BYTES 40960              // ring stage of the int8 weight stream: one cp.async.bulk each (the TMA unit of an SM
                                               // completes ~1 bulk copy per 440 cycles whatever its size: few, large copies)
#define UMMA_F32_STAGE_BYTES 20480             // ring stage of the float weight stream (rows of dense1 / zdense / output)
#define UMMA_MAX_RECS 120
#define UMMA_MAX_I8_CHUNKS 48
#define UMMA_MAX_F32_CHUNKS 24
enum { UB_CUR = 0, UB_PREV1 = 1, UB_PREV2 = 2, UB_HQ_RD = 3, UB_HQ_WR = 4 };
enum { UR_STAGE_FIRST = 1, UR_STAGE_LAST = 2, UR_ZERO_FIRST = 4 };
// One record per weight image: nk k-blocks (32 bytes of K each) of all rows of an int8 matrix in the canonical K-major operand
// layout (8-row group stride = nk * 256 B), packed back to back into the ring stages.  The list is fixed at compile time
// (umma_program.h); this run-time form exists for the debug hook and the CPU emulation test of the program.
struct UmmaRec {
  unsigned short a_off16;      // offset of the image inside its ring stage, in 16-byte units
  unsigned short tile_step;    // rows between the starts of consecutive M = 128 tiles
  unsigned short b_kb;         // first k-block of the B operand inside its buffer
  unsigned char nk;            // 1..8 (one tile), 1..4 (two tiles), 1..3 (three tiles)
  unsigned char n_tiles;
  unsigned char b_buf;         // UB_*
  unsigned char flags;         // UR_*
  unsigned char d_blk;         // accumulator column block (x NS columns) of tile 0
  unsigned char d_tile_stride; // column blocks between consecutive tiles
  signed char dep;             // act_ready barrier to wait for before the first MMA of this record, -1 = none
  signed char commit;          // acc_full barrier to commit after the last MMA of this record, -1 = none
  unsigned short pad;
};
struct UmmaCodecDev {
  const unsigned char *i8_stream; const ChunkDesc *i8_chunks; int n_i8_chunks;          // per step: one chunk = one ring stage = one bulk copy
  const unsigned char *f32_stream; const ChunkDesc *f32_chunks; int n_f32_chunks, n_f32_prologue;
};
struct CoreWeightsDev {
  F32LayerDev enc_dense1, enc_zdense, dec_dense1, dec_output;
  I8LayerDev enc_gru_in[5], enc_gru_rec[5], enc_conv[5];
  I8LayerDev dec_gru_in[5], dec_gru_rec[5], dec_glu[5], dec_conv[5];
  CodecStreamDev enc_stream, dec_stream;
  UmmaCodecDev enc_umma, dec_umma;
  int full_tiles;          // host-side hint for the launchers: keep full tiles and leave SMs free when other kernels run beside the codec
  float one;               // 1.0f as a RUN-TIME value: acc = fma(round(w x), one, acc) is an exact packed add that ptxas cannot contract with the multiply
  int float_fma;           // MEASUREMENT ONLY (RADE_B200_DEBUG_FLOAT_FMA=1): wide float layers with fused multiply-add — not bit-exact, 9 % faster encoder
  long long *trace;        // debug: clock64() stamps of CTA 0's warp roles (rade_b200_debug_trace_*), nullptr in production
  int enc_z_tanh;          // bottleneck 1 (model05): tanh on the latents, src/rade_enc.c:107-113; 0 for bottleneck 3
};

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "libradae_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
  return -1; } } while (0)
#define CUDA_CHECK_FATAL(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "libradae_b200: fatal CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
  exit(1); } } while (0)

// ---- DSP constant tables (device pointers), built on the host exactly the way RADAE.__init__ / acquisition.__init__ /
//      receiver_one.__init__ / complex_bpf.__init__ derive them (radae/radae.py:172-219, radae/dsp.py:40-61,153-176,401-412)
struct DspTables {
  const float2 *Winv;      // [30][160]  exp(+j n w_c)/M
  const float2 *Wfwd;      // [160][30]  exp(-j n w_c)
  const float2 *P;         // [30] pilots (freq domain), Pend [30]
  const float2 *Pend;
  const float2 *p;         // [160] time-domain pilot symbol, pend [160]
  const float2 *pend;
  const float2 *p_w;       // [160][40] coarse-frequency-shifted pilots (reference table; kept for the debug hook)
  const float2 *cs_tab;    // [160][24] (cos, sin)(2*pi*2.5k*n/Fs), k = 0..20 (+3 zero pads): the coarse grid is +-2.5k Hz
  const unsigned char *acq_tab;  // AcqTables (below): cs_tab + pilot tables packed for ONE TMA bulk copy into shared memory
  const float2 *Pmat;      // [30][2][3] LS projectors
  const float2 *eq_rot;    // [30] exp(-j w_c a)
  const float *bpf_h;      // [101]
  const float2 *bpf_exp;   // [1152] exp(-j alpha (i+1)), float32 argument (rx reads [0,1120), the TX filter up to the EOO frame)
  const float2 *eoo_base;  // [1152] P E 0 0 0 E frame after the PA limiter
  const float *fcoarse;    // [40]
  float pilot_gain;
  float p0_abs;            // |P[0]|
};

// constant block the acquisition kernels keep resident in shared memory (rx_detect / rx_track), filled by one bulk copy
#define RADE_SRANK 6                           // rank of the coarse-grid basis (each of the cos and the sin family), residual < 1e-8
struct __align__(128) AcqTables {
  // Coarse grid in a low-rank basis.  The 40 grid frequencies are +-2.5 k Hz, k = 0..20, over a 160-sample window: time-bandwidth
  // product 2, so cos(w_k (m + 1/2)) and sin(w_k (m + 1/2)), m = 0..79 (the window folded about its centre), are spanned to
  // 1e-9 by six vectors each (singular values fall by ~100x per index).  basis[m] = {bc[m][0..3]}, {bc[m][4..5], 0, 0},
  // {bs[m][0..3]}, {bs[m][4..5], 0, 0};  expand[k] = {ac[k][0..3]}, {ac[k][4..5], as[k][0..1]}, {as[k][2..5]} with
  // cos(w_k (m + 1/2)) = sum_r ac[k][r] bc[m][r], sin(...) = sum_r as[k][r] bs[m][r]   (tables.cpp, one-sided Jacobi SVD)
  float4 basis[RADE_M / 2][4];
  float4 expand[21][3];
  // fine search around the tracked frequency (refine_moments in ofdm_rx.cu): bk[n][k] = ((n - 79.5) / 80)^k / k!, k = 0..8, and
  // phd[pos][i] = exp(-j 2 pi (-1 + 0.1 i) / Fs * (79.5 + 960 pos)), the part of the centre / next-frame phase that does not
  // depend on the stream
  double bk[RADE_M][10];
  double2 phd[2][24];
  double2 phd10[2][80];          // first fix: exp(-j 2 pi (-10 + 0.25 i) / Fs * (79.5 + 960 pos))
  float4 ps4[RADE_M];            // (p.x, p.y, p.y, -p.x): conj(x)*p = x.x*(.x,.y) + x.y*(.z,.w) with two packed FMAs
  double2 pcd[RADE_M];           // conj(p) widened to complex128 (refine steering vectors)
  float2 pend[RADE_M];           // end-of-over pilot symbol
};

// per-stream receiver control block (everything radae_rx keeps between calls, radae_rxe.py:128-142)
struct __align__(16) RxCtl {
  double fmax, foff_err, snr_est, rx_phase_re, rx_phase_im;
  unsigned long long detect_key;
  float2 bpf_phase;
  float Dthresh, Dtmax12, Dtmax12_eoo, pad0;
  int state, nin, tmax, tmax_candidate;
  int valid_count, synced_count, n_check, bpf_first;
  int ring_head, candidate, endofover, valid_output;
  int uw_fail, ret, ran_sync, tracking;       // tracking: in sync when the current call began (rx_bpf)
};

// state of the optional TX band-pass filter (radae_tx(txbpf_en=True)); all zero == a new complex_bpf object
struct TxBpfState {
  float2 mem[RADE_BPF_MEM];
  float2 phase;
  int started, pad;
};

// per-stream channel-simulator state
struct __align__(16) ChanState {
  double phase;            // carrier phase accumulated so far (radians)
  long long t;             // samples processed
  float2 delay[64];        // last tx samples (two-path delay line)
};
