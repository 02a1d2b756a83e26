// The per-step MMA program of the tcgen05 codec kernels, computed at COMPILE TIME (constexpr) from the layer dimensions and shared
// by the host (weights.cpp bakes the int8 weight stream in exactly this order and packing) and the device (core_codec_umma.cu
// instantiates one straight-line block of tcgen05.mma per record: every descriptor offset, accumulator column, dependency and
// commit is an immediate).  Why compile time: a record list interpreted at run time by the single issuer thread cost ~480 cycles
// per record (decode, dispatch, uniform-datapath latency) next to ~30 per MMA — the issuer, not the tensor core, bounded the kernel.
//
// A record = one weight image: nk k-blocks (32 bytes of K each) of ALL rows of an int8 matrix in the canonical K-major operand
// layout (8-row group stride = nk * 256 B), packed back to back into ring stages of <= UMMA_I8_STAGE_BYTES (one bulk copy each).
#pragma once
#include "rade_common.h"

enum { UM_ENC_GRU_IN = 0, UM_ENC_GRU_REC = 5, UM_ENC_CONV = 10, UM_DEC_GRU_IN = 15, UM_DEC_GRU_REC = 20, UM_DEC_GLU = 25, UM_DEC_CONV = 30 };

struct UmmaRecC {
  int mat, kb;                 // source: matrix id (UM_* + layer) and first k-block inside it
  int n_rows, K;               // the matrix is int8 [n_rows][K]
  int a_off16;                 // offset of the image inside its ring stage, in 16-byte units
  int tile_step;               // rows between the starts of consecutive M = 128 tiles
  int b_kb;                    // first k-block of the B operand inside its buffer
  int nk, n_tiles;
  int b_buf;                   // UB_*
  int flags;                   // UR_*
  int d_blk, d_tile_stride;    // accumulator column block (x NS columns) of tile 0 / between tiles
  int dep, commit;             // act_ready barrier to wait for before / acc_full barrier to commit after this record (-1 = none)
};
struct UmmaProgC {
  int n;
  UmmaRecC r[UMMA_MAX_RECS];
  int n_stages;
  int stage_bytes[UMMA_MAX_I8_CHUNKS];
  int cur_bytes;               // builder state: fill of the open stage
  bool ok;
};

constexpr void umma_prog_close(UmmaProgC &P) {
  if (P.cur_bytes == 0) return;
  P.r[P.n - 1].flags |= UR_STAGE_LAST;
  if (P.n_stages >= UMMA_MAX_I8_CHUNKS) { P.ok = false; return; }
  P.stage_bytes[P.n_stages++] = P.cur_bytes;
  P.cur_bytes = 0;
}
// k-blocks [kb_lo, kb_hi) of matrix `mat`: as many k-blocks per image as fit the open stage (at most 8).  The k-blocks from
// kb_lo + fresh_from on depend on the layer before (act_ready[dep]) and start an image of their own.
constexpr void umma_prog_add(UmmaProgC &P, int mat, int N, int K, int kb_lo, int kb_hi, int tile_step, int n_tiles, int b_buf, int b_kb0,
                             int d_blk, int d_tile_stride, int zero_first, int dep, int fresh_from, int commit) {
  if (dep >= 0 && fresh_from > 0 && kb_lo + fresh_from < kb_hi) {
    umma_prog_add(P, mat, N, K, kb_lo, kb_lo + fresh_from, tile_step, n_tiles, b_buf, b_kb0, d_blk, d_tile_stride, zero_first, -1, 0, -1);
    umma_prog_add(P, mat, N, K, kb_lo + fresh_from, kb_hi, tile_step, n_tiles, b_buf, b_kb0 + fresh_from, d_blk, d_tile_stride, 0, dep, 0, commit);
    return;
  }
  if (n_tiles < 1 || n_tiles > 3 || N * 32 > UMMA_I8_STAGE_BYTES) { P.ok = false; return; }
  for (int kb = kb_lo; kb < kb_hi;) {
    const int space = UMMA_I8_STAGE_BYTES - P.cur_bytes;
    int nk = kb_hi - kb;
    if (nk > 8) nk = 8;
    if (nk > space / (N * 32)) nk = space / (N * 32);
    if (nk < 1) { umma_prog_close(P); continue; }
    if (P.n >= UMMA_MAX_RECS) { P.ok = false; return; }
    UmmaRecC &r = P.r[P.n++];
    r.mat = mat; r.kb = kb; r.n_rows = N; r.K = K;
    r.a_off16 = P.cur_bytes / 16; r.tile_step = tile_step; r.b_kb = b_kb0 + (kb - kb_lo); r.nk = nk; r.n_tiles = n_tiles;
    r.b_buf = b_buf; r.d_blk = d_blk; r.d_tile_stride = d_tile_stride;
    r.flags = (P.cur_bytes == 0 ? UR_STAGE_FIRST : 0) | ((zero_first && kb == kb_lo) ? UR_ZERO_FIRST : 0);
    r.dep = kb == kb_lo ? dep : -1;
    r.commit = kb + nk == kb_hi ? commit : -1;
    P.cur_bytes += N * nk * 32;
    kb += nk;
  }
}

// Accumulator column blocks (x NS columns): encoder GRU l at 4 (l & 1) + {0: [z;r] input, 1: [z;r] recurrent, 2: [n;-] input,
// 3: [n;-] recurrent}, conv l at 8 + (l & 1); decoder GRU l at 6 (l & 1) + {0..5: z, r, n tiles x (input, recurrent)}, GLU l at
// 12 + (l & 1), conv l at 14 + (l & 1).  act_ready / acc_full indices: encoder 2 l (GRU), 2 l + 1 (conv); decoder 3 l, 3 l + 1, 3 l + 2.
constexpr UmmaProgC umma_make_enc_prog() {
  UmmaProgC P{}; P.ok = true;
  const int dil[5] = {1, 2, 2, 2, 2};
  int off = 64;
  for (int l = 0; l < 5; l++) {
    const int gs = (l & 1) * 4, cs = 8 + (l & 1);
    // GRU: tiles [z; r] (rows 0..127) and [n; -] (rows 128..); fresh input = conv l-1's 96 outputs (3 k-blocks)
    umma_prog_add(P, UM_ENC_GRU_IN + l, 192, off, 0, off / 32, 128, 2, UB_CUR, 0, gs, 2, 1, l ? 2 * (l - 1) + 1 : -1, l ? off / 32 - 3 : 0, -1);
    umma_prog_add(P, UM_ENC_GRU_REC + l, 192, 64, 0, 2, 128, 2, UB_PREV1, off / 32, gs + 1, 2, 1, -1, 0, 2 * l);
    off += ENC_GRU;
    // conv (k = 2): tap 0 = concat prefix of step t - dilation, tap 1 = current prefix whose last 2 k-blocks are GRU l's outputs
    umma_prog_add(P, UM_ENC_CONV + l, 96, 2 * off, 0, off / 32, 0, 1, dil[l] == 1 ? UB_PREV1 : UB_PREV2, 0, cs, 0, 1, -1, 0, -1);
    umma_prog_add(P, UM_ENC_CONV + l, 96, 2 * off, off / 32, 2 * off / 32, 0, 1, UB_CUR, 0, cs, 0, 0, 2 * l, off / 32 - 2, 2 * l + 1);
    off += ENC_CONV;
  }
  umma_prog_close(P);
  return P;
}
constexpr UmmaProgC umma_make_dec_prog() {
  UmmaProgC P{}; P.ok = true;
  int off = 96;
  for (int l = 0; l < 5; l++) {
    const int gs = (l & 1) * 6, us = 12 + (l & 1), cs = 14 + (l & 1);
    // GRU: three overlapping tiles starting at rows 0, 96, 192 (z, r, n of unit u in TMEM lane u); fresh input = conv l-1 (1 k-block)
    umma_prog_add(P, UM_DEC_GRU_IN + l, 288, off, 0, off / 32, DEC_GRU, 3, UB_CUR, 0, gs, 2, 1, l ? 3 * (l - 1) + 2 : -1, l ? off / 32 - 1 : 0, -1);
    umma_prog_add(P, UM_DEC_GRU_REC + l, 288, 96, 0, 3, DEC_GRU, 3, UB_HQ_RD, 3 * l, gs + 1, 2, 1, -1, 0, 3 * l);
    umma_prog_add(P, UM_DEC_GLU + l, 96, 96, 0, 3, 0, 1, UB_HQ_WR, 3 * l, us, 0, 1, 3 * l, 0, 3 * l + 1);
    off += DEC_GRU;
    umma_prog_add(P, UM_DEC_CONV + l, 32, 2 * off, 0, off / 32, 0, 1, UB_PREV1, 0, cs, 0, 1, -1, 0, -1);
    umma_prog_add(P, UM_DEC_CONV + l, 32, 2 * off, off / 32, 2 * off / 32, 0, 1, UB_CUR, 0, cs, 0, 0, 3 * l + 1, off / 32 - 3, 3 * l + 2);
    off += DEC_CONV;
  }
  umma_prog_close(P);
  return P;
}

static constexpr UmmaProgC kUmmaEncProg = umma_make_enc_prog();
static constexpr UmmaProgC kUmmaDecProg = umma_make_dec_prog();
static_assert(kUmmaEncProg.ok && kUmmaDecProg.ok, "MMA program does not fit its tables");
