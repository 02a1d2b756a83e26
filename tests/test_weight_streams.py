"""CPU checks of the codec's per-step weight streams as the library's host code builds them (rade_b200_debug_codec_stream /
rade_b200_debug_codec_program, no device involved).  mma.sync kernels: every chunk is walked in consumption order and the
int8 / float matrices are reconstructed.  tcgen05 kernels: the issuer's MMA program is EXECUTED on the CPU (numpy emulation of
the descriptor addressing, the ring stages incl. stale bytes, the accumulator column blocks) against random int8 activations and
compared with plain matrix products; the float stream is walked in the float warps' order."""
import ctypes as C
import numpy as np
import pytest
from radae_b200 import _capi, rdw

STAGE = 32768


def get_stream(lib, which, umma):
    n_chunks, n_pro = C.c_int(0), C.c_int(0)
    n = lib.rade_b200_debug_codec_stream(which, umma, None, 0, None, 0, C.byref(n_chunks), C.byref(n_pro))
    assert n > 0
    buf = np.zeros(n, np.uint8); ch = np.zeros(2 * n_chunks.value, np.uint32)
    assert lib.rade_b200_debug_codec_stream(which, umma, buf.ctypes.data, n, ch.ctypes.data, n_chunks.value, C.byref(n_chunks), C.byref(n_pro)) == n
    return buf, ch.reshape(-1, 2), n_pro.value


class Walker:
    def __init__(self, buf, chunks):
        self.buf, self.chunks, self.i = buf, chunks, 0
    def next(self):
        off, n = self.chunks[self.i]; self.i += 1
        assert n % 16 == 0 and 0 < n <= STAGE
        return self.buf[off:off + n]
    def f32_rows(self, nout, noutp, nrows):
        rpc = (STAGE // (noutp * 4)) & ~3
        rows = []
        for r0 in range(0, nrows, rpc):
            n = min(rpc, nrows - r0)
            c = self.next().view(np.float32).reshape(n, noutp)
            assert not c[:, nout:].any()
            rows.append(c[:, :nout])
        return np.concatenate(rows)
    def i8_fragments(self, N, K, kb_lo, kb_hi):
        ntl = N // 8; kbc = max(STAGE // (ntl * 256), 1)
        W = np.zeros((N, (kb_hi - kb_lo) * 32), np.int8)
        for kb0 in range(kb_lo, kb_hi, kbc):
            nk = min(kbc, kb_hi - kb0)
            t = self.next().view(np.int8).reshape(nk, ntl, 32, 2, 4)         # [kb][nt][lane]{b0, b1}, 4 bytes each
            for lane in range(32):
                g, tig = lane >> 2, lane & 3
                for kb in range(nk):
                    col = (kb0 - kb_lo + kb) * 32 + tig * 4
                    W[np.arange(ntl) * 8 + g, col:col + 4] = t[kb, :, lane, 0]
                    W[np.arange(ntl) * 8 + g, col + 16:col + 20] = t[kb, :, lane, 1]
        return W
    def i8_umma(self, N, K, kb_lo, kb_hi, tile_step, n_tiles):
        span = max(tile_step * (n_tiles - 1) + 128, N); nk_max = STAGE // (span * 32)
        assert nk_max >= 1
        W = np.zeros((N, (kb_hi - kb_lo) * 32), np.int8)
        r, b = np.meshgrid(np.arange(N), np.arange(nk_max * 32), indexing="ij")
        for kb0 in range(kb_lo, kb_hi, nk_max):
            nk = min(nk_max, kb_hi - kb0); kbytes = nk * 32
            c = self.next()
            assert c.size == N * kbytes and span * kbytes <= STAGE             # the tiles' reads stay inside the stage
            rr, bb = r[:, :kbytes], b[:, :kbytes]
            W[:, (kb0 - kb_lo) * 32:(kb0 - kb_lo) * 32 + kbytes] = c[(rr // 8) * (kbytes * 8) + (bb // 16) * 128 + (rr % 8) * 16 + bb % 16].view(np.int8)
        return W


@pytest.fixture(scope="module")
def lib():
    return _capi.lib()


def test_mma_sync_streams_hold_every_weight(lib):
    umma = 0
    A = rdw.read_rdw(rdw.default_weights_path())
    def i8(w, name, N, K, lo, hi, step, tiles):
        got = w.i8_fragments(N, K, lo, hi)
        assert np.array_equal(got, A[name + ".w8"][:, lo * 32:hi * 32]), name
    # ---- encoder
    buf, chunks, n_pro = get_stream(lib, 0, umma)
    w = Walker(buf, chunks)
    assert np.array_equal(w.f32_rows(64, 64, 84), A["enc_dense1.wf"]) and w.i == n_pro
    assert np.array_equal(w.f32_rows(80, 80, 64), A["enc_zdense.wf"][:64])
    off = 64
    for l in range(1, 6):
        i8(w, f"enc_gru{l}_input", 192, off, 0, off // 32, 64, 3)
        i8(w, f"enc_gru{l}_recurrent", 192, 64, 0, 2, 64, 3)
        assert np.array_equal(w.f32_rows(80, 80, 64), A["enc_zdense.wf"][off:off + 64]); off += 64
        if l == 5:
            assert np.array_equal(w.f32_rows(64, 64, 84), A["enc_dense1.wf"])
        i8(w, f"enc_conv{l}", 96, 2 * off, 0, off // 32, 0, 1)
        i8(w, f"enc_conv{l}", 96, 2 * off, off // 32, 2 * off // 32, 0, 1)
        assert np.array_equal(w.f32_rows(80, 80, 96), A["enc_zdense.wf"][off:off + 96]); off += 96
    assert w.i == len(chunks) and off == 864
    # ---- decoder
    buf, chunks, n_pro = get_stream(lib, 1, umma)
    w = Walker(buf, chunks)
    assert np.array_equal(w.f32_rows(96, 96, 80), A["dec_dense1.wf"]) and w.i == n_pro
    assert np.array_equal(w.f32_rows(84, 96, 96), A["dec_output.wf"][:96])
    off = 96
    for l in range(1, 6):
        i8(w, f"dec_gru{l}_input", 288, off, 0, off // 32, 96, 3)
        i8(w, f"dec_gru{l}_recurrent", 288, 96, 0, 3, 96, 3)
        i8(w, f"dec_glu{l}", 96, 96, 0, 3, 0, 1)
        assert np.array_equal(w.f32_rows(84, 96, 96), A["dec_output.wf"][off:off + 96]); off += 96
        if l == 5:
            assert np.array_equal(w.f32_rows(96, 96, 80), A["dec_dense1.wf"])
        i8(w, f"dec_conv{l}", 32, 2 * off, 0, off // 32, 0, 1)
        i8(w, f"dec_conv{l}", 32, 2 * off, off // 32, 2 * off // 32, 0, 1)
        assert np.array_equal(w.f32_rows(84, 96, 32), A["dec_output.wf"][off:off + 32]); off += 32
    assert w.i == len(chunks) and off == 736


# ------------------------------------------------------------------------------------------------ tcgen05 formulation
UB_CUR, UB_PREV1, UB_PREV2, UB_HQ_RD, UB_HQ_WR = range(5)
UR_STAGE_FIRST, UR_STAGE_LAST, UR_ZERO_FIRST = 1, 2, 4
I8_STAGE, F32_STAGE, I8_NST = 40960, 22528, 3
RECF = ["a_off16", "tile_step", "b_kb", "nk", "n_tiles", "b_buf", "flags", "d_blk", "d_tile_stride", "dep", "commit", "p0", "p1"]


def get_program(lib, which):
    n = lib.rade_b200_debug_codec_program(which, None, 0)
    assert 0 < n <= 120
    a = np.zeros((n, 13), np.int32)
    assert lib.rade_b200_debug_codec_program(which, a.ctypes.data, n) == n
    return [dict(zip(RECF, map(int, r))) for r in a]


def b_layout(x):
    """[NS = 8 streams][K] int8 -> the B-operand image: offset(n, k) = (k / 16) * 128 + n * 16 + k % 16"""
    K = x.shape[1]
    img = np.zeros(K * 8, np.int8)
    n, k = np.meshgrid(np.arange(8), np.arange(K), indexing="ij")
    img[(k // 16) * 128 + n * 16 + k % 16] = x
    return img


def run_program(recs, buf, chunks, bufs, n_blocks, rng, stop_after_commit=None):
    """emulate issuer_step on the CPU: the ring (stages keep stale bytes, an image may read past its stage into the next one),
    descriptor addressing, accumulator column blocks.  Returns the blocks [n_blocks][128 lanes][8 streams] and the event list."""
    ring = rng.integers(-128, 128, I8_NST * I8_STAGE + 65536).astype(np.int8)          # + what follows the ring in shared memory
    acc = rng.integers(-1000, 1000, (n_blocks, 128, 8)).astype(np.int64)                # TMEM starts with garbage too
    ci, stage, events, base = 0, -1, [], None
    lane = np.arange(128)
    for rec in recs:
        if rec["flags"] & UR_STAGE_FIRST:
            assert base is None
            stage = (stage + 1) % I8_NST; base = stage * I8_STAGE
            off, n = chunks[ci]; ci += 1
            assert 0 < n <= I8_STAGE and n % 16 == 0
            ring[base:base + n] = buf[off:off + n].view(np.int8)
        assert base is not None
        if rec["dep"] >= 0:
            events.append(("wait", rec["dep"]))
        nk, tiles = rec["nk"], rec["n_tiles"]
        assert 1 <= nk <= 8 and 1 <= tiles <= 3
        B = bufs[rec["b_buf"]]
        asbo = nk * 256
        a0 = base + rec["a_off16"] * 16
        for k in range(nk):
            kb = rec["b_kb"] + k
            bt = B[kb * 256:(kb + 1) * 256].reshape(2, 8, 16).transpose(1, 0, 2).reshape(8, 32).astype(np.int64)       # [stream][32 k]
            for g in range(tiles):
                r = lane + g * rec["tile_step"]
                kk = np.arange(32)
                addr = a0 + (r[:, None] // 8) * asbo + k * 256 + (kk[None, :] // 16) * 128 + (r[:, None] % 8) * 16 + kk[None, :] % 16
                assert addr.max() < ring.size, "tile reads past shared memory"
                d = ring[addr].astype(np.int64) @ bt.T
                blk = rec["d_blk"] + g * rec["d_tile_stride"]
                acc[blk] = d if ((rec["flags"] & UR_ZERO_FIRST) and k == 0) else acc[blk] + d
        if rec["flags"] & UR_STAGE_LAST:
            base = None
        if rec["commit"] >= 0:
            events.append(("commit", rec["commit"]))
            if stop_after_commit == rec["commit"]:
                return acc, events
    assert ci == len(chunks) and base is None
    return acc, events


def test_umma_encoder_program_computes_every_layer(lib):
    A = rdw.read_rdw(rdw.default_weights_path())
    rng = np.random.default_rng(5)
    buf, chunks, _ = get_stream(lib, 0, 1)
    recs = get_program(lib, 0)
    x = {b: rng.integers(-127, 128, (8, 864)).astype(np.int8) for b in (UB_CUR, UB_PREV1, UB_PREV2)}
    acc, ev = run_program(recs, buf, chunks, {b: b_layout(v) for b, v in x.items()}, 10, rng)
    off = 64; dil = [1, 2, 2, 2, 2]
    for l in range(5):
        gs, cs = (l & 1) * 4, 8 + (l & 1)
        Wi = A[f"enc_gru{l + 1}_input.w8"].astype(np.int64); Wr = A[f"enc_gru{l + 1}_recurrent.w8"].astype(np.int64)
        gi = Wi @ x[UB_CUR][:, :off].astype(np.int64).T; gr = Wr @ x[UB_PREV1][:, off:off + 64].astype(np.int64).T     # [192][8]
        off += 64
        Wc = A[f"enc_conv{l + 1}.w8"].astype(np.int64)
        old = x[UB_PREV1] if dil[l] == 1 else x[UB_PREV2]
        cv = Wc[:, :off] @ old[:, :off].astype(np.int64).T + Wc[:, off:] @ x[UB_CUR][:, :off].astype(np.int64).T
        off += 96
        if l >= 3:      # slots are reused by layer parity: the final contents belong to layers 3 (odd) and 4 (even)
            assert np.array_equal(acc[gs][:128], gi[:128]) and np.array_equal(acc[gs + 1][:128], gr[:128])           # [z; r]
            assert np.array_equal(acc[gs + 2][:64], gi[128:]) and np.array_equal(acc[gs + 3][:64], gr[128:])          # [n; -]
            assert np.array_equal(acc[cs][:96], cv)
    # hand-over protocol: every accumulator is committed once, in layer order; the fresh k-blocks wait for the layer before
    assert [e[1] for e in ev if e[0] == "commit"] == list(range(10))
    assert [e[1] for e in ev if e[0] == "wait"] == [0, 1, 2, 3, 4, 5, 6, 7, 8]
    for i, e in enumerate(ev):
        if e[0] == "wait":
            assert ("commit", e[1]) in ev[:i]


def test_umma_encoder_program_slot_reuse_is_ordered(lib):
    """the accumulators of layers l and l + 2 share TMEM columns: stopping the program after layer 1's conv must leave the sums of
    layers 0 and 1 in place"""
    A = rdw.read_rdw(rdw.default_weights_path())
    rng = np.random.default_rng(6)
    buf, chunks, _ = get_stream(lib, 0, 1)
    recs = get_program(lib, 0)
    x = {b: rng.integers(-127, 128, (8, 864)).astype(np.int8) for b in (UB_CUR, UB_PREV1, UB_PREV2)}
    acc, _ = run_program(recs, buf, chunks, {b: b_layout(v) for b, v in x.items()}, 10, rng, stop_after_commit=3)
    Wi = A["enc_gru1_input.w8"].astype(np.int64)
    assert np.array_equal(acc[0][:128], (Wi @ x[UB_CUR][:, :64].astype(np.int64).T)[:128])
    Wc = A["enc_conv2.w8"].astype(np.int64); off = 288
    cv = Wc[:, :off] @ x[UB_PREV2][:, :off].astype(np.int64).T + Wc[:, off:] @ x[UB_CUR][:, :off].astype(np.int64).T
    assert np.array_equal(acc[9][:96], cv)


def test_umma_decoder_program_computes_every_layer(lib):
    A = rdw.read_rdw(rdw.default_weights_path())
    rng = np.random.default_rng(7)
    buf, chunks, _ = get_stream(lib, 1, 1)
    recs = get_program(lib, 1)
    x = {b: rng.integers(-127, 128, (8, 736)).astype(np.int8) for b in (UB_CUR, UB_PREV1)}
    h = {b: rng.integers(-127, 128, (8, 480)).astype(np.int8) for b in (UB_HQ_RD, UB_HQ_WR)}
    bufs = {b: b_layout(v) for b, v in {**x, **h}.items()}
    acc, ev = run_program(recs, buf, chunks, bufs, 16, rng)
    off = 96
    for l in range(5):
        gs, us, cs = (l & 1) * 6, 12 + (l & 1), 14 + (l & 1)
        Wi = A[f"dec_gru{l + 1}_input.w8"].astype(np.int64); Wr = A[f"dec_gru{l + 1}_recurrent.w8"].astype(np.int64)
        gi = Wi @ x[UB_CUR][:, :off].astype(np.int64).T; gr = Wr @ h[UB_HQ_RD][:, 96 * l:96 * l + 96].astype(np.int64).T  # [288][8]
        gl = A[f"dec_glu{l + 1}.w8"].astype(np.int64) @ h[UB_HQ_WR][:, 96 * l:96 * l + 96].astype(np.int64).T
        off += 96
        Wc = A[f"dec_conv{l + 1}.w8"].astype(np.int64)
        cv = Wc[:, :off] @ x[UB_PREV1][:, :off].astype(np.int64).T + Wc[:, off:] @ x[UB_CUR][:, :off].astype(np.int64).T
        off += 32
        if l >= 3:
            for g in range(3):        # gate g of unit u in lane u of tile g
                assert np.array_equal(acc[gs + 2 * g][:96], gi[96 * g:96 * g + 96]) and np.array_equal(acc[gs + 2 * g + 1][:96], gr[96 * g:96 * g + 96])
            assert np.array_equal(acc[us][:96], gl) and np.array_equal(acc[cs][:32], cv)
    assert [e[1] for e in ev if e[0] == "commit"] == list(range(15))
    assert [e[1] for e in ev if e[0] == "wait"] == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]


def test_umma_float_streams_in_the_float_warps_order(lib):
    A = rdw.read_rdw(rdw.default_weights_path())

    class Rows:                                   # the FloatCursor of core_codec_umma.cu
        def __init__(self, buf, chunks, n_pro):
            self.buf, self.chunks, self.n_pro, self.ci, self.p, self.left = buf, chunks, n_pro, 0, 0, 0
        def take(self, nout, noutp, nrows):
            out, rb = [], noutp * 4
            while nrows:
                if self.left == 0:
                    self.p, self.left = (int(v) for v in self.chunks[self.ci]); self.ci += 1
                    assert 0 < self.left <= F32_STAGE
                n = min(nrows, self.left // rb)
                assert n >= 4 and n % 4 == 0
                c = self.buf[self.p:self.p + n * rb].view(np.float32).reshape(n, noutp)
                assert not c[:, nout:].any()
                out.append(c[:, :nout]); self.p += n * rb; self.left -= n * rb; nrows -= n
            return np.concatenate(out)

    buf, chunks, n_pro = get_stream(lib, 0, 2)
    w = Rows(buf, chunks, n_pro)
    assert np.array_equal(w.take(64, 64, 84), A["enc_dense1.wf"]) and w.ci == n_pro and w.left == 0
    assert np.array_equal(w.take(80, 80, 64), A["enc_zdense.wf"][:64])
    assert np.array_equal(w.take(64, 64, 84), A["enc_dense1.wf"])
    off = 64
    for l in range(5):
        for n in (64, 96):
            assert np.array_equal(w.take(80, 80, n), A["enc_zdense.wf"][off:off + n]); off += n
    assert w.ci == len(chunks) and w.left == 0 and off == 864 and len(chunks) <= 24
    buf, chunks, n_pro = get_stream(lib, 1, 2)
    w = Rows(buf, chunks, n_pro)
    assert np.array_equal(w.take(96, 96, 80), A["dec_dense1.wf"]) and w.ci == n_pro and w.left == 0
    assert np.array_equal(w.take(84, 96, 96), A["dec_output.wf"][:96])
    assert np.array_equal(w.take(96, 96, 80), A["dec_dense1.wf"])
    off = 96
    for l in range(5):
        for n in (96, 32):
            assert np.array_equal(w.take(84, 96, n), A["dec_output.wf"][off:off + n]); off += n
    assert w.ci == len(chunks) and w.left == 0 and off == 736 and len(chunks) <= 24
