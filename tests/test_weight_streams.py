"""CPU check of the codec's per-step weight streams as the library's host code builds them (rade_b200_debug_codec_stream, no
device involved): every chunk is walked in the kernels' consumption order and the int8 / float matrices are reconstructed and
compared with the RDW arrays — for today's mma.sync fragment order and for the tcgen05 operand layout of the round-2 plan."""
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


@pytest.mark.parametrize("umma", [0, 1])
def test_encoder_and_decoder_streams_hold_every_weight(lib, umma):
    A = rdw.read_rdw(rdw.default_weights_path())
    def i8(w, name, N, K, lo, hi, step, tiles):
        got = w.i8_umma(N, K, lo, hi, step, tiles) if umma else w.i8_fragments(N, K, lo, hi)
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
