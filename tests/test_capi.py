"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the public headers declare,
reports the reference's sizes, and builds the same DSP constant tables as the oracle (no compute, no GPU)."""
import os, re
import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from radae_b200.build import build
    build()                                   # nvcc cross-compiles without a GPU
    from radae_b200 import _capi
    return _capi.lib()


def declared_symbols():
    names = []
    for h in ("rade_api.h", "rade_b200.h"):
        src = open(os.path.join(REPO, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"RADE_EXPORT\s+[\w\s\*]+?\b(rade_\w+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from radae_b200 import _capi
    names = declared_symbols()
    assert len(names) >= 18 + 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature"


def test_reference_symbol_set_is_complete(lib):
    ref = ["rade_initialize", "rade_finalize", "rade_open", "rade_close", "rade_version", "rade_n_tx_out",
           "rade_n_tx_eoo_out", "rade_nin_max", "rade_n_features_in_out", "rade_n_eoo_bits", "rade_tx",
           "rade_tx_set_eoo_bits", "rade_tx_eoo", "rade_nin", "rade_rx", "rade_sync", "rade_freq_offset",
           "rade_snrdB_3k_est"]                 # the 18 RADE_EXPORTs of src/rade_api.h:71-129
    for n in ref:
        assert hasattr(lib, n)
    assert lib.rade_version() == 1
    # size getters do not touch the device: the reference's constants
    assert (lib.rade_n_tx_out(None), lib.rade_n_tx_eoo_out(None), lib.rade_nin_max(None),
            lib.rade_n_features_in_out(None), lib.rade_n_eoo_bits(None)) == (960, 1152, 1120, 432, 180)


def test_embedded_weights_are_the_rdw_file(lib):
    import ctypes
    from radae_b200 import rdw
    n = ctypes.c_size_t(0)
    p = lib.rade_b200_default_weights_blob(ctypes.byref(n))
    blob = ctypes.string_at(p, n.value)
    assert blob == open(rdw.default_weights_path(), "rb").read()
    arrays = rdw.read_rdw(blob)
    assert arrays["enc_gru1_input.w8"].shape == (192, 64) and arrays["dec_output.wf"].shape == (736, 84)


def test_dsp_tables_match_oracle(lib):
    from oracle import dsp as od
    c = od.consts()

    def cget(which, n):
        b = np.zeros(2 * n, np.float32)
        assert lib.rade_b200_debug_tables(which, b.ctypes.data, 2 * n) == n
        return b.view(np.complex64)

    for which, ref, tol in [(0, c.Winv, 1e-8), (1, c.Wfwd, 1e-7), (2, c.p, 1e-7), (3, c.pend, 1e-7), (4, c.p_w, 1e-7),
                            (5, c.Pmat, 1e-6), (6, c.eq_rot, 1e-7), (8, c.eoo_base, 1e-6)]:
        assert np.max(np.abs(cget(which, ref.size) - ref.ravel())) < tol, which
    f = od.ComplexBPF()
    assert np.array_equal(cget(7, 1152), f.phase_vec_exp[:1152])     # receive filter reads [0,1120), TX filter up to the EOO frame
    h = np.zeros(101, np.float32); lib.rade_b200_debug_tables(16, h.ctypes.data, 101)
    assert np.max(np.abs(h - f.h)) < 1e-8
    k = np.zeros(5, np.float32); assert lib.rade_b200_debug_tables(18, k.ctypes.data, 5) == 5
    assert k[1] == c.bpf_bw and k[2] == c.bpf_centre and k[3] == f.alpha and abs(k[0] - c.pilot_gain) < 1e-5


def test_coarse_grid_basis_spans_the_reference_grid(lib, golden):
    """rx_detect / rx_track evaluate the 40-point coarse frequency grid (radae/dsp.py:163-173, :204-205) in a rank-6 + 6 basis of
    the folded window.  (1) the float32 tables reproduce cos / sin of every grid point to 2e-7; (2) on a recorded signal the
    float32 low-rank evaluation is as close to the exact (float64) |Dt| as the reference's own float32 matrix product"""
    from oracle import dsp as od
    k = np.zeros(5, np.float32); assert lib.rade_b200_debug_tables(18, k.ctypes.data, 5) == 5
    basis = np.zeros(80 * 16, np.float32); assert lib.rade_b200_debug_tables(19, basis.ctypes.data, basis.size) == basis.size
    expand = np.zeros(21 * 12, np.float32); assert lib.rade_b200_debug_tables(20, expand.ctypes.data, expand.size) == expand.size
    basis = basis.reshape(80, 16); expand = expand.reshape(21, 12)
    bc, bs = basis[:, 0:6], basis[:, 8:14]
    ac, as_ = expand[:, 0:6], expand[:, 6:12]
    assert not basis[:, 6:8].any() and not basis[:, 14:16].any()
    m = np.arange(80) + 0.5
    w = 2 * np.pi * 2.5 * np.arange(21) / od.FS
    res = max(np.abs(ac.astype(np.float64) @ bc.astype(np.float64).T - np.cos(np.outer(w, m))).max(),
              np.abs(as_.astype(np.float64) @ bs.astype(np.float64).T - np.sin(np.outer(w, m))).max())
    assert res < 2e-7 and abs(res - k[4]) < 1e-8, (res, k[4])
    # a received window: |Dt1| over 960 timing offsets x 40 frequencies, three ways
    c = od.consts()
    rx = golden("rx_mpp_3dB")["rx_in"][5000:5000 + od.RXBUF].astype(np.complex64)
    n = np.arange(od.M)
    Y32 = (np.conj(rx[np.arange(960)[:, None] + n[None, :]]) * c.p.astype(np.complex64)[None, :]).astype(np.complex64)
    E = np.exp(2j * np.pi * np.outer(n, np.arange(-50, 50, 2.5)) / od.FS)           # the grid of radae/dsp.py:163-165
    truth = np.abs(Y32.astype(np.complex128) @ E)
    direct = np.abs(Y32 @ E.astype(np.complex64)).astype(np.float32)
    ye = (Y32[:, 80:] + Y32[:, 79::-1]).astype(np.complex64); yo = (Y32[:, 80:] - Y32[:, 79::-1]).astype(np.complex64)
    A = ((ye @ bc).astype(np.complex64) @ ac.T).astype(np.complex64); B = ((yo @ bs).astype(np.complex64) @ as_.T).astype(np.complex64)
    low = np.zeros((960, 40), np.float32)
    low[:, 20:40] = np.abs(A + 1j * B)[:, 0:20]; low[:, 19::-1] = np.abs(A - 1j * B)[:, 1:21]
    e_direct = np.sqrt(np.mean((direct - truth) ** 2)) / truth.mean(); e_low = np.sqrt(np.mean((low - truth) ** 2)) / truth.mean()
    assert e_low < 5e-7 and e_low < 1.5 * e_direct, (e_low, e_direct)
    assert np.unravel_index(low.argmax(), low.shape) == np.unravel_index(truth.argmax(), truth.shape)


def test_product_fails_loudly_without_a_gpu(lib):
    """no CPU fallback: on a machine without a CUDA device the batch constructor raises"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from radae_b200 import RadeBatch
    with pytest.raises(RuntimeError):
        RadeBatch(4)


def test_product_does_not_import_the_oracle():
    import subprocess, sys
    code = "import sys, radae_b200, radae_b200.batch, radae_b200.streaming, radae_b200.rdw; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'product imports oracle'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=REPO)
    # ... and no source file of the product imports, links, dlopens or executes anything under oracle/ (comments may cite it)
    import re
    forbidden = re.compile(r"^\s*(import|from)\s+oracle\b|oracle/_ref|librade_ref|libcore_oracle|core_oracle\.c|nnet_shim\.c\s*\"|"
                           r"CDLL\([^)]*oracle|dlopen\([^)]*oracle|-loracle", re.M)
    n_files = 0
    for root, _, files in os.walk(os.path.join(REPO, "radae_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".S")):
                n_files += 1
                m = forbidden.search(open(os.path.join(root, f)).read())
                assert m is None, (f, m.group(0))
    assert n_files > 15


def test_weight_blobs_are_validated_before_use(lib):
    """rade_open(model_file) / rade_b200_open(weights, len) take caller-supplied bytes — truncated or crafted
    containers must be rejected as a whole (host-only hook, the same parser rade_b200_open uses)"""
    import struct
    from radae_b200 import rdw
    chk = lambda b: lib.rade_b200_debug_check_weights(bytes(b), len(b))
    good = open(rdw.default_weights_path(), "rb").read()
    assert chk(good) == 0
    assert chk(open(rdw.model05_weights_path(), "rb").read()) == 0
    assert chk(good[:len(good) // 2]) == -1 and chk(good[:70]) == -1 and chk(b"RADEB200") == -1 and chk(b"") == -1
    n = struct.unpack_from("<I", good, 12)[0]
    for i in (0, n // 2, n - 1):                       # payload shorter than rows x cols, offset overflow, offset into the table
        e = 64 + 80 * i
        nm, dt, rows, cols, _, off, nbytes = struct.unpack_from("<48sIIIIQQ", good, e)
        for patch in ((off, nbytes - 16), (2 ** 64 - 8, nbytes), (len(good) - 4, nbytes), (64, nbytes)):
            bad = bytearray(good); struct.pack_into("<QQ", bad, e + 64, *patch)
            assert chk(bad) == -1, (i, patch)
        bad = bytearray(good); struct.pack_into("<I", bad, e + 52, rows + 1)
        assert chk(bad) == -1
    blob_path = "/root/reference/bin/model19_check3.bin"
    if os.path.exists(blob_path):                      # the reference's DNNw blob (not on the GPU box: CPU suite only)
        blob = open(blob_path, "rb").read()
        assert chk(blob) == 0
        assert chk(blob[:-64]) == -1 and chk(blob[:len(blob) // 3]) == -1
        # an index list that claims more blocks than the file holds, or points outside the matrix
        off = 0
        while off < len(blob):
            size, block = struct.unpack_from("<ii", blob, off + 12)
            name = blob[off + 20:off + 64].split(b"\0")[0]
            if name.endswith(b"_weights_idx"):
                bad = bytearray(blob); struct.pack_into("<i", bad, off + 64, 10 ** 6); assert chk(bad) == -1
                bad = bytearray(blob); struct.pack_into("<i", bad, off + 68, 10 ** 6); assert chk(bad) == -1
                bad = bytearray(blob); struct.pack_into("<i", bad, off + 12, size - 4); assert chk(bad) == -1
                break
            off += 64 + block
