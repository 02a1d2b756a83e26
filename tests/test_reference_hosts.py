"""The reference's own C host programs (src/radae_tx.c, src/radae_rx.c) must build against include/rade_api.h and link
against libradae_b200.so UNCHANGED.  Where the reference tree is absent (GPU box) an equivalent minimal host written
against the same header stands in (tests/hosts/), so the C call sequence is exercised on the device either way."""
import os, subprocess, sys
import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
LIBDIR = os.path.join(REPO, "radae_b200", "lib")
HOSTS = os.path.join(REPO, "tests", "hosts")


def build_host(name, outdir):
    from radae_b200.build import build
    build()
    src = os.path.join(REF_SRC, name + ".c")
    if not os.path.exists(src):
        src = os.path.join(HOSTS, name + ".c")
    exe = os.path.join(outdir, name)
    subprocess.run(["gcc", "-O2", "-I", os.path.join(REPO, "include"), src, "-L", LIBDIR, "-lradae_b200",
                    f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True)
    return exe, src


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present")
def test_reference_hosts_compile_and_link_unchanged(tmp_path):
    for name in ("radae_tx", "radae_rx"):
        exe, src = build_host(name, str(tmp_path))
        assert src.startswith(REF_SRC)
        out = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
        assert "rade_open" in out and ("rade_tx" in out or "rade_rx" in out)


def test_stand_in_hosts_compile(tmp_path):
    for name in ("radae_tx", "radae_rx", "multi_loopback"):
        src = os.path.join(HOSTS, name + ".c")
        exe = os.path.join(str(tmp_path), name)
        from radae_b200.build import build
        build()
        subprocess.run(["gcc", "-O2", "-I", os.path.join(REPO, "include"), src, "-L", LIBDIR, "-lradae_b200",
                        f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True)


@pytest.mark.gpu
def test_c_hosts_pipe_tx_into_rx_on_device(tmp_path, golden):
    """features.f32 | radae_tx | radae_rx > features_out.f32 — compare the C-host path with the Python mirror"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.core import synth_features
    tx_exe, _ = build_host("radae_tx", str(tmp_path))
    rx_exe, _ = build_host("radae_rx", str(tmp_path))
    F = 14
    feats = synth_features(1, 12 * F, seed=9)[0].astype(np.float32)
    fin = tmp_path / "features_in.f32"; feats.tofile(fin)
    txf = tmp_path / "tx.f32"; fout = tmp_path / "features_out.f32"
    subprocess.run(f"{tx_exe} < {fin} > {txf}", shell=True, check=True, cwd=tmp_path, stderr=subprocess.DEVNULL)
    tx = np.fromfile(txf, np.complex64)
    assert len(tx) == 960 * F + 1152                      # F modem frames + the EOO frame (src/radae_tx.c:49-52)
    (np.concatenate([tx, np.zeros(2000, np.complex64)])).tofile(txf)
    subprocess.run(f"{rx_exe} < {txf} > {fout}", shell=True, check=True, cwd=tmp_path, stderr=subprocess.DEVNULL)
    got = np.fromfile(fout, np.float32).reshape(-1, 432)
    # same thing through the Python mirror of the reference classes
    from radae_b200 import radae_tx, radae_rx
    t = radae_tx(); out = np.zeros(960, np.complex64); sig = []
    for f in range(F):
        t.do_radae_tx(feats[12 * f:12 * (f + 1)].reshape(-1), out); sig.append(out.copy())
    eoo = np.zeros(1152, np.complex64); t.do_eoo(eoo); sig.append(eoo); sig.append(np.zeros(2000, np.complex64))
    sig = np.concatenate(sig)
    assert np.array_equal(sig[:len(tx)], tx)
    r = radae_rx(v=0, reset_decoder_on_sync=False); o = 0; ref = []; fl = np.zeros(432, np.float32)      # the C host runs rade_rx with the C decoder: no reset
    while o + r.get_nin() <= len(sig):
        n = r.get_nin(); ret = r.do_radae_rx(sig[o:o + n], fl); o += n
        if ret & 1: ref.append(fl.copy())
    assert len(got) == len(ref) and len(ref) >= F - 7
    assert np.array_equal(got, np.array(ref))
    assert os.path.getsize(tmp_path / "eoo_rx.f32") == 180 * 4          # one EOO frame detected (src/radae_rx.c:49-51)
