"""GPU: the C-level multi-device context (rade_b200_open_devices / rade_b200_open_multi) — streams split into
contiguous blocks, one context per device, calls routed by block with one host thread per device.  Results must be identical,
stream by stream, to one context holding all streams.  On a one-GPU box the device list names device 0 twice (two contexts on
one GPU exercise the same routing); with two or more GPUs visible the blocks really live on different devices."""
import numpy as np
import pytest
from gpu_util import need_gpu

pytestmark = pytest.mark.gpu


def test_multi_device_context_equals_single_context():
    torch = need_gpu()
    from radae_b200 import RadeBatch
    from radae_b200.batch import RadeMulti
    from oracle.core import synth_features
    S, F = 21, 10
    feats = np.ascontiguousarray(synth_features(S, 12 * F, seed=8).reshape(S, F, 432))
    ndev = torch.cuda.device_count()
    devices = [0, 1, 0] if ndev >= 2 else [0, 0, 0]
    m = RadeMulti(S, devices=devices)
    assert m.n_devices == 3 and m.blocks() == [(0, 7), (7, 7), (14, 7)]
    b = RadeBatch(S)
    rng = np.random.default_rng(1)
    for k in range(F):
        tx_m, tx_b = m.tx(feats[:, k]), b.tx(feats[:, k])
        assert np.array_equal(tx_m, tx_b), k
        nin_m, nin_b = m.nin(), b.nin()
        assert np.array_equal(nin_m, nin_b), k
        x = np.zeros((S, 1120), np.complex64)
        # loop the transmit frame back with a little noise (960 samples per frame: feed what each stream asks for, cyclically)
        for s in range(S):
            seg = np.resize(tx_b[s], int(nin_b[s]))
            x[s, :nin_b[s]] = seg + 0.05 * (rng.standard_normal(len(seg)) + 1j * rng.standard_normal(len(seg))).astype(np.complex64)
        fm, rm, em = m.rx(x); fb, rb, eb = b.rx(x)
        assert np.array_equal(rm, rb) and np.array_equal(fm, fb), k
    m.close(); b.close()


def test_open_multi_by_mask():
    need_gpu()
    from radae_b200.batch import RadeMulti
    m = RadeMulti(16, device_mask=0x1)
    assert m.n_devices == 1 and m.blocks() == [(0, 16)]
    m.close()
    with pytest.raises(RuntimeError):
        RadeMulti(16, device_mask=1 << 40)


def test_c_host_drives_the_multi_device_context(tmp_path):
    """a plain C program (tests/hosts/multi_loopback.c) opens rade_b200_open_multi / rade_b200_open_devices, loops every transmit
    frame back into the receiver and compares with one single-device context, bit for bit"""
    torch = need_gpu()
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "radae_b200", "lib")
    exe = str(tmp_path / "multi_loopback")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "hosts", "multi_loopback.c"),
                    "-L", lib, "-lradae_b200", "-Wl,-rpath," + lib, "-o", exe], check=True)
    devs = ["0", "1", "0"] if torch.cuda.device_count() >= 2 else ["0", "0", "0"]
    for args in (["13", "12"], ["13", "12"] + devs):               # every visible device by mask 0; an explicit device list
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.stdout, r.stderr)
        w = r.stdout.split()
        assert int(w[1]) >= 13 * 3 and int(w[3]) == 0, r.stdout
    assert int(w[5]) == 3
