"""GPU: channel simulator. Explicit form vs oracle.dsp.channel (<= 1e-5 rel RMS); generator form statistically."""
import numpy as np
import pytest
from gpu_util import need_gpu, relrms
from oracle import dsp as od

pytestmark = pytest.mark.gpu


def test_channel_apply_vs_oracle():
    torch = need_gpu()
    from radae_b200 import RadeBatch
    S, n = 5, 1920
    rng = np.random.default_rng(2)
    c = lambda: ((rng.standard_normal((S, n)) + 1j * rng.standard_normal((S, n))) / np.sqrt(2)).astype(np.complex64)
    tx, G1, G2, nz = c(), c(), c(), c()
    b = RadeBatch(S)
    dev = [torch.view_as_real(torch.tensor(a)).cuda() for a in (tx, G1, G2, nz)]
    out = torch.zeros((S, n, 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    b.channel_apply_dev(out.data_ptr(), dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(),
                        n, 16, 0.9, -11.0, 0.3, 0.5, 0.7)
    b.synchronize()
    got = torch.view_as_complex(out).cpu().numpy()
    ref = np.array([od.channel(tx[s], G1[s], G2[s], 16, 0.9, -11.0, 0.3, 0.5, nz[s], gain=0.7) for s in range(S)])
    assert relrms(got, ref) < 1e-5
    b.close()


def test_channel_generator_statistics():
    torch = need_gpu()
    from radae_b200 import RadeBatch
    S, F = 64, 50
    b = RadeBatch(S)
    EbNodB = 3.0
    b.channel_config(EbNodB=EbNodB, freq_offset_hz=0.0, doppler_spread_hz=1.0, delay_samples=16, gain=1.0, seed=9)
    tx = torch.zeros((S, 960, 2), dtype=torch.float32, device="cuda")
    tx[:, :, 0] = 1.0                                   # constant carrier: rx = G1 + G2 (delayed) + noise
    rx = torch.zeros_like(tx)
    torch.cuda.synchronize()
    acc = []
    for _ in range(F):
        b.channel_dev(rx.data_ptr(), tx.data_ptr()); b.synchronize()
        acc.append(torch.view_as_complex(rx).cpu().numpy().copy())
    y = np.concatenate(acc, axis=1)                     # [S, F*960]
    sigma = od.ebno_sigma(EbNodB)
    # high-pass part is the noise: first difference of a slowly varying fading process + white noise
    d = np.diff(y, axis=1)
    assert abs(np.mean(np.abs(d) ** 2) / (2 * sigma ** 2) - 1) < 0.05
    # total power = fading power (normalised to 1 on average over streams) + sigma^2
    p = np.mean(np.abs(y) ** 2) - sigma ** 2
    assert 0.6 < p < 1.4
    # AWGN only: exact carrier + noise
    b.channel_config(EbNodB=10.0, freq_offset_hz=25.0, doppler_spread_hz=0.0, seed=3)
    b.channel_dev(rx.data_ptr(), tx.data_ptr()); b.synchronize()
    y = torch.view_as_complex(rx).cpu().numpy()
    ph = np.exp(1j * 2 * np.pi * 25.0 / 8000 * np.arange(1, 961))
    res = y - ph[None, :]
    assert abs(np.mean(np.abs(res) ** 2) / od.ebno_sigma(10.0) ** 2 - 1) < 0.05
    b.close()


def test_channel_apply_host_with_fading_file(tmp_path):
    """the reference's g-file workflow (inference.py --g_file): write / read a fading file, run the explicit channel on host
    arrays, compare with the oracle"""
    need_gpu()
    from radae_b200 import RadeBatch, gfile
    S, n = 3, 2400
    rng = np.random.default_rng(5)
    G1, G2, hf_gain, d = gfile.multipath_samples("mpp", nseconds=1, seed=2)
    p = str(tmp_path / "g_mpp.f32"); gfile.write_g(p, G1, G2, hf_gain)
    mp_gain, G = gfile.read_g(p, n_samples=n)
    c = lambda: ((rng.standard_normal((S, n)) + 1j * rng.standard_normal((S, n))) / np.sqrt(2)).astype(np.complex64)
    tx, nz = c(), c()
    g1 = np.tile(G[:, 0], (S, 1)); g2 = np.tile(G[:, 1], (S, 1))
    b = RadeBatch(S)
    sigma = od.ebno_sigma(3.0)
    got = b.channel_apply(tx, g1, g2, nz, delay=d, mp_gain=1.0, freq_offset_hz=-11.0, sigma=sigma)
    ref = np.array([od.channel(tx[s], g1[s], g2[s], d, 1.0, -11.0, 0.0, sigma, nz[s]) for s in range(S)])
    assert relrms(got, ref) < 1e-5
    b.close()


def test_channel_apply_vs_reference_forward_fixture(golden):
    """the CUDA explicit channel (rade_b200_channel_apply_drift through the host-pointer C ABI) against the rate-Fs channel of the
    reference's RADAE.forward itself (tests/golden/channel.npz, tools/make_golden_channel.py): multipath + power normalisation,
    frequency offset with drift, phase offset, AWGN with the reference's noise draw, gain; <= 1e-5 relative RMS"""
    need_gpu()
    from radae_b200 import RadeBatch
    g = golden("channel")
    b = RadeBatch(1)
    for name in g["names"]:
        EbNodB, f0, df_dt, ph0, gain, d = g[f"{name}_params"]
        tx, G, noise = g[f"{name}_tx"], g[f"{name}_G"], g[f"{name}_noise"]
        mpg = od.mp_gain_of(tx, G[:, 0], G[:, 1], int(d))
        rx = b.channel_apply(tx[None], G[None, :, 0], G[None, :, 1], noise[None], delay=int(d), mp_gain=mpg, freq_offset_hz=float(f0),
                             phase0=float(ph0), sigma=od.ebno_sigma(EbNodB), gain=float(gain), df_dt=float(df_dt))[0]
        ref = g[f"{name}_rx"]
        err = np.sqrt(np.mean(np.abs(rx - ref) ** 2) / np.mean(np.abs(ref) ** 2))
        assert err < 1e-5, (name, err)
    b.close()
