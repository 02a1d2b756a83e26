"""CPU: fading-file tooling: file format round trip exactly as inference.py:160-171 reads it, and the
statistics of the doppler_spread.m / multipath_samples.m restatement."""
import numpy as np
from radae_b200 import gfile


def test_g_file_round_trip_matches_reference_reader(tmp_path):
    rng = np.random.default_rng(0)
    n = 5000
    G1 = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    G2 = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    p = str(tmp_path / "g.f32")
    gfile.write_g(p, G1, G2, 0.625)
    # the reference's own reader, verbatim semantics (inference.py:161-166)
    G = np.reshape(np.fromfile(p, dtype=np.csingle), (1, -1, 2))
    mp_gain = np.real(G[:, 0, 0]); ref = mp_gain * G[:, 1:, :]
    g, got = gfile.read_g(p)
    assert g == 0.625 and np.array_equal(got, ref[0].astype(np.complex64))
    assert np.array_equal(got[:, 0], (0.625 * G1).astype(np.complex64))
    g2, short = gfile.read_g(p, n_samples=100)
    assert short.shape == (100, 2)
    try:
        gfile.read_g(p, n_samples=n + 1); assert False
    except ValueError:
        pass


def test_multipath_samples_statistics():
    G1, G2, hf_gain, d = gfile.multipath_samples("mpp", fs=8000, nseconds=120, seed=3)
    assert d == 16 and len(G1) == 960000
    assert abs(hf_gain ** 2 * (np.var(G1) + np.var(G2)) - 1) < 1e-6           # multipath_samples.m:27-31
    # Gaussian Doppler spectrum, sigma = spread/2 = 0.5 Hz: rms frequency of the process ~ 0.5 Hz
    x = G1[::80]                                                                # 100 Hz
    X = np.abs(np.fft.fft(x * np.hanning(len(x)))) ** 2
    f = np.fft.fftfreq(len(x), 1 / 100.0)
    rms = np.sqrt(np.sum(X * f ** 2) / np.sum(X))
    assert 0.3 < rms < 0.7, rms
    # slow process: adjacent 8 kHz samples are almost equal
    assert np.mean(np.abs(np.diff(G1)) ** 2) < 1e-6 * np.mean(np.abs(G1) ** 2)
    for ch, (spread, delay_s) in gfile.CHANNELS.items():
        assert gfile.multipath_samples(ch, nseconds=2, seed=1)[3] == round(delay_s * 8000)
