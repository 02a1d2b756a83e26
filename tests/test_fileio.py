"""CPU: file-format helpers against the reference scripts' semantics (int16tof32.py:40-52, f32toint16.py)."""
import numpy as np
from radae_b200 import fileio


def test_int16_f32_round_trips_like_the_reference_scripts():
    x = np.array([0, 1, -1, 32767, -32768, 1234], np.int16)
    y = fileio.int16_to_f32(x)
    assert y.dtype == np.float32 and np.array_equal(y, x.astype(np.float32))
    z = fileio.int16_to_f32(x, zeropad=True)
    assert len(z) == 12 and np.array_equal(z[::2], y) and not z[1::2].any()
    iq = fileio.wav_to_iq(x)
    assert iq.dtype == np.complex64 and np.array_equal(iq.real, y) and not iq.imag.any()
    f = np.array([0.5, -0.25, 1.0, -1.0, 0.99999, 0.0], np.float32)
    q = fileio.f32_to_int16(f)
    assert np.array_equal(q, (f * np.float32(32767.0)).astype(np.int16))
    assert np.array_equal(fileio.f32_to_int16(f, real=True), q[::2])
    assert np.array_equal(fileio.f32_to_int16(f, scale=100.0), np.array([50, -25, 100, -100, 99, 0], np.int16))


def test_feature_and_iq_files(tmp_path):
    rng = np.random.default_rng(1)
    feats = rng.standard_normal((5, 432)).astype(np.float32)
    p = str(tmp_path / "features.f32")
    np.concatenate([feats.reshape(-1), np.ones(100, np.float32)]).tofile(p)       # trailing partial frame
    got = fileio.read_features(p)
    assert got.shape == (5, 432) and np.array_equal(got, feats)
    assert fileio.used_features(got).shape == (5, 12, 20)
    assert np.array_equal(fileio.used_features(got)[2, 3], feats[2].reshape(12, 36)[3, :20])
    fileio.write_features(p, feats); assert np.array_equal(fileio.read_features(p), feats)
    x = (rng.standard_normal(960) + 1j * rng.standard_normal(960)).astype(np.complex64)
    q = str(tmp_path / "rx.f32"); fileio.write_iq(q, x)
    assert np.array_equal(fileio.read_iq(q), x)
    assert np.array_equal(np.fromfile(q, np.float32)[::2], x.real)               # interleaved I, Q float32
