"""File formats either side of the C ABI — host-side helpers, numpy only.

  features .f32   36 float32 per 10 ms vector (20 used + 16 ignored), 12 vectors = 432 floats per modem frame
                  (radae_txe.py:49-50, src/radae_tx.c:30-40, src/radae_rx.c:30-50)
  IQ .f32 / .c64  interleaved float32 I,Q = complex64, 960 samples per modem frame at 8 kHz (rade_tx output, rade_rx input)
  latents z .f32  80 float32 per 40 ms step (inference.py --write_latent)
  int16 <-> f32   int16tof32.py:40-52 (--zeropad: real int16 -> complex with Q = 0) and f32toint16.py (--scale, --real)
"""
import numpy as np

NB_TOTAL_FEATURES = 36
NUM_USED_FEATURES = 20
FRAME_FEATURES = 12 * NB_TOTAL_FEATURES           # one modem frame at the API (rade_n_features_in_out)


def read_features(path, whole_frames=True):
    """-> [n_frames, 432] float32 (API layout) — trailing partial frame dropped like src/radae_tx.c's fread loop"""
    x = np.fromfile(path, dtype=np.float32)
    n = len(x) // FRAME_FEATURES
    return x[:n * FRAME_FEATURES].reshape(n, FRAME_FEATURES) if whole_frames else x.reshape(-1, NB_TOTAL_FEATURES)


def write_features(path, feats):
    np.ascontiguousarray(feats, np.float32).tofile(path)


def used_features(feats):
    """API layout [..., 432] -> [..., 12, 20]: the part of the feature vectors the encoder reads / the decoder writes"""
    f = np.asarray(feats, np.float32)
    return f.reshape(f.shape[:-1] + (12, NB_TOTAL_FEATURES))[..., :NUM_USED_FEATURES]


def read_iq(path):
    return np.fromfile(path, dtype=np.complex64)


def write_iq(path, x):
    np.ascontiguousarray(x, np.complex64).tofile(path)


def int16_to_f32(x, zeropad=False):
    """int16tof32.py: samples keep their integer scale; --zeropad interleaves Q = 0 so a real file becomes IQ"""
    y = np.asarray(x, np.int16).astype(np.float32)
    if zeropad:
        z = np.zeros(2 * len(y), np.float32); z[::2] = y; y = z
    return y


def f32_to_int16(x, scale=32767.0, real=False):
    """f32toint16.py: multiply, truncate toward zero (numpy astype), optionally keep only the I channel of an IQ stream"""
    y = (np.asarray(x, np.float32) * np.float32(scale)).astype(np.int16)
    return y[::2] if real else y


def wav_to_iq(samples_int16):
    """the receive chain of the reference's off-air tests: 8 kHz s16 mono -> (x, 0) complex (int16tof32.py --zeropad)"""
    return int16_to_f32(samples_int16, zeropad=True).view(np.complex64)
