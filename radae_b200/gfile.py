"""Rate-Fs fading files ("g files") for the channel simulator — SURVEY.md §8 f3.

Format (reference: inference.py:160-171, multipath_samples.m): a flat complex64 file; the FIRST pair holds hf_gain in its
first real part, then one (G1, G2) pair of complex Doppler-spread samples per 8 kHz sample.  The reference scales the pairs
by hf_gain when it loads them (`G = mp_gain*G[:,1:,:]`).

`doppler_spread` / `multipath_samples` restate doppler_spread.m:7-52 and multipath_samples.m:7-31 with scipy (Gaussian
Doppler spectrum shaped by a 100-tap fir2 filter at a low sample rate, linear interpolation up to Fs).  Octave's
`randn('seed', 1)` stream cannot be reproduced, so files made here are statistically — not sample — identical to the
reference's g_mpp.f32 etc.; anything read from an existing file is used exactly as the reference uses it.

Host-side tooling only (numpy / scipy): the device side is `rade_b200_channel_apply[_dev]`, which takes the G1, G2 arrays."""
import numpy as np

CHANNELS = {"mpg": (0.1, 0.5e-3), "mpp": (1.0, 2e-3), "mpd": (2.0, 4e-3)}      # Doppler spread Hz, path delay s


def read_g(path, n_samples=None):
    """-> (mp_gain, G[n, 2] complex64 already scaled by mp_gain, like inference.py:161-166)"""
    raw = np.fromfile(path, dtype=np.complex64).reshape(-1, 2)
    mp_gain = float(np.real(raw[0, 0]))
    G = (mp_gain * raw[1:]).astype(np.complex64)
    if n_samples is not None:
        if len(G) < n_samples:
            raise ValueError("Multipath Doppler spread file too short")          # inference.py:167-169
        G = G[:n_samples]
    return mp_gain, G


def write_g(path, G1, G2, hf_gain):
    G1 = np.asarray(G1, np.complex64); G2 = np.asarray(G2, np.complex64)
    out = np.empty((len(G1) + 1, 2), np.complex64)
    out[0] = (hf_gain, 0)
    out[1:, 0] = G1; out[1:, 1] = G2
    out.tofile(path)


def doppler_spread(spread_hz, fs, nsam, rng):
    """doppler_spread.m: complex Gaussian process with a Gaussian Doppler spectrum (sigma = spread/2) at rate fs"""
    from scipy import signal
    sigma = spread_hz / 2.0
    low_fs = float(np.ceil(10 * spread_hz))
    ntaps = 100
    M = fs / low_fs
    if M != np.floor(M):
        M = np.floor(M); low_fs = fs / M
    M = int(M)
    nsam_low = max(int(np.ceil(nsam / M)), 2)
    x = np.linspace(0.0, low_fs / 2, 51)                       # 0:lowFs/100:lowFs/2
    y = (1 / (sigma * np.sqrt(2 * np.pi))) * np.exp(-(x ** 2) / (2 * sigma * sigma))
    y[-1] = 0.0                                               # even-length (type II) FIR: zero at Nyquist (it is ~1e-22 anyway)
    b = signal.firwin2(ntaps, np.linspace(0.0, 1.0, 51), y)   # fir2(Ntaps-1, ...) = Ntaps coefficients
    noise = rng.standard_normal(nsam_low + ntaps) + 1j * rng.standard_normal(nsam_low + ntaps)
    low = signal.lfilter(b, 1, noise)[ntaps:]
    t_low = 1 + M * np.arange(nsam_low)                       # interp1((1:M:Nsam_low*M), ..., 1:Nsam, "extrap")
    t = np.arange(1, nsam + 1)
    re = np.interp(t, t_low, low.real); im = np.interp(t, t_low, low.imag)
    beyond = t > t_low[-1]                                    # linear extrapolation past the last low-rate sample
    if beyond.any():
        slope = (low[-1] - low[-2]) / M
        ext = low[-1] + slope * (t[beyond] - t_low[-1])
        re[beyond] = ext.real; im[beyond] = ext.imag
    return (re + 1j * im).astype(np.complex64)


def multipath_samples(ch, fs=8000, nseconds=10, seed=1):
    """multipath_samples.m:7-31 -> (G1, G2, hf_gain, delay_samples)"""
    spread, delay_s = CHANNELS[ch]
    rng = np.random.default_rng(seed)
    n = int(fs * nseconds)
    G1 = doppler_spread(spread, fs, n, rng); G2 = doppler_spread(spread, fs, n, rng)
    hf_gain = 1.0 / np.sqrt(np.var(G1) + np.var(G2))
    return G1, G2, float(hf_gain), int(round(delay_s * fs))
