"""CPU tests: the numpy DSP oracle against golden vectors produced by the Python reference
(radae_txe.radae_tx, radae_rxe.radae_rx, radae.complex_bpf — see tools/make_golden.py)."""
import math
import numpy as np
import pytest
from oracle import dsp as od
from oracle.core import CoreOraclePort

SCENARIOS = ["awgn_clean", "awgn_1dB", "mpp_3dB", "slip_plus", "slip_minus", "offair_long_qso", "foff_test",
             "dfdt", "noise_only", "sine_noise", "mpd_fading"]


def relrms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / np.sqrt(np.mean(np.abs(b) ** 2)))


def test_constants():
    c = od.consts()
    assert (od.NMF, od.NEOO, od.NIN_MAX, od.RXBUF, od.N_FEATURES, od.N_EOO_BITS) == (960, 1152, 1120, 2112, 432, 180)
    assert abs(c.pilot_gain - 23.2038) < 1e-3
    assert abs(float(c.bpf_centre) - 1475.0) < 0.01 and abs(float(c.bpf_bw) - 1740.0) < 0.01
    assert c.fcoarse[0] == -50.0 and c.fcoarse[-1] == 47.5 and len(c.fcoarse) == 40


def test_transmitter_vs_reference(golden):
    g = golden("tx")
    tx = np.array([od.transmitter_one(z) for z in g["z"]])
    assert relrms(tx, g["tx"]) < 1e-5
    assert relrms(od.eoo_frame(), g["eoo_nobits"]) < 1e-5
    assert relrms(od.eoo_frame(g["eoo_bits"]), g["eoo_withbits"]) < 1e-5
    # size-independent properties: cyclic prefix is the tail copy, PA limiter bounds |tx| < 1
    fr = tx[0].reshape(5, 192)
    assert np.allclose(fr[:, :32], fr[:, -32:])
    assert np.abs(tx).max() < 1.0


def test_bpf_chunked_vs_reference_including_memory_quirk(golden):
    g = golden("bpf")
    f = od.ComplexBPF()
    o = 0; ys = []
    for n in g["chunks"]:
        ys.append(f.bpf(g["x"][o:o + n])); o += n
    assert relrms(np.concatenate(ys), g["y"]) < 1e-5


def test_tx_bandpass_and_clip_vs_reference(golden):
    """radae_tx(txbpf_en=True): six frames + EOO through one filter object (state carries), tools/make_golden_txbpf.py"""
    g = golden("tx_bpf")
    tx = od.RadaeTx(None, txbpf_en=True)
    out = np.array([tx.do_radae_tx_from_z(z) for z in g["z"]])
    assert relrms(out, g["tx"]) < 1e-5
    tx.set_eoo_bits(g["eoo_bits"])
    assert relrms(tx.do_eoo(), g["eoo"]) < 1e-5
    assert np.abs(out).max() <= 1.0 + 1e-6
    # the filter matters: the unfiltered frames of tx.npz (same z) are far away
    assert relrms(golden("tx")["tx"], g["tx"]) > 0.1


@pytest.mark.parametrize("name", SCENARIOS)
def test_streaming_receiver_vs_reference(golden, name):
    g = golden("rx_" + name)
    # foff_test: the RADE_FOFF_TEST mode of the C API (radae_rx(foff_err=10)): first sync is knocked 10 Hz off, the
    # unique word fails, the receiver drops out and re-acquires
    rx = od.RadaeRx(CoreOraclePort(n_streams=1), foff_err=10.0 if name == "foff_test" else 0.0)
    o = 0
    tr = {k: [] for k in ("nin", "ret", "state", "tmax", "fmax", "snr", "uw_errors")}
    zs, fs, eo = [], [], []
    while o + rx.nin <= len(g["rx_in"]):
        nin = rx.nin
        ret, feat, eoo = rx.do_radae_rx(g["rx_in"][o:o + nin]); o += nin
        for k, v in (("nin", nin), ("ret", ret), ("state", rx.state), ("tmax", rx.tmax), ("fmax", rx.fmax),
                     ("snr", rx.receiver.snrdB_3k_est), ("uw_errors", rx.uw_errors)):
            tr[k].append(v)
        if ret & 1:
            zs.append(rx.z_hat.reshape(-1)); fs.append(feat)
        if ret & 2:
            eo.append(eoo)
    for k in ("nin", "ret", "state", "tmax", "uw_errors"):            # framing / state machine: bit exact
        assert np.array_equal(np.array(tr[k]), g[k]), k
    assert np.max(np.abs(np.array(tr["fmax"]) - g["fmax"])) < 1e-6
    assert np.max(np.abs(np.array(tr["snr"]) - g["snr"])) < 1e-3
    if name in ("noise_only", "sine_noise"):                           # ctests acq_noise / acq_sine: must not acquire
        assert not (g["state"] == 2).any() and len(zs) == 0
        return
    assert relrms(np.array(zs).reshape(-1, 240), g["z_hat"]) < 1e-5     # PSK symbols: 1e-5 relative rms
    if len(eo):
        assert np.max(np.abs(np.array(eo) - g["eoo"])) < 1e-3


def test_arange_restatement_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        f = float(rng.uniform(-60, 60))
        assert np.array_equal(od.arange_like_numpy(f - 1, f + 1, 0.1), np.arange(f - 1, f + 1, 0.1))
    for f in np.arange(-50, 50, 2.5):
        assert np.array_equal(od.arange_like_numpy(f - 10, f + 10, 0.25), np.arange(f - 10, f + 10, 0.25))


def test_channel_oracle_vs_reference_forward(golden):
    """oracle.dsp.channel / ebno_sigma / mp_gain_of against the rate-Fs channel of RADAE.forward itself
    (radae/radae.py:529-599; fixture: tools/make_golden_channel.py): two-path multipath with the power normalisation, frequency
    offset with drift (cumsum phase), phase offset, AWGN with the reference's own noise draw, gain"""
    g = golden("channel")
    for name in g["names"]:
        EbNodB, f0, df_dt, ph0, gain, d = g[f"{name}_params"]
        tx, G, noise = g[f"{name}_tx"], g[f"{name}_G"], g[f"{name}_noise"]
        sigma = od.ebno_sigma(EbNodB)
        assert abs(sigma - float(g[f"{name}_sigma"])) < 1e-6 * sigma
        mpg = od.mp_gain_of(tx, G[:, 0], G[:, 1], int(d))
        rx = od.channel(tx, G[:, 0], G[:, 1], int(d), mpg, f0, ph0, sigma, noise, gain=gain, df_dt=df_dt)
        ref = g[f"{name}_rx"]
        err = np.sqrt(np.mean(np.abs(rx - ref) ** 2) / np.mean(np.abs(ref) ** 2))
        assert err < 2e-6, (name, err)          # float32 phase accumulation (torch.cumsum) vs the closed form


@pytest.mark.parametrize("span, step, nk", [(1.0, 0.1, 9), (10.0, 0.25, 18)])
def test_refine_moments_form_equals_direct_complex128(golden, span, step, nk):
    """rx_track / rx_finish evaluate acquisition.refine (radae/dsp.py:233-270) as Taylor moments of the window about the tracked
    frequency (refine_moments: 9 terms for +-1 Hz; refine_first_fix: 18 terms for +-10 Hz) instead of one complex128 steering
    vector per frequency.  The formulation, restated in numpy float64, must agree with the reference's direct sums far below
    the csingle rounding the reference applies afterwards — for every (t, f) of the search, on a recorded signal."""
    c = od.consts()
    rx = golden("rx_mpp_3dB")["rx_in"][5000:5000 + od.RXBUF].astype(np.complex64).astype(np.complex128)
    p = c.p.astype(np.complex128)
    n = np.arange(od.M); nc = n - 79.5
    f0, tmax = -11.3, 400
    fgrid = np.arange(f0 - span, f0 + span, step)
    w0 = 2 * np.pi * f0 / od.FS
    q = np.conj(p) * np.exp(-1j * w0 * nc)
    bk = np.array([(nc / 80.0) ** k / math.factorial(k) for k in range(nk)])            # [nk][160]
    worst = 0.0
    for t in range(tmax - 8, tmax + 8):
        for pos in (0, 1):
            x = rx[t + pos * od.NMF: t + pos * od.NMF + od.M]
            Mk = bk @ (x * q)                                                            # the moments
            for f in fgrid:
                w = 2 * np.pi * f / od.FS
                direct = np.dot(x, np.exp(-1j * w * n) * np.conj(p)) * (np.exp(-1j * w * od.NMF) if pos else 1.0)
                al = 2 * np.pi * (f - f0) / od.FS * 80.0
                D = 0j
                for k in range(nk - 1, -1, -1):
                    D = D * (-1j * al) + Mk[k]
                mom = D * np.exp(-1j * w * (79.5 + od.NMF * pos))
                worst = max(worst, abs(mom - direct) / (np.abs(x).sum() * np.abs(p).max()))
    assert worst < 2e-15, worst
