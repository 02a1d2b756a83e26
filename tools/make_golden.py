"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported from /root/reference) on seeded inputs.

Run only where /root/reference exists:   python tools/make_golden.py
The committed .npz files are what tests compare against on machines without the reference (the GPU box).

Determinism: radae_base.n() -> identity (tools/refenv.py); np.random.randint inside acquisition.check_pilots
(radae/dsp.py:293, unseeded in the reference) is replaced by oracle.dsp.refresh_rows' deterministic schedule — the
random rows only feed sigma_r, so any schedule is reference-conformant; ours is what the CUDA path implements.
"""
import os, sys
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import refenv
from oracle import dsp as od
from oracle.core import CoreOracleRef, CoreOraclePort, synth_features, pack_enc_input

GOLD = os.path.join(REPO, "tests", "golden")


class RowSchedule:
    """stands in for np.random inside the reference's check_pilots"""
    def __init__(self):
        self.n_call = 0; self.i = 0
    def randint(self, high):
        assert high == od.NMF
        t = int(od.refresh_rows(self.n_call)[self.i])
        self.i += 1
        if self.i == od.NUPDATE:
            self.i = 0; self.n_call += 1
        return t


def make_rx_input(scn, tx_frames, eoo):
    """Assemble a receive signal: lead-in noise, modem frames through the oracle channel, EOO, tail."""
    rng = np.random.default_rng(scn["seed"])
    tx = np.concatenate(tx_frames + ([eoo] if scn.get("eoo", True) else []))
    if scn.get("resample", 1.0) != 1.0:
        r = scn["resample"]
        n_out = int(len(tx) / r)
        pos = np.arange(n_out) * r
        i0 = np.floor(pos).astype(int); fr = (pos - i0).astype(np.float32)
        i1 = np.minimum(i0 + 1, len(tx) - 1)
        tx = ((1 - fr) * tx[i0] + fr * tx[i1]).astype(np.complex64)
    T = len(tx)
    if scn.get("multipath"):
        # slowly rotating two-path channel, delay 16 samples (radae/radae.py:530-534)
        n = np.arange(T)
        G1 = (0.8 * np.exp(1j * 2 * np.pi * 0.3 * n / od.FS)).astype(np.complex64)
        G2 = (0.6 * np.exp(-1j * (1.0 + 2 * np.pi * 0.5 * n / od.FS))).astype(np.complex64)
    else:
        G1 = np.ones(T, np.complex64); G2 = np.zeros(T, np.complex64)
    noise = ((rng.standard_normal(T) + 1j * rng.standard_normal(T)) / np.sqrt(2)).astype(np.complex64)
    sigma = od.ebno_sigma(scn["EbNodB"])
    rx = od.channel(tx, G1, G2, 16, 1.0, scn["freq_offset"], 0.3, sigma, noise, gain=scn.get("gain", 1.0))
    if scn.get("df_dt"):
        # linear frequency drift on top of the fixed offset (inference.py --df_dt, radae/radae.py:542-553): phase pi*df_dt*t^2
        t = np.arange(T, dtype=np.float64) / od.FS
        rx = (rx * np.exp(1j * np.pi * scn["df_dt"] * t * t)).astype(np.complex64)
    lead = scn["lead"]
    nl = ((rng.standard_normal(lead) + 1j * rng.standard_normal(lead)) / np.sqrt(2)).astype(np.complex64)
    tail = scn.get("tail", 2 * od.NMF)
    nt = ((rng.standard_normal(tail) + 1j * rng.standard_normal(tail)) / np.sqrt(2)).astype(np.complex64)
    g = np.float32(scn.get("gain", 1.0) * sigma)
    return np.concatenate([g * nl, rx, g * nt]).astype(np.complex64)


def make_no_signal_input(scn):
    """what the reference's acq_noise / acq_sine ctests feed (CMakeLists.txt:188-208): noise, optionally plus a carrier
    at 1 kHz -- the receiver must never reach sync"""
    rng = np.random.default_rng(scn["seed"])
    T = scn["n_mf"] * od.NMF
    x = ((rng.standard_normal(T) + 1j * rng.standard_normal(T)) / np.sqrt(2)).astype(np.complex64) * np.float32(scn["noise_rms"])
    if scn.get("sine_amp"):
        x = x + (scn["sine_amp"] * np.exp(2j * np.pi * scn["sine_freq"] * np.arange(T) / od.FS)).astype(np.complex64)
    return x.astype(np.complex64)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="regenerate just this receiver scenario (e.g. foff_test); the other fixtures are left alone")
    only = ap.parse_args().only
    radae = refenv.load()
    import torch
    sys.path.insert(0, refenv.REF)
    os.makedirs(GOLD, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(refenv.REF)                       # the reference resolves its checkpoint relative to CWD
    with refenv.quiet():
        import radae_txe, radae_rxe
        from radae import RADAE
    ck = "model19_check3/checkpoints/checkpoint_epoch_100.pth"

    bits = np.sign(np.random.default_rng(65647).random(od.N_EOO_BITS) - 0.5).astype(np.float32)
    if not only:
        bits = make_static_fixtures(radae, radae_txe, RADAE, torch, ck)
    make_rx_scenarios(radae_rxe, ck, bits, only)
    os.chdir(cwd)


def make_static_fixtures(radae, radae_txe, RADAE, torch, ck):
    # ---------------------------------------------------------------- core codec (a1, a2)
    S, T = 2, 24
    feats36 = synth_features(S, 4 * T, seed=1234)
    x = pack_enc_input(feats36)
    ref = CoreOracleRef("int8", S)
    z_c = ref.encode(x)
    f_c = ref.decode(z_c)
    reff = CoreOracleRef("f32", S)
    z_cf = reff.encode(x)
    f_cf = reff.decode(z_cf)
    with refenv.quiet():
        model = RADAE(21, 80, EbNodB=100, rate_Fs=True, pilots=True, pilot_eq=True, eq_mean6=False,
                      cyclic_prefix=0.004, coarse_mag=True, time_offset=-16, bottleneck=3)
    model.load_state_dict(torch.load(ck, map_location="cpu", weights_only=True)["state_dict"], strict=False)
    model.core_encoder_statefull_load_state_dict(); model.core_decoder_statefull_load_state_dict(); model.eval()
    z_py = np.zeros_like(z_c); f_py = np.zeros_like(f_c); f_py_from_c = np.zeros_like(f_c)
    with torch.inference_mode():
        for s in range(S):
            for mod in (model.core_encoder_statefull.module, model.core_decoder_statefull.module):
                for name, m in mod.named_modules():
                    if hasattr(m, "reset") and m is not mod: m.reset()
            for t in range(T):
                z_py[s, t] = model.core_encoder_statefull(torch.tensor(x[s, t].reshape(1, 4, 21))).numpy()[0, 0]
            for t in range(T):
                f_py[s, t] = model.core_decoder_statefull(torch.tensor(z_py[s:s+1, t:t+1])).numpy().reshape(84)
            model.core_decoder_statefull.module.reset()
            for t in range(T):
                f_py_from_c[s, t] = model.core_decoder_statefull(torch.tensor(z_c[s:s+1, t:t+1])).numpy().reshape(84)
    # the reference's own acceptance metric (loss.py / radae_base.distortion_loss)
    from radae.radae_base import distortion_loss
    def loss(a, b):
        return float(distortion_loss(torch.tensor(a.reshape(S, 4 * T, 21)), torch.tensor(b.reshape(S, 4 * T, 21))).mean())
    target = x.reshape(S, 4 * T, 21)
    l_py = loss(target, f_py); l_c = loss(target, ref_dec_of(ref, z_c, S)) if False else loss(target, f_c)
    print(f"core: loss(py float)={l_py:.4f} loss(C int8)={l_c:.4f} delta={abs(l_py-l_c):.4f}  (reference bar: < 0.01)")
    np.savez_compressed(os.path.join(GOLD, "core_codec.npz"), features36=feats36, z_c_int8=z_c, f_c_int8=f_c,
                        z_c_f32=z_cf, f_c_f32=f_cf, z_py=z_py, f_py=f_py, loss_py=l_py, loss_c_int8=l_c)

    # ---------------------------------------------------------------- transmitter (a4, a5)
    with refenv.quiet():
        tx_ref = radae_txe.radae_tx(ck, bypass_enc=True)
    n_mf = 6
    S1 = 1
    feats = synth_features(1, 12 * n_mf, seed=77)
    enc = CoreOracleRef("int8", 1)
    z_all = enc.encode(pack_enc_input(feats))[0].reshape(n_mf, 240)
    tx_out = np.zeros((n_mf, od.NMF), np.complex64)
    buf = np.zeros(od.NMF, np.csingle)
    for i in range(n_mf):
        tx_ref.do_radae_tx(z_all[i].copy(), buf); tx_out[i] = buf
    eoo0 = np.zeros(od.NEOO, np.csingle); tx_ref.do_eoo(eoo0); eoo0 = eoo0.copy()
    bits = np.sign(np.random.default_rng(65647).random(od.N_EOO_BITS) - 0.5).astype(np.float32)
    with refenv.quiet():
        tx_ref.set_eoo_bits(bits)
    eoo1 = np.zeros(od.NEOO, np.csingle); tx_ref.do_eoo(eoo1)
    np.savez_compressed(os.path.join(GOLD, "tx.npz"), features36=feats, z=z_all, tx=tx_out, eoo_nobits=eoo0,
                        eoo_bits=bits, eoo_withbits=eoo1)
    print("tx: rms", float(np.sqrt(np.mean(np.abs(tx_out) ** 2))))

    # ---------------------------------------------------------------- BPF (a7)
    rng = np.random.default_rng(5)
    sig = ((rng.standard_normal(4000) + 1j * rng.standard_normal(4000))).astype(np.complex64)
    c = od.consts()
    bpf = radae.complex_bpf(101, od.FS, c.bpf_bw, c.bpf_centre, od.FS)
    chunks = [960, 1120, 800, 960]
    outs = []; o = 0
    for n in chunks:
        outs.append(np.array(bpf.bpf(sig[o:o + n])).astype(np.complex64)); o += n
    np.savez_compressed(os.path.join(GOLD, "bpf.npz"), x=sig, chunks=np.array(chunks), y=np.concatenate(outs))

    return bits


def make_rx_scenarios(radae_rxe, ck, bits, only=""):
    # ---------------------------------------------------------------- streaming receiver scenarios (a7-a11)
    scenarios = {
        "awgn_clean": dict(seed=11, EbNodB=20.0, freq_offset=13.0, lead=2 * od.NMF + 300, n_mf=16),
        "awgn_1dB":   dict(seed=12, EbNodB=1.0, freq_offset=13.0, lead=od.NMF + 555, n_mf=24),
        "mpp_3dB":    dict(seed=13, EbNodB=6.0, freq_offset=-11.0, lead=od.NMF + 100, n_mf=24, multipath=True, gain=0.5),
        "slip_plus":  dict(seed=14, EbNodB=20.0, freq_offset=5.0, lead=od.NMF + 424, n_mf=36, resample=0.995),
        "slip_minus": dict(seed=15, EbNodB=20.0, freq_offset=-3.0, lead=od.NMF + 944, n_mf=36, resample=1.005),
        # RADE_FOFF_TEST (src/rade_api.c:263-264 -> radae_rx(foff_err=10), radae_rxe.py:271-273): 10 Hz is added to fmax on
        # the first sync, the decoder sees garbage, the unique word fails and the receiver has to drop sync and re-acquire
        "foff_test":  dict(seed=16, EbNodB=10.0, freq_offset=7.0, lead=od.NMF + 200, n_mf=48, foff_err=10.0),
        # frequency drift (ctest radae_rx_dfdt, CMakeLists.txt:363-371, there 0.1 Hz/s over minutes; here 0.5 Hz/s over 6 s so
        # that the tracked fmax visibly has to move) at the low-SNR operating point
        "dfdt":       dict(seed=17, EbNodB=1.0, freq_offset=13.0, df_dt=0.5, lead=od.NMF + 700, n_mf=48),
        # no signal present (ctests acq_noise / acq_sine, CMakeLists.txt:188-208): the receiver must stay out of sync
        "noise_only": dict(seed=18, no_signal=True, noise_rms=0.7, n_mf=40),
        "sine_noise": dict(seed=19, no_signal=True, noise_rms=0.5, sine_amp=1.0, sine_freq=1000.0, n_mf=40),
    }
    # a real off-air RADE V1 recording shipped with the reference (8 kHz s16); fed the way the
    # reference's ctest does: int16 -> (x, 0) complex, unscaled (int16tof32.py --zeropad, CMakeLists.txt:400-406)
    import wave
    with wave.open(os.path.join(refenv.REF, "wav/long_qso.wav")) as w:
        assert w.getframerate() == 8000 and w.getnchannels() == 1 and w.getsampwidth() == 2
        offair = np.frombuffer(w.readframes(12 * 8000), np.int16).copy()
    scenarios["offair_long_qso"] = dict(offair=True)
    for name, scn in scenarios.items():
        if only and name != only:
            continue
        if scn.get("offair"):
            rx_in = offair.astype(np.float32).astype(np.complex64)
        elif scn.get("no_signal"):
            rx_in = make_no_signal_input(scn)
        else:
            n_mf = scn["n_mf"]
            feats = synth_features(1, 12 * n_mf, seed=100 + scn["seed"])
            enc = CoreOracleRef("int8", 1)
            z_all = enc.encode(pack_enc_input(feats))[0].reshape(n_mf, 240)
            frames = [od.transmitter_one(z_all[i]) for i in range(n_mf)]
            rx_in = make_rx_input(scn, frames, od.eoo_frame(bits))
        with refenv.quiet():
            rxr = radae_rxe.radae_rx(ck, bypass_dec=True, v=0, foff_err=scn.get("foff_err", 0))
        sched = RowSchedule()
        import radae.dsp as rdsp
        class _NP:                       # np proxy whose random.randint is our schedule
            def __getattr__(self, k): return getattr(np, k)
        proxy = _NP(); proxy.random = sched
        rdsp.np = proxy
        dec = CoreOracleRef("int8", 1)
        trace = dict(nin=[], ret=[], state=[], tmax=[], fmax=[], snr=[], Dthresh=[], Dtmax12=[], Dtmax12_eoo=[],
                     uw_errors=[], valid_count=[])
        zs, feats_out, eoos = [], [], []
        o = 0
        floats = np.zeros(rxr.get_n_floats_out(), np.float32)
        while o + rxr.get_nin() <= len(rx_in):
            nin = rxr.get_nin()
            ret = rxr.do_radae_rx(rx_in[o:o + nin].copy(), floats); o += nin
            trace["nin"].append(nin); trace["ret"].append(ret)
            trace["state"].append({"search": 0, "candidate": 1, "sync": 2}[rxr.state])
            trace["tmax"].append(int(rxr.tmax) if hasattr(rxr, "tmax") else 0)
            trace["fmax"].append(float(rxr.fmax) if hasattr(rxr, "fmax") else 0.0)
            trace["snr"].append(float(rxr.receiver.snrdB_3k_est))
            trace["Dthresh"].append(float(rxr.acq.Dthresh) if hasattr(rxr.acq, "Dthresh") else 0.0)
            trace["Dtmax12"].append(float(rxr.acq.Dtmax12) if hasattr(rxr.acq, "Dtmax12") else 0.0)
            trace["Dtmax12_eoo"].append(float(rxr.acq.Dtmax12_eoo))
            if ret & 1:
                z = floats[:240].copy(); zs.append(z)
                f = dec.decode(z.reshape(1, 3, 80))[0].reshape(12, 21)      # what src/rade_api.c:494-513 does
                rxr.sum_uw_errors(int(np.sum(f[0::4, 20] > 0)))
                out = np.zeros((12, 36), np.float32); out[:, :20] = f[:, :20]; feats_out.append(out.reshape(-1))
            if ret & 2:
                eoos.append(floats[:od.N_EOO_BITS].copy())
            trace["uw_errors"].append(int(rxr.uw_errors)); trace["valid_count"].append(int(rxr.valid_count))
        rdsp.np = np
        n_valid = sum(1 for r in trace["ret"] if r & 1)
        ber = float(np.mean(eoos[0] * bits < 0)) if eoos else -1
        print(f"{name}: calls={len(trace['ret'])} valid={n_valid} eoo={len(eoos)} eoo_ber={ber:.3f} final_state={trace['state'][-1]} "
              f"nin set={sorted(set(trace['nin']))} fmax_end={trace['fmax'][-1]:.2f}")
        rx_store = dict(rx_in_int16=offair) if scn.get("offair") else dict(rx_in=rx_in)
        np.savez_compressed(os.path.join(GOLD, f"rx_{name}.npz"), eoo_bits=bits, **rx_store,
                            z_hat=np.array(zs, np.float32).reshape(-1, 240), features=np.array(feats_out, np.float32).reshape(-1, 432),
                            eoo=np.array(eoos, np.float32).reshape(-1, od.N_EOO_BITS),
                            **{k: np.array(v) for k, v in trace.items()})


if __name__ == "__main__":
    main()
