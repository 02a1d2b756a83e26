"""tests/golden/channel.npz: the rate-Fs channel simulator of the REFERENCE (RADAE.forward, radae/radae.py:529-599) run as it is,
imported from /root/reference, on seeded inputs — pins oracle.dsp.channel / ebno_sigma and, through them, the
CUDA explicit-form channel (rade_b200_channel_apply).      python tools/make_golden_channel.py     (only where /root/reference exists)

Cases: two-path multipath with time-varying gains G (delay 2 ms = 16 samples) and the whole-tensor power normalisation mp_gain;
frequency offset with drift df_dt (cumsum phase); phase offset; AWGN at finite Eb/No (sigma for bottleneck 3) with the
reference's own torch.randn_like draw captured; user gain.  One stream per forward call: the reference normalises mp_gain over
the whole batch tensor (radae.py:536-539), the batched implementations do it per stream."""
import os, sys
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tools"))
import refenv
from oracle.core import synth_features

GOLD = os.path.join(REPO, "tests", "golden")


def main():
    refenv.load()
    import torch
    cwd = os.getcwd(); os.chdir(refenv.REF)
    with refenv.quiet():
        from radae import RADAE
    cases = [dict(name="awgn_foff", EbNodB=3.0, freq_offset=-11.0, df_dt=0.0, phase_offset=0.0, gain=1.0, mp=False),
             dict(name="mp_clean", EbNodB=100.0, freq_offset=0.0, df_dt=0.0, phase_offset=0.0, gain=1.0, mp=True),
             dict(name="mp_drift_gain", EbNodB=6.0, freq_offset=13.0, df_dt=-0.5, phase_offset=0.7, gain=0.25, mp=True)]
    out = {}
    n_feat = 12 * 10                                                   # 10 modem frames = 9600 samples
    feats = torch.tensor(synth_features(1, n_feat, seed=777)[:, :, :20])
    feats = torch.cat([feats, -torch.ones(1, n_feat, 1)], dim=2)       # aux symbol
    for c in cases:
        with refenv.quiet():
            m = RADAE(21, 80, EbNodB=c["EbNodB"], rate_Fs=True, pilots=True, pilot_eq=True, eq_mean6=False, cyclic_prefix=0.004,
                      time_offset=-16, coarse_mag=True, bottleneck=3, freq_offset=c["freq_offset"], df_dt=c["df_dt"],
                      phase_offset=c["phase_offset"], gain=c["gain"])
            ck = torch.load(refenv.CKPT, map_location="cpu", weights_only=True)
            m.load_state_dict(ck["state_dict"], strict=False)
        m.eval()
        T = m.num_timesteps_at_rate_Rs(n_feat)
        T_pil = T + T // m.Ns
        n_fs = T_pil * (m.M + m.Ncp)
        rng = np.random.default_rng(11)
        n = np.arange(n_fs)
        if c["mp"]:
            G1 = 0.8 * np.exp(1j * 2 * np.pi * 0.7 * n / 8000.0) * (1 + 0.2 * np.cos(2 * np.pi * 0.3 * n / 8000.0))
            G2 = 0.6 * np.exp(-1j * (1.0 + 2 * np.pi * 1.1 * n / 8000.0))
        else:
            G1 = np.ones(n_fs); G2 = np.zeros(n_fs)
        G = torch.tensor(np.stack([G1, G2], axis=1)[None].astype(np.complex64))
        H = torch.ones(1, T, m.Nc)
        drawn = {}
        real_randn_like = torch.randn_like

        def capture(t, *a, **k):                                       # the reference's own noise draw, recorded
            v = real_randn_like(t, *a, **k)
            if t.dtype == torch.complex64 and t.numel() == n_fs:
                drawn["noise"] = v.clone()
            return v
        torch.manual_seed(4321)
        torch.randn_like = capture
        try:
            with torch.no_grad():
                res = m.forward(feats, H, G)
        finally:
            torch.randn_like = real_randn_like
        tx = res["tx"].numpy()[0].astype(np.complex64); rx = res["rx"].numpy()[0].astype(np.complex64)
        assert tx.shape == (n_fs,) and "noise" in drawn
        out[c["name"] + "_tx"] = tx; out[c["name"] + "_rx"] = rx
        out[c["name"] + "_G"] = G.numpy()[0]; out[c["name"] + "_noise"] = drawn["noise"].numpy()[0].astype(np.complex64)
        out[c["name"] + "_sigma"] = np.float64(np.asarray(res["sigma"]).reshape(-1)[0])
        out[c["name"] + "_params"] = np.array([c["EbNodB"], c["freq_offset"], c["df_dt"], c["phase_offset"], c["gain"], float(m.d_samples)])
    os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "channel.npz"), names=np.array([c["name"] for c in cases]), **out)
    print("wrote tests/golden/channel.npz:", {k: v.shape for k, v in out.items() if k.endswith("_rx")})


if __name__ == "__main__":
    main()
