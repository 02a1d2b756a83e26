"""Generate tests/golden/tx_bpf.npz by running the REFERENCE's transmitter with its optional TX band-pass filter.

  radae_txe.radae_tx(ckpt, bypass_enc=True, txbpf_en=True)     (/root/reference/radae_txe.py:47-144)

fed with the latents of tests/golden/tx.npz (same z, so tx.npz holds the unfiltered frames for the same input):
six modem frames through do_radae_tx, then set_eoo_bits + do_eoo — the filter state (102-sample memory quirk and
mixer phase, radae/dsp.py:96-99) carries through all seven calls.  Run only where /root/reference exists:
    python tools/make_golden_txbpf.py
"""
import os, sys
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import refenv

GOLD = os.path.join(REPO, "tests", "golden")


def main():
    refenv.load()
    sys.path.insert(0, refenv.REF)
    os.chdir(refenv.REF)                       # the reference resolves its checkpoint relative to CWD
    with refenv.quiet():
        import radae_txe
        tx_ref = radae_txe.radae_tx("model19_check3/checkpoints/checkpoint_epoch_100.pth", bypass_enc=True, txbpf_en=True)
    g = np.load(os.path.join(GOLD, "tx.npz"))
    z = g["z"]
    n_mf = z.shape[0]
    Nmf, Neoo = tx_ref.get_Nmf(), tx_ref.get_Neoo()
    out = np.zeros((n_mf, Nmf), np.complex64)
    buf = np.zeros(Nmf, np.csingle)
    for i in range(n_mf):
        tx_ref.do_radae_tx(z[i].copy(), buf); out[i] = buf
    with refenv.quiet():
        tx_ref.set_eoo_bits(g["eoo_bits"])
    eoo = np.zeros(Neoo, np.csingle)
    tx_ref.do_eoo(eoo)
    np.savez_compressed(os.path.join(GOLD, "tx_bpf.npz"), z=z, eoo_bits=g["eoo_bits"], tx=out, eoo=eoo)
    print("tx_bpf: rms %.4f peak %.4f (unfiltered rms %.4f)" % (np.sqrt(np.mean(np.abs(out) ** 2)), np.abs(out).max(),
                                                                 np.sqrt(np.mean(np.abs(g["tx"]) ** 2))))


if __name__ == "__main__":
    main()
