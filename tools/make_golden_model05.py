"""Generate tests/golden/core_codec_model05.npz: the reference's SECOND core codec configuration —
model05: 20 features per 10 ms vector, no auxiliary symbol (80-wide encoder input / decoder output), bottleneck 1 (tanh on
z), weights in /root/reference/bin/model05.bin — as exercised by the reference's ctests c_encoder_model5 / c_decoder_model5
(CMakeLists.txt:519-545: `test_rade_enc 1 0 bin/model05.bin`, `test_rade_dec 0 bin/model05.bin`).

  * z_c / f_c: the reference's own rade_enc.c / rade_dec.c (oracle/_ref, int8 path) with the blob loaded through parse_weights
  * z_py / f_py: the PyTorch stateful modules of the checkpoint (radae_base.n() -> identity), for the reference's own
    acceptance bar: loss < 0.2 and C-vs-Python |delta loss| small
Run only where /root/reference exists:   python tools/make_golden_model05.py
"""
import os, sys
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
import refenv
from oracle.core import CoreOracleRef, synth_features

GOLD = os.path.join(REPO, "tests", "golden")


def main():
    refenv.load()
    import torch
    sys.path.insert(0, refenv.REF)
    os.chdir(refenv.REF)
    with refenv.quiet():
        from radae import RADAE
        from radae.radae_base import distortion_loss
        model = RADAE(20, 80, 100.0)                                   # stateful_encoder.py:58 (bottleneck 1 is the default)
    ck = torch.load("model05/checkpoints/checkpoint_epoch_100.pth", map_location="cpu", weights_only=True)
    model.load_state_dict(ck["state_dict"], strict=False)
    model.core_encoder_statefull_load_state_dict(); model.core_decoder_statefull_load_state_dict(); model.eval()
    assert model.bottleneck == 1

    S, T = 2, 24
    feats36 = synth_features(S, 4 * T, seed=505)
    x = np.ascontiguousarray(feats36[:, :, :20].reshape(S, T, 80))
    blob = open(os.path.join(refenv.REF, "bin/model05.bin"), "rb").read()
    ref = CoreOracleRef("int8", S, blob=blob, input_dim=80, output_dim=80, bottleneck=1)
    z_c = ref.encode(x); f_c = ref.decode(z_c)
    z_py = np.zeros_like(z_c); f_py = np.zeros_like(f_c)
    with torch.inference_mode():
        for s in range(S):
            for mod in (model.core_encoder_statefull.module, model.core_decoder_statefull.module):
                for name, m in mod.named_modules():
                    if hasattr(m, "reset") and m is not mod: m.reset()
            for t in range(T):
                z_py[s, t] = model.core_encoder_statefull(torch.tensor(x[s, t].reshape(1, 4, 20))).numpy()[0, 0]
            for t in range(T):
                f_py[s, t] = model.core_decoder_statefull(torch.tensor(z_py[s:s + 1, t:t + 1])).numpy().reshape(80)

    def loss(a, b):
        return float(distortion_loss(torch.tensor(a.reshape(S, 4 * T, 20)), torch.tensor(b.reshape(S, 4 * T, 20))).mean())
    l_py, l_c = loss(x, f_py), loss(x, f_c)
    print(f"model05: loss(py float)={l_py:.4f} loss(C int8)={l_c:.4f} delta={abs(l_py - l_c):.4f}; "
          f"z rms diff {np.sqrt(np.mean((z_c - z_py) ** 2)):.4f} on |z|<=1; f rms diff {np.sqrt(np.mean((f_c - f_py) ** 2)):.4f}")
    np.savez_compressed(os.path.join(GOLD, "core_codec_model05.npz"), features36=feats36, z_c_int8=z_c, f_c_int8=f_c,
                        z_py=z_py, f_py=f_py, loss_py=l_py, loss_c_int8=l_c)


if __name__ == "__main__":
    main()
