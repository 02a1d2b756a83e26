"""Produce radae_b200/weights/model19_check3.rdw from the reference's shipped weights.

Source of truth: /root/reference/bin/model19_check3.bin (the DNNw blob the reference C path loads,
src/test_rade_enc.c:49-67; identical to the arrays compiled into src/rade_{enc,dec}_data.c).
Cross-check (--verify-pth): re-derive int8 weights / scales / biases from the PyTorch checkpoint with the
exporter's formulas (weight-exchange/wexchange/c_export/common.py:132-137 quantize_weight, :180-194
compute_scaling, :267 final scale, :360-368 GRU gate reorder r,z,n -> z,r,n, :307-311 conv (out,in,k) ->
(k*in,out)) and require exact equality with the blob.

Run where /root/reference exists:   python tools/export_weights.py --verify-pth
"""
import argparse, os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radae_b200 import rdw


def scaling(w_in_out):
    """per-output scale for a [in][out] float matrix (common.py:180-194)"""
    mx = np.max(np.abs(w_in_out), axis=0) / 127
    sm = np.max(np.abs(w_in_out[0::2] + w_in_out[1::2]), axis=0) / 129
    return np.maximum(mx, sm)


def derive_from_pth(ckpt_path):
    import torch
    sd = torch.load(ckpt_path, map_location="cpu", weights_only=True)["state_dict"]
    g = lambda k: sd[k].detach().numpy().astype(np.float32).copy()
    out = {}

    def lin_q(name, w_in_out, bias):
        sc = scaling(w_in_out)
        q = np.round(w_in_out / sc).astype(np.int64)
        assert q.max() <= 127 and q.min() > -128
        out[f"{name}.w8"] = q.T.astype(np.int8)
        out[f"{name}.scale"] = (sc / 127).astype(np.float32)
        out[f"{name}.bias"] = bias.astype(np.float32)

    def gru(prefix, name):
        wi, wh, bi, bh = (g(f"{prefix}.weight_ih_l0"), g(f"{prefix}.weight_hh_l0"),
                          g(f"{prefix}.bias_ih_l0"), g(f"{prefix}.bias_hh_l0"))
        N = wi.shape[0] // 3
        for x in (wi, wh, bi, bh):       # r,z,n -> z,r,n
            t = x[0:N].copy(); x[0:N] = x[N:2 * N]; x[N:2 * N] = t
        lin_q(f"{name}_input", wi.T, bi)
        lin_q(f"{name}_recurrent", wh.T, bh)

    def conv(prefix, name):
        w = np.transpose(g(f"{prefix}.weight"), (2, 1, 0))          # (k, in, out), tap 0 = oldest
        lin_q(name, w.reshape(-1, w.shape[-1]), g(f"{prefix}.bias"))

    def dense_f(prefix, name):
        out[f"{name}.wf"] = g(f"{prefix}.weight").T.copy()
        out[f"{name}.bias"] = g(f"{prefix}.bias")

    def glu(prefix, name):
        # weight_norm parametrisation: original0 = g [out,1], original1 = v [out,in]; w = g * v/||v|| per output row
        # (older checkpoints such as model05 still carry the pre-parametrisation names weight_g / weight_v)
        if f"{prefix}.weight_g" in sd:
            gg, v = g(f"{prefix}.weight_g"), g(f"{prefix}.weight_v")
        else:
            gg, v = g(f"{prefix}.parametrizations.weight.original0"), g(f"{prefix}.parametrizations.weight.original1")
        import torch as _t
        w = _t._weight_norm(_t.tensor(v), _t.tensor(gg), 0).numpy()
        lin_q(name, w.T, np.zeros(w.shape[0], np.float32))

    dense_f("core_encoder.module.dense_1", "enc_dense1")
    dense_f("core_encoder.module.z_dense", "enc_zdense")
    dense_f("core_decoder.module.dense_1", "dec_dense1")
    dense_f("core_decoder.module.output", "dec_output")
    for i in range(1, 6):
        gru(f"core_encoder.module.gru{i}", f"enc_gru{i}")
        gru(f"core_decoder.module.gru{i}", f"dec_gru{i}")
        conv(f"core_encoder.module.conv{i}.conv", f"enc_conv{i}")
        conv(f"core_decoder.module.conv{i}.conv", f"dec_conv{i}")
        glu(f"core_decoder.module.glu{i}.gate", f"dec_glu{i}")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blob", default="/root/reference/bin/model19_check3.bin")
    ap.add_argument("--pth", default="/root/reference/model19_check3/checkpoints/checkpoint_epoch_100.pth")
    ap.add_argument("--out", default=rdw.default_weights_path())
    ap.add_argument("--verify-pth", action="store_true")
    args = ap.parse_args()

    arrays = rdw.dnnw_to_arrays(open(args.blob, "rb").read())
    if args.verify_pth:
        ref = derive_from_pth(args.pth)
        worst = 0.0
        for k, a in arrays.items():
            b = ref[k]
            assert a.shape == b.shape, (k, a.shape, b.shape)
            if a.dtype == np.int8:
                assert np.array_equal(a, b), f"{k}: int8 mismatch"
            else:
                d = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))
                worst = max(worst, d)
                assert d <= 1e-6 * max(1.0, float(np.abs(a).max())), f"{k}: float mismatch {d}"
        print(f"verify-pth: {len(arrays)} arrays, int8 exact, worst float abs diff {worst:.3g}")
    rdw.write_rdw(args.out, arrays, model_name=os.path.splitext(os.path.basename(args.blob))[0])
    back = rdw.read_rdw(args.out)
    assert all(np.array_equal(back[k], arrays[k]) for k in arrays)
    print(f"wrote {args.out}: {len(arrays)} arrays, {os.path.getsize(args.out)} bytes")


if __name__ == "__main__":
    main()
