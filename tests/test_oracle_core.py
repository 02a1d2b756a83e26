"""CPU tests: the core-codec oracle against the golden vectors produced by the reference (tools/make_golden.py)."""
import numpy as np
import pytest
from oracle.core import CoreOraclePort, CoreOracleRef, pack_enc_input, synth_features


def test_port_matches_reference_c_golden(golden):
    """our C port == the reference's own rade_enc.c/rade_dec.c (+shim) outputs committed as golden, bit for bit"""
    g = golden("core_codec")
    x = pack_enc_input(g["features36"])
    o = CoreOraclePort(n_streams=x.shape[0])
    z = o.encode(x)
    assert np.array_equal(z, g["z_c_int8"])
    f = o.decode(g["z_c_int8"])
    assert np.array_equal(f, g["f_c_int8"])


def test_synth_features_reproducible(golden):
    g = golden("core_codec")
    assert np.array_equal(synth_features(2, 96, seed=1234), g["features36"])


def test_int8_path_within_reference_acceptance_band(golden):
    """the reference accepts its C path when |loss(C) - loss(Python)| < 0.01 (CMakeLists.txt:521-556, loss.py:107-112)"""
    g = golden("core_codec")
    assert abs(float(g["loss_py"]) - float(g["loss_c_int8"])) < 0.01
    # and the int8 encoder output sits ~2 % rms from the float PyTorch encoder (SURVEY.md §8c sanity number)
    rel = np.sqrt(np.mean((g["z_c_int8"] - g["z_py"]) ** 2)) / np.sqrt(np.mean(g["z_py"] ** 2))
    assert rel < 0.05


def test_float_variant_pins_layouts_against_pytorch(golden):
    """float-weight build of the reference C code + shim vs the PyTorch stateful modules: only the tanh/sigmoid
    rational approximation (~5e-5) separates them, so layouts / gate order / conv taps / state handling are right"""
    g = golden("core_codec")
    rel_z = np.sqrt(np.mean((g["z_c_f32"] - g["z_py"]) ** 2)) / np.sqrt(np.mean(g["z_py"] ** 2))
    assert rel_z < 5e-4


@pytest.mark.skipif(not CoreOracleRef.available("int8"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("S,T,seed", [(3, 30, 99), (48, 12, 11), (64, 40, 2024)])
def test_port_equals_ref_build_on_fresh_inputs(S, T, seed):
    x = pack_enc_input(synth_features(S, 4 * T, seed=seed))
    p, r = CoreOraclePort(n_streams=S), CoreOracleRef("int8", S)
    zp, zr = p.encode(x, nthreads=2), r.encode(x, nthreads=2)
    assert np.array_equal(zp, zr)
    assert np.array_equal(p.decode(zr), r.decode(zr))
    assert r.max_abs_acc() < 2 ** 24          # float accumulation in the generic C gemv stayed exact


def test_streams_are_independent_and_chunking_is_stateful():
    S, T = 2, 12
    x = pack_enc_input(synth_features(S, 4 * T, seed=5))
    a = CoreOraclePort(n_streams=S).encode(x)
    b0 = CoreOraclePort(n_streams=1).encode(x[:1]); b1 = CoreOraclePort(n_streams=1).encode(x[1:])
    assert np.array_equal(a, np.concatenate([b0, b1]))
    c = CoreOraclePort(n_streams=S)
    parts = [c.encode(x[:, i:i + 3]) for i in range(0, T, 3)]
    assert np.array_equal(a, np.concatenate(parts, axis=1))


def test_model05_reference_c_path_within_the_reference_bar(golden):
    """second codec configuration of the reference: model05 = 80-wide input / output, bottleneck 1 (tanh on z),
    weights loaded from the reference's DNNw blob the way src/test_rade_enc.c does.  Golden: tools/make_golden_model05.py.
    The reference's own ctests accept loss < 0.2 (c_encoder_model5 / c_decoder_model5, CMakeLists.txt:519-545)."""
    g = golden("core_codec_model05")
    assert float(g["loss_c_int8"]) < 0.2 and float(g["loss_py"]) < 0.2
    assert abs(float(g["loss_c_int8"]) - float(g["loss_py"])) < 0.01
    assert np.abs(g["z_c_int8"]).max() <= 1.0                                   # bottleneck 1: tanh-limited latents
    assert np.sqrt(np.mean((g["z_c_int8"] - g["z_py"]) ** 2)) < 0.06


MODEL05_BLOB = "/root/reference/bin/model05.bin"


@pytest.mark.skipif(not (CoreOracleRef.available("int8") and __import__("os").path.exists(MODEL05_BLOB)),
                    reason="needs oracle/_ref and the reference's bin/model05.bin (CPU container only)")
def test_model05_blob_through_the_shim_reproduces_golden(golden):
    g = golden("core_codec_model05")
    x = np.ascontiguousarray(g["features36"][:, :, :20].reshape(2, -1, 80))
    with open(MODEL05_BLOB, "rb") as f:
        r = CoreOracleRef("int8", 2, blob=f.read(), input_dim=80, output_dim=80, bottleneck=1)
    z = r.encode(x)
    assert np.array_equal(z, g["z_c_int8"])
    assert np.array_equal(r.decode(z), g["f_c_int8"])


def test_port_runs_model05_bit_exact_vs_reference_golden(golden):
    """the C port on radae_b200/weights/model05.rdw (80-wide rows, bottleneck 1) == the reference C sources on bin/model05.bin"""
    from radae_b200 import rdw
    g = golden("core_codec_model05")
    x = np.ascontiguousarray(g["features36"][:, :, :20].reshape(2, -1, 80))
    o = CoreOraclePort(rdw.model05_weights_path(), n_streams=2, bottleneck=1)
    assert (o.in_dim, o.out_dim) == (80, 80)
    z = o.encode(x)
    assert np.array_equal(z, g["z_c_int8"])
    assert np.array_equal(o.decode(z), g["f_c_int8"])
