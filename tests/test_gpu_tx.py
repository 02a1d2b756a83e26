"""GPU parity: transmitter (core encoder + OFDM modulator + EOO) through the C ABI vs the oracle and the
reference's golden output.  Tolerance: tx samples <= 1e-5 relative RMS (complex64 re-association), z bit exact."""
import numpy as np
import pytest
from gpu_util import need_gpu, relrms
from oracle import dsp as od
from oracle.core import CoreOraclePort

pytestmark = pytest.mark.gpu


def test_rade_tx_single_stream_vs_reference_golden(golden):
    need_gpu()
    from radae_b200 import radae_tx
    g = golden("tx")
    tx = radae_tx()
    assert (tx.get_n_features_in(), tx.get_Nmf(), tx.get_Neoo(), tx.get_Neoo_bits()) == (432, 960, 1152, 180)
    feats = g["features36"][0].reshape(-1, 432)
    out = np.zeros(960, np.complex64)
    for i in range(feats.shape[0]):
        tx.do_radae_tx(feats[i], out)
        assert relrms(out, g["tx"][i]) < 1e-5, i            # golden = reference transmitter_one on the C-encoder z
    eoo = np.zeros(1152, np.complex64)
    tx.do_eoo(eoo)
    assert relrms(eoo, g["eoo_nobits"]) < 1e-5
    tx.set_eoo_bits(g["eoo_bits"])
    tx.do_eoo(eoo)
    assert relrms(eoo, g["eoo_withbits"]) < 1e-5
    tx.close()


def test_batched_tx_vs_oracle_many_streams():
    need_gpu()
    from radae_b200 import RadeBatch
    from oracle.core import synth_features
    S, F = 33, 4
    feats = synth_features(S, 12 * F, seed=21).reshape(S, F, 432)
    b = RadeBatch(S)
    core = CoreOraclePort(n_streams=S)
    for f in range(F):
        tx = b.tx(feats[:, f])
        x = np.concatenate([feats[:, f].reshape(S, 12, 36)[:, :, :20], -np.ones((S, 12, 1), np.float32)], axis=2).reshape(S, 3, 84)
        z = core.encode(x, nthreads=8)
        ref = np.array([od.transmitter_one(z[s]) for s in range(S)])
        assert relrms(tx, ref) < 1e-5
        fr = tx.reshape(S, 5, 192)
        assert np.array_equal(fr[:, :, :32], fr[:, :, -32:])      # cyclic prefix is a copy of the tail
        assert np.abs(tx).max() <= 1.0 + 1e-6                      # PA model: tanh(|x|) saturates at 1
    bits = np.sign(np.random.default_rng(1).random((S, 180)) - 0.5).astype(np.float32)
    b.tx_set_eoo_bits(bits)
    eoo = b.tx_eoo()
    ref = np.array([od.eoo_frame(bits[s]) for s in range(S)])
    assert relrms(eoo, ref) < 1e-5
    b.close()


def test_bypass_enc_transmitter_equals_rade_tx(golden):
    """radae_tx(bypass_enc=True) fed with the latents of the stand-alone core encoder (the composition src/rade_api.c:411-436
    makes: C encoder -> do_radae_tx(z)) gives the very samples rade_tx produces from the features — bit for bit — and the
    reference's golden frames to 1e-5"""
    need_gpu()
    from radae_b200 import radae_tx, RadeBatch
    g = golden("tx")
    feats = g["features36"][0].reshape(-1, 432)
    enc = RadeBatch(1)
    x = np.concatenate([feats.reshape(-1, 12, 36)[:, :, :20], -np.ones((feats.shape[0], 12, 1), np.float32)], axis=2)
    z = enc.core_encode(x.reshape(1, -1, 84))[0].reshape(-1, 240)
    enc.close()
    assert np.array_equal(z, g["z"])
    full, byp = radae_tx(), radae_tx(bypass_enc=True)
    assert (byp.get_n_floats_in(), full.get_n_floats_in()) == (240, 432)
    a = np.zeros(960, np.complex64); c = np.zeros(960, np.complex64)
    for i in range(feats.shape[0]):
        full.do_radae_tx(feats[i], a); byp.do_radae_tx(z[i], c)
        assert np.array_equal(a, c), i
        assert relrms(c, g["tx"][i]) < 1e-5, i
    full.close(); byp.close()


def test_tx_bandpass_filter_vs_reference_golden(golden):
    """radae_tx(txbpf_en=True): golden = the reference's own filtered + clipped frames for the same latents
    (tools/make_golden_txbpf.py); six frames and the EOO frame run through one filter state"""
    need_gpu()
    from radae_b200 import radae_tx
    g = golden("tx_bpf")
    tx = radae_tx(txbpf_en=True, bypass_enc=True)
    out = np.zeros(960, np.complex64)
    for i, z in enumerate(g["z"]):
        tx.do_radae_tx(z, out)
        assert relrms(out, g["tx"][i]) < 1e-5, i
        assert np.abs(out).max() <= 1.0 + 1e-6
    tx.set_eoo_bits(g["eoo_bits"])
    eoo = np.zeros(1152, np.complex64)
    tx.do_eoo(eoo)
    assert relrms(eoo, g["eoo"]) < 1e-5
    tx.close()
    # with the encoder inside (features in): same frames, since tx.npz / tx_bpf.npz share features -> z
    tx = radae_tx(txbpf_en=True)
    feats = golden("tx")["features36"][0].reshape(-1, 432)
    for i in range(feats.shape[0]):
        tx.do_radae_tx(feats[i], out)
        assert relrms(out, g["tx"][i]) < 1e-5, i
    tx.close()


def test_batched_tx_bandpass_vs_oracle_and_reset():
    need_gpu()
    from radae_b200 import RadeBatch
    from oracle.core import synth_features
    S, F = 19, 3
    feats = synth_features(S, 12 * F, seed=5).reshape(S, F, 432)
    b = RadeBatch(S)
    b.tx_bpf_enable(True)
    core = CoreOraclePort(n_streams=S)
    refs = [od.RadaeTx(None, txbpf_en=True) for _ in range(S)]
    first = None
    for f in range(F):
        tx = b.tx(feats[:, f])
        x = np.concatenate([feats[:, f].reshape(S, 12, 36)[:, :, :20], -np.ones((S, 12, 1), np.float32)], axis=2).reshape(S, 3, 84)
        z = core.encode(x, nthreads=8)
        ref = np.array([refs[s].do_radae_tx_from_z(z[s]) for s in range(S)])
        assert relrms(tx, ref) < 1e-5, f
        assert relrms(b.tx_z(z.reshape(S, 240) * 0), np.array([refs[s].do_radae_tx_from_z(z[s] * 0) for s in range(S)])) < 1e-5
        first = tx if first is None else first
    bits = np.sign(np.random.default_rng(2).random((S, 180)) - 0.5).astype(np.float32)
    b.tx_set_eoo_bits(bits)
    eoo = b.tx_eoo()
    for s in range(S):
        refs[s].set_eoo_bits(bits[s])
    assert relrms(eoo, np.array([refs[s].do_eoo() for s in range(S)])) < 1e-5
    # the fused modulator+channel loop-back has no filter stage and must say so instead of skipping it
    with pytest.raises(RuntimeError):
        b.tx_channel_link_dev(0)
    # reset: encoder and filter start again -> frame 0 is reproduced exactly
    b.reset()
    assert np.array_equal(b.tx(feats[:, 0]), first)
    # filter off again: plain modulator output
    b.tx_bpf_enable(False); b.reset()
    core2 = CoreOraclePort(n_streams=S)
    x = np.concatenate([feats[:, 0].reshape(S, 12, 36)[:, :, :20], -np.ones((S, 12, 1), np.float32)], axis=2).reshape(S, 3, 84)
    assert relrms(b.tx(feats[:, 0]), np.array([od.transmitter_one(zz) for zz in core2.encode(x, nthreads=8)])) < 1e-5
    b.close()
