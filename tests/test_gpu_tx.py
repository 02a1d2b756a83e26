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
