"""EXPERIMENTAL (branch tcgen05-codec): the tcgen05 formulation of the core encoder and decoder (core_encoder_umma_kernel / core_decoder_umma_kernel, selected with
RADE_B200_CODEC_UMMA=1 for the whole process) must be bit-identical to the oracle like the mma.sync kernel.  The switch is read
once per process, so the check runs in a child process.  Not yet run on a GPU: gated like the other late additions."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import numpy as np, sys
sys.path.insert(0, %r)
from radae_b200 import RadeBatch
from oracle.core import CoreOraclePort, pack_enc_input, synth_features
for S, T in ((1, 5), (8, 3), (37, 7), (200, 4)):
    x = pack_enc_input(synth_features(S, 4 * T, seed=7 + S))
    o = CoreOraclePort(n_streams=S)
    zo = o.encode(x, nthreads=8)
    b = RadeBatch(S)
    zg = b.core_encode(x)
    assert np.array_equal(zg, zo), ("first call", S, T, int((zg != zo).sum()))
    fo = o.decode(zo, nthreads=8)
    fg = b.core_decode(zo)
    assert np.array_equal(fg, fo), ("decoder, first call", S, T, int((fg != fo).sum()))
    x2 = pack_enc_input(synth_features(S, 4 * 2, seed=99 + S))
    z2 = o.encode(x2, nthreads=8)
    assert np.array_equal(b.core_encode(x2), z2), ("encoder state carried into a second call", S)
    assert np.array_equal(b.core_decode(z2), o.decode(z2, nthreads=8)), ("decoder state carried into a second call", S)
    b.close()
print("UMMA-CODEC-OK")
"""


@pytest.mark.skipif(os.environ.get("RADE_B200_RUN_UNVALIDATED") != "1", reason="experimental tcgen05 encoder: compiled, never run; enable with RADE_B200_RUN_UNVALIDATED=1")
def test_umma_encoder_and_decoder_bit_exact_vs_oracle():
    from gpu_util import need_gpu
    need_gpu()
    env = dict(os.environ, RADE_B200_CODEC_UMMA="1", PYTHONPATH=REPO)
    r = subprocess.run(["timeout", "120", sys.executable, "-c", CHILD % REPO], capture_output=True, text=True, env=env, cwd=REPO)
    assert r.returncode == 0 and "UMMA-CODEC-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
