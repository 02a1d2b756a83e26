"""The library holds two kernel families for the core codec: the tcgen05 kernels (core_codec_umma.cu, the default — every other
GPU test runs them) and the mma.sync kernels of round 1 (core_codec.cu, RADE_B200_CODEC=mma, kept for A/B measurements).  The
switch is read once per process, so the legacy family is checked in a child process against the same oracle, bit for bit."""
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
print("CODEC-FAMILY-OK")
"""


@pytest.mark.parametrize("family", ["mma", "umma"])
def test_codec_family_bit_exact_vs_oracle(family):
    from gpu_util import need_gpu
    need_gpu()
    env = dict(os.environ, RADE_B200_CODEC=family, PYTHONPATH=REPO)
    r = subprocess.run(["timeout", "120", sys.executable, "-c", CHILD % REPO], capture_output=True, text=True, env=env, cwd=REPO)
    assert r.returncode == 0 and "CODEC-FAMILY-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
