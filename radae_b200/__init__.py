"""radae_b200 — B200-native (sm_100a) RADE V1 inference + OFDM modem hot path behind the reference's C ABI.

Everything numerical lives in lib/libradae_b200.so (hand-written CUDA, built by `python -m radae_b200.build`).
This package only holds the host-side mirror of the reference's interfaces (streaming.py: radae_tx / radae_rx),
the batched-context wrapper (batch.py) and the weight container tools (rdw.py).
"""
from .batch import RadeBatch          # noqa: F401
from .streaming import radae_tx, radae_rx   # noqa: F401
