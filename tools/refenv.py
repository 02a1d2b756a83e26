"""Import the Python reference from /root/reference for fixture generation (tools/ only).

Never imported by the product, tests or bench: /root/reference does not exist on the GPU box.
Makes the reference deterministic the way SURVEY.md §0 describes:
  * radae.radae_base.n (random "8-bit quantisation noise", radae/radae_base.py:80) -> identity
  * torch single-threaded
"""
import os, sys, io, contextlib

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present; golden fixtures can only be regenerated where /root/reference exists")
    for p in (os.path.join(HERE, "_mpl_stub"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    torch.set_num_threads(1)
    import radae.radae_base as rb
    rb.n = lambda x: x
    import radae
    return radae


@contextlib.contextmanager
def quiet():
    """the reference prints configuration chatter to stderr on construction"""
    old = sys.stderr
    sys.stderr = io.StringIO()
    try:
        yield
    finally:
        sys.stderr = old


CKPT = os.path.join(REF, "model19_check3/checkpoints/checkpoint_epoch_100.pth")
