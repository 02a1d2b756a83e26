"""Command-line streaming transmitter: `python -m radae_b200.radae_txe < features.f32 > tx.iq.f32`.

Same contract and flags as the reference's script (/root/reference/radae_txe.py:145-182): 432-float feature frames (or,
with --bypass_enc, 240-float latent frames) on stdin, 960 complex64 samples per frame on stdout, the 1152-sample
end-of-over frame after the last one.  All arithmetic runs in libradae_b200.so on the GPU (see streaming.radae_tx)."""
import argparse
import sys
import numpy as np
from .streaming import radae_tx

EOO_TEST_SEED = 65647          # both ends of the reference's EOO data test derive the bits from this seed


def eoo_test_bits(n):
    return np.sign(np.random.default_rng(EOO_TEST_SEED).random(n) - 0.5).astype(np.float32)


def frames(stream, n_bytes):
    """whole records only: a short read ends the stream, like the reference's loop"""
    while True:
        buf = stream.read(n_bytes)
        if len(buf) != n_bytes:
            return
        yield buf


def build_parser():
    p = argparse.ArgumentParser(description="RADE V1 streaming transmitter on libradae_b200: features.f32 on stdin, IQ.f32 on stdout")
    p.add_argument("--model_name", type=str, default="", help="RDW or DNNw weight file (default: the embedded model19_check3 weights)")
    p.add_argument("--noauxdata", dest="auxdata", action="store_false", help="not supported by the device path (raises)")
    p.add_argument("--txbpf", action="store_true", help="enable the TX band-pass filter + clip")
    p.add_argument("--bypass_enc", action="store_true", help="bypass the core encoder, read z (240 floats per frame) from stdin")
    p.add_argument("--eoo_data_test", action="store_true", help="send the seeded EOO test bits (also written to eoo_tx.f32)")
    p.set_defaults(auxdata=True)
    return p


def main(argv=None, stdin=None, stdout=None):
    args = build_parser().parse_args(argv)
    stdin = stdin or sys.stdin.buffer
    stdout = stdout or sys.stdout.buffer
    tx = radae_tx(model_name=args.model_name, auxdata=args.auxdata, txbpf_en=args.txbpf, bypass_enc=args.bypass_enc)
    if args.eoo_data_test:
        bits = eoo_test_bits(tx.get_Neoo_bits())
        tx.set_eoo_bits(bits)
        bits.tofile("eoo_tx.f32")
    out = np.zeros(tx.get_Nmf(), np.complex64)
    for rec in frames(stdin, 4 * tx.get_n_floats_in()):
        tx.do_radae_tx(np.frombuffer(rec, np.float32), out)
        stdout.write(out.tobytes())
    eoo = np.zeros(tx.get_Neoo(), np.complex64)
    tx.do_eoo(eoo)
    stdout.write(eoo.tobytes())
    if args.eoo_data_test:                       # trailing silence so the receiver can finish the EOO frame
        stdout.write(np.zeros(tx.get_Neoo(), np.complex64).tobytes())
    stdout.flush()
    tx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
