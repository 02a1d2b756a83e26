"""Command-line streaming receiver: `python -m radae_b200.radae_rxe < rx.iq.f32 > features.f32`.

Same contract and flags as the reference's script (/root/reference/radae_rxe.py:332-378): complex64 samples on stdin,
read nin (800 | 960 | 1120) at a time; one 432-float feature frame (or, with --bypass_dec, 240 latents) on stdout for
every modem frame decoded in sync; with --eoo_data_test the soft bits of the end-of-over frame are compared with the
seeded test pattern and "PASS" is printed for a bit error rate below 5 %."""
import argparse
import sys
import numpy as np
from .streaming import radae_rx
from .radae_txe import eoo_test_bits


def build_parser():
    p = argparse.ArgumentParser(description="RADE V1 streaming receiver on libradae_b200: IQ.f32 on stdin, features.f32 on stdout")
    p.add_argument("--model_name", type=str, default="", help="RDW or DNNw weight file (default: the embedded model19_check3 weights)")
    p.add_argument("--noauxdata", dest="auxdata", action="store_false", help="not supported by the device path (raises)")
    p.add_argument("-v", type=int, default=2, help="verbosity: 2 prints one status line per call to stderr")
    p.add_argument("--disable_unsync", type=float, default=0.0, help="reference test mode, not supported by the device path (raises when non-zero)")
    p.add_argument("--no_stdout", action="store_false", dest="use_stdout", help="do not write the decoded frames")
    p.add_argument("--foff_err", type=float, default=0.0, help="frequency error added on first sync (0 or 10 Hz = RADE_FOFF_TEST)")
    p.add_argument("--bypass_dec", action="store_true", help="write z_hat (240 floats per frame) instead of features")
    p.add_argument("--eoo_data_test", action="store_true", help="count bit errors in the EOO frame against the seeded test bits")
    p.set_defaults(auxdata=True, use_stdout=True)
    return p


def main(argv=None, stdin=None, stdout=None):
    args = build_parser().parse_args(argv)
    stdin = stdin or sys.stdin.buffer
    stdout = stdout or sys.stdout.buffer
    rx = radae_rx(model_name=args.model_name, auxdata=args.auxdata, v=args.v, disable_unsync=args.disable_unsync,
                  foff_err=args.foff_err, bypass_dec=args.bypass_dec, eoo_data_test=args.eoo_data_test)
    floats_out = np.zeros(rx.get_n_floats_out(), np.float32)
    n_call = 0
    while True:
        nin = rx.get_nin()
        buf = stdin.read(8 * nin)
        if len(buf) != 8 * nin:
            break
        ret = rx.do_radae_rx(np.frombuffer(buf, np.complex64), floats_out)
        n_call += 1
        if args.v >= 2:
            print(f"{n_call:4d} sync: {int(rx.get_sync())} nin: {rx.get_nin():4d} SNRdB: {rx.get_snrdB_3k_est():3d} ret: {ret}", file=sys.stderr)
        if (ret & 1) and args.use_stdout:
            stdout.write(floats_out.tobytes())
        if (ret & 2) and args.eoo_data_test:
            bits = eoo_test_bits(rx.get_Neoo_bits())
            n_err = int(np.sum(floats_out[:bits.size] * bits < 0))
            ber = n_err / bits.size
            print(f"EOO data n_bits: {bits.size} n_errors: {n_err} BER: {ber:5.2f}", file=sys.stderr)
            if ber < 0.05:
                print("PASS", file=sys.stderr)
    stdout.flush()
    rx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
