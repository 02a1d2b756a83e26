"""CPU numerics study for the round-2 plan: can the 48-row |Dt| refresh of acquisition.check_pilots
(/root/reference/radae/dsp.py:291-300; rx_track_kernel's FFMA2 loop today) run on the tf32 tensor cores?

Emulates tcgen05 kind::tf32 operands (fp32 with the low 13 mantissa bits ignored) on a real receive buffer from the golden
fixtures and compares the row sums sigma_r is built from against float64:
  fp32     element 2.9e-07  row sum 6.5e-08  mean 1.8e-09      <- what the CUDA path does today
  1xTF32   element 7.9e-04  row sum 8.8e-04  mean 6.9e-04      <- biased (truncation shrinks every product): unusable
  3xTF32   element 3.6e-07  row sum 4.0e-07  mean 2.6e-07      <- hi/lo split of both operands, 3 products: fp32-class
Run: python tools/tf32_refresh_study.py"""
import os, sys
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import dsp as od


def tf32(a):
    u = np.asarray(a, np.float32).copy().view(np.uint32)
    u &= np.uint32(0xFFFFE000)
    return u.view(np.float32)


def split(a):
    h = tf32(a)
    return h, tf32((a - h).astype(np.float32))


def main():
    c = od.consts()
    g = np.load(os.path.join(REPO, "tests", "golden", "rx_awgn_1dB.npz"))
    x = g["rx_in"][4000:4000 + od.RXBUF].astype(np.complex64)
    pw = np.asarray(c.p_w)                                               # [160][40]
    rows = np.array([np.conj(x[t:t + od.M]) for t in od.refresh_rows(7)])  # [48][160]
    truth = rows.astype(np.complex128) @ pw.astype(np.complex128)
    d = lambda p, q: p.astype(np.float64) @ q.astype(np.float64)         # exact products, wide accumulation
    kinds = {
        "fp32": lambda A, B: (A @ B).astype(np.float32),
        "1xTF32": lambda A, B: d(tf32(A), tf32(B)).astype(np.float32),
        "3xTF32": lambda A, B: (lambda ah, al, bh, bl: (d(ah, bh) + d(ah, bl) + d(al, bh)).astype(np.float32))(*split(A), *split(B)),
    }
    Ar, Ai, Br, Bi = rows.real.copy(), rows.imag.copy(), pw.real.copy(), pw.imag.copy()
    for name, mm in kinds.items():
        D = (mm(Ar, Br) - mm(Ai, Bi)).astype(np.float64) + 1j * (mm(Ar, Bi) + mm(Ai, Br)).astype(np.float64)
        rs, rt = np.abs(D).sum(1), np.abs(truth).sum(1)
        print(f"{name:7s} element {np.abs(D - truth).max() / np.abs(truth).max():.1e}  row sum {np.max(np.abs(rs - rt) / rt):.1e}  "
              f"mean {abs(rs.sum() - rt.sum()) / rt.sum():.1e}")


if __name__ == "__main__":
    main()
