#!/usr/bin/env python
"""Timeline of CTA 0 of the tcgen05 codec kernels (debug hook rade_b200_debug_trace_*): who waits for whom.
    python tools/codec_trace.py [enc|dec] [streams]        (needs a B200)
Prints, per step and layer, clock64 stamps relative to the first one: issuer (op start / dependency satisfied / last MMA issued),
epilogue warp 0 (accumulator ready / outputs published), float warp 0 (segment available / consumed), int8 producer (chunk issued)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radae_b200 import RadeBatch, _capi
from oracle.core import pack_enc_input, synth_features

which = sys.argv[1] if len(sys.argv) > 1 else "enc"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = 3
lib = _capi.lib()
b = RadeBatch(S)
x = pack_enc_input(synth_features(min(S, 64), 4 * T, seed=3))
x = np.ascontiguousarray(np.tile(x, ((S + 63) // 64, 1, 1))[:S])
z = b.core_encode(x)
for _ in range(2):
    b.core_encode(x) if which == "enc" else b.core_decode(z)
N = 16384
lib.rade_b200_debug_trace_enable(b.h, N)
b.core_encode(x) if which == "enc" else b.core_decode(z)
tr = np.zeros(N, np.int64)
assert lib.rade_b200_debug_trace_read(b.h, tr.ctypes.data, N) == 0
t0 = tr[:1024][tr[:1024] > 0].min() if (tr[:1024] > 0).any() else tr[2048:8192][tr[2048:8192] > 0].min()
t0 = min(t0, tr[2048:8192][tr[2048:8192] > 0].min())
rel = lambda v: int(v - t0) if v > 0 else -1
w = 0 if which == "enc" else 1
n_recs = lib.rade_b200_debug_codec_program(w, None, 0)
recs = np.zeros((n_recs, 13), np.int32); lib.rade_b200_debug_codec_program(w, recs.ctypes.data, n_recs)
nlayers = 10 if which == "enc" else 15
print(f"{which} S={S}: total span {int(tr[2048:8192].max() - t0)} cycles, {n_recs} records per step")
print("  kernel entry %d, set-up done %d, all roles done %d, state stored %d (same origin)" % tuple(rel(tr[8000 + i]) for i in range(4)))
for t in range(T):
    print(f"--- step {t}")
    for i in range(n_recs):
        if recs[i][9] >= 0:
            a, d = rel(tr[(t * 128 + i) * 2]), rel(tr[(t * 128 + i) * 2 + 1])
            print(f"  I rec{i:3d} waits for layer {recs[i][9]:2d}: reached {a:7d}, satisfied {d:7d} (idle {d - a})")
    print("  I accumulator commits:", [rel(tr[1024 + t * 16 + l]) for l in range(nlayers)])
    for l in range(nlayers):
        a, d = rel(tr[2048 + (t * 16 + l) * 2]), rel(tr[2048 + (t * 16 + l) * 2 + 1])
        print(f"  E layer{l:2d}: acc ready {a:7d}  published {d:7d}  ({d - a} cycles)")
    for j in range(10):
        a, d = rel(tr[4096 + (t * 16 + j) * 2]), rel(tr[4096 + (t * 16 + j) * 2 + 1])
        print(f"  F seg{j:2d}: available {a:7d}  consumed {d:7d}  ({d - a} cycles)")
    pc = [rel(tr[6144 + t * 64 + c]) for c in range(64)]
    print("  P int8 stage copy issue times:", [p for p in pc if p >= 0])
b.close()
