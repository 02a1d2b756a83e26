"""Timeline of the loop-back step on one GPU: which kernel runs when, on its own stream, nothing serialised
(rade_b200_timeline_begin / _read: CUDA event pairs on the launching streams).  usage: python tools/step_timeline.py [streams] [steps]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radae_b200 import RadeBatch, _capi

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = _capi.lib()
b = RadeBatch(S)
b.channel_config(EbNodB=3.0, freq_offset_hz=-11.0, freq_offset_spread_hz=0.0, doppler_spread_hz=1.0, delay_samples=16, gain=1.0, seed=77)
b.pipeline_enable(True)
rng = np.random.default_rng(0)
feats = np.zeros((4, S, 432), np.float32); feats.reshape(4, S, 12, 36)[..., :20] = rng.standard_normal((4, S, 12, 20)) * 0.5
d_in = [torch.tensor(feats[i]).cuda() for i in range(4)]
d_fo = torch.zeros((S, 432), device="cuda"); d_ret = torch.zeros(S, dtype=torch.int32, device="cuda"); d_eoo = torch.zeros((S, 180), device="cuda")
for k in range(45):                                  # acquisition + settle
    b.loopback_step_dev(d_in[k % 4].data_ptr(), d_fo.data_ptr(), d_ret.data_ptr(), d_eoo.data_ptr())
b.synchronize()
print("streams with valid output:", int((d_ret.cpu().numpy() & 1).sum()), "of", S)
lib.rade_b200_timeline_begin(b.h)
for k in range(K):
    b.loopback_step_dev(d_in[k % 4].data_ptr(), d_fo.data_ptr(), d_ret.data_ptr(), d_eoo.data_ptr())
cap = 64 * K
kid = np.zeros(cap, np.int32); t0 = np.zeros(cap, np.float32); t1 = np.zeros(cap, np.float32)
n = lib.rade_b200_timeline_read(b.h, kid.ctypes.data, t0.ctypes.data, t1.ctypes.data, cap)
lib.rade_b200_profile_kernel_name.restype = ctypes.c_char_p
rows = sorted((float(t0[i]), float(t1[i]), lib.rade_b200_profile_kernel_name(int(kid[i])).decode()) for i in range(n))
print("%-24s %9s %9s %9s   (us since the first step was submitted; 'start' = the launching stream reached the kernel)" % ("kernel", "start", "end", "length"))
for a, e, name in rows:
    print("%-24s %9.1f %9.1f %9.1f   %s" % (name, a * 1e3, e * 1e3, (e - a) * 1e3, " " * int(a * 1e3 / 8) + "#" * max(1, int((e - a) * 1e3 / 8))))
b.close()
