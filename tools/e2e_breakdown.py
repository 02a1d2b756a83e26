#!/usr/bin/env python
"""wall-clock breakdown of the host-buffer calls of the e2e leg (1024 streams): which call costs what, alone and with the other
side running concurrently.   python tools/e2e_breakdown.py   (needs a B200)"""
import sys, os, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radae_b200 import RadeBatch
from radae_b200.batch import HostLink
from oracle.core import synth_features

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
feats = pin((8, S, 432), torch.float32)
feats[...] = np.transpose(np.tile(synth_features(64, 96, seed=1).reshape(64, 8, 432), (S // 64 + 1, 1, 1))[:S], (1, 0, 2))
brx = RadeBatch(S); btx = RadeBatch(S); bch = RadeBatch(S)
for c in (btx, bch): c.channel_config(EbNodB=3.0, freq_offset_hz=-11.0, doppler_spread_hz=1.0, delay_samples=16, gain=1.0, seed=5)
txs = pin((3, S, 960, 2), torch.float32).view(np.complex64).reshape(3, S, 960)
link = HostLink(brx)
tx = pin((S, 960, 2), torch.float32).view(np.complex64).reshape(S, 960)
rxb = pin((S, 960, 2), torch.float32).view(np.complex64).reshape(S, 960)
def timeit(f, n=20):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
k = [0]
def do_tx(): btx.tx(feats[k[0] % 8], out=tx); k[0] += 1
def do_ch(): link.channel_push(btx, tx); link.rx()
for i in range(14): do_tx(); link.channel_push(btx, tx); link.rx()
print("S =", S)
print("rade_b200_tx (H2D features, encoder, modulator writes tx in place)   %.3f ms" % timeit(do_tx))
t_ch = []
def ch_only():
    t0 = time.perf_counter(); link.channel_push(btx, tx); t_ch.append(time.perf_counter() - t0); link.rx()
t_pair = timeit(ch_only)
print("rade_b200_channel_hostlink (reads tx, writes FIFOs in place)          %.3f ms" % (np.mean(t_ch[2:]) * 1e3))
print("rade_b200_hostlink_rx (reads FIFOs in place, receiver, decoder, D2H)  %.3f ms" % (t_pair - np.mean(t_ch[2:]) * 1e3))
print("rade_b200_channel (tx -> rx host arrays)                              %.3f ms" % timeit(lambda: btx.channel(tx, out=rxb)))
# device-only reference points
d_feat = torch.tensor(feats[0]).cuda(); d_tx = torch.empty((S, 960, 2), device="cuda")
def dev_tx(): btx.tx_dev(d_tx.data_ptr(), d_feat.data_ptr()); btx.synchronize()
print("tx_dev + sync (no PCIe)                                               %.3f ms" % timeit(dev_tx))
# copy-engine reference: 7.86 MB each way
h = pin((S, 960, 2), torch.float32); t_h = torch.from_numpy(h)
def h2d(): d_tx.copy_(t_h, non_blocking=True); torch.cuda.synchronize()
def d2h(): t_h.copy_(d_tx, non_blocking=True); torch.cuda.synchronize()
print("cudaMemcpyAsync H2D 7.86 MB + sync                                    %.3f ms" % timeit(h2d))
print("cudaMemcpyAsync D2H 7.86 MB + sync                                    %.3f ms" % timeit(d2h))
# both sides concurrently (two Python threads)
stop = [False]
def prod():
    while not stop[0]: do_tx()
th = threading.Thread(target=prod); th.start()
print("hostlink_rx + channel_hostlink while rade_b200_tx runs on another thread  %.3f ms" % timeit(ch_only))
stop[0] = True; th.join()
for name, ch in (("three threads", bch), ("two threads (tx + channel on one)", btx)):
    fo, ro = link.duplex_run(btx, ch, feats, 6, txs)
    t0 = time.perf_counter(); link.duplex_run(btx, ch, feats, 40, txs); dt = (time.perf_counter() - t0) / 40
    print("rade_b200_duplex_run, %s: %.3f ms per modem frame -> %.2f M F/s" % (name, dt * 1e3, S * 3 / dt / 1e6))
# per-kernel device times inside the host-buffer calls (library profiler: CUDA events around every launch)
for name, ctx, fn in (("rx side", brx, lambda: (link.channel_push(btx, tx), link.rx())), ("tx side", btx, lambda: (do_tx(), link.channel_push(btx, tx), link.rx()))):
    ctx.profile_enable(True)
    for _ in range(10): fn()
    prof = ctx.profile_read(); ctx.profile_enable(False)
    print(name, {k: round(ms / cnt, 4) for k, (ms, cnt) in prof.items()})
