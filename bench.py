#!/usr/bin/env python
"""bench.py — RADE hot path throughput on B200(s):  40 ms modem frames/s through
    features -> core encoder -> OFDM mod -> HF channel (MPP) -> link -> acquisition/demod/EQ -> core decoder -> features

    python bench.py --gpus N --steps K --warmup W                 our arm (CUDA, one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   CPU arm: the oracle port of the same pipeline on host cores

One step = one 120 ms modem frame (3 forty-ms frames F) for every stream of the batch.  Workload = BASELINE.json
configs[2] ("Full TX+OFDM+MPP-multipath+demod+RX pipeline, 1024 concurrent streams, 1xB200"), the configuration the
headline metric (enc->OFDM->chan->demod->dec frames/s) is quoted on; weak scaling: 1024 streams per GPU.
`--workload codec` runs configs[1] (CoreEncoder+CoreDecoder only, 8192 streams) instead.

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how every field is derived.
"""
import argparse, json, os, sys, threading, time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np

F_PER_STEP = 3
METRIC = "40ms_modem_frames_per_s_enc_ofdm_chan_demod_dec"
# algorithmic HBM bytes per stream per modem frame (= per launch per stream), DESIGN.md §Kernels
KERNEL_BYTES = {
    "core_encoder_kernel": 1728 + 960 + 2 * 3040,       # features in, z out, state load+store
    "ofdm_mod_kernel": 960 + 7680,
    "channel_stream_kernel": 7680 + 7680 + 2 * 528,
    "link_push_kernel": 7680 + 7680,
    "link_pop_kernel": 7680 + 7680,
    "rx_bpf_kernel": 7680 + 7680 + 2 * 816,
    "rx_detect_kernel": 16896 + 7680,                   # ring read once, row sums written
    "rx_track_kernel": 16896 + 7680 + 384,              # ring read once, row sums read + 48x2 refreshed
    "rx_demod_kernel": 9216 + 960,
    "rx_finish_kernel": 7680 + 256,
    "core_decoder_kernel": 960 + 1728 + 2 * 2672,
}
CODEC_BYTES_PER_F = {"core_encoder_kernel": 336 + 320, "core_decoder_kernel": 320 + 336}   # + state once per launch


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons during the timed region (nvidia-smi's clocks line, via NVML)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), 0
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {getattr(nv, n): n for n in dir(nv) if n.startswith("nvmlClocksThrottleReason") or n.startswith("nvmlClocksEventReason")}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if isinstance(bit, int) and bit and (r & bit) == bit and "None" not in n and "All" not in n:
                        self.reasons.add(n.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1)
        tidy = sorted({{"SwPowerCap": "sw_power_cap", "HwSlowdown": "hw_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown",
                        "SwThermalSlowdown": "sw_thermal_slowdown", "GpuIdle": "gpu_idle", "ApplicationsClocksSetting": "applications_clocks_setting",
                        "HwPowerBrakeSlowdown": "hw_power_brake_slowdown", "SyncBoost": "sync_boost",
                        "DisplayClockSetting": "display_clock_setting", "UserDefinedClocks": "applications_clocks_setting"}.get(r, r) for r in self.reasons})
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": float(self.sm_max) if self.sm_max else None,
                "reasons": [r for r in tidy if r != "gpu_idle"]}


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """one host process: n_streams independent streams of the oracle pipeline for n_frames modem frames"""
    seed, n_streams, n_frames, warm = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import dsp as od
    from oracle.core import CoreOraclePort, CoreOracleRef, synth_features
    Core = CoreOracleRef if CoreOracleRef.available("int8") else None
    rng = np.random.default_rng(seed)
    sigma = od.ebno_sigma(3.0)
    streams = []
    for s in range(n_streams):
        core = Core("int8", 1) if Core else CoreOraclePort(n_streams=1)
        streams.append((od.RadaeTx(core), od.RadaeRx(core), np.zeros(0, np.complex64)))
    feats = synth_features(n_streams, 12 * (n_frames + warm), seed=seed).reshape(n_streams, n_frames + warm, 432)
    t0 = None
    nz = lambda n: ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)).astype(np.complex64)
    for f in range(n_frames + warm):
        if f == warm:
            t0 = time.perf_counter()
        for s in range(n_streams):
            tx, rx, fifo = streams[s]
            y, _ = tx.do_radae_tx(feats[s, f])
            n = np.arange(960)
            g1 = (0.7 * np.exp(1j * 2 * np.pi * 0.5 * (n + 960 * f) / 8000)).astype(np.complex64)
            g2 = (0.7 * np.exp(-1j * 2 * np.pi * 0.3 * (n + 960 * f) / 8000)).astype(np.complex64)
            y = od.channel(y, g1, g2, 16, 1.0, -11.0, 0.0, sigma, nz(960))
            fifo = np.concatenate([fifo, y])
            if len(fifo) >= rx.nin:
                k = rx.nin
                rx.do_radae_rx(fifo[:k]); fifo = fifo[k:]
            streams[s] = (tx, rx, fifo)
    return time.perf_counter() - t0, "reference" if Core else "port"


def cpu_pipeline_rate(n_frames, warm=2, streams_per_proc=1, procs=None):
    """frames F per second of the CPU oracle pipeline using all host cores (one process per core, 1 thread each)"""
    import multiprocessing as mp
    procs = procs or os.cpu_count()
    ctx = mp.get_context("spawn")
    # one single-threaded process per core: the thread-pool sizes of numpy's BLAS / OpenMP are fixed when the child imports
    # numpy, so the limits must be in the environment the children inherit (setting them inside the worker is too late and
    # lets every process start one thread per core)
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"
    try:
        with ctx.Pool(procs) as pool:
            res = pool.map(_cpu_worker, [(1000 + i, streams_per_proc, n_frames, warm) for i in range(procs)])
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    slowest = max(r[0] for r in res)
    total_F = procs * streams_per_proc * n_frames * F_PER_STEP
    return total_F / slowest, procs, res[0][1], f"{procs} procs x {streams_per_proc} stream x {n_frames} modem frames (after {warm} warm-up frames, i.e. receiver in sync), MPP-like 2-path, Eb/No 3 dB, -11 Hz"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    per_step = []
    frames_per_sample = args.ref_frames
    for i in range(args.warmup + args.steps):
        rate, procs, kind, sample = cpu_pipeline_rate(frames_per_sample, warm=8)
        if i >= args.warmup:
            per_step.append(rate)
    v = float(np.mean(per_step))
    sample_units = cores * frames_per_sample * F_PER_STEP
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * sample_units / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8xint8->int32 + f32", "data": "synthetic",
            "config": {"workload": "full TX+OFDM+MPP+demod+RX pipeline (BASELINE configs[2]) — CPU oracle port on host cores; each step = a bounded sample",
                       "sample": sample},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": f"DSP = numpy restatement (oracle/dsp.py); core codec = {'reference rade_enc.c/rade_dec.c + nnet shim (oracle/_ref)' if kind == 'reference' else 'oracle/core_oracle.c'}"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    from radae_b200 import RadeBatch, _capi, rdw, multigpu
    # NCCL prints its version banner on stdout when the first communicator is created; stdout carries ONE JSON line, so
    # file descriptor 1 points at stderr while the process group comes up and the weights are broadcast
    sys.stdout.flush()
    saved_fd = os.dup(1); os.dup2(2, 1)
    try:
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # one NCCL broadcast of the weight blob at start-up (the only collective on this path)
        blob = open(rdw.default_weights_path(), "rb").read() if rank == 0 else None
        blob = multigpu.broadcast_weights(dist if world > 1 else None, rank, blob, device="cuda")
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
    finally:
        sys.stdout.flush(); os.dup2(saved_fd, 1); os.close(saved_fd)

    codec_only = args.workload == "codec"
    S = args.streams or (8192 if codec_only else 1024)
    K, W = args.steps, args.warmup
    b = RadeBatch(S, device=local, weights=blob)
    ext = torch.cuda.ExternalStream(b.cuda_stream, device=torch.device("cuda", local))
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")           # 2x the 126 MB L2

    from oracle.core import synth_features                                          # input generator only (numpy)
    n_feat_frames = 8
    base = synth_features(min(S, 64), 12 * n_feat_frames, seed=1234 + 1000 * rank).reshape(-1, n_feat_frames, 432)
    feats_host = np.ascontiguousarray(np.tile(base, ((S + base.shape[0] - 1) // base.shape[0], 1, 1))[:S])
    feats_host += (np.random.default_rng(rank).standard_normal(feats_host.shape) * 0.01).astype(np.float32) * (feats_host != 0)

    if codec_only:
        T = 3
        x = torch.tensor(np.ascontiguousarray(np.concatenate([feats_host.reshape(S, n_feat_frames, 12, 36)[..., :20],
                         -np.ones((S, n_feat_frames, 12, 1), np.float32)], axis=-1).reshape(S, n_feat_frames, 3, 84))).cuda()
        z = torch.empty((S, T, 80), device="cuda"); fo = torch.empty((S, T, 84), device="cuda")

        def step(k):
            b.core_encode_dev(z.data_ptr(), xs[k % n_feat_frames].data_ptr(), T)
            b.core_decode_dev(fo.data_ptr(), z.data_ptr(), T)
        xs = [x[:, i].contiguous() for i in range(n_feat_frames)]
    else:
        b.channel_config(EbNodB=3.0, freq_offset_hz=-11.0, freq_offset_spread_hz=0.0, doppler_spread_hz=1.0,
                         delay_samples=16, gain=1.0, seed=77 + rank)
        d_feats = [torch.tensor(feats_host[:, i]).cuda() for i in range(n_feat_frames)]
        d_tx = torch.empty((S, 960, 2), device="cuda"); d_ch = torch.empty((S, 960, 2), device="cuda")
        d_rxin = torch.zeros((S, 1120, 2), device="cuda"); d_act = torch.zeros(S, dtype=torch.uint8, device="cuda")
        d_fo = torch.zeros((S, 432), device="cuda"); d_ret = torch.zeros(S, dtype=torch.int32, device="cuda")
        d_eoo = torch.zeros((S, 180), device="cuda")

        # Software pipeline over frames: the transmitter side of frame k+1 (core encoder, modulator, channel -> link FIFO) runs on
        # a second CUDA stream concurrently with the receiver side of frame k (pop nin[s], DSP, core decoder); fork / join are
        # INSIDE every timed step, so a step still contains one full TX and one full RX pass for every stream.
        pipelined = not args.no_pipeline
        b.pipeline_enable(pipelined)

        def step(k):
            # = fork; tx_channel_link_dev (core encoder; OFDM modulator + HF channel -> per-stream FIFO in one kernel);
            # rx_link_dev (pop nin[s], receiver DSP, core decoder); join — 8 kernel launches
            b.loopback_step_dev(d_feats[(k + 1) % n_feat_frames].data_ptr(), d_fo.data_ptr(), d_ret.data_ptr(), d_eoo.data_ptr())
        # prime the FIFO with frame 0 so that the receiver always has the frame the transmitter produced one step earlier
        b.tx_dev(d_tx.data_ptr(), d_feats[0].data_ptr()); b.channel_link_dev(d_tx.data_ptr()); b.pipeline_join()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- acquisition phase (not timed): run until (almost) every stream is in sync so the timed region is steady state
    pre = 0
    if not codec_only:
        for k in range(40):
            step(pre); pre += 1
            if k >= 8 and k % 4 == 0:
                if np.mean([s.state == 2 for s in b.rx_status()]) > 0.98:
                    break
    for k in range(W):
        step(pre + k)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = b.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for k in range(K):
        with torch.cuda.stream(ext):
            l2_flush.zero_()                                    # evict the previous step's working set from L2
            ev[k][0].record(ext)
        step(pre + W + k)
        with torch.cuda.stream(ext):
            ev[k][1].record(ext)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.result()
    launches = b.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(c) for a, c in ev)
    tmax = torch.tensor([dev_ms], device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tmax.item())
    total_F = S * world * F_PER_STEP * K
    value = total_F / (dev_ms_max / 1000.0)
    sync_frac = None if codec_only else float(np.mean([s.state == 2 for s in b.rx_status()]))

    # ---- the same K steps without the frame pipeline (everything on one stream), for reference
    serial_ms = None
    if not codec_only and not args.no_pipeline:
        b.synchronize(); b.pipeline_enable(False)
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k in range(K):
            with torch.cuda.stream(ext):
                l2_flush.zero_(); ev2[k][0].record(ext)
            step(pre + W + K + k)
            with torch.cuda.stream(ext):
                ev2[k][1].record(ext)
        barrier()
        serial_ms = sum(a.elapsed_time(c) for a, c in ev2) / K
    # ---- per-kernel pass (CUDA events around every launch, inside the library, same steps; after the timed region)
    b.profile_enable(True)
    for k in range(K):
        with torch.cuda.stream(ext):
            l2_flush.zero_()
        step(pre + W + K + k)
    prof = b.profile_read()
    b.profile_enable(False)
    tot_prof = sum(ms for ms, _ in prof.values())
    hbm_peak, peak_src = measured_peaks()
    kernels = {}
    for name, (ms, cnt) in prof.items():
        per_launch_ms = ms / cnt
        if codec_only:
            byts = S * (CODEC_BYTES_PER_F[name] * 3 + (2 * 3040 if "encoder" in name else 2 * 2672))
        else:
            byts = S * KERNEL_BYTES.get(name, 0)
        kernels[name] = {"ms_per_launch": round(per_launch_ms, 4), "share": round(ms / tot_prof, 4),
                         "algorithmic_bytes_per_launch": byts, "achieved_gbs": round(byts / per_launch_ms / 1e6, 2),
                         "frac_of_hbm": round(byts / per_launch_ms / 1e6 / hbm_peak, 5)}
    dom = max(kernels, key=lambda n: kernels[n]["share"])
    traffic = None
    try:                                                  # DRAM bytes per launch from the committed ncu --set full capture
        import glob
        for fn in sorted(glob.glob(os.path.join(REPO, "profiles", "r*_traffic.json"))):
            traffic = json.load(open(fn))["bytes_per_launch"].get(dom, traffic)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_of_hbm"], "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                "share_of_step": kernels[dom]["share"],
                "note": "compute/latency-bound at this batch: algorithmic bytes are tiny next to the int8/fp32/fp64 math (DESIGN.md §Roofline)"}

    # ---- end to end through the host-buffer C ABI (pinned host in, host out, every step), wall clock
    e2e = None
    if (rank == 0 or world > 1) and not args.no_e2e:
        e2e = run_e2e(b, S, feats_host, codec_only, K, world, dist if world > 1 else None, torch,
                      n_ctx=args.e2e_contexts, weights=blob, device=local, serial=args.e2e_serial)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, procs, kind, sample = cpu_pipeline_rate(200, warm=8)
        cpu_base = {"value": rate, "unit": "frames/s", "cores": procs, "kind": "port", "sample": sample,
                    "note": "numpy DSP restatement + " + ("reference rade_enc.c/rade_dec.c on the nnet shim" if kind == "reference" else "C core port")}

    try:
        # the math-pipe figure SURVEY.md §8(d) asks for next to the HBM fraction: F/s x FLOP per F / peak, with the survey's own
        # work table (enc+dec 1 838 720 MAC per F; full synced pipeline ~2.95 M MAC per F) against the fp32 CUDA-core peak of
        # this device (SMs x 128 lanes x 2 flop x SM clock); per GPU, so it does not change with the number of ranks
        props = torch.cuda.get_device_properties(local)
        mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        peak_tf = props.multi_processor_count * 128 * 2 * float(mhz) * 1e6 / 1e12
        flop_per_F = 2 * (1838720 if codec_only else 2950000)
        ach_tf = (value / world) * flop_per_F / 1e12
        roofline["math"] = {"flop_per_frame": flop_per_F, "achieved_tflops": round(ach_tf, 2), "peak_tflops": round(peak_tf, 1),
                            "frac": round(ach_tf / peak_tf, 4),
                            "peak": "fp32 FMA peak of one GPU at the sampled SM clock; the codec's MACs actually run as int8 IMMA"}
    except Exception:
        pass
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int8xint8->int32 (codec) + f32/f64 (DSP)", "data": "synthetic",
                "config": {"workload": ("CoreEncoder+CoreDecoder only, %d streams x 3 steps per launch (BASELINE configs[1])" % S) if codec_only else
                           ("full TX+OFDM+MPP+demod+RX pipeline, %d concurrent streams per GPU (BASELINE configs[2]); MPP 1 Hz/2 ms, Eb/No 3 dB, -11 Hz" % S),
                           "streams_per_gpu": S, "frames_per_step_per_stream": F_PER_STEP, "l2": "flushed (256 MB memset) before every timed step",
                           "timing": "CUDA events on the library's stream around each step, summed; max over ranks",
                           "sync_fraction": sync_frac, "acquisition_steps_before_timing": pre,
                           "pipeline": None if codec_only else ("off (one stream)" if args.no_pipeline else
                                       "TX side of frame k+1 on a second CUDA stream, concurrent with the RX side of frame k; fork/join inside every timed step"),
                           "ms_per_step_unpipelined": serial_ms},
                "gpu_launches": int(launches), "wall_s": t_wall, "clocks": clocks, "roofline": roofline, "kernels": kernels,
                "e2e": e2e, "cpu_baseline": cpu_base}
        print(json.dumps(line))
    b.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(b, S, feats_host, codec_only, K, world, dist, torch, n_ctx=2, weights=None, device=0, serial=False):
    """same metric through the host-pointer C ABI: pinned host features in, host features out, all copies timed; every call
    is the synchronous reference-style call.  Full pipeline: like the reference's `radae_tx | ch | radae_rx` (two programs
    joined by a pipe) the transmitter side (rade_b200_tx, rade_b200_channel, FIFO push) and the receiver side (FIFO gather +
    rade_b200_rx) run on two host threads with a context each, joined by the pinned host sample FIFO; `--e2e-serial` runs
    the four calls back to back on one thread instead.  n_ctx > 1 splits the streams over several such pairs."""
    import threading
    from radae_b200 import RadeBatch
    from radae_b200.batch import HostLink
    # host threads of the C-side sample FIFOs: this rank's share of the cores (torchrun exports OMP_NUM_THREADS=1)
    duplex = not codec_only and not serial
    os.environ.setdefault("RADE_B200_HOST_THREADS", str(max(1, min(16, (os.cpu_count() or 1) // max(1, world) // max(1, n_ctx) // (2 if duplex else 1)))))
    n_feat_frames = feats_host.shape[1]
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    bounds = [(S * i) // n_ctx for i in range(n_ctx + 1)]
    ctxs = []
    for i in range(n_ctx):
        lo, hi = bounds[i], bounds[i + 1]
        n = hi - lo
        c = {"b": RadeBatch(n, device=device, weights=weights), "n": n}
        fh = feats_host[lo:hi]
        if codec_only:
            x = np.ascontiguousarray(np.concatenate([fh.reshape(n, n_feat_frames, 12, 36)[..., :20],
                                     -np.ones((n, n_feat_frames, 12, 1), np.float32)], axis=-1).reshape(n, n_feat_frames, 3, 84))
            c["x"] = [np.ascontiguousarray(x[:, j]) for j in range(n_feat_frames)]
        else:
            c["btx"] = RadeBatch(n, device=device, weights=weights) if duplex else c["b"]     # transmitter-side context
            c["btx"].channel_config(EbNodB=3.0, freq_offset_hz=-11.0, doppler_spread_hz=1.0, delay_samples=16, gain=1.0, seed=5 + i)
            c["link"] = HostLink(c["b"])
            c["tx"] = pin((n, 960, 2), torch.float32).view(np.complex64).reshape(n, 960)
            c["rx"] = pin((n, 960, 2), torch.float32).view(np.complex64).reshape(n, 960)
            c["feats"] = []
            for j in range(n_feat_frames):
                a = pin((n, 432), torch.float32); a[...] = fh[:, j]; c["feats"].append(a)
        ctxs.append(c)

    def step(c, k):
        if codec_only:
            z = c["b"].core_encode(c["x"][k % n_feat_frames])
            c["b"].core_decode(z)
        else:
            tx_side(c, k)
            c["link"].rx()                                              # gather nin[s] per stream, H2D rx_in, D2H features/ret/eoo/nin

    def tx_side(c, k):
        c["btx"].tx(c["feats"][k % n_feat_frames], out=c["tx"])        # H2D features (pinned), D2H tx samples
        c["btx"].channel(c["tx"], out=c["rx"])                          # H2D tx, D2H rx (the channel is a simulator outside rade_api.h)
        c["link"].push(c["rx"])                                         # pinned host FIFO, C/OpenMP

    def run(c, k0, n):
        if not duplex:
            for k in range(k0, k0 + n):
                step(c, k)
            return
        # two host threads joined by the sample FIFO; the transmitter may run at most two frames ahead of the receiver
        filled, space = threading.Semaphore(0), threading.Semaphore(2)

        def producer():
            for k in range(k0, k0 + n):
                space.acquire(); tx_side(c, k); filled.release()
        th = threading.Thread(target=producer); th.start()
        for k in range(k0, k0 + n):
            filled.acquire(); c["link"].rx(); space.release()
        th.join()

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=n_ctx)                      # persistent host threads, one per context

    def run_all(k0, n):
        if n_ctx == 1:
            run(ctxs[0], k0, n)
        else:
            list(pool.map(lambda c: run(c, k0, n), ctxs))

    if codec_only:
        h2d = S * (3 * 84 * 4 + 3 * 80 * 4); d2h = S * (3 * 80 * 4 + 3 * 84 * 4)
    else:
        h2d = S * (432 * 4 + 960 * 8 + 1120 * 8 + 1); d2h = S * (960 * 8 + 960 * 8 + 432 * 4 + 4 + 180 * 4 + 4)
    run_all(0, 12 if not codec_only else 2)                           # warm-up incl. acquisition
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    run_all(100, K)
    for c in ctxs:
        c["b"].synchronize()
        if "btx" in c: c["btx"].synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    pool.shutdown()
    for c in ctxs:
        if "link" in c: c["link"].close()
        if c.get("btx") is not None and c["btx"] is not c["b"]: c["btx"].close()
        c["b"].close()
    return {"value": S * world * F_PER_STEP * K / dt, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": K, "contexts": n_ctx, "host_threads": n_ctx * (2 if duplex else 1),
            "timing": ("host wall clock around K modem frames per stream; transmitter side (rade_b200_tx, rade_b200_channel, FIFO push) and "
                       "receiver side (FIFO gather, rade_b200_rx) on two host threads joined by the pinned host FIFO, like radae_tx | ch | radae_rx; "
                       "every call synchronous; max over ranks") if duplex else
                      "host wall clock around K synchronous steps (tx, channel, push, rx back to back on one host thread), pinned buffers, host FIFO in C; max over ranks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="full", choices=["full", "codec"])
    ap.add_argument("--streams", type=int, default=0, help="streams per GPU (default 1024 full / 8192 codec)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="development runs: skip the end-to-end leg (the line then has e2e = null)")
    ap.add_argument("--ref-frames", type=int, default=200, help="--impl reference: modem frames per process and step (bounded sample)")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e leg: tx, channel, push, rx back to back on ONE host thread")
    ap.add_argument("--no-pipeline", action="store_true", help="run TX and RX of a frame back to back on one stream")
    ap.add_argument("--e2e-contexts", type=int, default=1, help="host threads / contexts serving the streams in the e2e leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
