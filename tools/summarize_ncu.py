"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py r01
"""
import collections, csv, os, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(REPO, "profiles"); os.makedirs(OUT, exist_ok=True)
GO = os.path.join(REPO, "gpurun_out")

# 1. launch list: per-kernel launches, total and share (cold-cache, serialised: compare SHARES)
ll = os.path.join(GO, f"launches_{tag}.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("<unnamed>::", "")
        if name.startswith("void at::") or "elementwise" in name or "Memset" in name:
            name = "torch/" + name[:40]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
    tot = sum(v[1] for k, v in agg.items() if not k.startswith("torch/"))
    with open(os.path.join(OUT, f"{tag}_launch_list_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 600  python bench.py --steps 3 --warmup 3\n")
        f.write(f"# {len(rows)} launches captured; durations are cold-cache and serialised under the profiler: compare shares\n")
        f.write(f"{'kernel':34s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share_of_ours':>13s}\n")
        for k, (n, ns) in agg.items():
            sh = f"{ns / tot:13.3f}" if not k.startswith("torch/") else " " * 13
            f.write(f"{k:34s} {n:8d} {ns / 1e3:10.1f} {ns / 1e3 / n:9.2f} {sh}\n")
    import shutil
    shutil.copy(ll, os.path.join(OUT, f"{tag}_launches.csv"))

# 2. per-kernel --set full reports: key raw metrics + opcode / stall breakdown from the source page
KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg", "smsp__inst_executed.sum")
traffic = {}
for fn in sorted(os.listdir(GO)):
    if not (fn.endswith(f"_{tag}.ncu-rep") and fn.startswith("prof_")):
        continue
    rep = os.path.join(GO, fn)
    kname = fn[len("prof_"):-len(f"_{tag}.ncu-rep")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sr = list(csv.reader(src.splitlines()))
    with open(os.path.join(OUT, f"{tag}_{kname}_ncu_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kname} -s 20 -c 1  python bench.py --steps 3 --warmup 3\n")
        if len(rr) > 2:
            tb = 0.0
            for h, u, v in zip(rr[0], rr[1], rr[2]):
                if h in KEYS or h == "Kernel Name":
                    f.write(f"{h:80s} {u:14s} {v}\n")
                if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tb += float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic[kname] = tb
        if len(sr) > 2:
            hdr, data = sr[1], sr[2:]
            isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
            stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            tot = sum(int(r[isamp] or 0) for r in data) or 1
            byop, byin, st = collections.Counter(), collections.Counter(), collections.Counter()
            for r in data:
                t = r[isrc].split()
                op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
                byop[op] += int(r[isamp] or 0); byin[op] += int(r[iex] or 0)
                for i in stalls:
                    st[hdr[i]] += int(r[i] or 0)
            f.write(f"\n# warp-state samples by opcode (total {tot}); instructions executed (warp level, total {sum(byin.values())})\n")
            for op, c in byop.most_common(14):
                f.write(f"{op:12s} {c:8d} {100 * c / tot:5.1f}%   inst {byin[op]}\n")
            f.write("\n# stall reasons (all samples)\n")
            for k, c in st.most_common(8):
                f.write(f"{k:28s} {c}\n")
            f.write("\n# hottest SASS lines\n")
            for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:12]:
                f.write(f"{r[isamp]:>6s} {r[iex]:>9s}  {r[isrc][:110]}\n")
import json
if traffic:
    json.dump({"how": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, bench.py default workload (1024 streams)",
               "bytes_per_launch": traffic}, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
print("written:", sorted(os.listdir(OUT)))
