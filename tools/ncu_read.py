"""Summarise one kernel's ncu exports (tools/ncu_rx.sh): key metrics from the raw page, stall mix and hottest SASS instructions
from the source page.  usage: python tools/ncu_read.py gpurun_out/<tag>_<kernel>"""
import csv, sys
from collections import Counter
base = sys.argv[1]
rows = list(csv.reader(open(base + "_raw.csv")))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:70], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in hdr:
        if any(w in k for w in ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "warps_active.avg.pct", "issue_active.avg.pct", "inst_executed.sum", "pipe_fma_cycles", "pipe_fp64", "pipe_alu", "pipe_xu",
                                "pipe_fmaheavy", "pipe_fmalite", "lsu_wavefronts.avg.pct", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "idc", "imc", "registers_per_thread", "occupancy", "warp_issue_stalled"]):
            if d[k] not in ("", "0", "n/a"):
                print("  %-95s %s" % (k, d[k]))
rows = list(csv.reader(open(base + "_source.csv")))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("samples", tot, {k[6:]: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
ops = Counter(); opn = Counter()
for r in data:
    op = r[1].strip().split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    ops[op.split(".")[0]] += int(r[ix['# Samples']]); opn[op.split(".")[0]] += int(r[ix['Instructions Executed']] or 0)
print("samples by opcode:", ops.most_common(14))
print("executed by opcode:", opn.most_common(14))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}
    print(r[0][-5:], r[1].strip()[:64].ljust(64), r[ix['# Samples']].rjust(5), r[ix['Instructions Executed']].rjust(8), dict(sorted(st.items(), key=lambda x: -x[1])[:3]))
