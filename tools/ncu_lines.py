"""Per-CUDA-source-line stall samples from an .ncu-rep captured with --import-source on (kernels built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep [top_n]
"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines = [r for r in rows if r and r[0].isdigit() and len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp] or 0) for r in lines) or 1
print(f"total samples {tot}")
for r in sorted(lines, key=lambda r: -int(r[isamp] or 0))[:top]:
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stalls), reverse=True)[:3]
    print(f"{r[0]:>5s} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}%  inst {r[iex]:>9s}  {', '.join(f'{h}:{c}' for c, h in st if c)}  | {r[1].strip()[:90]}")
