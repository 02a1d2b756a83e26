#!/bin/bash
# quick codec iteration on a GPU box: bit-exact check (both families), timeline of CTA 0, codec-only timing at 1024 and 8192 streams
tag=${1:-r02}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_codec_families.py -m gpu -x -q 2>&1 | tail -12
python tools/codec_trace.py enc 1024 > gpurun_out/${tag}_trace_enc.txt 2>&1
python tools/codec_trace.py dec 1024 > gpurun_out/${tag}_trace_dec.txt 2>&1
head -1 gpurun_out/${tag}_trace_enc.txt gpurun_out/${tag}_trace_dec.txt
for s in 1024 8192; do
  timeout 120 python bench.py --workload codec --streams $s --no-cpu-baseline --no-e2e > gpurun_out/${tag}_codec${s}.json 2> gpurun_out/${tag}_codec${s}.err
done
python - <<PY
import json
for f in ["codec1024","codec8192"]:
    try:
        d=json.load(open("gpurun_out/${tag}_%s.json"%f)); print(f, "%.3g F/s"%d["value"], "%.4f ms"%d["ms_per_step"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/${tag}_%s.err"%f).read()[-1500:])
PY
