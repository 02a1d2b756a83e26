#!/bin/bash
# codec bring-up on a GPU box: bit-exact tests of both kernel families, then codec-only and full-pipeline timings (tag = $1)
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_core.py tests/test_gpu_codec_families.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest_codec.log
cat gpurun_out/${tag}_pytest_codec.log
for fam in umma mma; do
  RADE_B200_CODEC=$fam timeout 120 python bench.py --workload codec --streams 1024 --no-cpu-baseline > gpurun_out/${tag}_codec1024_$fam.json 2> gpurun_out/${tag}_codec1024_$fam.err
  RADE_B200_CODEC=$fam timeout 120 python bench.py --workload codec --no-cpu-baseline > gpurun_out/${tag}_codec8192_$fam.json 2> gpurun_out/${tag}_codec8192_$fam.err
done
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/${tag}_full_umma.json 2> gpurun_out/${tag}_full_umma.err
python - <<PY
import json
for f in ["codec1024_umma","codec1024_mma","codec8192_umma","codec8192_mma","full_umma"]:
    try:
        d=json.load(open("gpurun_out/${tag}_%s.json"%f)); print(f, "%.3g F/s"%d["value"], "%.4f ms"%d["ms_per_step"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()}, "e2e %.3g"%d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/${tag}_%s.err"%f).read()[-1500:])
PY
