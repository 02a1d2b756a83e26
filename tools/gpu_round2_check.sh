#!/bin/bash
# full GPU check of a round-2 state: all GPU tests, then the three bench workloads (tag = $1)
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest_gpu.log
cat gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err
timeout 300 python bench.py --workload codec > gpurun_out/${tag}_bench_codec.json 2> gpurun_out/${tag}_bench_codec.err
timeout 300 python bench.py --workload rx-search --no-cpu-baseline > gpurun_out/${tag}_bench_rxsearch.json 2> gpurun_out/${tag}_bench_rxsearch.err
python - <<PY
import json
for f in ["full","codec","rxsearch"]:
    try:
        d=json.load(open("gpurun_out/${tag}_bench_%s.json"%f))
        print(f, "value %.3g F/s"%d["value"], "%.4f ms/step"%d["ms_per_step"], "e2e %.3g"%(d["e2e"] or {}).get("value",0), {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
        print("   cpu_baseline", json.dumps(d.get("cpu_baseline"))[:400])
        print("   feat_rms_err", json.dumps(d.get("feat_rms_err"))[:700])
        print("   e2e", json.dumps(d.get("e2e"))[:500])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/${tag}_bench_%s.err"%f).read()[-2500:])
PY
