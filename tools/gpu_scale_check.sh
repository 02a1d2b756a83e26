#!/bin/bash
# weak-scaling check on one multi-GPU box (tag = $1): the driver's own launch line at N = 8, 4, 2, then the in-process multi-GPU test
tag=${1:-r02}
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_${n}gpu.json")); e=d["e2e"]
    print("N=$n value %.4g F/s  %.4f ms/step  e2e %.4g  pipe %.4g" % (d["value"], d["ms_per_step"], e["value"], (e.get("pipe") or {}).get("value", 0)))
except Exception as ex:
    print("N=$n ERR", ex); print(open("gpurun_out/${tag}_bench_${n}gpu.err").read()[-1500:])
PY
done
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
