#!/bin/bash
# 8- vs 16-stream codec tiles: bit-exact tests (the large-batch tests run the 16-stream kernels), then timings for both widths
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_core.py tests/test_gpu_codec_families.py -m gpu -x -q 2>&1 | tail -6
RADE_B200_CODEC_NS=16 timeout 300 python -m pytest tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -4
for ns in 8 16; do for s in 1024 2048 8192; do
  RADE_B200_CODEC_NS=$ns timeout 120 python bench.py --workload codec --streams $s --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NS=$ns S=$s %.3g F/s' % d['value'], {k:v['ms_per_launch'] for k,v in d['kernels'].items()})"
done; done
