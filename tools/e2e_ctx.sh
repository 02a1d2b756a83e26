#!/bin/bash
# e2e throughput vs host contexts/threads and OpenMP team size of the host FIFO code
for cfg in "1 16 active" "4 16 active" "4 4 passive" "4 2 passive" "8 2 passive" "2 8 passive"; do
  set -- $cfg
  OMP_NUM_THREADS=$2 OMP_WAIT_POLICY=$3 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-contexts $1 2>/tmp/b.err > /tmp/b.json || tail -3 /tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('contexts $1 omp $2 $3', 'e2e', round(d['e2e']['value']), 'value', round(d['value']))"
done
