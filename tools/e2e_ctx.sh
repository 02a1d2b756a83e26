#!/bin/bash
# e2e throughput vs number of host contexts/threads serving the 1024 streams (each context gets cores/contexts FIFO threads)
for c in 1 2 4; do
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-contexts $c 2>/tmp/b.err > /tmp/b.json || tail -3 /tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('contexts', $c, 'e2e', round(d['e2e']['value']), 'value', round(d['value']))"
done
