#!/bin/bash
# timing experiment: rx_track with the refresh (1) / refine (2) / both (3) compute phases skipped (results are wrong then)
for d in 0 1 2 3; do
  RADE_B200_TRACK_DEBUG=$d timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/tmp/b.err > /tmp/b.json || tail -3 /tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('dbg', $d, d['kernels']['rx_track_kernel']['ms_per_launch'], d['config']['sync_fraction'])"
done
