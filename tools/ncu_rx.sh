#!/bin/bash
# one ncu --set full capture of each named receiver kernel inside the default bench (tag = $1, kernels = $2...), with the metric
# summary and the per-instruction sampling view (SASS) exported as CSV
tag=${1:-r02}; shift
mkdir -p gpurun_out
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 30 -c 1 -o gpurun_out/${tag}_$k -f \
      python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/${tag}_${k}_run.log 2>&1
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_${k}_source.csv 2>/dev/null
  rm -f gpurun_out/${tag}_$k.ncu-rep
done
ls -la gpurun_out | grep ${tag}_
