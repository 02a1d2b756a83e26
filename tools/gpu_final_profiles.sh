#!/bin/bash
# re-capture of the kernels that changed after the r03c profile set (tag = $1): launch list + --set full of rx_demod, rx_finish, rx_detect
tag=${1:-r03g}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/${tag}_launchlist_run.log 2>&1
for k in rx_demod_kernel rx_finish_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -o gpurun_out/prof_${k}_${tag} -f \
      python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 3 --warmup 3 > gpurun_out/${tag}_${k}_run.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:rx_detect_kernel -s 6 -c 1 -o gpurun_out/prof_rx_detect_kernel_${tag} -f \
    python bench.py --workload rx-search --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/${tag}_rx_detect_run.log 2>&1
python tools/summarize_ncu.py ${tag} > gpurun_out/${tag}_summarize.log 2>&1
mkdir -p gpurun_out/profiles_${tag}
cp profiles/${tag}_* gpurun_out/profiles_${tag}/ 2>/dev/null
rm -f gpurun_out/prof_*_${tag}.ncu-rep
ls gpurun_out/profiles_${tag}/
