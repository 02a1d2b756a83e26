#!/bin/bash
# ncu evidence of one state of the code (tag = $1): the launch list of the default bench command and one --set full capture of
# each dominant kernel, summarised ON the GPU box by tools/summarize_ncu.py (the .ncu-rep files stay there); the text
# summaries come back in gpurun_out/profiles_<tag>/.
tag=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/${tag}_launchlist_run.log 2>&1
for k in core_encoder_umma_kernel core_decoder_umma_kernel rx_refresh_kernel rx_track_kernel rx_demod_kernel rx_bpf_kernel channel_stream_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -o gpurun_out/prof_${k}_${tag} -f \
      python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 3 --warmup 3 > gpurun_out/${tag}_${k}_run.log 2>&1
done
# rx_detect does its real work only while streams search: capture it in the rx-search workload
ncu --set full --clock-control none --import-source on -k regex:rx_detect_kernel -s 6 -c 1 -o gpurun_out/prof_rx_detect_kernel_${tag} -f \
    python bench.py --workload rx-search --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/${tag}_rx_detect_run.log 2>&1
mkdir -p profiles_tmp && rm -rf profiles_save && cp -r profiles profiles_save
python tools/summarize_ncu.py ${tag} > gpurun_out/${tag}_summarize.log 2>&1
mkdir -p gpurun_out/profiles_${tag}
cp profiles/${tag}_* gpurun_out/profiles_${tag}/ 2>/dev/null
rm -f gpurun_out/prof_*_${tag}.ncu-rep
ls -la gpurun_out/profiles_${tag}/
