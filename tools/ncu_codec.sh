#!/bin/bash
# one ncu --set full capture of the codec kernels (tag = $1), summary CSV of the metrics that decide what binds
tag=${1:-r02}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:core_.*coder_umma_kernel -s 4 -c 2 -o gpurun_out/${tag}_codec_umma -f \
    python bench.py --workload codec --streams 1024 --no-cpu-baseline --no-e2e --steps 3 --warmup 3 > gpurun_out/${tag}_ncu_run.log 2>&1
ncu -i gpurun_out/${tag}_codec_umma.ncu-rep --page raw --csv > gpurun_out/${tag}_codec_umma_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${tag}_codec_umma_raw.csv")))
hdr=rows[0]
want=["Kernel Name","gpu__time_duration.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed","sm__inst_executed_pipe_lsu.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","sm__warps_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_bytes.sum","l1tex__m_xbar2l1tex_read_bytes.sum","smsp__inst_executed.sum","l1tex__data_pipe_lsu_wavefronts.sum","l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed","sm__cycles_elapsed.max","smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct","smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct","smsp__warp_issue_stalled_barrier_per_warp_active.pct","smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct","smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print("==", d.get("Kernel Name","")[:60])
    for k in hdr:
        if any(w in k for w in ["time_duration.sum","wavefronts","pipe_tensor","issue_active.avg.pct","bank_conflicts","warps_active.avg.pct","dram__bytes_read.sum","dram__bytes_write.sum","xbar2l1tex_read_bytes.sum","smsp__inst_executed.sum","warp_issue_stalled","sm__cycles_elapsed.max","shared_op","lsu_mem_shared"]):
            if d[k] not in ("","0","n/a"): print("  %-90s %s"%(k,d[k]))
PY
# per-instruction sampling (stall reasons) of the encoder kernel, SASS view
ncu -i gpurun_out/${tag}_codec_umma.ncu-rep --page source --csv --print-source sass --kernel-name regex:core_encoder_umma_kernel > gpurun_out/${tag}_enc_source.csv 2>/dev/null
ncu -i gpurun_out/${tag}_codec_umma.ncu-rep --page source --csv --print-source sass --kernel-name regex:core_decoder_umma_kernel > gpurun_out/${tag}_dec_source.csv 2>/dev/null
rm -f gpurun_out/${tag}_codec_umma.ncu-rep
ls -la gpurun_out/ | tail -5
