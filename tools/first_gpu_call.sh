#!/bin/bash
# First GPU call of the next round: run everything that was written after the GPU budget of round 1
# was spent, in one box session.  Usage (from the repo root, under gpurun):
#   gpurun --timeout 600 -- 'bash tools/first_gpu_call.sh'
# Outputs land in gpurun_out/first_call_*.log.  Every step is bounded by `timeout`; a failing step does not stop the others.
set -u
mkdir -p gpurun_out
echo "== gated GPU tests (foff_test, dfdt, noise_only, sine_noise, mpd_fading, CLI pipe)"
RADE_B200_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/first_call_pytest.log
echo "== tcgen05 probes"
cd tools/microbench
for p in umma_i8_swapab umma_gru_layer umma_conv_layer umma_gru_chain umma_tf32_probe umma_tf32_refresh; do
  extra=""
  case $p in umma_gru_layer|umma_conv_layer|umma_gru_chain) extra="-Xcompiler -ffp-contract=off";; esac
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo $extra -o $p $p.cu 2>&1 | grep -i error
  timeout 60 ./$p 2>&1 | tee ../../gpurun_out/first_call_$p.log
done
