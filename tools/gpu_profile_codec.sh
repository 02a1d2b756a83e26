tag=r03n
mkdir -p gpurun_out
for k in core_encoder_umma_kernel core_decoder_umma_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -o gpurun_out/prof_${k}_${tag} -f \
      python bench.py --no-cpu-baseline --no-e2e --no-pipeline --steps 3 --warmup 3 > gpurun_out/${tag}_${k}_run.log 2>&1
done
python tools/summarize_ncu.py ${tag} > gpurun_out/${tag}_summarize.log 2>&1
mkdir -p gpurun_out/profiles_${tag}
cp profiles/${tag}_* gpurun_out/profiles_${tag}/ 2>/dev/null
rm -f gpurun_out/prof_*_${tag}.ncu-rep
ls gpurun_out/profiles_${tag}/
