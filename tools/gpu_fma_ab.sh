for m in exact fma; do
  f=0; [ $m = fma ] && f=1
  RADE_B200_DEBUG_FLOAT_FMA=$f timeout 120 python bench.py --workload codec --streams 1024 --no-cpu-baseline --no-e2e > gpurun_out/r02r_codec1024_$m.json 2>/dev/null
  RADE_B200_DEBUG_FLOAT_FMA=$f timeout 120 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02r_full_$m.json 2>/dev/null
done
RADE_B200_DEBUG_FLOAT_FMA=1 python tools/codec_trace.py enc 1024 > gpurun_out/r02r_trace_enc_fma.txt 2>&1
python - <<PY
import json
for f in ["codec1024_exact","codec1024_fma","full_exact","full_fma"]:
    d=json.load(open("gpurun_out/r02r_%s.json"%f)); print(f, "%.3g F/s"%d["value"], "%.4f ms"%d["ms_per_step"], {k:v["ms_per_launch"] for k,v in d["kernels"].items() if "core" in k})
PY
grep "commits" gpurun_out/r02r_trace_enc_fma.txt
