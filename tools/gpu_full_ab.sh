for fam in umma mma; do
  RADE_B200_CODEC=$fam timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02p_full_$fam.json 2> gpurun_out/r02p_full_$fam.err
done
python - <<PY
import json
for f in ["umma","mma"]:
    d=json.load(open("gpurun_out/r02p_full_%s.json"%f)); print(f, "%.3g F/s"%d["value"], "%.4f ms"%d["ms_per_step"], "unpipelined %.4f"%d["config"]["ms_per_step_unpipelined"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
PY
