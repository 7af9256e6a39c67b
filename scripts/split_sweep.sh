for c in 100000 512 256 128; do
  OOD_SPLIT_MIN_C=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split_$c.json 2> /dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_split_$c.json"))
k=d["kernels"]
print("min_c=$c", round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], "conv", round(k["conv3x3_tc"]["ms_per_step"],3), k["conv3x3_tc"]["launches_per_step"], "ew", round(k["alignnet_ew"]["ms_per_step"],3), "stats", round(k["in_stats"]["ms_per_step"],3))
PY
done
