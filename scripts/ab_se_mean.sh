# A/B of OOD_SE_MEAN_FROM_CONV (encoder SE means from the conv epilogue) on one box: two short bench runs.
run() {
  env "$@" timeout 65 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_tmp.json 2> /dev/null
  python - "$*" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_tmp.json")); k = d["kernels"]
print(sys.argv[1] or "default", round(d["value"], 1), "img/s", round(d["e2e"]["value"], 1), "e2e", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"], "MHz |",
      " ".join(f"{n} {k[n]['ms_per_step']:.2f}" for n in ("conv3x3_tc", "encoder_conv", "alignnet_ew", "in_stats", "se_residual")))
PY
}
run OOD_SE_MEAN_FROM_CONV=0
run OOD_SE_MEAN_FROM_CONV=1
