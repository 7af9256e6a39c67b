run() {
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_tmp.json 2> /dev/null
  python - "$*" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_tmp.json")); k = d["kernels"]
print(sys.argv[1] or "default", round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"], "MHz |",
      " ".join(f"{n} {k[n]['ms_per_step']:.2f}" for n in ("conv3x3_tc", "encoder_conv", "alignnet_ew", "in_stats", "torgb")))
PY
}
run OOD_EPW8=0 OOD_FUSED_T_MAX_CO=0
run OOD_EPW8=0
run OOD_EPW8=1
run OOD_EPW8=1 OOD_SPLIT_MIN_C=128
run OOD_EPW8=0 OOD_FUSED_T_MAX_CO=0
