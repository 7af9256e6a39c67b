# same-box A/B of the AlignNet fusions: images/s of the default bench with each switch off, then all on
run() {
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_tmp.json 2> /dev/null
  python - "$*" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_tmp.json")); k = d["kernels"]
print(sys.argv[1] or "default", round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"], "MHz |",
      " ".join(f"{n} {k[n]['ms_per_step']:.2f}" for n in ("conv3x3_tc", "alignnet_ew", "in_stats", "torgb", "tap_sum")))
PY
}
run OOD_SPLIT_MIN_C=100000 OOD_CONV_STATS=0 OOD_FOLD_SHORTCUT=0
run OOD_CONV_STATS=0 OOD_FOLD_SHORTCUT=0
run OOD_FOLD_SHORTCUT=0
run OOD_CONV_STATS=0
run A=1
run OOD_SPLIT_MIN_C=100000 OOD_CONV_STATS=0 OOD_FOLD_SHORTCUT=0
