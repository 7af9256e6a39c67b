#!/bin/bash
# One `ncu --set full` capture per kernel family (one launch each) of a bench step / micro-benchmark; exports the raw page
# as CSV into gpurun_out/ (the .ncu-rep stays in /tmp unless small).  Usage: scripts/ncu_capture_all.sh
set -u
mkdir -p gpurun_out
cap() {   # name regex skip command...
  local name=$1 regex=$2 skip=$3; shift 3
  ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s "$skip" -c 1 \
      -f -o /tmp/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_r01_${name}_raw.csv 2>/dev/null
  sz=$(stat -c %s /tmp/ncu_$name.ncu-rep 2>/dev/null || echo 99999999)
  if [ "$sz" -lt 6000000 ]; then cp /tmp/ncu_$name.ncu-rep gpurun_out/ncu_r01_$name.ncu-rep; fi
}
B="python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline"
cap conv256 'conv_tc_kernel<.int.256' 20 $B
cap blurrows 'blur_rows_kernel' 4 $B
cap torgb 'torgb_mma_kernel<.int.1' 0 $B
cap ew 'alignnet_ew_kernel<.*1>' 7 $B
cap warp 'warp_mix_kernel' 7 $B
cap rows 'conv_rows_kernel<.int.32' 0 $B
cap instats 'in_partial_kernel<.*5>' 7 $B
cap maskblend 'mask_blend_kernel' 0 $B
