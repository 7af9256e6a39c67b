#!/bin/bash
# Round evidence in one GPU call: bench lines, ncu launch list of one step, full captures of the dominant kernels, sweeps.
# Everything lands in gpurun_out/; the files quoted in profiles/README.md are copied to profiles/ afterwards.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_value.json 2> /dev/null
python bench.py --no-graph --no-cpu-baseline > gpurun_out/bench_eager.json 2> /dev/null
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cap() {   # name regex skip command...
  local name=$1 regex=$2 skip=$3; shift 3
  ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s "$skip" -c 1 \
      -f -o /tmp/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_r01_${name}_raw.csv 2>/dev/null
  sz=$(stat -c %s /tmp/ncu_$name.ncu-rep 2>/dev/null || echo 99999999)
  if [ "$sz" -lt 6000000 ]; then cp /tmp/ncu_$name.ncu-rep gpurun_out/ncu_r01_$name.ncu-rep; fi
}
B="python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline"
# AlignNet second convolution, 1024 -> 1024 channels at 64 px, full K, fused output statistics (STATS variant, 3rd launch)
cap conv256 'conv_tc_kernel<.int.256, .int.64, .bool.0, .bool.1' 2 $B
# fused-phase transposed convolution of the 1024 px layer (8 epilogue warps, N = 128: 64 -> 32 channels at 512 -> 1025 px)
cap convt 'conv_tc_kernel<.int.128, .int.64, .bool.0, .bool.0, .int.8' 0 $B
python scripts/kernel_sweep.py > gpurun_out/sweep.log 2>&1
python scripts/bench_extra.py > gpurun_out/bench_extra.log 2>&1
python scripts/seed_bench.py > gpurun_out/seed_bench.log 2>&1
python scripts/convT_bench.py > gpurun_out/convT_bench.log 2>&1
tail -n 3 gpurun_out/bench_default.err
cat gpurun_out/bench_default.json | cut -c1-400
