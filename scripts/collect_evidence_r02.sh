#!/bin/bash
# Round-2 evidence in one GPU call:  /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/collect_evidence_r02.sh'
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r02_default_run.json 2> gpurun_out/bench_r02_default_run.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cap() {   # name regex skip command...
  local name=$1 regex=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s "$skip" -c 1 \
      -f -o /tmp/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_r02_${name}_raw.csv 2>/dev/null
  sz=$(stat -c %s /tmp/ncu_$name.ncu-rep 2>/dev/null || echo 99999999)
  if [ "$sz" -lt 6000000 ]; then cp /tmp/ncu_$name.ncu-rep gpurun_out/ncu_r02_$name.ncu-rep; fi
}
B=16 cap rows32 'conv_rows_kernel<.int.32, .int.32' 4 python scripts/rows_bench.py
B=16 cap rows64 'conv_rows_kernel<.int.64, .int.64' 4 python scripts/rows_bench.py
cap conv256_pair 'conv_tc_kernel' 3 python scripts/modconv_ncu.py alignnet
cap modconv_plain 'conv_tc_kernel' 3 python scripts/modconv_ncu.py plain
cap modconv_up 'conv_tc_kernel' 9 python scripts/modconv_ncu.py transposed
cap wgrad 'wgrad_tc_kernel' 3 python scripts/modconv_ncu.py wgrad
cap convt_rows 'convt_rows_kernel' 3 python scripts/convT_bench.py
python scripts/kernel_sweep.py --out gpurun_out/sweep_r02.json > gpurun_out/sweep_r02.log 2>&1
python scripts/rows_bench.py > gpurun_out/rows_bench_r02.txt 2>&1
python scripts/convT_bench.py >> gpurun_out/rows_bench_r02.txt 2>&1
python scripts/step_profile.py > gpurun_out/step_profile_r02.txt 2>&1
python scripts/inversion_profile.py > gpurun_out/inversion_profile_r02.txt 2>&1
python scripts/se_bench.py > gpurun_out/se_bench_r02.txt 2>&1
tail -3 gpurun_out/bench_r02_default_run.err
cut -c1-400 gpurun_out/bench_r02_default_run.json
ls -la gpurun_out | tail -30
