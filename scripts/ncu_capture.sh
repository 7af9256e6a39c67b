#!/bin/bash
# usage: scripts/ncu_capture.sh <name> <kernel-regex> <launch-skip> <launch-count>
# One `ncu --set full` capture of selected launches of one bench step (bench.py --ncu brackets exactly one step with
# cudaProfilerStart/Stop); exports the raw + details pages as CSV into gpurun_out/ and drops the (large) .ncu-rep.
set -e
name=$1; regex=$2; skip=$3; count=$4
mkdir -p gpurun_out
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s "$skip" -c "$count" \
    -f -o /tmp/ncu_$name python bench.py --ncu --steps 1 --warmup 3 > gpurun_out/ncu_$name.log 2>&1
ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv
ncu -i /tmp/ncu_$name.ncu-rep --page details --csv > gpurun_out/ncu_${name}_details.csv 2>/dev/null || true
ncu -i /tmp/ncu_$name.ncu-rep --page source --csv > gpurun_out/ncu_${name}_source.csv 2>/dev/null || true
ls -la /tmp/ncu_$name.ncu-rep
sz=$(stat -c %s /tmp/ncu_$name.ncu-rep)
if [ "$sz" -lt 12000000 ]; then cp /tmp/ncu_$name.ncu-rep gpurun_out/; fi
