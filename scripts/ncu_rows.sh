#!/bin/bash
# ncu --set full of the row-sliding conv kernel (1024 px 32->32 layer, batch 16) with source-level stall sampling
set -u
mkdir -p gpurun_out
B=16 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_rows_kernel" -s 4 -c 1 \
    -f -o /tmp/ncu_rows python scripts/rows_bench.py > gpurun_out/ncu_rows.log 2>&1
ncu -i /tmp/ncu_rows.ncu-rep --page raw --csv > gpurun_out/ncu_r02_rows_raw.csv
ncu -i /tmp/ncu_rows.ncu-rep --page source --csv > gpurun_out/ncu_r02_rows_source.csv 2>/dev/null || true
ls -la /tmp/ncu_rows.ncu-rep
sz=$(stat -c %s /tmp/ncu_rows.ncu-rep); if [ "$sz" -lt 30000000 ]; then cp /tmp/ncu_rows.ncu-rep gpurun_out/ncu_r02_rows.ncu-rep; fi
tail -3 gpurun_out/ncu_rows.log
