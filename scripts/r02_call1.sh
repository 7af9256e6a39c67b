#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/diag_blend_grad.py > gpurun_out/r02_diag_blend.txt 2>&1
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu.log
python bench.py --u8-io > gpurun_out/r02_bench_u8.json 2> gpurun_out/r02_bench_u8.err
cat gpurun_out/r02_diag_blend.txt gpurun_out/r02_pytest_gpu.log
cut -c1-1500 gpurun_out/r02_bench_u8.json
tail -5 gpurun_out/r02_bench_u8.err
