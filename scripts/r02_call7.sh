#!/bin/bash
set -u
mkdir -p gpurun_out
OOD_ROWS_MIN_STRIPS=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "row_sliding or fused_torgb_epilogue or encoder_epilogues_on_wide" 2>&1 | tail -5
timeout 120 python scripts/rows_bench.py
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
