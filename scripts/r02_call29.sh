#!/bin/bash
set -u
OOD_ROWS_MIN_STRIPS=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "fused_torgb_epilogue or row_sliding" 2>&1 | tail -3
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --no-extra-legs --no-cpu-baseline --no-u8-io 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], d['kernels']['torgb'], d['kernels']['conv3x3_tc'])"
OOD_FUSE_RGB_ROWS=0 python bench.py --no-extra-legs --no-cpu-baseline --no-u8-io 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unfused: value', d['value'], 'e2e', d['e2e']['value'], d['clocks'], d['kernels']['torgb'], d['kernels']['conv3x3_tc'])"
