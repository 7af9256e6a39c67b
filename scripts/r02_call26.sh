#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "transposed" 2>&1 | tail -5
timeout 120 python scripts/convT_bench.py 2>&1 | head -4
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --no-extra-legs --no-cpu-baseline --no-u8-io 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])"
OOD_CONVT_ROWS=0 python bench.py --no-extra-legs --no-cpu-baseline --no-u8-io 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('without convt_rows: value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])"
