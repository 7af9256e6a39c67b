#!/bin/bash
set -u
mkdir -p gpurun_out
OOD_ROWS_MIN_STRIPS=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "row_sliding or fused_torgb_epilogue or encoder_epilogues_on_wide or conv3x3" 2>&1 | tail -3
timeout 120 python scripts/rows_bench.py
timeout 120 python scripts/convT_bench.py
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err
tail -3 gpurun_out/r02_bench9.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench9.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'u8', d.get('e2e_u8',{}).get('value'), 'clk', d['clocks'])
print({k:(round(v['ms_per_step'],3), v['launches_per_step'], round(v['achieved'])) for k,v in d['kernels'].items()})
P
