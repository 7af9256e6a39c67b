#!/bin/bash
set -u
mkdir -p gpurun_out
OOD_ROWS_MIN_STRIPS=1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "row_sliding or fused_torgb_epilogue or encoder_epilogues_on_wide or glue or thumbnail or conv1x1_stride2 or f16_storage" 2>&1 | tail -5
python scripts/rows_bench.py
python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err
tail -3 gpurun_out/r02_bench6.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench6.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'u8', d.get('e2e_u8',{}).get('value'))
print({k:(round(v['ms_per_step'],3), v['launches_per_step'], round(v['achieved'])) for k,v in d['kernels'].items()})
P
python scripts/aten_ops_in_step.py 2>&1 | tail -12
