#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s 2>&1 | grep -E "bf16 OOD pipeline|FastEncoder|passed|failed|^FAILED|^E  " | tail -30 > gpurun_out/r02_pytest_gpu5.log
cat gpurun_out/r02_pytest_gpu5.log
B=16 python scripts/diag_bf16_budget.py 2>&1 | tail -14
python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
tail -3 gpurun_out/r02_bench5.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench5.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'u8', d.get('e2e_u8',{}).get('value'))
print({k:(round(v['ms_per_step'],3), v['launches_per_step'], round(v['achieved'])) for k,v in d['kernels'].items()})
P
