#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s 2>&1 | grep -E "bf16 OOD pipeline|FastEncoder|gen\+blend|final W\+|passed|failed|Error|error" | tail -30 > gpurun_out/r02_pytest_gpu2.log
cat gpurun_out/r02_pytest_gpu2.log
( time python bench.py ) > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
tail -5 gpurun_out/r02_bench2.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench2.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','e2e','e2e_u8','config3','config4','gpu_reference','parity'):
    print(k, json.dumps(d.get(k))[:700])
print('cpu_baseline', {k:v for k,v in d.get('cpu_baseline',{}).items() if k!='ops'})
P
