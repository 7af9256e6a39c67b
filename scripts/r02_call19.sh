#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/r02_bench19.json 2> gpurun_out/r02_bench19.err
tail -3 gpurun_out/r02_bench19.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench19.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'u8', d.get('e2e_u8',{}).get('value'), 'clk', d['clocks'])
print({k:(round(v['ms_per_step'],3), v['launches_per_step'], round(v['achieved'])) for k,v in d['kernels'].items()})
P
OOD_SE_FUSED=0 python bench.py --no-extra-legs --no-cpu-baseline --no-u8-io 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SE unfused: value', d['value'], 'e2e', d['e2e']['value'])"
