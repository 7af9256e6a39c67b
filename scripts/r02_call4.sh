#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu4.log
cat gpurun_out/r02_pytest_gpu4.log
python bench.py --no-extra-legs --no-cpu-baseline > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
tail -3 gpurun_out/r02_bench4.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench4.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'u8', d.get('e2e_u8',{}).get('value'))
print({k:(round(v['ms_per_step'],3), v['launches_per_step'], round(v['achieved'])) for k,v in d['kernels'].items()})
P
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_r02.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    n=r[ki]; agg[n][0]+=1; agg[n][1]+=v
tot=sum(v[1] for v in agg.values())
ours=sum(v[1] for k,v in agg.items() if k.startswith('ood::') or 'ood::' in k)
print('total us', tot/1e3, 'ours share', ours/tot)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    if 'ood::' not in k: print(v[0], round(v[1]/1e3,1), k[:110])
P
