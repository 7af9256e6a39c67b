#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "se_tail or glue or fast_encoder or ood_pipeline" 2>&1 | tail -3
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -k "regex:se_tail|head_weights" --log-file gpurun_out/launches_small.csv \
    python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'P'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/launches_small.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    n=re.sub(r'\(.*','',r[ki]); agg[n][0]+=1; agg[n][1]+=v
for k,v in agg.items(): print(v[0], round(v[1]/1e3,1), 'us', k[:80])
P
python bench.py --no-extra-legs --no-cpu-baseline --no-u8-io 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])"
