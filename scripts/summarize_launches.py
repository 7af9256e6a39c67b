"""gpurun_out/launches_r01.csv (ncu gpu__time_duration.sum per launch of ONE step) -> per-kernel totals and shares."""
import collections
import csv
import sys

src = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_r01.csv'
dst = sys.argv[2] if len(sys.argv) > 2 else 'profiles/launches_r01_summary.csv'
lines = [l for l in open(src) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    us = v / 1000 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1000)
    k = row['Kernel Name'].split('(')[0][:100]
    agg[k][0] += 1
    agg[k][1] += us
    tot += us
with open(dst, 'w') as f:
    f.write('kernel,launches,total_us,share_of_step\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{v[0]},{v[1]:.1f},{v[1] / tot:.4f}\n')
    mine = sum(v[1] for k, v in agg.items() if 'ood::' in k or 'rows::' in k)
    f.write(f'"TOTAL (serialised, cold-cache ncu replay)",{sum(v[0] for v in agg.values())},{tot:.1f},1.0\n')
    f.write(f'"of which kernels of libood_b200.so",,{mine:.1f},{mine / tot:.4f}\n')
print(open(dst).read()[:1800])
