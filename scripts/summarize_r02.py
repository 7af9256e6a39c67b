"""profiles/ summaries of round 2 from the files scripts/collect_evidence_r02.sh leaves in gpurun_out/:
ncu_r02_summary.json (one row per ncu --set full capture) and launches_r02_summary.csv (launch list by kernel)."""
import collections, csv, glob, json, os, re, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def summ(path):
    rows = list(csv.reader(open(path)))
    hdr, units, r = rows[0], rows[1], rows[2]
    g = lambda name: next((f'{r[i]} {units[i]}'.strip() for i, h in enumerate(hdr) if h == name), None)
    return dict(kernel=g('Kernel Name'), time=g('gpu__time_duration.sum'), dram_read=g('dram__bytes_read.sum'), dram_write=g('dram__bytes_write.sum'),
                tensor_pipe_active_pct=g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), issue_active_pct=g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                warps_active_pct=g('sm__warps_active.avg.pct_of_peak_sustained_active'), dram_throughput_pct=g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                l2_hit_pct=g('lts__t_sector_hit_rate.pct'), warp_instructions=g('smsp__inst_executed.sum'), regs=g('launch__registers_per_thread'),
                sm_clock=g('sm__cycles_elapsed.avg.per_second'), grid=g('launch__grid_size'), block=g('launch__block_size'))


old = json.load(open(os.path.join(P, 'ncu_r02_summary.json'))) if os.path.exists(os.path.join(P, 'ncu_r02_summary.json')) else {}
out = {k: v for k, v in old.items() if k.endswith('_before_elect')}
for p in sorted(glob.glob(os.path.join(G, 'ncu_r02_*_raw.csv'))):
    name = os.path.basename(p)[len('ncu_r02_'):-len('_raw.csv')]
    if name == 'rows':
        continue
    try:
        out[name] = summ(p)
        shutil.copy(p, os.path.join(P, os.path.basename(p)))
    except Exception as e:
        out[name] = dict(error=str(e))
json.dump(out, open(os.path.join(P, 'ncu_r02_summary.json'), 'w'), indent=1)

rows = [r for r in csv.reader(open(os.path.join(G, 'launches_r02.csv'))) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    n = re.sub(r'\(.*', '', re.sub(r'^void ', '', r[ki]))
    agg[n][0] += 1
    agg[n][1] += v
tot, n_l = sum(v[1] for v in agg.values()), sum(v[0] for v in agg.values())
ours = lambda k: k.startswith(('ood::', 'rows::', 'wg::', 'trows::')) or 'ood::' in k
t_o, n_o = sum(v[1] for k, v in agg.items() if ours(k)), sum(v[0] for k, v in agg.items() if ours(k))
with open(os.path.join(P, 'launches_r02_summary.csv'), 'w') as f:
    f.write('# one bf16 step (batch 16) under ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
    f.write(f'# total {tot / 1e6:.3f} ms in {n_l} launches; kernels of libood_b200.so: {t_o / 1e6:.3f} ms = {t_o / tot:.4f} of the time, {n_o} launches; '
            f'the rest: {n_l - n_o} launches (the noise draws)\n')
    f.write('launches,ms,share,kernel\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'{v[0]},{v[1] / 1e6:.4f},{v[1] / tot:.4f},"{k}"\n')
shutil.copy(os.path.join(G, 'launches_r02.csv'), os.path.join(P, 'launches_r02.csv'))
for name in ('bench_r02_default_run.json', 'sweep_r02.json', 'rows_bench_r02.txt', 'step_profile_r02.txt', 'inversion_profile_r02.txt', 'se_bench_r02.txt', 'ncu_r02_rows32.ncu-rep', 'ncu_r02_convt_rows.ncu-rep'):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(P, name))
print(open(os.path.join(P, 'launches_r02_summary.csv')).read()[:1500])
print(json.dumps(out.get('convt_rows'), indent=0))
