python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_value.json 2> /dev/null
python bench.py --no-graph --no-cpu-baseline > gpurun_out/bench_eager.json 2> /dev/null
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python scripts/bench_extra.py > gpurun_out/bench_extra.log 2>&1
grep config4_inversion_ours gpurun_out/bench_extra.log | cut -c1-160
cut -c1-200 gpurun_out/bench_default.json
