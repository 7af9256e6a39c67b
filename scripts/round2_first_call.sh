#!/bin/bash
# First GPU call of the next round: what the third session of round 1 could not measure (its GPU budget ended).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# 1. the whole GPU suite (the third session re-ran only the cases its changes touched; tests/test_zz_samm_bwd_gpu.py has never run);
# 2. default bench + the byte-format serving loop (imgio.ByteServing: uint8 frames across PCIe, `e2e_u8`);
# 3. ncu --set full of the rewritten mask_blend and of the two image-format kernels; refreshed launch list;
# 4. the SAMM micro-benchmarks (mask_blend 0.77 / warp_mix) on this box.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.log
python bench.py --u8-io > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --ncu --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cap() {   # name regex command...
  local name=$1 regex=$2; shift 2
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" -s 3 -c 1 \
      -f -o /tmp/ncu_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_r02_${name}_raw.csv 2>/dev/null
}
cap maskblend 'mask_blend_kernel' python scripts/samm_bench.py
cap img2tensor 'img2tensor_u8_kernel' python scripts/imgio_bench.py
cap tensor2img 'tensor2img_u8_kernel' python scripts/imgio_bench.py
python scripts/samm_bench.py > gpurun_out/r02_samm_bench.txt 2>&1
python scripts/imgio_bench.py > gpurun_out/r02_imgio_bench.txt 2>&1
cat gpurun_out/r02_pytest_gpu.log gpurun_out/r02_samm_bench.txt gpurun_out/r02_imgio_bench.txt
cut -c1-600 gpurun_out/r02_bench_default.json
