#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 120 python scripts/convT_bench.py
python - <<'P'
import os, sys, torch, time
sys.path.insert(0, '.')
from ood_gan_inversion_b200 import kernels as K
def timeit(fn, warm=3, rep=7):
    for _ in range(warm): fn()
    best = 1e9
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
w = torch.randn(512, 512, 3, 3, device='cuda'); wp = K.pack_conv_weight(w, torch.bfloat16, False)
for res in (16, 32, 64, 128, 256):
    x = torch.randn(16, res, res, 512, device='cuda').bfloat16()
    time.sleep(0.3)
    ms = timeit(lambda: K.conv3x3(x, wp, 512, transposed=True))
    fl = 2.0 * 16 * 512 * 512 * 9 * res * res
    print(f'modconv3x3_up 512->512 res {res}: {ms*1e3:.1f} us {fl/ms/1e9:.0f} TFLOP/s ({fl/ms/1e9/1671.7:.3f} of burst peak)')
P
python bench.py --no-extra-legs --no-cpu-baseline --no-u8-io 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], d['kernels']['conv3x3_tc'])"
