"""BASELINE config 5: kernel sweep on one B200 -- 512-channel 3x3 modulated conv (plain + upsample) and upfirdn2d
up2 / down2 / blur at every resolution 4..1024, against the measured tensor and HBM rooflines.

    python scripts/kernel_sweep.py [--out profiles/sweep_rNN.json]

CUDA events on the launching stream, 0.5 s idle + 3 warm-up + best-of-5 timed launches per case (inputs of successive cases
differ, and every case above 64 px exceeds L2).  Algorithmic work per SURVEY.md section 8(d).
"""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402


def timeit(fn, warm=3, rep=5):
    torch.cuda.synchronize()
    time.sleep(0.5)            # every case starts from an idle GPU: the peaks it is compared with are burst figures (kernel timed alone)
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='')
    ap.add_argument('--max-gb', type=float, default=6.0)
    args = ap.parse_args()
    pk = json.load(open('MEASURED_PEAKS.json')) if os.path.exists('MEASURED_PEAKS.json') else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
    dev = 'cuda'
    rows = []
    torch.manual_seed(0)
    w = torch.randn(512, 512, 3, 3, device=dev)
    wp = K.pack_conv_weight(w, torch.bfloat16, False)
    k4 = torch.tensor([1., 3., 3., 1.], device=dev)
    k2d = torch.outer(k4, k4) / 64
    for res in [4, 8, 16, 32, 64, 128, 256, 512, 1024]:
        # --- modulated conv 512 -> 512, bf16 tcgen05; batch chosen to keep the activation under --max-gb
        for transposed in (False, True):
            out_px = (2 * res + 1) ** 2 if transposed else res * res
            b = 16
            while b > 1 and b * max(out_px, res * res) * 512 * 2 / 1e9 > args.max_gb:
                b //= 2
            if b * max(out_px, res * res) * 512 * 2 / 1e9 > args.max_gb:
                continue
            x = torch.randn(b, res, res, 512, device=dev).bfloat16()
            d = torch.rand(b, 512, device=dev) + 0.5
            if transposed:
                ms = timeit(lambda: K.conv3x3(x, wp, 512, transposed=True, impl=0))
            else:
                nz = torch.randn(b, 1, res, res, device=dev)
                nw, bias = torch.tensor([0.1], device=dev), torch.randn(512, device=dev)
                ms = timeit(lambda: K.conv3x3(x, wp, 512, impl=0, d=d, noise=nz, noise_w=nw, bias=bias, act=True))
            fl = 2.0 * b * 512 * 512 * 9 * res * res
            tf = fl / (ms * 1e-3) / 1e12
            rows.append(dict(op='modconv3x3_up' if transposed else 'modconv3x3', res=res, batch=b, ms=ms, tflops=tf,
                             frac_of_bf16_peak=tf / pk['bf16_tflops']))
            del x
        # --- upfirdn2d on NCHW planes (fp32 and bf16): blur pad(1,1) on 2r+1, up2 pad(2,1), down2 pad(1,1)
        for dt in (torch.float32, torch.bfloat16):
            es = 4 if dt == torch.float32 else 2
            c = 512
            b = 4
            while b * c * (2 * res + 1) ** 2 * es / 1e9 > args.max_gb and c > 8:
                c //= 2
            cases = [('blur', torch.randn(b, c, 2 * res + 1, 2 * res + 1, device=dev).to(dt), k2d * 4, 1, 1, (1, 1)),
                     ('up2', torch.randn(b, c, res, res, device=dev).to(dt), k2d * 4, 2, 1, (2, 1)),
                     ('down2', torch.randn(b, c, 2 * res, 2 * res, device=dev).to(dt), k2d, 1, 2, (1, 1))]
            for name, x, k, up, down, pad in cases:
                y = K.upfirdn2d_nchw(x, k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
                ms = timeit(lambda: K.upfirdn2d_nchw(x, k, up, up, down, down, pad[0], pad[1], pad[0], pad[1]))
                gb = (x.numel() + y.numel()) * es / 1e9
                rows.append(dict(op=f'upfirdn2d_{name}', dtype=str(dt).split('.')[-1], res=res, planes=b * c, ms=ms, gbs=gb / (ms * 1e-3),
                                 frac_of_hbm_peak=gb / (ms * 1e-3) / pk['hbm_gbs']))
            del cases
    # --- the bf16 pipeline's own FIR path: NHWC fused blur + demod + noise + bias + lrelu + next style at the generator's sizes
    taps = K.fir_taps(gain=2.0)
    for res, c in [(64, 512), (128, 256), (256, 128), (512, 64), (1024, 32)]:
        b = 16
        t = torch.randn(b, res + 1, res + 1, c, device=dev).bfloat16()
        dd, ss = torch.rand(b, c, device=dev) + 0.5, torch.rand(b, c, device=dev) + 0.5
        nz, nw, bias = torch.randn(b, 1, res, res, device=dev), torch.tensor([0.1], device=dev), torch.randn(c, device=dev)
        ms = timeit(lambda: K.blur_act(t, taps, d=dd, noise=nz, noise_w=nw, bias=bias, s_next=ss, act=True, want_y=False, want_ys=True))
        gb = (b * c * ((res + 1) ** 2 + res * res) * 2 + b * res * res * 4) / 1e9
        rows.append(dict(op='blur_act_nhwc', dtype='bfloat16', res=res, channels=c, batch=b, ms=ms, gbs=gb / (ms * 1e-3),
                         frac_of_hbm_peak=gb / (ms * 1e-3) / pk['hbm_gbs']))
        del t
    for r in rows:
        print(json.dumps(r))
    if args.out:
        json.dump(dict(peaks=pk, rows=rows), open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
