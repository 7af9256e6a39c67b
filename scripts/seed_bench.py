"""Per-level timing of the AlignNet first convolution: whole (K = 9*2C) vs the split form (enc-only half -> fp32 seed, cur half
seeded with it) at batch 16, plus the res0 pass with and without the fused statistics.  CUDA events, L2 flushed between runs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ood_gan_inversion_b200 import kernels as K  # noqa: E402

DEV = 'cuda'
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def timed(fn, rep=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


B = int(os.environ.get('B', 16))
for c, r in [(512, 32), (512, 64), (256, 128), (128, 256)]:
    torch.manual_seed(0)
    x2 = torch.randn(B, r, r, 2 * c, device=DEV).bfloat16()
    lo, hi = x2[..., :c].contiguous(), x2[..., c:].contiguous()
    w = 0.02 * torch.randn(2 * c, 2 * c, 3, 3, device=DEV)
    wf = K.pack_conv_weight(w, torch.bfloat16, False)
    wl = K.pack_conv_weight(w[:, :c].contiguous(), torch.bfloat16, False)
    wh = K.pack_conv_weight(w[:, c:].contiguous(), torch.bfloat16, False)
    slope = torch.full((2 * c,), 0.25, device=DEV)
    seed, _ = K.conv3x3(hi, wh, 2 * c, out_f32=True)
    fl = 2.0 * B * r * r * (2 * c) * (2 * c) * 9
    t_full = timed(lambda: K.conv3x3(x2, wf, 2 * c, prelu=slope))
    t_half = timed(lambda: K.conv3x3(lo, wl, 2 * c, prelu=slope))
    t_seedout = timed(lambda: K.conv3x3(hi, wh, 2 * c, out_f32=True))
    t_seeded = timed(lambda: K.conv3x3(lo, wl, 2 * c, prelu=slope, acc_in=seed))
    seed_t, _ = K.conv3x3(hi, wh, 2 * c, out_f32=True, tiled=True)
    t_tout = timed(lambda: K.conv3x3(hi, wh, 2 * c, out_f32=True, tiled=True))
    t_tseeded = timed(lambda: K.conv3x3(lo, wl, 2 * c, prelu=slope, acc_in=seed_t, tiled=True))
    print(f'   tile order: fp32-out {t_tout:7.1f} ({fl / 2 / t_tout / 1e6:6.0f}) | seeded {t_tseeded:7.1f} ({fl / 2 / t_tseeded / 1e6:6.0f}) | '
          f'2 cycles split {t_tout + 2 * t_tseeded:7.1f}', flush=True)
    print(f'C={c} R={r}: full {t_full:7.1f} us ({fl / t_full / 1e6:6.0f} TF/s) | half bf16-out {t_half:7.1f} ({fl / 2 / t_half / 1e6:6.0f}) | '
          f'half fp32-out {t_seedout:7.1f} ({fl / 2 / t_seedout / 1e6:6.0f}) | half seeded {t_seeded:7.1f} ({fl / 2 / t_seeded / 1e6:6.0f}) | '
          f'2 cycles: whole {2 * t_full:7.1f} vs split {t_seedout + 2 * t_seeded:7.1f}', flush=True)
    cur, enc = lo, hi
    st6 = K.in_stats(cur, enc)
    st2 = K.in_stats(x2)
    wn, bn = torch.ones(2 * c, device=DEV), torch.zeros(2 * c, device=DEV)
    t_res = timed(lambda: K.alignnet_res0(x2, st2, wn, bn, cur, enc, st6))
    t_st = timed(lambda: K.in_stats(x2))
    t_rs = timed(lambda: K.alignnet_res0_stats(x2, st2, wn, bn, cur, enc, st6))
    gb = B * r * r * c * 2 * 6 / 1e3
    print(f'          res0 {t_res:7.1f} us ({gb / t_res:5.0f} GB/s) + in_stats {t_st:7.1f} = {t_res + t_st:7.1f} | res0_stats {t_rs:7.1f} ({gb / t_rs:5.0f} GB/s)', flush=True)
