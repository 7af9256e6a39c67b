"""GPU probe: MN-major tcgen05 operand descriptors of the weight-gradient kernel (LBO / SBO candidates via OOD_WGRAD_LBO / _SBO)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import torch.nn.functional as F
from ood_gan_inversion_b200 import kernels as K
torch.backends.cudnn.allow_tf32 = False
for (b, h, w, ci, co, taps) in [(2, 16, 16, 64, 128, 9), (1, 8, 64, 64, 128, 1), (2, 32, 32, 256, 256, 9), (1, 16, 128, 128, 128, 9)]:
    g = torch.randn(b, co, h, w, generator=torch.Generator().manual_seed(1)).bfloat16().float().cuda()
    x = torch.randn(b, ci, h, w, generator=torch.Generator().manual_seed(2)).bfloat16().float().cuda()
    k = 3 if taps == 9 else 1
    wt = torch.zeros(co, ci, k, k, device='cuda', requires_grad=True)
    F.conv2d(x.double(), wt.double(), padding=k // 2).backward(g.double())
    ref = wt.grad.float()
    out = K.conv_wgrad(g.permute(0, 2, 3, 1).contiguous().bfloat16(), x.permute(0, 2, 3, 1).contiguous().bfloat16(), taps)
    torch.cuda.synchronize()
    err = float((out - ref).abs().max()); sc = float(ref.abs().max())
    print((b, h, w, ci, co, taps), 'max err %%.4g of %%.4g' %% (err, sc), 'OK' if err < 1e-3 * sc + 1e-2 else 'MISMATCH')
''' % ROOT
for lbo, sbo in [(None, None), (1024, 8192), (8192, 2048), (128, 1024), (1024, 1024), (8192, 8192), (1024, 128)]:
    env = dict(os.environ)
    if lbo is not None:
        env['OOD_WGRAD_LBO'], env['OOD_WGRAD_SBO'] = str(lbo), str(sbo)
    print('=== LBO', lbo, 'SBO', sbo, flush=True)
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout[-1500:], r.stderr[-600:] if r.returncode else '', flush=True)
    if lbo is None and 'MISMATCH' not in r.stdout and r.returncode == 0:
        break
