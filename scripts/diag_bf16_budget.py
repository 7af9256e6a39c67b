"""Where does the bf16 pipeline's image error come from?  (round 2: max-abs 0.017-0.021 at batch 16 against the 2e-2 bound)
Runs the 1024 px pipeline against the fp32 oracle with single stages switched between bf16 and fp32."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.set_grad_enabled(False)
import ood_gan_inversion_b200.stylegan as sg
from ood_gan_inversion_b200.arch import ood_faceGAN_e4e
from oracle import ood as oood
DEV = 'cuda'
B = int(os.environ.get('B', 8))
sd = oood.synthetic_ood_state(1024, seed=0)
sdd = {k: v.to(DEV) for k, v in sd.items()}
net = ood_faceGAN_e4e(out_size=1024, style_dim=512, encoder='E4E', enable_modulation=True, warp_scale=0.08, cycle_align=2, blend_with_gen=True, ModSize=256)
net.load_state_dict(sd, strict=True)
net = net.to(DEV).eval()
net.strict_rng = True
x = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(2))
x = F.interpolate(x, (1024, 1024), mode='bicubic', align_corners=False).clamp(-1, 1).to(DEV)
torch.manual_seed(123)
ref, rlats, raligns = oood.ood_forward(sdd, x)
w_ref, f_ref = oood.e4e_encoder(sdd, F.interpolate(x, (256, 256), mode='bilinear'))


def report(tag, out, lats, aligns):
    e = (out - ref).abs()
    psnr = 10 * math.log10(4.0 / float(((out - ref) ** 2).mean()))
    fl = max(float((aligns[k][:, :2] - raligns[k][:, :2]).abs().max()) for k in (1, 2, 3, 4))
    al = max(float((aligns[k][:, 2:] - raligns[k][:, 2:]).abs().max()) for k in (1, 2, 3, 4))
    print(f'{tag:46s} out max {float(e.max()):.4g} p99.99 {float(e.flatten()[::7].quantile(0.9999)):.4g} mean {float(e.mean()):.3g} PSNR {psnr:.1f}; '
          f'lats {float((lats - rlats).abs().max()):.3g}; flow {fl:.3g} alpha {al:.3g} mask {float((aligns[1024] - raligns[1024]).abs().max()):.3g}', flush=True)


def run(tag, precision, encode=None):
    sg.set_precision(precision)
    orig = net.encode
    if encode is not None:
        net.encode = encode
    try:
        torch.manual_seed(123)
        out, lats = net(x)
        report(tag, out, lats, dict(net.aligns))
    finally:
        net.encode = orig
        sg.set_precision('bf16')
    return out


run('A all bf16', 'bf16')
run('A2 all fp32', 'fp32')
enc_bf16 = None
sg.set_precision('bf16')
wb, fb = net.encode(x)
run('B fp32-oracle encoder -> bf16 gen+SAMM', 'bf16', encode=lambda t: (w_ref, [f.to(torch.bfloat16) for f in f_ref]))
run('B2 oracle w, bf16-encoder feats', 'bf16', encode=lambda t: (w_ref, fb))
run('B3 bf16-encoder w, oracle feats', 'bf16', encode=lambda t: (wb, [f.to(torch.bfloat16) for f in f_ref]))
run('C bf16 encoder -> fp32 gen+SAMM', 'fp32', encode=lambda t: (wb.float(), [f.float() for f in fb]))
# generator alone (no SAMM): bf16 vs fp32 with oracle lats
lat = rlats
sg.set_precision('bf16')
g16, _ = net.generator(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
sg.set_precision('fp32')
g32, _ = net.generator(lat, input_is_tensor=True, input_is_latent=True, randomize_noise=False)
sg.set_precision('bf16')
print('generator alone bf16 vs fp32 (oracle lats): max', float((g16 - g32).abs().max()), 'mean', float((g16 - g32).abs().mean()), 'range', float(g32.min()), float(g32.max()))
# where is the error: against |x - gen| and alpha error
torch.manual_seed(123)
out, _ = net(x)
e = (out - ref).abs()
idx = e.flatten().topk(5).indices
for i in idx:
    b, c, yy, xx = [int(v) for v in torch.unravel_index(i, e.shape)]
    print('top err', float(e[b, c, yy, xx]), 'at', (b, c, yy, xx), 'alpha ours/ref', float(net.aligns[1024][b, 0, yy, xx]), float(raligns[1024][b, 0, yy, xx]),
          'x', float(x[b, c, yy, xx]), 'ref out', float(ref[b, c, yy, xx]))
