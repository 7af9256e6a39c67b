"""GPU parity tests: every C-ABI kernel against the oracle (oracle/) on seeded inputs.  Run with -m gpu on a B200."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ops as oops, samm as osamm, stylegan as ostyle

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _row_kernel_on_small_problems(monkeypatch):
    # the row-sliding convolution kernel declines launches with fewer strips than SMs; the parity cases are small on purpose
    monkeypatch.setenv('OOD_ROWS_MIN_STRIPS', '1')


@pytest.fixture(scope='module', autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def K():
    from ood_gan_inversion_b200 import kernels
    return kernels


def g(seed):
    return torch.Generator().manual_seed(seed)


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=g(seed))


def nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype).to(DEV)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous().cpu()


# ----------------------------------------------------------------------------------------------- upfirdn2d
def test_upfirdn2d_golden_cases(golden):
    for c in golden('ops.pt')['upfirdn2d']:
        p = c['pad']
        y = K().upfirdn2d_nchw(c['x'].to(DEV), c['k'].to(DEV), c['up'], c['up'], c['down'], c['down'], p[0], p[1], p[0], p[1])
        torch.testing.assert_close(y.cpu(), c['y'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('cfg', [
    dict(shape=(2, 3, 65, 65), up=1, down=1, pad=(1, 1), gain=4.0),       # conv-up blur, odd width
    dict(shape=(2, 3, 32, 32), up=2, down=1, pad=(2, 1), gain=4.0),       # rgb skip upsample
    dict(shape=(1, 4, 64, 48), up=1, down=2, pad=(1, 1), gain=1.0),       # down2
    dict(shape=(1, 2, 100, 130), up=1, down=1, pad=(2, 1), gain=1.0),     # multi-tile, ragged
    dict(shape=(1, 2, 19, 23), up=3, down=2, pad=(4, 3), gain=9.0),       # generic path
    dict(shape=(1, 1, 1, 1), up=2, down=1, pad=(2, 1), gain=4.0),         # smallest input
    dict(shape=(2, 2, 131, 1025), up=1, down=1, pad=(1, 1), gain=4.0),    # wide planes: row-streaming kernel, odd width, 2+ strips
    dict(shape=(1, 3, 300, 260), up=1, down=1, pad=(2, 1), gain=1.0),     # row-streaming blur, pad (2,1), one ragged strip
    dict(shape=(1, 2, 140, 1100), up=1, down=2, pad=(1, 1), gain=1.0),    # row-streaming down2, ragged second strip
    dict(shape=(2, 1, 70, 600), up=2, down=1, pad=(2, 1), gain=4.0),      # row-streaming up2 (polyphase), two strips
    dict(shape=(3, 2, 64, 64), up=2, down=1, pad=(2, 1), gain=4.0),       # narrowest strip (64 columns)
    dict(shape=(2, 2, 128, 120), up=1, down=2, pad=(1, 1), gain=1.0),     # down2 on a 60-wide output
])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_upfirdn2d_vs_oracle(cfg, dtype):
    x = rnd(*cfg['shape'], seed=3).to(dtype)
    k = oops.fir_kernel([1, 3, 3, 1], cfg['gain'])
    ref = oops.upfirdn2d(x.float(), k, cfg['up'], cfg['down'], cfg['pad'])
    p = cfg['pad']
    y = K().upfirdn2d_nchw(x.to(DEV), k.to(DEV), cfg['up'], cfg['up'], cfg['down'], cfg['down'], p[0], p[1], p[0], p[1])
    assert y.dtype == dtype and y.shape == ref.shape
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(y.float().cpu(), ref, **tol)


def test_upfirdn2d_asymmetric_kernel_and_axes():
    x = rnd(2, 2, 11, 9, seed=5)
    k = rnd(3, 5, seed=6)
    ref = oops.upfirdn2d_xy(x, k, 2, 1, 1, 2, 3, 1, 0, 2)
    y = K().upfirdn2d_nchw(x.to(DEV), k.to(DEV), 2, 1, 1, 2, 3, 1, 0, 2)
    torch.testing.assert_close(y.cpu(), ref, rtol=1e-5, atol=1e-5)


def test_upfirdn2d_linearity_at_full_size():
    # size-independent property at a BASELINE size: up2(a*x + y) == a*up2(x) + up2(y)
    x, y = torch.randn(1, 3, 512, 512, device=DEV), torch.randn(1, 3, 512, 512, device=DEV)
    k = oops.fir_kernel([1, 3, 3, 1], 4.0).to(DEV)
    f = lambda t: K().upfirdn2d_nchw(t, k, 2, 2, 1, 1, 2, 1, 2, 1)
    torch.testing.assert_close(f(2.5 * x + y), 2.5 * f(x) + f(y), rtol=1e-4, atol=1e-4)
    assert f(x).shape == (1, 3, 1024, 1024)
    # DC gain: a constant image stays constant away from the border
    c = f(torch.ones(1, 1, 64, 64, device=DEV))
    torch.testing.assert_close(c[..., 4:-4, 4:-4], torch.ones_like(c[..., 4:-4, 4:-4]))


# ----------------------------------------------------------------------------------------------- bias act
@pytest.mark.parametrize('shape', [(2, 6, 5, 7), (3, 8), (2, 4, 16, 16), (1, 3, 1, 1)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_fused_bias_act(shape, dtype):
    x = rnd(*shape, seed=1).to(dtype)
    b = rnd(shape[1], seed=2)
    ref = oops.fused_leaky_relu(x.float(), b)
    y = K().fused_bias_act(x.to(DEV), b.to(DEV), None, 0, 0.2, math.sqrt(2))
    tol = dict(rtol=1e-6, atol=1e-6) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(y.float().cpu(), ref, **tol)
    if dtype == torch.float32:
        gout = rnd(*shape, seed=4)
        gx_ref, gb_ref = oops.fused_leaky_relu_backward(gout, ref)
        gx = K().fused_bias_act(gout.to(DEV), None, y, 1, 0.2, math.sqrt(2))
        torch.testing.assert_close(gx.cpu(), gx_ref, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(K().bias_grad(gx).cpu(), gb_ref, rtol=1e-4, atol=1e-5)
        assert float(K().fused_bias_act(gout.to(DEV), None, y, 2, 0.2, 1.0).abs().max()) == 0.0


def test_fused_bias_act_golden(golden):
    for c in golden('ops.pt')['fused_leaky_relu']:
        y = K().fused_bias_act(c['x'].to(DEV), c['b'].to(DEV), None, 0, 0.2, math.sqrt(2))
        torch.testing.assert_close(y.cpu(), c['y'], rtol=1e-6, atol=1e-6)


# ----------------------------------------------------------------------------------------------- layout / modulation
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_layout_roundtrip_and_scale(dtype):
    x = rnd(3, 40, 5, 7, seed=1)
    s = rnd(3, 40, seed=2)
    y = K().nchw_to_nhwc(x.to(DEV), s.to(DEV), dtype)
    ref = (x * s[:, :, None, None]).permute(0, 2, 3, 1)
    tol = dict(rtol=1e-6, atol=1e-6) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(y.float().cpu(), ref, **tol)
    back = K().nhwc_to_nchw(K().nchw_to_nhwc(x.to(DEV), None, dtype))
    torch.testing.assert_close(back.cpu(), x.to(dtype).float())
    const = rnd(1, 64, 4, 4, seed=3)
    yb = K().nchw_to_nhwc(const.to(DEV), rnd(5, 64, seed=4).to(DEV), dtype, batch=5)
    assert yb.shape == (5, 4, 4, 64)
    z = K().nhwc_scale(y, s.to(DEV))
    torch.testing.assert_close(z.float().cpu(), (y.float().cpu() * s[:, None, None, :]).to(dtype).float(), **tol)


def test_modulation_and_demod():
    b, D, ci, co = 3, 512, 96, 64
    lat = rnd(b, 18, D, seed=1).to(DEV)
    mw, mb, w = rnd(ci, D, seed=2), 1 + 0.1 * rnd(ci, seed=3), rnd(co, ci, 3, 3, seed=4)
    wsq = K().weight_sumsq(w.to(DEV))
    torch.testing.assert_close(wsq.cpu(), w.pow(2).sum([2, 3]), rtol=1e-5, atol=1e-5)
    cs = 1 / math.sqrt(ci * 9)
    s, d = K().modulation(lat[:, 5], mw.to(DEV), mb.to(DEV), wsq, cs, co)
    s_ref = ostyle.equal_linear(lat[:, 5].cpu(), mw, mb)
    wb = cs * w[None] * s_ref[:, None, :, None, None]
    d_ref = cs * torch.rsqrt(wb.pow(2).sum([2, 3, 4]) + 1e-8)
    torch.testing.assert_close(s.cpu(), s_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(d.cpu(), d_ref, rtol=1e-4, atol=1e-7)
    _, d1 = K().modulation(lat[:, 5], mw.to(DEV), mb.to(DEV), None, cs, co)
    torch.testing.assert_close(d1.cpu(), torch.full((b, co), cs))
    for ci_major in (False, True):
        pk = K().pack_conv_weight(w.to(DEV), torch.float32, ci_major).cpu()
        ref = w.reshape(co, ci, 9).permute(2, 1, 0) if ci_major else w.reshape(co, ci, 9).permute(2, 0, 1)
        torch.testing.assert_close(pk, ref.contiguous())


# ----------------------------------------------------------------------------------------------- convolution
CONV_CASES = [
    dict(b=2, h=16, w=16, ci=64, co=64),
    dict(b=3, h=4, w=4, ci=128, co=64),        # several samples per 128-pixel tile
    dict(b=1, h=32, w=32, ci=128, co=256),     # BN = 256
    dict(b=2, h=20, w=12, ci=64, co=128),      # ragged patch grid
    dict(b=1, h=8, w=136, ci=32, co=32),       # BK = 32 (64B swizzle), 128-wide row tiles + remainder
    dict(b=5, h=8, w=8, ci=64, co=96),         # BN = 32 path with 3 N tiles, batch remainder
]


def conv_ref(x, w, transposed):
    if transposed:
        return F.conv_transpose2d(x, w.transpose(0, 1), stride=2)
    return F.conv2d(x, w, padding=1)


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('transposed', [False, True])
def test_conv3x3_simt_raw(case, transposed):
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    if ci % 16 or co % 4:
        pytest.skip('simt alignment')
    x, w = rnd(b, ci, h, w_, seed=1), rnd(co, ci, 3, 3, seed=2)
    wp = K().pack_conv_weight(w.to(DEV), torch.float32, True)
    y, _ = K().conv3x3(nhwc(x, torch.float32), wp, co, transposed=transposed, impl=1)
    ref = conv_ref(x.double(), w.double(), transposed).float()
    torch.testing.assert_close(nchw(y), ref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('transposed', [False, True])
def test_conv3x3_tcgen05_raw(case, transposed):
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    x, w = rnd(b, ci, h, w_, seed=1).bfloat16().float(), rnd(co, ci, 3, 3, seed=2).bfloat16().float()
    wp = K().pack_conv_weight(w.to(DEV), torch.bfloat16, False)
    y, _ = K().conv3x3(nhwc(x, torch.bfloat16), wp, co, transposed=transposed, impl=0, out_f32=True)
    ref = conv_ref(x.double(), w.double(), transposed).float()      # same bf16-rounded operands, exact products
    torch.testing.assert_close(nchw(y), ref, rtol=1e-4, atol=2e-3)
    yb, _ = K().conv3x3(nhwc(x, torch.bfloat16), wp, co, transposed=transposed, impl=0)
    assert yb.dtype == torch.bfloat16
    torch.testing.assert_close(nchw(yb), ref, rtol=1e-2, atol=0.15)


@pytest.mark.parametrize('impl', [0, 1])
def test_conv3x3_fused_epilogue(impl):
    b, h, w_, ci, co = 3, 16, 16, 64, 128
    dt = torch.bfloat16 if impl == 0 else torch.float32
    x, w = rnd(b, ci, h, w_, seed=1).to(dt).float(), rnd(co, ci, 3, 3, seed=2).to(dt).float()
    d, bias, s_next = 0.05 * (1 + rnd(b, co, seed=3).abs()), rnd(co, seed=4), 1 + 0.3 * rnd(b, co, seed=5)
    noise, nw = rnd(b, 1, h, w_, seed=6), torch.tensor([0.37])
    wp = K().pack_conv_weight(w.to(DEV), dt, impl == 1)
    y, ys = K().conv3x3(nhwc(x, dt), wp, co, impl=impl, d=d.to(DEV), noise=noise.to(DEV), noise_w=nw.to(DEV),
                        bias=bias.to(DEV), s_next=s_next.to(DEV), act=True, want_y=True, want_ys=True)
    ref = F.conv2d(x.double(), w.double(), padding=1).float() * d[:, :, None, None] + nw * noise
    ref = oops.fused_leaky_relu(ref, bias)
    tol = dict(rtol=2e-2, atol=3e-2) if impl == 0 else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(nchw(y), ref, **tol)
    torch.testing.assert_close(nchw(ys), ref * s_next[:, :, None, None], **tol)
    # shared noise plane (registered buffers, batch stride 0)
    y1, _ = K().conv3x3(nhwc(x, dt), wp, co, impl=impl, d=d.to(DEV), noise=noise[:1].contiguous().to(DEV), noise_w=nw.to(DEV),
                        bias=bias.to(DEV), act=True)
    ref1 = oops.fused_leaky_relu(F.conv2d(x.double(), w.double(), padding=1).float() * d[:, :, None, None] + nw * noise[:1], bias)
    torch.testing.assert_close(nchw(y1), ref1, **tol)


def test_conv3x3_tc_matches_simt_large():
    # tensor-core path against the fp32 SIMT path on the device at a generator-sized layer (64x64, 512 channels)
    b, h, ci, co = 2, 64, 512, 512
    x = torch.randn(b, h, h, ci, device=DEV).bfloat16()
    w = torch.randn(co, ci, 3, 3, device=DEV).bfloat16().float()
    y_tc, _ = K().conv3x3(x, K().pack_conv_weight(w, torch.bfloat16, False), co, impl=0, out_f32=True)
    y_si, _ = K().conv3x3(x.float(), K().pack_conv_weight(w, torch.float32, True), co, impl=1)
    torch.testing.assert_close(y_tc, y_si, rtol=1e-3, atol=2e-2)
    t_tc, _ = K().conv3x3(x, K().pack_conv_weight(w, torch.bfloat16, False), co, transposed=True, impl=0, out_f32=True)
    t_si, _ = K().conv3x3(x.float(), K().pack_conv_weight(w, torch.float32, True), co, transposed=True, impl=1)
    torch.testing.assert_close(t_tc, t_si, rtol=1e-3, atol=2e-2)


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('case', [dict(b=2, h=16, w=16, ci=64, co=64), dict(b=3, h=2, w=2, ci=128, co=128), dict(b=16, h=4, w=4, ci=64, co=256),
                                  dict(b=1, h=9, w=13, ci=32, co=32), dict(b=2, h=64, w=64, ci=64, co=32)])
def test_stride2_pad1_conv(impl, case):
    """transposed=3: the encoder's stride-2 pad-1 3x3 conv (psp_encoders.py:41-48) with fused bias + LeakyReLU(0.01) (as PReLU)."""
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    dt = torch.bfloat16 if impl == 0 else torch.float32
    x, w = rnd(b, ci, h, w_, seed=1).to(dt).float(), (0.2 * rnd(co, ci, 3, 3, seed=2)).to(dt).float()
    bias, slope = rnd(co, seed=3), torch.full((co,), 0.01)
    y, _ = K().conv3x3(nhwc(x, dt), K().pack_conv_weight(w.to(DEV), dt, impl == 1), co, transposed=3, impl=impl, bias=bias.to(DEV),
                       prelu=slope.to(DEV))
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), stride=2, padding=1), 0.01).float()
    assert y.shape[1:3] == ((h - 1) // 2 + 1, (w_ - 1) // 2 + 1)
    tol = dict(rtol=1e-4, atol=2e-3) if impl == 1 else dict(rtol=2e-2, atol=3e-2)
    torch.testing.assert_close(nchw(y), ref, **tol)


@pytest.mark.parametrize('case', [dict(g=3, b=2, h=8, ci=64, co=128, shared=True), dict(g=4, b=16, h=2, ci=128, co=256, shared=False),
                                  dict(g=2, b=3, h=32, ci=64, co=64, shared=True), dict(g=5, b=2, h=4, ci=64, co=64, shared=False)])
def test_grouped_stride2_conv(case):
    """Grouped form of ood_conv3x3 (the encoder's style heads, one launch per depth): per-group weights / bias / slopes,
    optionally one input shared by every group."""
    g_, b, h, ci, co = case['g'], case['b'], case['h'], case['ci'], case['co']
    dt = torch.bfloat16
    xin = rnd(b if case['shared'] else g_ * b, ci, h, h, seed=1).to(dt).float()
    ws = [(0.2 * rnd(co, ci, 3, 3, seed=10 + i)).to(dt).float() for i in range(g_)]
    bias, slope = rnd(g_, co, seed=3), 0.05 + 0.2 * torch.rand(g_, co, generator=g(4))
    wp = torch.cat([K().pack_conv_weight(w.to(DEV), dt, False) for w in ws], 0)
    y, _ = K().conv3x3(nhwc(xin, dt), wp, co, transposed=3, bias=bias.to(DEV), prelu=slope.to(DEV), groups=g_, in_shared=case['shared'])
    assert y.shape == (g_ * b, (h - 1) // 2 + 1, (h - 1) // 2 + 1, co)
    for i in range(g_):
        xi = xin if case['shared'] else xin[i * b:(i + 1) * b]
        r = torch.nn.functional.conv2d(xi.double(), ws[i].double(), bias[i].double(), stride=2, padding=1)
        ref = torch.where(r > 0, r, r * slope[i].double().reshape(1, -1, 1, 1)).float()
        torch.testing.assert_close(nchw(y[i * b:(i + 1) * b]), ref, rtol=2e-2, atol=3e-2)


@pytest.mark.parametrize('impl', [0, 1])
def test_conv1x1_and_tap_sum(impl):
    """transposed=4 (1x1 form) + ood_tap_sum == a 3x3 pad-1 convolution to 3 channels (the AlignNet 2C -> 3 head)."""
    b, ci, h, w_ = 2, 64, 11, 9
    dt = torch.bfloat16 if impl == 0 else torch.float32
    x, w = rnd(b, ci, h, w_, seed=1).to(dt).float(), (0.2 * rnd(3, ci, 3, 3, seed=2)).to(dt).float()
    w27 = torch.zeros(32, ci)
    w27[:27] = w.permute(2, 3, 0, 1).reshape(27, ci)
    proj, _ = K().conv3x3(nhwc(x, dt), K().pack_conv1x1_weight(w27.to(DEV), dt, impl == 1), 32, transposed=4, impl=impl, out_f32=True)
    assert proj.shape == (b, h, w_, 32) and proj.dtype == torch.float32
    ref1 = torch.einsum('bchw,nc->bhwn', x.double(), w27.double()).float()
    torch.testing.assert_close(proj.cpu(), ref1, rtol=1e-4, atol=2e-3)
    out = K().tap_sum(proj)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1).float()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=5e-3)
    # 1x1 conv with bias + PReLU epilogue, 256 output channels
    w2, bias, slope = (0.2 * rnd(256, ci, seed=5)).to(dt).float(), rnd(256, seed=6), 0.1 + 0.1 * torch.rand(256, generator=g(7))
    if impl == 0:
        y, _ = K().conv3x3(nhwc(x, dt), K().pack_conv1x1_weight(w2.to(DEV), dt, False), 256, transposed=4, bias=bias.to(DEV), prelu=slope.to(DEV))
        r = torch.einsum('bchw,nc->bnhw', x.double(), w2.double()) + bias.double().reshape(1, -1, 1, 1)
        ref2 = torch.where(r > 0, r, r * slope.double().reshape(1, -1, 1, 1)).float()
        torch.testing.assert_close(nchw(y), ref2, rtol=2e-2, atol=3e-2)


def test_modulated_conv_identity_vs_oracle():
    # (W*s*d) (*) x == d . (W (*) (s . x)): the kernels' formulation against the reference's (oracle) formulation
    b, ci, co, h = 2, 64, 32, 12
    x, style = rnd(b, ci, h, h, seed=1), rnd(b, 512, seed=2)
    w, mw, mb = rnd(1, co, ci, 3, 3, seed=3), rnd(ci, 512, seed=4), torch.ones(ci)
    ref = ostyle.modulated_conv2d(x, style, w, mw, mb)
    wsq = K().weight_sumsq(w[0].to(DEV))
    s, d = K().modulation(style.to(DEV), mw.to(DEV), mb.to(DEV), wsq, 1 / math.sqrt(ci * 9), co)
    xs = K().nchw_to_nhwc(x.to(DEV), s, torch.float32)
    y, _ = K().conv3x3(xs, K().pack_conv_weight(w[0].to(DEV), torch.float32, True), co, impl=1, d=d)
    torch.testing.assert_close(nchw(y), ref, rtol=1e-4, atol=1e-4)
    ref_up = ostyle.modulated_conv2d(x, style, w, mw, mb, upsample=True, blur_k=oops.fir_kernel([1, 3, 3, 1], 4.0), blur_pad=(1, 1))
    t, _ = K().conv3x3(xs, K().pack_conv_weight(w[0].to(DEV), torch.float32, True), co, transposed=True, impl=1)
    img, _, _ = K().blur_act(t, K().fir_taps(gain=2.0), d=d, act=False, want_img=True)
    torch.testing.assert_close(nchw(img), ref_up, rtol=1e-4, atol=1e-4)


# ----------------------------------------------------------------------------------------------- blur + epilogue, torgb
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(2, 32, 9, 9), (1, 64, 33, 17), (2, 8, 65, 65), (2, 32, 67, 65), (1, 64, 150, 131)])
def test_blur_act(dtype, shape):
    b, c, ih, iw = shape
    if dtype == torch.float32 and c % 4 or dtype == torch.bfloat16 and c % 8:
        pytest.skip('vector width')
    t = rnd(b, c, ih, iw, seed=1).to(dtype).float()
    d, bias, s_next = 0.5 + rnd(b, c, seed=2).abs(), rnd(c, seed=3), 1 + 0.3 * rnd(b, c, seed=4)
    noise, nw = rnd(b, 1, ih - 1, iw - 1, seed=5), torch.tensor([0.21])
    img_ref = oops.upfirdn2d(t, oops.fir_kernel([1, 3, 3, 1], 4.0), pad=(1, 1)) * d[:, :, None, None]
    y_ref = oops.fused_leaky_relu(img_ref + nw * noise, bias)
    img, y, ys = K().blur_act(nhwc(t, dtype), K().fir_taps(gain=2.0), d=d.to(DEV), noise=noise.to(DEV), noise_w=nw.to(DEV),
                              bias=bias.to(DEV), s_next=s_next.to(DEV), act=True, want_img=True, want_y=True, want_ys=True)
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(nchw(img), img_ref, **tol)
    torch.testing.assert_close(nchw(y), y_ref, **tol)
    torch.testing.assert_close(nchw(ys), y_ref * s_next[:, :, None, None], **tol)
    if dtype == torch.bfloat16:   # fp32 accumulators in, bf16 activations out
        _, y2, _ = K().blur_act(nhwc(t, torch.float32), K().fir_taps(gain=2.0), d=d.to(DEV), noise=noise.to(DEV),
                                noise_w=nw.to(DEV), bias=bias.to(DEV), act=True, dtype=torch.bfloat16)
        assert y2.dtype == torch.bfloat16
        torch.testing.assert_close(nchw(y2), y_ref, **tol)
    y3, ys3 = K().noise_act(nhwc(img_ref, dtype), noise.to(DEV), nw.to(DEV), bias.to(DEV), s_next.to(DEV), True, True)
    torch.testing.assert_close(nchw(y3), y_ref, **tol)
    torch.testing.assert_close(nchw(ys3), y_ref * s_next[:, :, None, None], **tol)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('c,h', [(32, 16), (512, 8), (64, 34)])
def test_torgb(dtype, c, h):
    b = 2
    y = rnd(b, c, h, h, seed=1).to(dtype).float()
    w, s, bias, skip = rnd(3, c, seed=2), 1 + 0.3 * rnd(b, c, seed=3), rnd(3, seed=4), rnd(b, 3, h // 2, h // 2, seed=5)
    wrgb = K().torgb_weight(w.to(DEV), s.to(DEV))
    ref = torch.einsum('bchw,bkc->bkhw', y, w[None] * s[:, None, :] / math.sqrt(c)) + bias.reshape(1, 3, 1, 1)
    out0 = K().torgb(nhwc(y, dtype), wrgb, bias.to(DEV))
    tol = dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out0.cpu(), ref, **tol)
    ref = ref + oops.upfirdn2d(skip, oops.fir_kernel([1, 3, 3, 1], 4.0), up=2, pad=(2, 1))
    out = K().torgb(nhwc(y, dtype), wrgb, bias.to(DEV), skip.to(DEV))
    torch.testing.assert_close(out.cpu(), ref, **tol)


# ----------------------------------------------------------------------------------------------- SAMM kernels
@pytest.mark.parametrize('r', [9, 32, 70])
def test_alignnet_tail_and_folded_field_step(r):
    """ood_alignnet_tail: PReLU -> conv3x3(3->3) -> InstanceNorm plus InstanceNorm(shortcut), returned as an affine form that
    ood_field_step applies on load (bottleneck_IR tail, e4e/encoders/helpers.py:426-448)."""
    b = 2
    res, sc = rnd(b, 3, r, r, seed=1), 0.5 + 2 * rnd(b, 3, r, r, seed=2)
    slope, w = 0.25 + 0.1 * rnd(3, seed=3), 0.3 * rnd(3, 3, 3, 3, seed=4)
    wr, br, ws, bs = 1 + 0.2 * rnd(3, seed=5), rnd(3, seed=6), 1 + 0.2 * rnd(3, seed=7), rnd(3, seed=8)
    F_ = torch.nn.functional
    ref = F_.instance_norm(F_.conv2d(F_.prelu(res, slope), w, padding=1), weight=wr, bias=br, eps=1e-5) + \
        F_.instance_norm(sc, weight=ws, bias=bs, eps=1e-5)
    d = lambda t: t.to(DEV)
    r2, coef = K().alignnet_tail(d(res), d(sc), d(slope), d(w), d(wr), d(br), d(ws), d(bs), 1e-5)
    z = r2 * coef[:, :, 0, None, None] + d(sc) * coef[:, :, 1, None, None] + coef[:, :, 2, None, None]
    torch.testing.assert_close(z.cpu(), ref, rtol=1e-4, atol=1e-4)
    k = oops.fir_kernel([1, 3, 3, 1])
    direct = K().field_step(z, None, None, 0.08)
    folded = K().field_step(r2, None, None, 0.08, z2=d(sc), coef=coef)
    torch.testing.assert_close(folded, direct, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('r', [12, 32, 50])
def test_field_step(r):
    b, scale = 2, 0.08
    z, prev = rnd(b, 3, r, r, seed=1), None
    k = oops.fir_kernel([1, 3, 3, 1])

    def heads(z):
        return torch.cat([torch.tanh(z[:, 0:1]) * scale, torch.tanh(z[:, 1:2]) * scale, torch.sigmoid(z[:, 2:])], 1)

    f1 = oops.upfirdn2d(heads(z), k, pad=(2, 1))
    a1 = K().field_step(z.to(DEV), None, None, scale)
    torch.testing.assert_close(a1.cpu(), f1, rtol=1e-5, atol=1e-6)
    z2 = rnd(b, 3, r, r, seed=2)
    f2 = oops.upfirdn2d(heads(z2), k, pad=(2, 1))
    acc = torch.cat([torch.clip(f1[:, 0:1] + f2[:, 0:1], -scale, scale), torch.clip(f1[:, 1:2] + f2[:, 1:2], -scale, scale),
                     torch.clip(osamm.prm(f1[:, 2:], f2[:, 2:]), 0, 1)], 1)
    coarse = torch.rand(b, 3, r // 2, r // 2, generator=g(3))
    acc_c = torch.cat([acc[:, :2], torch.clip(osamm.prm(coarse[:, 2:], acc[:, 2:]), 0, 1)], 1)
    a2 = K().field_step(z2.to(DEV), a1, None, scale)
    torch.testing.assert_close(a2.cpu(), acc, rtol=1e-5, atol=1e-6)
    a3 = K().field_step(z2.to(DEV), a1, coarse.to(DEV), scale)
    torch.testing.assert_close(a3.cpu(), acc_c, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(2, 16, 12, 12), (1, 64, 32, 20)])
def test_warp_mix(dtype, shape):
    b, c, h, w = shape
    gen = rnd(b, c, h, w, seed=1).to(dtype).float()
    field = torch.cat([0.3 * rnd(b, 2, h, w, seed=2), torch.rand(b, 1, h, w, generator=g(3))], 1)   # large flow: hits the border
    ref = osamm.warp_mix(gen, field)
    out = K().warp_mix(nhwc(gen, dtype), field.to(DEV))
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(nchw(out), ref, **tol)


@pytest.mark.parametrize('size', [64, 128])      # 64: generic taps; 128: every level <= size/4 (shared-tap fast path)
def test_mask_blend(size):
    b = 2
    fields = [torch.rand(b, 3, r, r, generator=g(r)) for r in (4, 8, 16, 32)]
    x, gen = rnd(b, 3, size, size, seed=1), rnd(b, 3, size, size, seed=2)
    alpha = osamm.compose_masks(fields, size)
    ref = osamm.blend(alpha, x, gen)
    out, a = K().mask_blend([f.to(DEV) for f in fields], x.to(DEV), gen.to(DEV))
    torch.testing.assert_close(a.cpu(), alpha, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)
    out1, a1 = K().mask_blend([fields[1].to(DEV)], x.to(DEV), gen.to(DEV))
    torch.testing.assert_close(a1.cpu(), osamm.compose_masks(fields[1:2], size), rtol=1e-5, atol=1e-5)
    # idempotence-style property: blending an image with itself returns it
    same, _ = K().mask_blend([f.to(DEV) for f in fields], x.to(DEV), x.to(DEV))
    torch.testing.assert_close(same.cpu(), x, rtol=1e-6, atol=1e-6)


# ----------------------------------------------------------------------------------------------- image byte formats
def _frames(b, h, w, seed):
    return torch.randint(0, 256, (b, h, w, 3), generator=g(seed), dtype=torch.uint8)


def test_imgio_golden_cases(golden):
    """ood_img2tensor_u8 / ood_tensor2img_u8 against the outputs of the reference's own img2tensor / tensor2img: bit-exact."""
    from ood_gan_inversion_b200 import imgio
    G = golden('imgio.pt')
    for f, t in zip(G['frames'], G['frame_tensors']):
        assert torch.equal(imgio.img2tensor(f.to(DEV))[0].cpu(), t)
    for t, f, f01 in zip(G['tensors'], G['tensor_frames'], G['tensor_frames_rgb01']):
        assert torch.equal(imgio.tensor2img(t.to(DEV), min_max=(-1, 1))[0].cpu(), f)
        assert torch.equal(imgio.tensor2img(t.to(DEV), rgb2bgr=False, min_max=(0, 1))[0].cpu(), f01)


@pytest.mark.parametrize('shape', [(2, 16, 24), (3, 5, 7), (1, 33, 31), (2, 64, 64)])     # vector path and the odd-size scalar path
def test_imgio_vs_oracle(shape):
    from oracle import imgio as oimg
    b, h, w = shape
    fr = _frames(b, h, w, seed=h)
    out = K().img2tensor_u8(fr.to(DEV))
    ref = torch.stack([oimg.frame_to_tensor(f.numpy()) for f in fr])
    assert torch.equal(out.cpu(), ref)
    keep = K().img2tensor_u8(fr.to(DEV), swap_rb=False)
    assert torch.equal(keep.cpu(), ref.flip(1))
    t = rnd(b, 3, h, w, seed=w) * 1.5
    q = K().tensor2img_u8(t.to(DEV), lo=-1.0, hi=1.0)
    refq = torch.stack([torch.from_numpy(oimg.tensor_to_frame(x)) for x in t])
    assert torch.equal(q.cpu(), refq)
    # an unaligned view forces the scalar path: same bytes
    big = torch.zeros(b * h * w * 3 + 1, dtype=torch.uint8, device=DEV)
    big[1:] = fr.to(DEV).reshape(-1)
    assert torch.equal(K().img2tensor_u8(big[1:].view(b, h, w, 3)).cpu(), ref)


def test_imgio_round_trip_full_size():
    """Size-independent property at the bench size: frame -> tensor -> frame is the identity for every byte value."""
    fr = _frames(2, 1024, 1024, seed=3).to(DEV)
    fr[0, 0, :256, 0] = torch.arange(256, dtype=torch.uint8, device=DEV)
    back = K().tensor2img_u8(K().img2tensor_u8(fr), lo=-1.0, hi=1.0)
    assert torch.equal(back, fr)


def test_imgio_byte_serving_loop():
    """imgio.ByteServing under graphs.PipelinedForward (uint8 frames across PCIe both ways): captured, replayed, bit-exact."""
    from oracle import imgio as oimg
    from ood_gan_inversion_b200 import imgio
    from ood_gan_inversion_b200.graphs import PipelinedForward
    net = imgio.ByteServing(lambda x: (x * 0.75 + 0.1, None))
    reqs = [_frames(2, 32, 48, seed=s).pin_memory() for s in range(3)]
    outs = [torch.empty(2, 32, 48, 3, dtype=torch.uint8).pin_memory() for _ in reqs]
    pipe = PipelinedForward(net, reqs[0].to(DEV), depth=2)
    for x, o in zip(reqs, outs):
        pipe.submit(x, o)
    pipe.synchronize()
    for x, o in zip(reqs, outs):
        ref = torch.stack([torch.from_numpy(oimg.tensor_to_frame(oimg.frame_to_tensor(f.numpy()) * 0.75 + 0.1)) for f in x])
        assert torch.equal(o, ref)


def test_imgio_errors_are_loud():
    with pytest.raises(RuntimeError):
        K().img2tensor_u8(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))                       # CPU tensor
    with pytest.raises(RuntimeError):
        K().img2tensor_u8(torch.zeros(1, 4, 4, 3, device=DEV))                              # not uint8
    with pytest.raises(RuntimeError):
        K().tensor2img_u8(torch.zeros(1, 3, 4, 4, device=DEV), lo=1.0, hi=1.0)              # empty range


def test_errors_are_loud():
    with pytest.raises(RuntimeError):
        K().upfirdn2d_nchw(torch.zeros(1, 1, 4, 4), torch.ones(2, 2), 1, 1, 1, 1, 0, 0, 0, 0)      # CPU tensor
    with pytest.raises(RuntimeError):
        K().upfirdn2d_nchw(torch.zeros(1, 1, 2, 2, device=DEV), torch.ones(4, 4, device=DEV), 1, 1, 1, 1, 0, 0, 0, 0)  # empty
    with pytest.raises(RuntimeError):
        K().conv3x3(torch.zeros(1, 4, 4, 24, device=DEV, dtype=torch.bfloat16), torch.zeros(9, 32, 24, device=DEV,
                    dtype=torch.bfloat16), 32, impl=0)                                              # cin % 32


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_bicubic_up_add(dtype):
    x, y = rnd(2, 16, 8, 8, seed=1).to(dtype).float(), rnd(2, 16, 16, 16, seed=2).to(dtype).float()
    ref = F.interpolate(x, size=(16, 16), mode='bicubic', align_corners=True) + y
    out = K().bicubic_up_add(nhwc(x, dtype), nhwc(y, dtype))
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(nchw(out), ref, **tol)
    x2, y2 = rnd(1, 8, 5, 7, seed=3).to(dtype).float(), rnd(1, 8, 9, 16, seed=4).to(dtype).float()     # non-integer ratio
    ref2 = F.interpolate(x2, size=(9, 16), mode='bicubic', align_corners=True) + y2
    torch.testing.assert_close(nchw(K().bicubic_up_add(nhwc(x2, dtype), nhwc(y2, dtype))), ref2, **tol)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_alignnet_norm_kernels(dtype):
    b, c, h, w = 2, 32, 50, 70
    cur, enc = (3 * rnd(b, c, h, w, seed=1) + 1).to(dtype).float(), (0.5 * rnd(b, c, h, w, seed=2) - 2).to(dtype).float()
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    a, e = osamm.instance_norm(cur), osamm.instance_norm(enc)
    z0 = torch.cat([a - e, e], 1)
    w0, b0 = 1 + 0.1 * rnd(2 * c, seed=3), 0.1 * rnd(2 * c, seed=4)
    st6 = K().in_stats(nhwc(cur, dtype), nhwc(enc, dtype))
    torch.testing.assert_close(st6[..., 0].cpu(), cur.mean((2, 3)), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(st6[..., 1].cpu(), torch.rsqrt(cur.var((2, 3), unbiased=False) + 1e-5), rtol=1e-4, atol=1e-4)
    front = K().alignnet_front(nhwc(cur, dtype), nhwc(enc, dtype), st6, w0.to(DEV), b0.to(DEV))
    torch.testing.assert_close(nchw(front), osamm.instance_norm(z0, w0, b0), **tol)
    t = rnd(b, 2 * c, h, w, seed=5).to(dtype).float()
    st2 = K().in_stats(nhwc(t, dtype))
    res = K().alignnet_res0(nhwc(t, dtype), st2, w0.to(DEV), b0.to(DEV), nhwc(cur, dtype), nhwc(enc, dtype), st6)
    torch.testing.assert_close(nchw(res), osamm.instance_norm(t, w0, b0) + z0, **tol)
    app = K().in_apply(nhwc(t, dtype), st2, w0.to(DEV), b0.to(DEV))
    torch.testing.assert_close(nchw(app), osamm.instance_norm(t, w0, b0), **tol)


@pytest.mark.parametrize('impl', [0, 1])
def test_conv3x3_prelu_epilogue(impl):
    b, h, ci, co = 2, 12, 64, 64
    dt = torch.bfloat16 if impl == 0 else torch.float32
    x, w = rnd(b, ci, h, h, seed=1).to(dt).float(), (0.1 * rnd(co, ci, 3, 3, seed=2)).to(dt).float()
    slope = 0.25 + 0.1 * rnd(co, seed=3)
    y, _ = K().conv3x3(nhwc(x, dt), K().pack_conv_weight(w.to(DEV), dt, impl == 1), co, impl=impl, prelu=slope.to(DEV))
    ref = F.prelu(F.conv2d(x.double(), w.double(), padding=1).float(), slope)
    tol = dict(rtol=2e-2, atol=2e-2) if impl == 0 else dict(rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(nchw(y), ref, **tol)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_spm_warp_fused_route_vs_reference_golden(golden, precision):
    """Two-level SPM_Warp (AlignNet on the fused NHWC kernels, channel-padded C=8 -> granule) vs the unmodified reference."""
    import ood_gan_inversion_b200.stylegan as sgm
    from ood_gan_inversion_b200.samm import StyledscaleNshfitBlock
    sgm.set_precision(precision)
    G = golden('samm.pt')
    blk = StyledscaleNshfitBlock(8, 8, 512, scale=0.08, btn=None, cycle_align=2, diff_fAndg=True).to(DEV)
    blk.load_state_dict(G['sd'])
    tol = dict(rtol=1e-3, atol=2e-4) if precision == 'fp32' else dict(rtol=5e-2, atol=2e-2)
    a1, f1 = blk(G['enc'].to(DEV), None, image=G['gen'].to(DEV), aligned=None)
    torch.testing.assert_close(f1.cpu(), G['field'], **tol)
    torch.testing.assert_close(a1.cpu(), G['aligned'], **(tol if precision == 'fp32' else dict(rtol=5e-2, atol=5e-2)))
    a2, f2 = blk(G['enc2'].to(DEV), None, image=G['gen2'].to(DEV), aligned=f1)
    torch.testing.assert_close(f2.cpu(), G['field2'], **tol)
    sgm.set_precision('bf16')


@pytest.mark.parametrize('case', [dict(b=2, h=40, w=128, ci=64, co=64), dict(b=1, h=33, w=256, ci=32, co=32),
                                  dict(b=3, h=5, w=384, ci=64, co=32), dict(b=1, h=70, w=128, ci=32, co=32)])
def test_conv3x3_row_sliding_kernel(case):
    """Row-sliding tcgen05 kernel (W % 128 == 0, Ci, Co in {32, 64}): strips of 32 rows, remainders, several column blocks."""
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    x, w = rnd(b, ci, h, w_, seed=1).bfloat16().float(), rnd(co, ci, 3, 3, seed=2).bfloat16().float()
    d, bias, s_next = 0.05 * (1 + rnd(b, co, seed=3).abs()), rnd(co, seed=4), 1 + 0.3 * rnd(b, co, seed=5)
    noise, nw = rnd(b, 1, h, w_, seed=6), torch.tensor([0.37])
    wp = K().pack_conv_weight(w.to(DEV), torch.bfloat16, False)
    raw, _ = K().conv3x3(nhwc(x, torch.bfloat16), wp, co, impl=0)
    ref_raw = F.conv2d(x.double(), w.double(), padding=1).float()
    torch.testing.assert_close(nchw(raw), ref_raw, rtol=1e-2, atol=0.15)
    y, ys = K().conv3x3(nhwc(x, torch.bfloat16), wp, co, impl=0, d=d.to(DEV), noise=noise.to(DEV), noise_w=nw.to(DEV),
                        bias=bias.to(DEV), s_next=s_next.to(DEV), act=True, want_y=True, want_ys=True)
    ref = oops.fused_leaky_relu(ref_raw * d[:, :, None, None] + nw * noise, bias)
    torch.testing.assert_close(nchw(y), ref, rtol=2e-2, atol=3e-2)
    torch.testing.assert_close(nchw(ys), ref * s_next[:, :, None, None], rtol=2e-2, atol=3e-2)


@pytest.mark.parametrize('case', [dict(b=2, h=16, w=16, ci=64, co=128), dict(b=1, h=32, w=32, ci=128, co=256),   # generic tiles, BN == Co
                                  dict(b=2, h=34, w=128, ci=64, co=64), dict(b=1, h=6, w=256, ci=32, co=32)])   # row-sliding kernel
def test_conv3x3_fused_torgb_epilogue(case):
    """ToRGB (1x1 modconv + bias + up-FIR skip) computed in the conv epilogue == separate torgb kernel == oracle formula."""
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    x, w = rnd(b, ci, h, w_, seed=1).bfloat16().float(), (0.2 * rnd(co, ci, 3, 3, seed=2)).bfloat16().float()
    d, bias, s_next = 0.3 * (1 + rnd(b, co, seed=3).abs()), rnd(co, seed=4), 1 + 0.3 * rnd(b, co, seed=5)
    noise, nw = rnd(b, 1, h, w_, seed=6), torch.tensor([0.37])
    wr, sr, rb, skip = rnd(3, co, seed=7), 1 + 0.3 * rnd(b, co, seed=8), rnd(3, seed=9), rnd(b, 3, h // 2, w_ // 2, seed=10)
    wrgb = K().torgb_weight(wr.to(DEV), sr.to(DEV))
    wp = K().pack_conv_weight(w.to(DEV), torch.bfloat16, False)
    args = dict(impl=0, d=d.to(DEV), noise=noise.to(DEV), noise_w=nw.to(DEV), bias=bias.to(DEV), s_next=s_next.to(DEV), act=True)
    y, ys, rgb = K().conv3x3(nhwc(x, torch.bfloat16), wp, co, want_y=True, want_ys=True,
                             rgb=(wrgb, rb.to(DEV), skip.to(DEV), K().fir_taps(gain=2.0)), **args)
    yref = oops.fused_leaky_relu(F.conv2d(x.double(), w.double(), padding=1).float() * d[:, :, None, None] + nw * noise, bias)
    rgb_ref = torch.einsum('bchw,bkc->bkhw', yref, wr[None] * sr[:, None, :] / math.sqrt(co)) + rb.reshape(1, 3, 1, 1) + \
        oops.upfirdn2d(skip, oops.fir_kernel([1, 3, 3, 1], 4.0), up=2, pad=(2, 1))
    torch.testing.assert_close(nchw(y), yref, rtol=2e-2, atol=3e-2)
    torch.testing.assert_close(rgb.cpu(), rgb_ref, rtol=1e-2, atol=3e-2)
    # rgb only (no activation written), no skip
    y2, ys2, rgb2 = K().conv3x3(nhwc(x, torch.bfloat16), wp, co, want_y=False, want_ys=False,
                                rgb=(wrgb, rb.to(DEV), None, K().fir_taps(gain=2.0)), **args)
    assert y2 is None and ys2 is None
    torch.testing.assert_close(rgb2.cpu(), rgb_ref - oops.upfirdn2d(skip, oops.fir_kernel([1, 3, 3, 1], 4.0), up=2, pad=(2, 1)),
                               rtol=1e-2, atol=3e-2)


# ----------------------------------------------------------------------------------------------- encoder glue kernels
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('case', [dict(c=64, h=6, w=10, stride=1, f32=False), dict(c=128, h=5, w=7, stride=2, f32=True),
                                  dict(c=512, h=4, w=4, stride=1, f32=True)])
def test_se_gate_and_residual(dtype, case):
    """ood_se_gate / ood_se_residual vs SEModule + residual sum + next BatchNorm (e4e/encoders/helpers.py:59-76,476-501)."""
    b, c, h, w_, st = 2, case['c'], case['h'], case['w'], case['stride']
    stream_f32 = case['f32'] and dtype == torch.bfloat16
    v = rnd(b, c, h, w_, seed=1).to(dtype).float()
    sc_full = rnd(b, c, h * st, w_ * st, seed=2)
    sc_full = sc_full if stream_f32 else sc_full.to(dtype).float()
    w1, w2 = 0.2 * rnd(c // 16, c, seed=3), 0.2 * rnd(c, c // 16, seed=4)
    g2, h2 = 1 + 0.1 * rnd(c, seed=5), 0.1 * rnd(c, seed=6)
    gate_ref = torch.sigmoid(torch.relu(v.mean((2, 3)) @ w1.t()) @ w2.t())
    out_ref = v * gate_ref[:, :, None, None] + sc_full[:, :, ::st, ::st]
    tn_ref = out_ref * g2[None, :, None, None] + h2[None, :, None, None]
    vd = nhwc(v, dtype)
    gate = K().se_gate(K().in_stats(vd), w1.to(DEV), w2.to(DEV))
    torch.testing.assert_close(gate.cpu(), gate_ref, rtol=1e-4, atol=1e-4)
    scd = nhwc(sc_full, torch.float32 if stream_f32 else dtype)
    out, tn = K().se_residual(vd, gate, scd, st, g2.to(DEV), h2.to(DEV), out_f32=stream_f32)
    assert out.dtype == (torch.float32 if (stream_f32 or dtype == torch.float32) else torch.bfloat16) and tn.dtype == dtype
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(nchw(out), out_ref, **(dict(rtol=1e-4, atol=1e-4) if stream_f32 else tol))
    torch.testing.assert_close(nchw(tn), tn_ref, **tol)
    # affine-only form (first block's BatchNorm): no gate, no shortcut, no `out`
    o2, t2 = K().se_residual(vd, bn_g=g2.to(DEV), bn_h=h2.to(DEV), want_out=False)
    assert o2 is None
    torch.testing.assert_close(nchw(t2), v * g2[None, :, None, None] + h2[None, :, None, None], **tol)


@pytest.mark.parametrize('case', [dict(b=2, h=24, w=20, ci=64, co=128), dict(b=3, h=16, w=16, ci=128, co=256),
                                  dict(b=1, h=9, w=37, ci=64, co=512)])
def test_conv3x3_seeded_accumulator(case):
    """ood_conv3x3_args.acc_in: a convolution over cat[xa, xb] as conv(xa; W[:, :C]) seeded with the fp32 accumulators of
    conv(xb; W[:, C:]) -- the split AlignNet.raw_nhwc uses across alignment cycles (SAMM/helpers.py:96-101,154-166)."""
    b, h, w, ci, co = (case[k] for k in ('b', 'h', 'w', 'ci', 'co'))
    dt = torch.bfloat16
    xa, xb = rnd(b, ci, h, w, seed=1).to(dt).float(), rnd(b, ci, h, w, seed=2).to(dt).float()
    wt = (0.05 * rnd(co, 2 * ci, 3, 3, seed=3)).to(dt).float()
    slope = 0.25 + 0.1 * rnd(co, seed=4)
    pack = lambda t: K().pack_conv_weight(t.contiguous().to(DEV), dt, False)
    seed, _ = K().conv3x3(nhwc(xb, dt), pack(wt[:, ci:]), co, impl=0, out_f32=True)
    assert seed.dtype == torch.float32
    torch.testing.assert_close(nchw(seed), F.conv2d(xb.double(), wt[:, ci:].double(), padding=1).float(), rtol=1e-4, atol=1e-4)
    y, _ = K().conv3x3(nhwc(xa, dt), pack(wt[:, :ci]), co, impl=0, prelu=slope.to(DEV), acc_in=seed)
    ref = F.prelu(F.conv2d(torch.cat([xa, xb], 1).double(), wt.double(), padding=1).float(), slope)
    torch.testing.assert_close(nchw(y), ref, rtol=1e-2, atol=1e-2)
    # against the unsplit kernel: same fp32 sum in another order, then one bf16 rounding
    full, _ = K().conv3x3(nhwc(torch.cat([xa, xb], 1), dt), pack(wt), co, impl=0, prelu=slope.to(DEV))
    diff = (y.float() - full.float()).abs()
    assert float((diff > 2 ** -7 * full.float().abs().clamp_min(2 ** -6)).float().mean()) == 0.0
    # the same seed in the kernel's tile order (what AlignNet.raw_nhwc uses): identical arithmetic, coalesced fp32 access
    seed_t, _ = K().conv3x3(nhwc(xb, dt), pack(wt[:, ci:]), co, impl=0, out_f32=True, tiled=True)
    assert seed_t.dim() == 1 and seed_t.numel() >= seed.numel()
    y_t, _ = K().conv3x3(nhwc(xa, dt), pack(wt[:, :ci]), co, impl=0, prelu=slope.to(DEV), acc_in=seed_t, tiled=True)
    assert torch.equal(y_t, y)
    # no activation, bias and d on top of the seed
    d, bias = 0.5 + torch.rand(b, co, generator=g(5)), 0.1 * rnd(co, seed=6)
    y2, _ = K().conv3x3(nhwc(xa, dt), pack(wt[:, :ci]), co, impl=0, d=d.to(DEV), bias=bias.to(DEV), acc_in=seed)
    ref2 = F.conv2d(torch.cat([xa, xb], 1).double(), wt.double(), padding=1).float() * d[:, :, None, None] + bias[None, :, None, None]
    torch.testing.assert_close(nchw(y2), ref2, rtol=1e-2, atol=1e-2)
    with pytest.raises(RuntimeError):          # the seed is a tcgen05-path feature
        K().conv3x3(nhwc(xa, torch.float32), K().pack_conv_weight(wt[:, :ci].contiguous().to(DEV), torch.float32, True), co, impl=1,
                    acc_in=seed)


@pytest.mark.parametrize('case', [dict(b=2, h=64, w=128, ci=128, co=256, form=0), dict(b=4, h=64, w=64, ci=64, co=128, form=0),
                                  dict(b=2, h=63, w=50, ci=64, co=512, form=0), dict(b=4, h=128, w=128, ci=128, co=128, form=3),
                                  dict(b=8, h=64, w=32, ci=256, co=256, form=4), dict(b=2, h=64, w=128, ci=64, co=256, form=0, f16=True)])
def test_conv3x3_cta_pair_kernel(case, monkeypatch):
    """conv_tc_kernel<CTA2>: two CTAs of a cluster run one tcgen05.mma.cta_group::2 (M = 256: each CTA's own 128 pixels, half of the
    weight tile per CTA, multicast commits, the leader's barriers).  Same arithmetic as the single-CTA kernel -> bit-identical
    outputs for every epilogue variant (plain with d / noise / bias / activation / s_next, fp32 tile-order seed, fused statistics),
    and the fp64 formula within the storage rounding."""
    b, h, w_, ci, co, form = (case[k] for k in ('b', 'h', 'w', 'ci', 'co', 'form'))
    dt = torch.float16 if case.get('f16') else torch.bfloat16
    k = 1 if form == 4 else 3
    x = rnd(b, ci, h, w_, seed=1).to(dt).float()
    wt = (rnd(co, ci, k, k, seed=2) / (ci * k * k) ** 0.5).to(dt).float()
    from ood_gan_inversion_b200 import kernels as KK
    wp = KK.pack_conv1x1_weight(wt.to(DEV), dt, False) if form == 4 else KK.pack_conv_weight(wt.to(DEV), dt, False)
    xn = nhwc(x, dt)
    stride = 2 if form == 3 else 1
    torch.backends.cudnn.allow_tf32 = False          # launches of >= 100 tiles (narrower tiles below that): the fp32 formula on the device
    ref = F.conv2d(x.to(DEV), wt.to(DEV), stride=stride, padding=0 if form == 4 else 1).cpu()
    oh, ow = ref.shape[2], ref.shape[3]
    d, bias, s_next = 0.5 + torch.rand(b, co, generator=g(3)), rnd(co, seed=4), 1 + 0.3 * rnd(b, co, seed=5)
    noise, nw = rnd(b, 1, oh, ow, seed=6), torch.tensor([0.37])
    slope = 0.25 + 0.1 * rnd(co, seed=7)
    full = dict(transposed=form, impl=0, d=d.to(DEV), noise=noise.to(DEV), noise_w=nw.to(DEV), bias=bias.to(DEV), s_next=s_next.to(DEV), act=True,
                want_y=True, want_ys=True)

    def run():
        out = {}
        out['y'], out['ys'] = KK.conv3x3(xn, wp, co, **full)
        out['f32'], _ = KK.conv3x3(xn, wp, co, transposed=form, impl=0, out_f32=True)
        if form != 3:       # a second convolution seeded with the first one's accumulators (tile order and NHWC)
            seed_t, _ = KK.conv3x3(xn, wp, co, transposed=form, impl=0, out_f32=True, tiled=True)
            out['seeded_t'], _ = KK.conv3x3(xn, wp, co, transposed=form, impl=0, prelu=slope.to(DEV), acc_in=seed_t, tiled=True)
            out['seeded'], _ = KK.conv3x3(xn, wp, co, transposed=form, impl=0, prelu=slope.to(DEV), acc_in=out['f32'])
        if KK.conv3x3_stats_ok(xn, co, form):
            out['sy'], _, out['st'] = KK.conv3x3(xn, wp, co, transposed=form, impl=0, bias=bias.to(DEV), stats_eps=1e-5)
            out['ty'], _, out['sums'] = KK.conv3x3(xn, wp, co, transposed=form, impl=0, bias=bias.to(DEV), tile_sums=True)
        torch.cuda.synchronize()
        return out

    monkeypatch.setenv('OOD_CTA2', '0')
    single = run()
    monkeypatch.setenv('OOD_CTA2', '2')
    monkeypatch.setenv('OOD_CTA2_MIN_TILES', '2')
    pair = run()
    assert set(single) == set(pair) and 'y' in pair
    for name in single:
        if name in ('st', 'sums'):      # the pair form sums the statistics in another (fixed) order: transposed through shared memory
            torch.testing.assert_close(single[name], pair[name], rtol=1e-5, atol=1e-4 if name == 'sums' else 1e-6)
        else:
            assert torch.equal(single[name], pair[name]), name
    if 'st' in pair:
        again = run()
        assert torch.equal(again['st'], pair['st']) and torch.equal(again['sums'], pair['sums'])          # deterministic
    yref = oops.fused_leaky_relu(ref * d[:, :, None, None] + nw * noise, bias)
    tol = dict(rtol=2e-2, atol=3e-2) if dt == torch.bfloat16 else dict(rtol=4e-3, atol=4e-3)
    torch.testing.assert_close(nchw(pair['y']), yref, **tol)
    torch.testing.assert_close(nchw(pair['ys']), yref * s_next[:, :, None, None], **tol)
    torch.testing.assert_close(nchw(pair['f32']), ref, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize('case', [dict(b=2, h=64, w=64, ci=128, co=256), dict(b=2, h=32, w=64, ci=64, co=512), dict(b=3, h=40, w=24, ci=64, co=256)])
def test_conv_transposed_cta_pair_kernel(case, monkeypatch):
    """The stride-2 transposed form (four parity phases, model.py:246-258) on CTA pairs: the two CTAs of a pair take M-adjacent tiles of ONE
    phase, so every phase needs an even number of M tiles (the third case has odd phases and must fall back).  Bit-identical to the
    single-CTA launch and equal to conv_transpose2d."""
    b, h, w_, ci, co = (case[k] for k in ('b', 'h', 'w', 'ci', 'co'))
    x, w = rnd(b, ci, h, w_, seed=1).bfloat16().float(), (0.2 * rnd(co, ci, 3, 3, seed=2)).bfloat16().float()
    wp = K().pack_conv_weight(w.to(DEV), torch.bfloat16, False)
    xn = nhwc(x, torch.bfloat16)
    monkeypatch.setenv('OOD_CTA2', '0')
    y0, _ = K().conv3x3(xn, wp, co, transposed=True, impl=0, out_f32=True)
    monkeypatch.setenv('OOD_CTA2', '2')
    monkeypatch.setenv('OOD_CTA2_MIN_TILES', '2')
    y1, _ = K().conv3x3(xn, wp, co, transposed=True, impl=0, out_f32=True)
    yb, _ = K().conv3x3(xn, wp, co, transposed=True, impl=0)
    torch.cuda.synchronize()
    assert y1.shape == (b, 2 * h + 1, 2 * w_ + 1, co) and torch.equal(y0, y1)
    ref = conv_ref(x.double(), w.double(), True).float()
    torch.testing.assert_close(nchw(y1), ref, rtol=1e-4, atol=2e-3)
    torch.testing.assert_close(nchw(yb), ref, rtol=2e-2, atol=5e-2)


@pytest.mark.parametrize('dtype,c', [(torch.float32, 32), (torch.bfloat16, 64), (torch.bfloat16, 128)])
def test_alignnet_split_front_and_fused_statistics(dtype, c):
    b, h, w = 2, 37, 53
    cur, enc = (2 * rnd(b, c, h, w, seed=1) + 1).to(dtype).float(), (0.5 * rnd(b, c, h, w, seed=2) - 1).to(dtype).float()
    w0, b0 = (1 + 0.1 * rnd(2 * c, seed=3)).to(DEV), (0.1 * rnd(2 * c, seed=4)).to(DEV)
    cu, en = nhwc(cur, dtype), nhwc(enc, dtype)
    st6 = K().in_stats(cu, en)
    front = K().alignnet_front(cu, en, st6, w0, b0)
    lo, hi = K().alignnet_front_split(cu, en, st6, w0, b0)
    assert torch.equal(lo, front[..., :c]) and torch.equal(hi, front[..., c:])
    lo2, none = K().alignnet_front_split(cu, en, st6, w0, b0, want_hi=False)
    assert none is None and torch.equal(lo2, lo)
    t = nhwc(rnd(b, 2 * c, h, w, seed=5), dtype)
    st2 = K().in_stats(t)
    res = K().alignnet_res0(t, st2, w0, b0, cu, en, st6)
    res_s, st = K().alignnet_res0_stats(t, st2, w0, b0, cu, en, st6)
    assert torch.equal(res_s, res)
    ref = K().in_stats(res)
    torch.testing.assert_close(st[..., 0], ref[..., 0], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(st[..., 1], ref[..., 1], rtol=1e-4, atol=1e-5)
    x = res.float()
    torch.testing.assert_close(st[..., 0], x.mean((1, 2)), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(st[..., 1], torch.rsqrt(x.var((1, 2), unbiased=False) + 1e-5), rtol=1e-4, atol=1e-4)


def test_alignnet_cycle_carry_matches_unsplit_route(monkeypatch):
    """AlignNet.raw_nhwc with the per-level carry (enc-only half of the first convolution computed once, both cycles seeded
    with it) against the route that runs the whole convolution every cycle, and against the torch module (raw)."""
    import ood_gan_inversion_b200.stylegan as sgm
    from ood_gan_inversion_b200 import samm
    from ood_gan_inversion_b200.samm import AlignNet
    monkeypatch.setattr(samm, '_SPLIT_MIN_C', 64)          # the pipeline splits from C = 256 up; exercise it on a small level
    sgm.set_precision('bf16')
    torch.manual_seed(0)
    c, r, b = 64, 24, 2
    net = AlignNet(c, 3, scale=0.08).to(DEV)
    for m in net.modules():
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.xavier_normal_(m.weight)
    enc, cur1, cur2 = (nhwc(rnd(b, c, r, r, seed=s), torch.bfloat16) for s in (1, 2, 3))
    carry = {}
    with torch.no_grad():
        a1 = net.raw_nhwc(cur1, enc, carry=carry)
        assert 'seed' in carry and carry['seed'].dtype == torch.float32
        a2 = net.raw_nhwc(cur2, enc, carry=carry)
        p1, p2 = net.raw_nhwc(cur1, enc), net.raw_nhwc(cur2, enc)
        t2 = net.raw(cur2.permute(0, 3, 1, 2), enc.permute(0, 3, 1, 2))
    torch.testing.assert_close(a1, p1, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(a2, p2, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(a2, t2, rtol=5e-2, atol=5e-2)


@pytest.mark.parametrize('h,w,cp', [(70, 45, 32), (8, 32, 28), (33, 65, 27), (1, 1, 32), (256, 256, 32)])
def test_tap_sum_tiled_and_scalar_forms(h, w, cp):
    """ood_tap_sum (tiled shared-memory form when Cp % 4 == 0, scalar form otherwise) vs nine shifted slices."""
    b = 2
    proj = rnd(b, h, w, cp, seed=h + w).to(DEV)
    out = K().tap_sum(proj)
    pad = F.pad(proj, (0, 0, 1, 1, 1, 1))
    ref = torch.zeros(b, 3, h, w, device=DEV)
    for t in range(9):
        ref += pad[:, t // 3:t // 3 + h, t % 3:t % 3 + w, 3 * t:3 * t + 3].permute(0, 3, 1, 2)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    if cp >= 30:
        out2, sc = K().tap_sum(proj, shortcut=True)
        assert torch.equal(out2, out)
        assert torch.equal(sc, proj[..., 27:30].permute(0, 3, 1, 2).contiguous())


@pytest.mark.parametrize('case', [dict(b=2, h=24, w=20, ci=64, co=128), dict(b=3, h=32, w=32, ci=128, co=256),
                                  dict(b=2, h=16, w=8, ci=64, co=512), dict(b=1, h=40, w=72, ci=64, co=256, one=True)])
def test_conv3x3_fused_output_statistics(case):
    """ood_conv3x3_args.stats_out: (mean, rstd) per (image, output channel) of the stored bf16 output, from the epilogue,
    against ood_in_stats over the same tensor and against torch moments; the output itself is unchanged."""
    b, h, w, ci, co = (case[k] for k in ('b', 'h', 'w', 'ci', 'co'))
    dt = torch.bfloat16
    x = nhwc(rnd(b, ci, h, w, seed=1) + 0.3, dt)
    form = 4 if case.get('one') else 0
    wt = (0.05 * rnd(co, ci, 1, 1, seed=3) if form == 4 else 0.05 * rnd(co, ci, 3, 3, seed=3))
    pk = K().pack_conv1x1_weight(wt.reshape(co, ci).to(DEV), dt, False) if form == 4 else K().pack_conv_weight(wt.to(DEV), dt, False)
    assert K().conv3x3_stats_ok(x, co, form)
    y0, _ = K().conv3x3(x, pk, co, transposed=form, impl=0)
    y, _, st = K().conv3x3(x, pk, co, transposed=form, impl=0, stats_eps=1e-5)
    assert torch.equal(y, y0)
    ref = K().in_stats(y)
    torch.testing.assert_close(st[..., 0], ref[..., 0], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(st[..., 1], ref[..., 1], rtol=1e-4, atol=1e-5)
    yf = y.float()
    torch.testing.assert_close(st[..., 0], yf.mean((1, 2)), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(st[..., 1], torch.rsqrt(yf.var((1, 2), unbiased=False) + 1e-5), rtol=1e-3, atol=1e-4)
    # with an activation and a bias in front of the store
    slope, bias = (0.25 + 0.1 * rnd(co, seed=4)).to(DEV), (0.1 * rnd(co, seed=5)).to(DEV)
    y2, _, st2 = K().conv3x3(x, pk, co, transposed=form, impl=0, prelu=slope, bias=bias, stats_eps=1e-5)
    ref2 = K().in_stats(y2)
    torch.testing.assert_close(st2, ref2, rtol=1e-4, atol=1e-5)
    assert not K().conv3x3_stats_ok(nhwc(rnd(1, 64, 8, 8), dt), 128)        # 64 pixels: several images per tile
    if form == 0 and h % 2 == 0 and w % 2 == 0 and (h // 2) * (w // 2) >= 128:   # the encoder's stride-2 pad-1 form
        assert K().conv3x3_stats_ok(x, co, 3)
        y3, _, st3 = K().conv3x3(x, pk, co, transposed=3, impl=0, bias=bias, stats_eps=1e-5)
        torch.testing.assert_close(st3, K().in_stats(y3), rtol=1e-4, atol=1e-5)



@pytest.mark.parametrize('case', [dict(b=2, h=8, w=8, ci=64, co=32), dict(b=1, h=13, w=21, ci=64, co=32), dict(b=2, h=16, w=16, ci=128, co=64),
                                  dict(b=1, h=5, w=40, ci=32, co=32), dict(b=3, h=32, w=32, ci=64, co=128), dict(b=1, h=1, w=1, ci=32, co=32)])
def test_conv_transposed_fused_phases(case):
    """conv3x3 form 5 (four output-parity phases in one GEMM, zero-padded shift weights) == form 1 == conv_transpose2d
    (model.py:246-256), including odd sizes and the last output row / column that only the even phases own."""
    b, h, w, ci, co = (case[k] for k in ('b', 'h', 'w', 'ci', 'co'))
    dt = torch.bfloat16
    x = rnd(b, ci, h, w, seed=1).to(dt).float()
    wt = (0.1 * rnd(co, ci, 3, 3, seed=2)).to(dt).float()
    w9 = K().pack_conv_weight(wt.to(DEV), dt, False)
    wf = K().pack_convt_fused(w9)
    assert wf.shape == (4, 4 * co, ci)
    t5, _ = K().conv3x3(nhwc(x, dt), wf, co, transposed=5, impl=0)
    t1, _ = K().conv3x3(nhwc(x, dt), w9, co, transposed=1, impl=0)
    assert t5.shape == (b, 2 * h + 1, 2 * w + 1, co)
    ref = F.conv_transpose2d(x.double(), wt.double().transpose(0, 1), stride=2).float()
    torch.testing.assert_close(nchw(t5), ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(t5.float(), t1.float(), rtol=1e-2, atol=1e-2)
    f5, _ = K().conv3x3(nhwc(x, dt), wf, co, transposed=5, impl=0, out_f32=True)
    torch.testing.assert_close(nchw(f5), ref, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize('case', [dict(b=2, h=20, w=128, ci=64, co=64), dict(b=1, h=9, w=256, ci=64, co=32), dict(b=2, h=7, w=128, ci=32, co=32)])
def test_conv3x3_encoder_epilogues_on_wide_images(case):
    """The encoder's epilogues (bottleneck_IR_SE, e4e/encoders/helpers.py:476-501) at small channel counts on 128 / 256 px wide
    images: PReLU(conv) and PReLU(conv + bias) on the generic tiles, conv + folded-BatchNorm bias on the row-sliding kernel."""
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    dt = torch.bfloat16
    x, w = rnd(b, ci, h, w_, seed=1).to(dt).float(), (0.1 * rnd(co, ci, 3, 3, seed=2)).to(dt).float()
    slope, bias = 0.25 + 0.1 * rnd(co, seed=3), 0.2 * rnd(co, seed=4)
    wp = K().pack_conv_weight(w.to(DEV), dt, False)
    raw = F.conv2d(x.double(), w.double(), padding=1).float()
    y1, _ = K().conv3x3(nhwc(x, dt), wp, co, impl=0, prelu=slope.to(DEV))
    torch.testing.assert_close(nchw(y1), F.prelu(raw, slope), rtol=2e-2, atol=2e-2)
    y2, _ = K().conv3x3(nhwc(x, dt), wp, co, impl=0, prelu=slope.to(DEV), bias=bias.to(DEV))
    torch.testing.assert_close(nchw(y2), F.prelu(raw + bias[None, :, None, None], slope), rtol=2e-2, atol=2e-2)
    y3, _ = K().conv3x3(nhwc(x, dt), wp, co, impl=0, bias=bias.to(DEV))
    torch.testing.assert_close(nchw(y3), raw + bias[None, :, None, None], rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('case', [dict(b=2, h=16, w=16, ci=64, co=128), dict(b=3, h=9, w=13, ci=128, co=256), dict(b=1, h=128, w=128, ci=64, co=128),
                                  dict(b=16, h=2, w=2, ci=256, co=512)])
def test_conv1x1_stride2(impl, case):
    """ood_conv3x3(transposed=6): out[y,x] = W . in[2y,2x] + bias == F.conv2d(kernel 1, stride 2): the shortcut convolution of the
    encoder's down-sampling bottlenecks (e4e helpers.py:483-486) on the TMA element strides (odd sizes: last row / column kept)."""
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    dt = torch.bfloat16 if impl == 0 else torch.float32
    x = rnd(b, ci, h, w_, seed=1).to(dt).float()
    w = (rnd(co, ci, 1, 1, seed=2) / ci ** 0.5).to(dt).float()
    bias = rnd(co, seed=3)
    from ood_gan_inversion_b200 import kernels as K
    y, _ = K.conv3x3(x.permute(0, 2, 3, 1).contiguous().to(dt).to(DEV), K.pack_conv1x1_weight(w.to(DEV), dt, impl == 1), co, transposed=6,
                     impl=impl, bias=bias.to(DEV), out_f32=(impl == 0))
    ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), stride=2).float()
    assert y.shape == (b, (h - 1) // 2 + 1, (w_ - 1) // 2 + 1, co)
    torch.testing.assert_close(y.float().permute(0, 3, 1, 2).cpu(), ref, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize('hw,size', [((1024, 1024), (256, 256)), ((300, 420), (256, 256)), ((256, 256), (256, 256)), ((100, 64), (256, 256))])
def test_thumbnail_nhwc(hw, size):
    """ood_thumbnail_nhwc == F.interpolate(x, size, mode='bilinear') (OOD_faceGAN_e4e_arch.py:256) in NHWC with channels 3..31 zero:
    down-scaling (the 2x2 centre mean at 4:1), non-integer ratios, identity and up-scaling; fp32 exact to rounding, bf16 = its cast."""
    from ood_gan_inversion_b200 import kernels as K
    x = rnd(2, 3, *hw, seed=5).to(DEV)
    ref = F.interpolate(x, size, mode='bilinear')
    for dt in (torch.float32, torch.bfloat16):
        out = K.thumbnail_nhwc(x, size, 32, dt)
        assert out.shape == (2, size[0], size[1], 32) and out.dtype == dt
        assert float(out[..., 3:].abs().max()) == 0.0
        got = out[..., :3].float().permute(0, 3, 1, 2)
        if dt == torch.float32:
            torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
        else:
            torch.testing.assert_close(got, ref.to(dt).float(), rtol=1e-2, atol=1e-2)
    if hw == (1024, 1024):
        blocks = x.reshape(2, 3, 256, 4, 256, 4)[:, :, :, 1:3, :, 1:3].mean((3, 5))
        torch.testing.assert_close(K.thumbnail_nhwc(x, size, 32, torch.float32)[..., :3].permute(0, 3, 1, 2), blocks, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('case', [dict(b=2, h=16, w=16, ci=64, co=128, form=0), dict(b=2, h=17, w=13, ci=128, co=64, form=3),
                                  dict(b=3, h=8, w=8, ci=64, co=256, form=4), dict(b=2, h=12, w=12, ci=32, co=32, form=0)])
def test_conv3x3_f16_storage(case):
    """OOD_F16 operands on the tcgen05 path (the encoder's storage type): same kernel, A/B format field of the instruction
    descriptor = f16, half-precision epilogue pack (saturating); out_dtype switches the output between f16 and bf16."""
    from ood_gan_inversion_b200 import kernels as K
    b, h, w_, ci, co, form = (case[k] for k in ('b', 'h', 'w', 'ci', 'co', 'form'))
    k = 1 if form == 4 else 3
    x = rnd(b, ci, h, w_, seed=1).half().float()
    w = (rnd(co, ci, k, k, seed=2) / (ci * k * k) ** 0.5).half().float()
    bias = rnd(co, seed=3)
    xp = x.permute(0, 2, 3, 1).contiguous().half().to(DEV)
    wp = K.pack_conv1x1_weight(w.to(DEV), torch.float16, False) if form == 4 else K.pack_conv_weight(w.to(DEV), torch.float16, False)
    ref = F.conv2d(x.double(), w.double(), bias.double(), stride=2 if form == 3 else 1, padding=0 if form == 4 else 1).float()
    y32, _ = K.conv3x3(xp, wp, co, transposed=form, bias=bias.to(DEV), out_f32=True)
    torch.testing.assert_close(y32.permute(0, 3, 1, 2).cpu(), ref, rtol=1e-4, atol=1e-4)
    y16, _ = K.conv3x3(xp, wp, co, transposed=form, bias=bias.to(DEV))
    assert y16.dtype == torch.float16
    torch.testing.assert_close(y16.float().permute(0, 3, 1, 2).cpu(), ref, rtol=2e-3, atol=2e-3)
    yb, _ = K.conv3x3(xp, wp, co, transposed=form, bias=bias.to(DEV), out_dtype=torch.bfloat16)
    assert yb.dtype == torch.bfloat16
    torch.testing.assert_close(yb.float().permute(0, 3, 1, 2).cpu(), ref, rtol=1e-2, atol=1e-2)
    # saturation instead of inf
    big, _ = K.conv3x3(xp, wp, co, transposed=form, bias=torch.full((co,), 1e6, device=DEV))
    assert torch.isfinite(big.float()).all() and float(big.float().max()) == 65504.0


@pytest.mark.parametrize('case', [dict(b=2, h=40, w=128, dt='f16', act=2), dict(b=1, h=19, w=256, dt='f16', act=0),
                                  dict(b=3, h=9, w=128, dt='bf16', act=2), dict(b=5, h=128, w=128, dt='f16', act=2),
                                  dict(b=2, h=37, w=256, dt='f16', act=2, ci=32), dict(b=1, h=64, w=128, dt='bf16', act=2, ci=32)])
def test_conv3x3_row_sliding_kernel_encoder_variant(case, monkeypatch):
    """The encoder's 64 -> 64 convolutions (psp_encoders.py:128-131, helpers.py:114-119 at 256 / 128 px) on the row-sliding kernel:
    f16 operands and outputs, PReLU or bias-only epilogue, strips of 32 / 16 / 8 rows (a strip per SM when the image allows), against
    the fp64 formula."""
    from ood_gan_inversion_b200 import kernels as K
    monkeypatch.setenv('OOD_ROWS_MIN_STRIPS', '1')
    b, h, w_, act = case['b'], case['h'], case['w'], case['act']
    dt = torch.float16 if case['dt'] == 'f16' else torch.bfloat16
    ci = case.get('ci', 64)              # 32: the encoder's input layer (3 image channels padded to 32 -> 64)
    x = rnd(b, ci, h, w_, seed=1).to(dt).float()
    w = (rnd(64, ci, 3, 3, seed=2) / 24.0).to(dt).float()
    bias, slope = rnd(64, seed=3), 0.25 + 0.2 * rnd(64, seed=4)
    xp = x.permute(0, 2, 3, 1).contiguous().to(dt).to(DEV)
    wp = K.pack_conv_weight(w.to(DEV), dt, False)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    if act == 2:
        ref = torch.where(ref > 0, ref, ref * slope.double().reshape(1, -1, 1, 1))
    y, _ = K.conv3x3(xp, wp, 64, bias=bias.to(DEV), **(dict(prelu=slope.to(DEV)) if act == 2 else {}))
    assert y.dtype == dt
    tol = dict(rtol=2e-3, atol=2e-3) if dt == torch.float16 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(y.float().permute(0, 3, 1, 2).cpu(), ref.float(), **tol)
    if dt == torch.float16:          # the bf16 copy the alignment consumes (arch._feats_conv_nhwc)
        yb, _ = K.conv3x3(xp, wp, 64, bias=bias.to(DEV), out_dtype=torch.bfloat16, **(dict(prelu=slope.to(DEV)) if act == 2 else {}))
        assert yb.dtype == torch.bfloat16
        torch.testing.assert_close(yb.float().permute(0, 3, 1, 2).cpu(), ref.float(), rtol=1e-2, atol=1e-2)


def test_round2_entry_points_reject_bad_arguments():
    """The entry points added in round 2 fail loudly outside their envelopes (no silent fallback): ood_se_apply (channel counts, tile-sum shape),
    conv3x3 tile sums / statistics (narrow or grouped launches), ood_act_bwd_fused (no incoming gradient, mismatched shapes)."""
    from ood_gan_inversion_b200 import kernels as K
    v = rnd(2, 8, 8, 96, seed=1).half().to(DEV)
    sums = torch.zeros(2, 1, 96, 2, device=DEV)
    w1, w2 = rnd(6, 96, seed=2).to(DEV), rnd(96, 6, seed=3).to(DEV)
    with pytest.raises(RuntimeError):                       # 96 channels: not 64 / 128 / 256 / 512
        K.se_apply(v, sums, w1, w2, v.float())
    v2 = rnd(2, 8, 8, 128, seed=1).half().to(DEV)
    with pytest.raises(AssertionError):                     # tile sums of another tensor
        K.se_apply(v2, sums, rnd(8, 128, seed=2).to(DEV), rnd(128, 8, seed=3).to(DEV), v2.float())
    x = rnd(1, 8, 8, 64, seed=4).bfloat16().to(DEV)
    wp = K.pack_conv_weight(rnd(64, 64, 3, 3, seed=5).to(DEV), torch.bfloat16, False)
    assert not K.conv3x3_stats_ok(x, 64)                     # 64 output channels: no 128-wide tile
    with pytest.raises(RuntimeError):
        K.conv3x3(x, wp, 64, tile_sums=True)
    y = rnd(2, 8, 8, 32, seed=6).bfloat16().to(DEV)
    d = torch.ones(2, 32, device=DEV)
    with pytest.raises(AssertionError):                     # neither g_in nor a ToRGB gradient
        K.act_bwd_fused(None, None, None, y, d, None, None, None)
    with pytest.raises(AssertionError):                     # g_in of another shape
        K.act_bwd_fused(rnd(2, 8, 4, 32, seed=7).bfloat16().to(DEV), None, None, y, d, None, None, None)
    torch.cuda.synchronize()


def test_encoder_glue_kernels():
    """ood_latent_assemble (psp_encoders.py:199-214 + e4e_arch.py:261), ood_alignnet_head_weights (the InstanceNorm folded into the
    AlignNet head's projection) and se_residual's out_lp copy against their torch formulas."""
    from ood_gan_inversion_b200 import kernels as K
    n, b, dim = 18, 3, 512
    heads, avg, delta = rnd(n, b, dim, seed=1).to(DEV), rnd(1, dim, seed=2).to(DEV), rnd(1, n, dim, seed=3).to(DEV)
    for stage in (0, 2, 6, 17):
        w = [heads[0]] * n
        for i in range(1, stage + 1):
            w[i] = heads[0] + heads[i]
        ref = torch.stack(w, 1)
        torch.testing.assert_close(K.latent_assemble(heads, stage), ref, rtol=0, atol=0)
        torch.testing.assert_close(K.latent_assemble(heads, stage, avg, delta[0]), ref + avg.reshape(1, 1, -1) + delta, rtol=1e-6, atol=1e-6)
    c = 256
    st = torch.stack([rnd(b, c, seed=4), rnd(b, c, seed=5).abs() + 0.5], -1).contiguous().to(DEV)
    in_w, in_b = rnd(c, seed=6).to(DEV), rnd(c, seed=7).to(DEV)
    w27 = torch.zeros(32, c)
    w27[:27] = rnd(27, c, seed=8)
    w27, w1 = w27.to(DEV), rnd(3, c, seed=9).to(DEV)
    g = st[..., 1] * in_w
    h = in_b - st[..., 0] * g
    ref_w = w27.unsqueeze(0) * g.unsqueeze(1)
    for fold in (False, True):
        wps, bias = K.alignnet_head_weights(st, in_w, in_b, w27, w1 if fold else None, dtype=torch.float32)
        r = ref_w.clone()
        if fold:
            r[:, 27:30] = w1
        torch.testing.assert_close(wps, r, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(bias, h @ w27.t(), rtol=1e-4, atol=1e-4)
    wps16, _ = K.alignnet_head_weights(st, in_w, in_b, w27, w1)
    assert wps16.dtype == torch.bfloat16
    torch.testing.assert_close(wps16.float(), r.to(torch.bfloat16).float(), rtol=0, atol=0)
    for dt in (torch.bfloat16, torch.float16):
        v = rnd(2, 6, 5, 64, seed=10).to(dt).to(DEV)
        sc = rnd(2, 6, 5, 64, seed=11).to(DEV)
        gate = torch.rand(2, 64, generator=torch.Generator().manual_seed(12)).to(DEV)
        out, t, lp = K.se_residual(v, gate, sc, 1, torch.ones(64, device=DEV), torch.zeros(64, device=DEV), out_f32=True, want_lp=True)
        ref_o = v.float() * gate[:, None, None, :] + sc
        torch.testing.assert_close(out, ref_o, rtol=1e-6, atol=1e-6)
        assert lp.dtype == dt and torch.equal(lp, ref_o.to(dt)) and torch.equal(t, lp)


@pytest.mark.parametrize('dt', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('case', [dict(b=3, h=8, w=8, c=256, stride=1, sc_f32=True, tap=True), dict(b=2, h=6, w=10, c=64, stride=2, sc_f32=False, tap=False),
                                  dict(b=16, h=4, w=4, c=512, stride=1, sc_f32=True, tap=False), dict(b=1, h=33, w=17, c=128, stride=2, sc_f32=True, tap=True)])
def test_se_tail_cluster_kernel(dt, case):
    """ood_se_tail (one launch, a cluster of 8 CTAs per image, channel sums through distributed shared memory) == in_stats -> se_gate ->
    se_residual, and == the torch formula of SEModule + residual (helpers.py:59-76, 494-501)."""
    from ood_gan_inversion_b200 import kernels as K
    b, h, w, c, stride = (case[k] for k in ('b', 'h', 'w', 'c', 'stride'))
    cr = c // 16
    v = rnd(b, h, w, c, seed=1).to(dt).to(DEV)
    sc = rnd(b, h * stride, w * stride, c, seed=2)
    sc = (sc if case['sc_f32'] else sc.to(dt)).to(DEV)
    w1, w2 = (rnd(cr, c, seed=3) / c ** 0.5).to(DEV), (rnd(c, cr, seed=4) / cr ** 0.5).to(DEV)
    bn_g, bn_h = (1 + 0.1 * rnd(c, seed=5)).to(DEV), (0.1 * rnd(c, seed=6)).to(DEV)
    out, tn, lp = K.se_tail(v, w1, w2, sc, stride, bn_g, bn_h, want_lp=case['tap'])
    mean = v.float().mean((1, 2))
    gate = torch.sigmoid(torch.relu(mean @ w1.t()) @ w2.t())
    ref = v.float() * gate[:, None, None, :] + sc.float()[:, ::stride, ::stride]
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(tn.float(), (ref * bn_g + bn_h).to(dt).float(), rtol=1e-2, atol=1e-2)
    if case['tap']:
        torch.testing.assert_close(lp.float(), ref.to(dt).float(), rtol=1e-2, atol=1e-2)
    g2 = K.se_gate(K.in_stats(v), w1, w2)
    r = K.se_residual(v, g2, sc, stride, bn_g, bn_h, out_f32=True)
    torch.testing.assert_close(out, r[0], rtol=1e-5, atol=1e-5)
    assert torch.equal(out, K.se_tail(v, w1, w2, sc, stride, bn_g, bn_h)[0])          # deterministic


@pytest.mark.parametrize('per', ['2', '4'])
@pytest.mark.parametrize('dt', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('case', [dict(b=3, h=32, w=32, ci=256, co=256, form=0, sc_f32=True, tap=True), dict(b=2, h=33, w=17, ci=128, co=128, form=3, sc_f32=False, tap=False),
                                  dict(b=2, h=16, w=16, ci=512, co=512, form=0, sc_f32=True, tap=False), dict(b=5, h=64, w=64, ci=64, co=128, form=3, sc_f32=True, tap=True)])
def test_se_apply_from_the_convolution_tile_sums(case, dt, per, monkeypatch):
    """The encoder's squeeze-excite tail with the pooling in the producing convolution's epilogue (helpers.py:59-76, 494-501): conv3x3(...,
    tile_sums=True) leaves per-tile channel sums of v AS STORED (f16 or bf16), ood_se_apply turns them into the gate and applies it in one
    streaming pass.  Checked against the torch formula on the stored v, against the cluster kernel (ood_se_tail), for determinism, and
    the tile sums against a direct sum."""
    from ood_gan_inversion_b200 import kernels as K
    monkeypatch.setenv('OOD_SE_APPLY_PER', per)
    b, h, w_, ci, c, form = (case[k] for k in ('b', 'h', 'w', 'ci', 'co', 'form'))
    u = rnd(b, h, w_, ci, seed=1).to(dt).to(DEV)
    wt = (rnd(c, ci, 3, 3, seed=2) / (3 * ci ** 0.5)).to(dt).float()
    wp = K.pack_conv_weight(wt.to(DEV), dt, False)
    bias = rnd(c, seed=7).to(DEV)
    assert K.conv3x3_stats_ok(u, c, form)
    v, _, sums = K.conv3x3(u, wp, c, transposed=form, bias=bias, tile_sums=True)
    v0, _ = K.conv3x3(u, wp, c, transposed=form, bias=bias)
    assert torch.equal(v, v0) and v.dtype == dt
    oh, ow = v.shape[1], v.shape[2]
    torch.testing.assert_close(sums[..., 0].sum(1), v.float().sum((1, 2)), rtol=1e-4, atol=1e-2)
    stride = 2 if form == 3 else 1
    cr = c // 16
    sc = rnd(b, oh * stride, ow * stride, c, seed=3)
    sc = (sc if case['sc_f32'] else sc.to(dt)).to(DEV)
    w1, w2 = (rnd(cr, c, seed=4) / c ** 0.5).to(DEV), (rnd(c, cr, seed=5) / cr ** 0.5).to(DEV)
    bn_g, bn_h = (1 + 0.1 * rnd(c, seed=6)).to(DEV), (0.1 * rnd(c, seed=8)).to(DEV)
    out, tn, lp = K.se_apply(v, sums, w1, w2, sc, stride, bn_g, bn_h, want_lp=case['tap'])
    gate = torch.sigmoid(torch.relu(v.float().mean((1, 2)) @ w1.t()) @ w2.t())
    ref = v.float() * gate[:, None, None, :] + sc.float()[:, ::stride, ::stride]
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(tn.float(), (ref * bn_g + bn_h).to(dt).float(), rtol=1e-2, atol=1e-2)
    if case['tap']:
        torch.testing.assert_close(lp.float(), ref.to(dt).float(), rtol=1e-2, atol=1e-2)
    else:
        assert lp is None
    torch.testing.assert_close(out, K.se_tail(v, w1, w2, sc, stride, bn_g, bn_h)[0], rtol=1e-5, atol=1e-5)
    assert torch.equal(out, K.se_apply(v, sums, w1, w2, sc, stride, bn_g, bn_h)[0])          # deterministic
    out2, tn2, _ = K.se_apply(v, sums, w1, w2, sc, stride)                                  # last block: no following BatchNorm
    assert tn2 is None and torch.equal(out2, out)


@pytest.mark.parametrize('case', [dict(b=2, h=64, w=64, ci=128, co=128), dict(b=1, h=16, w=32, ci=64, co=256), dict(b=3, h=32, w=16, ci=64, co=64),
                                  dict(b=1, h=128, w=128, ci=64, co=64)])
def test_conv_transposed_split_into_exact_tiles(case, monkeypatch):
    """Form 1 on a power-of-two input runs as interior + last row + last column (conv_common.cuh: make_geom_transposed_part): every one
    of the (2h+1) x (2w+1) outputs is written exactly once and equals conv_transpose2d(stride 2); identical to the unsplit route."""
    b, h, w_, ci, co = case['b'], case['h'], case['w'], case['ci'], case['co']
    x, w = rnd(b, ci, h, w_, seed=1).bfloat16().float(), (0.2 * rnd(co, ci, 3, 3, seed=2)).bfloat16().float()
    wp = K().pack_conv_weight(w.to(DEV), torch.bfloat16, False)
    xn = nhwc(x, torch.bfloat16)
    y, _ = K().conv3x3(xn, wp, co, transposed=True, impl=0, out_f32=True)
    assert y.shape == (b, 2 * h + 1, 2 * w_ + 1, co)
    ref = conv_ref(x.double(), w.double(), True).float()
    torch.testing.assert_close(nchw(y), ref, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize('case', [dict(b=2, h=32, w=128), dict(b=1, h=40, w=256), dict(b=3, h=7, w=128), dict(b=1, h=64, w=128)])
def test_conv_transposed_row_streaming_kernel(case, monkeypatch):
    """convt_rows.cu (64 -> 32 channels, W % 128 == 0): phases in the MMA N dimension, every input row loaded once, two output rows per
    position row through bulk stores; interior by this kernel, last row / column by the generic tiles; equals conv_transpose2d(stride 2)."""
    monkeypatch.setenv('OOD_ROWS_MIN_STRIPS', '1')
    b, h, w_, ci, co = case['b'], case['h'], case['w'], 64, 32
    x, w = rnd(b, ci, h, w_, seed=1).bfloat16().float(), (0.2 * rnd(co, ci, 3, 3, seed=2)).bfloat16().float()
    wp = K().pack_conv_weight(w.to(DEV), torch.bfloat16, False)
    y, _ = K().conv3x3(nhwc(x, torch.bfloat16), wp, co, transposed=True, impl=0)
    assert y.shape == (b, 2 * h + 1, 2 * w_ + 1, co) and y.dtype == torch.bfloat16
    ref = conv_ref(x.double(), w.double(), True).float()
    torch.testing.assert_close(nchw(y), ref, rtol=1e-2, atol=0.08)
    # same values as the generic tiles (both round fp32 accumulators of the same products to bf16)
    monkeypatch.setenv('OOD_ROWS_MIN_STRIPS', '100000000')
    y2, _ = K().conv3x3(nhwc(x, torch.bfloat16), wp, co, transposed=True, impl=0)
    torch.testing.assert_close(y.float(), y2.float(), rtol=1e-2, atol=2e-2)
