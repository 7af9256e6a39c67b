// Shared description of a 3x3 convolution as "phases" of an implicit GEMM.
//
// A phase is a dense problem  out[b, oy*sy+py, ox*sx+px, :] = sum_t  W[tap_w[t]] . in[b, oy+dy[t], ox+dx[t], :]
// with zero padding outside the input.  The stride-1 pad-1 convolution (model.py:268-272) is one phase with nine
// taps; the stride-2 transposed convolution (model.py:246-256) is four phases (output row/column parity) with
// 4 + 2 + 2 + 1 = 9 taps in total -- the same FLOPs as the reference, no multiplications by inserted zeros.
#pragma once
#include "common.cuh"

namespace ood {

struct ConvPhase {
    int oh, ow;           // phase output grid
    int oy0, ox0;         // origin of the grid (a phase may cover a sub-rectangle: grid point (oy0 + i, ox0 + j))
    int py, px;           // output offset
    int ntaps;
    int dy[9], dx[9], wt[9];
    int m_total;          // batch * oh * ow
};

struct ConvGeom {
    int batch, h, w, cin, cout;
    int OH, OW;           // full output size
    int sy, sx;           // output stride of a phase (1 or 2)
    int isy, isx;         // input stride: input pixel = (oy*isy + dy, ox*isx + dx)   (2 for the strided data-gradient form)
    int nphases;
    ConvPhase ph[4];
};

static inline ConvGeom make_geom(int batch, int h, int w, int cin, int cout, int transposed) {
    ConvGeom g{};
    g.batch = batch; g.h = h; g.w = w; g.cin = cin; g.cout = cout;
    g.isy = g.isx = 1;
    if (transposed == 2) {
        // stride-2 "valid" 3x3 conv: out[y,x] = sum in[2y+ky, 2x+kx] * W[ky,kx] -- the data gradient of the stride-2
        // transposed conv (input = gradient of the (2h+1)x(2w+1) tensor, output h x w)
        g.OH = (h - 1) / 2; g.OW = (w - 1) / 2; g.sy = g.sx = 1; g.isy = g.isx = 2; g.nphases = 1;
        ConvPhase &p = g.ph[0];
        p.oh = g.OH; p.ow = g.OW; p.py = p.px = 0; p.ntaps = 9;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
                const int t = ky * 3 + kx;
                p.dy[t] = ky; p.dx[t] = kx; p.wt[t] = t;
            }
        p.m_total = batch * p.oh * p.ow;
    } else if (transposed == 5) {
        // the stride-2 transposed conv with its four output-parity phases fused into ONE implicit GEMM: a tile is 128 input
        // positions (oy, ox) of the (h+1) x (w+1) grid, N = 4*Co (phase-major columns), K = 4 input shifts x Ci.  Shift
        // (dy, dx) in {0,-1}^2 feeds every phase that has a tap there (weights zero-padded elsewhere: ood_b200.h), so the
        // input patch is read 4 times instead of 9 and a tile's outputs are 2x2 pixel blocks (full 128-byte runs).  Built for
        // the small-channel 512 / 1024 px layers, which are bound by per-tile overheads and HBM, not by the tensor pipe.
        g.OH = 2 * h + 1; g.OW = 2 * w + 1; g.sy = g.sx = 2; g.nphases = 1;
        ConvPhase &p = g.ph[0];
        p.oh = h + 1; p.ow = w + 1; p.py = p.px = 0; p.ntaps = 4;
        for (int t = 0; t < 4; ++t) { p.dy[t] = -(t >> 1); p.dx[t] = -(t & 1); p.wt[t] = t; }
        p.m_total = batch * p.oh * p.ow;
    } else if (transposed == 4) {
        // 1x1 convolution (one tap, no padding): the encoder's lateral / feature convolutions and the per-tap projection
        // of the AlignNet's 2C -> 3 head (see ood_tap_sum)
        g.OH = h; g.OW = w; g.sy = g.sx = 1; g.nphases = 1;
        ConvPhase &p = g.ph[0];
        p.oh = h; p.ow = w; p.py = p.px = 0; p.ntaps = 1;
        p.dy[0] = p.dx[0] = 0; p.wt[0] = 0;
        p.m_total = batch * h * w;
    } else if (transposed == 6) {
        // 1x1 stride-2 convolution (one tap, no padding): out[y,x] = W . in[2y, 2x] -- the shortcut convolutions of the encoder's
        // down-sampling bottlenecks (helpers.py:483-486: Conv2d(in, depth, (1,1), stride) + BatchNorm2d)
        g.OH = (h - 1) / 2 + 1; g.OW = (w - 1) / 2 + 1; g.sy = g.sx = 1; g.isy = g.isx = 2; g.nphases = 1;
        ConvPhase &p = g.ph[0];
        p.oh = g.OH; p.ow = g.OW; p.py = p.px = 0; p.ntaps = 1;
        p.dy[0] = p.dx[0] = 0; p.wt[0] = 0;
        p.m_total = batch * p.oh * p.ow;
    } else if (transposed == 3) {
        // stride-2 pad-1 3x3 conv: out[y,x] = sum in[2y+ky-1, 2x+kx-1] * W[ky,kx] -- the down-sampling convolutions of the
        // E4E encoder (GradualStyleBlock psp_encoders.py:41-48, bottleneck_IR_SE helpers.py:488-491)
        g.OH = (h - 1) / 2 + 1; g.OW = (w - 1) / 2 + 1; g.sy = g.sx = 1; g.isy = g.isx = 2; g.nphases = 1;
        ConvPhase &p = g.ph[0];
        p.oh = g.OH; p.ow = g.OW; p.py = p.px = 0; p.ntaps = 9;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
                const int t = ky * 3 + kx;
                p.dy[t] = ky - 1; p.dx[t] = kx - 1; p.wt[t] = t;
            }
        p.m_total = batch * p.oh * p.ow;
    } else if (!transposed) {
        g.OH = h; g.OW = w; g.sy = g.sx = 1; g.nphases = 1;
        ConvPhase &p = g.ph[0];
        p.oh = h; p.ow = w; p.py = p.px = 0; p.ntaps = 9;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
                const int t = ky * 3 + kx;
                p.dy[t] = ky - 1; p.dx[t] = kx - 1; p.wt[t] = t;
            }
        p.m_total = batch * h * w;
    } else {
        // out[Y,X] += in[y,x] * W[ky,kx] with Y = 2y+ky, X = 2x+kx  (conv_transpose2d, stride 2, no padding)
        g.OH = 2 * h + 1; g.OW = 2 * w + 1; g.sy = g.sx = 2; g.nphases = 4;
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                ConvPhase &p = g.ph[py * 2 + px];
                p.oh = h + 1 - py; p.ow = w + 1 - px; p.py = py; p.px = px; p.ntaps = 0;
                for (int ky = py; ky < 3; ky += 2)          // ky parity == Y parity
                    for (int kx = px; kx < 3; kx += 2) {
                        const int t = p.ntaps++;
                        p.dy[t] = (ky == 2) ? -1 : 0;       // Y = 2*oy (+py): ky=0 -> y=oy, ky=2 -> y=oy-1, ky=1 -> y=oy
                        p.dx[t] = (kx == 2) ? -1 : 0;
                        p.wt[t] = ky * 3 + kx;
                    }
                p.m_total = batch * p.oh * p.ow;
            }
    }
    return g;
}

// The stride-2 transposed convolution (form 1) in three parts whose phase grids are exact powers of two.  The four parity phases of a
// 2^k input have (2^k + 1 - py) x (2^k + 1 - px) outputs: rectangular 128-pixel M tiles quantise 65 columns to 128 (half-empty tiles:
// the 512-channel layers ran at 0.43-0.48 of the tensor peak at 32-128 px).  part 0: every phase restricted to its 2^k x 2^k
// interior; part 1: the last output row Y = 2h (phases py = 0 at oy = h); part 2: the last output column X = 2w without the corner
// (phases px = 0 at ox = w, oy < h).  Same taps, same weights, same arithmetic per output as make_geom(transposed = 1).
static inline ConvGeom make_geom_transposed_part(int batch, int h, int w, int cin, int cout, int part) {
    const ConvGeom full = make_geom(batch, h, w, cin, cout, 1);
    ConvGeom g = full;
    g.nphases = 0;
    for (int i = 0; i < 4; ++i) {
        ConvPhase p = full.ph[i];
        if (part == 0) { p.oh = h; p.ow = w; }
        else if (part == 1) { if (p.py != 0) continue; p.oy0 = h; p.oh = 1; }                     // ow stays w + 1 - px: the corner lives here
        else { if (p.px != 0) continue; p.ox0 = w; p.ow = 1; p.oh = h; }
        p.m_total = batch * p.oh * p.ow;
        g.ph[g.nphases++] = p;
    }
    return g;
}

// Fused StyledConv epilogue parameters (any pointer may be null).
struct ConvEpilogue {
    void *out_y, *out_ys;
    const float *d, *noise, *noise_w, *bias, *s_next;
    int64_t noise_bstride;
    int act;                // 0 none | 1 leaky-ReLU(0.2)*sqrt2 | 2 PReLU(prelu[c])
    int out_f32;
    const float *prelu;     // [Co], act == 2
    const float *acc_in;    // fp32 NHWC [B,OH,OW,Co] added to the accumulator first (tcgen05 path), or null
    float *stat_partial;    // STATS kernels: [B][m tiles per image][Co][2] partial moments of the stored output
    int stat_sums_only;     // STATS kernels: only the sums (the second moment's transpose-reduce is skipped, its slot is written as 0)
    int tiled;              // acc_in / fp32 out_y in tile order: float4 index ((tile*(BN/32) + chunk)*8 + j)*128 + row
    // fused ToRGB (model.py:363-372): rgb_out[b,k,Y,X] = sum_o y[o]*rgb_w[b,k,o] + rgb_bias[k] + up2fir(rgb_skip)[b,k,Y,X]
    const float *rgb_w, *rgb_bias, *rgb_skip;
    float *rgb_out;
    float rgb_k[4];         // flipped 1-D up-FIR taps
};

static inline ConvEpilogue make_epilogue(const ood_conv3x3_args &a, int out_f32) {
    ConvEpilogue e{};
    e.out_y = a.out_y; e.out_ys = a.out_ys; e.d = a.d; e.noise = a.noise; e.noise_w = a.noise_w; e.bias = a.bias;
    e.s_next = a.s_next; e.noise_bstride = a.noise_bstride; e.act = a.act; e.out_f32 = out_f32; e.prelu = a.prelu_slope; e.acc_in = a.acc_in; e.tiled = a.tiled;
    e.rgb_w = a.rgb_w; e.rgb_bias = a.rgb_bias; e.rgb_skip = a.rgb_skip; e.rgb_out = a.rgb_out;
    for (int i = 0; i < 4; ++i) e.rgb_k[i] = a.rgb_taps[3 - i];
    return e;
}

// bias + 2x2-tap polyphase up-FIR of the previous level's RGB (upfirdn2d up=2, pad (2,1)) for colour plane k of pixel (Y, X)
__device__ __forceinline__ float rgb_finish(const ConvEpilogue &e, float partial, int b, int k, int Y, int X, int H, int W) {
    float v = partial + __ldg(e.rgb_bias + k);
    if (e.rgb_skip) {
        const int h2 = H >> 1, w2 = W >> 1;
        const float *sp = e.rgb_skip + ((int64_t)b * 3 + k) * h2 * w2;
        const int ky0 = Y & 1, kx0 = X & 1;
        const int ra = (Y + ky0 - 2) >> 1, ca = (X + kx0 - 2) >> 1;
        const float wy[2] = {ky0 ? e.rgb_k[1] : e.rgb_k[0], ky0 ? e.rgb_k[3] : e.rgb_k[2]};
        const float wx[2] = {kx0 ? e.rgb_k[1] : e.rgb_k[0], kx0 ? e.rgb_k[3] : e.rgb_k[2]};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int rr = ra + dy;
            if (rr < 0 || rr >= h2) continue;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int cc = ca + dx;
                if (cc < 0 || cc >= w2) continue;
                v = fmaf(wy[dy] * wx[dx], __ldg(sp + (int64_t)rr * w2 + cc), v);
            }
        }
    }
    return v;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == 1) return lrelu_sqrt2(v);
    if (act == 2) return v > 0.f ? v : v * slope;
    return v;
}

}  // namespace ood
