// Backward of the SAMM gather / blend kernels ("grid_sample grads"; 8b: warp_alpha_bwd,
// mask_blend_bwd): the per-thread bodies, written without CUDA-only intrinsics so that the SAME source is
//   * the body of the __global__ kernels in samm_bwd.cu (atomics = atomicAdd), and
//   * compiled by g++ into a host emulation (tests/emu/samm_bwd_emu.cpp, atomics = plain adds, every thread coordinate
//     visited in a loop) that tests/test_samm_bwd_cpu.py checks against torch.autograd through the oracle
//     (oracle/samm.py: warp_mix, compose_masks, blend) without a GPU.
// First correct path: one work item per (pixel, channel group) / per pixel, partial sums merged with atomics; no shared
// memory, no shuffles.  Reference arithmetic: src/ops/SAMM/helpers.py:168-177 (grid_sample bilinear / zeros /
// align_corners=False on a linspace(-1, 1) base grid, alpha mix) and src/archs/OOD_faceGAN_e4e_arch.py:315-347
// (bilinear mask pyramid, y*x + x*(1-x) composition, clip, blend), differentiated by hand.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define OOD_HD __host__ __device__ __forceinline__
#else
#define OOD_HD inline
#endif

namespace ood_bwd {

struct DeviceAdd {       // samm_bwd.cu
#ifdef __CUDACC__
    __device__ __forceinline__ void operator()(float *p, float v) const { atomicAdd(p, v); }
#endif
};
struct HostAdd {         // host emulation: work items run one after another
    void operator()(float *p, float v) const { *p += v; }
};

template <typename T> OOD_HD float ld(const T *p);
template <> OOD_HD float ld<float>(const float *p) { return *p; }

// torch.linspace(-1, 1, n)[i]: start + i*step in the first half, end - (n-1-i)*step in the second (same as warp_mix_kernel)
OOD_HD float linspace_m1_1(int i, int n) {
    const float step = n > 1 ? 2.f / (float)(n - 1) : 0.f;
    return (i < n / 2) ? (-1.f + step * i) : (1.f - step * (n - 1 - i));
}

// ---------------------------------------------------------------------------------------------- warp + alpha mix, backward
// forward (warp_mix_kernel):  out[p,c] = alpha * smp[c] + (1 - alpha) * gen[p,c],   smp[c] = sum_q w_q * gen[tap_q, c]
//   with (ix, iy) = (((lx + dx + 1) * W - 1) / 2, ((ly + dy + 1) * H - 1) / 2), the four bilinear taps around it, zeros outside.
// One work item = (image b, pixel pix, channel group g of G): channels c = g, g + G, g + 2G, ...
//   ggen   [B,H,W,C] fp32, ZERO-INITIALISED by the caller:  += (1 - alpha) * go  at p,  += alpha * w_q * go  at every in-bounds tap
//   gfield [B,3,H,W] fp32, ZERO-INITIALISED by the caller:  d/d dx = (W / 2) * sum_c go * alpha * d smp / d ix,  likewise dy,
//                                                           d/d alpha = sum_c go * (smp - gen[p])
// (ATen's grid_sampler_2d_backward: out-of-bounds taps contribute neither value nor gradient.)
template <typename T, typename Add>
OOD_HD void warp_mix_bwd_item(const T *gen, const float *field, const T *gout, float *ggen, float *gfield, int b, int pix, int g,
                              int G, int H, int W, int C, Add add) {
    const int64_t P = (int64_t)H * W;
    const int y = pix / W, x = pix - y * W;
    const float *fb = field + (int64_t)b * 3 * P;
    const float f0 = fb[pix], f1 = fb[P + pix], alpha = fb[2 * P + pix];
    const float gx = linspace_m1_1(x, W) + f0, gy = linspace_m1_1(y, H) + f1;
    const float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    bool ok[4];
    int64_t off[4];
    float w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int xx = x0 + (q & 1), yy = y0 + (q >> 1);
        ok[q] = xx >= 0 && xx < W && yy >= 0 && yy < H;
        off[q] = ok[q] ? ((int64_t)yy * W + xx) * C : 0;
        w[q] = ((q & 1) ? wx1 : wx0) * ((q >> 1) ? wy1 : wy0);
    }
    const T *gb = gen + (int64_t)b * P * C;
    const T *gob = gout + ((int64_t)b * P + pix) * C;
    float *ggb = ggen + (int64_t)b * P * C;
    float s_ix = 0.f, s_iy = 0.f, s_al = 0.f;
    for (int c = g; c < C; c += G) {
        const float go = ld<T>(gob + c);
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = ok[q] ? ld<T>(gb + off[q] + c) : 0.f;
        const float center = ld<T>(gb + (int64_t)pix * C + c);
        const float smp = ((w[0] * t[0] + w[1] * t[1]) + w[2] * t[2]) + w[3] * t[3];
        add(ggb + (int64_t)pix * C + c, (1.f - alpha) * go);
        const float ga = go * alpha;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (ok[q]) add(ggb + off[q] + c, ga * w[q]);
        s_ix += ga * ((t[1] - t[0]) * wy0 + (t[3] - t[2]) * wy1);
        s_iy += ga * ((t[2] - t[0]) * wx0 + (t[3] - t[1]) * wx1);
        s_al += go * (smp - center);
    }
    float *gf = gfield + (int64_t)b * 3 * P;
    add(gf + pix, s_ix * (0.5f * (float)W));
    add(gf + P + pix, s_iy * (0.5f * (float)H));
    add(gf + 2 * P + pix, s_al);
}

// ---------------------------------------------------------------------------------------------- mask compose + blend, backward
struct MaskBwdParams {
    const float *f[4];     // level fields [B,3,r,r] (alpha = channel 2), ascending
    float *gf[4];          // their gradients [B,3,r,r], ZERO-INITIALISED by the caller; only channel 2 receives
    int r[4];
    float scale[4];        // float(r) / float(S) (ATen area_pixel_compute_scale, align_corners=False)
    int n;
};

struct BilinearTap {
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
};
OOD_HD BilinearTap bilinear_tap(int r, float scale, int y, int x) {       // ATen upsample_bilinear2d, align_corners=False
    BilinearTap t;
    const float sy = fmaxf(((float)y + 0.5f) * scale - 0.5f, 0.f), sx = fmaxf(((float)x + 0.5f) * scale - 0.5f, 0.f);
    t.y0 = (int)sy; t.x0 = (int)sx;
    t.y1 = t.y0 + (t.y0 < r - 1); t.x1 = t.x0 + (t.x0 < r - 1);
    t.ly1 = sy - t.y0; t.lx1 = sx - t.x0; t.ly0 = 1.f - t.ly1; t.lx0 = 1.f - t.lx1;
    return t;
}

// forward (mask_blend_kernel):  u_k = up(alpha_k);  A_1 = u_1;  A_k = u_k * A_{k-1} + A_{k-1} * (1 - A_{k-1});  Ac = clip(A_n, 0, 1);
//                               out_c = Ac * x_c + gen_c * (1 - Ac)
// One work item = one pixel of one image.  gx / ggen [B,3,S,S] are written (either may be NULL); the level gradients are merged
// with atomics (S^2 / r^2 pixels share a mask cell).  torch.clip passes the gradient where 0 <= A_n <= 1 (bounds included).
template <typename Add>
OOD_HD void mask_blend_bwd_item(const MaskBwdParams &mp, const float *xin, const float *gen, const float *gout, float *gx,
                                float *ggen, int b, int y, int x, int S, Add add) {
    const int64_t P = (int64_t)S * S;
    const int64_t pix = (int64_t)y * S + x;
    BilinearTap tap[4];
    float u[4], A[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                  // fixed trip count: the per-level arrays stay in registers
        if (k >= mp.n) break;
        const int r = mp.r[k];
        const float *a = mp.f[k] + ((int64_t)b * 3 + 2) * r * r;
        tap[k] = bilinear_tap(r, mp.scale[k], y, x);
        const BilinearTap &t = tap[k];
        u[k] = t.ly0 * (t.lx0 * a[(int64_t)t.y0 * r + t.x0] + t.lx1 * a[(int64_t)t.y0 * r + t.x1]) +
               t.ly1 * (t.lx0 * a[(int64_t)t.y1 * r + t.x0] + t.lx1 * a[(int64_t)t.y1 * r + t.x1]);
        A[k] = (k == 0) ? u[k] : (u[k] * A[k - 1] + A[k - 1] * (1.f - A[k - 1]));
    }
    float An = A[0];
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (k < mp.n) An = A[k];
    const float Ac = fminf(fmaxf(An, 0.f), 1.f);
    float gA = 0.f;
    for (int c = 0; c < 3; ++c) {
        const int64_t o = ((int64_t)b * 3 + c) * P + pix;
        const float go = gout[o];
        gA += go * (xin[o] - gen[o]);
        if (gx) gx[o] = Ac * go;
        if (ggen) ggen[o] = (1.f - Ac) * go;
    }
    if (!(An >= 0.f && An <= 1.f)) gA = 0.f;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        if (k >= mp.n) continue;
        const float gu = (k == 0) ? gA : gA * A[k - 1];
        if (k > 0) gA = gA * (u[k] + 1.f - 2.f * A[k - 1]);
        const int r = mp.r[k];
        float *ga = mp.gf[k] + ((int64_t)b * 3 + 2) * r * r;
        const BilinearTap &t = tap[k];
        add(ga + (int64_t)t.y0 * r + t.x0, gu * t.ly0 * t.lx0);
        add(ga + (int64_t)t.y0 * r + t.x1, gu * t.ly0 * t.lx1);
        add(ga + (int64_t)t.y1 * r + t.x0, gu * t.ly1 * t.lx0);
        add(ga + (int64_t)t.y1 * r + t.x1, gu * t.ly1 * t.lx1);
    }
}

// ---------------------------------------------------------------------------------------------- field step (heads + FIR + PRM), backward
// forward (field_step_kernel, SAMM/helpers.py:62-77,149-166):
//   h = [tanh(z0) * scale, tanh(z1) * scale, sigmoid(z2)];   f = FIR(h)   (4x4 separable correlation, pad (2, 1), zeros outside)
//   prev:    acc01 = clip(prev01 + f01, -scale, scale);   a1 = f2 * p2 + p2 * (1 - p2);   acc2 = clip(a1, 0, 1)      (else acc = f)
//   coarse:  u = bicubic_up(coarse alpha, align_corners=True);   a2 = acc2 * u + u * (1 - u);   acc2 = clip(a2, 0, 1)
// Two passes, no atomics except the bicubic scatter:
//   pass 1 (item = output pixel): recompute f, push gacc through the clips / PRMs -> gf (gradient at the FIR output, workspace),
//                                 gprev (written), gcoarse[:, 2] (+=, ZERO-INITIALISED by the caller)
//   pass 2 (item = input pixel):  gz = head'(z) * FIR^T(gf)  (gather over the 4x4 outputs whose window holds the pixel)
struct FieldBwdArgs {
    const float *z, *prev, *coarse;      // [B,3,R,R], [B,3,R,R] or NULL, [B,3,Rc,Rc] or NULL
    float kf[4];                         // the FIR taps as the forward applies them (flipped, upfirdn2d.py:179)
    float scale;
    int R, Rc;
};

OOD_HD float fs_head(const FieldBwdArgs &a, int b, int ch, int y, int x) {
    if (y < 0 || y >= a.R || x < 0 || x >= a.R) return 0.f;
    const float t = a.z[(((int64_t)b * 3 + ch) * a.R + y) * a.R + x];
    return ch < 2 ? tanhf(t) * a.scale : 1.f / (1.f + expf(-t));
}

OOD_HD float fs_cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
OOD_HD float fs_cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

template <typename Add>
OOD_HD void field_step_bwd_pass1_item(const FieldBwdArgs &a, const float *gacc, float *gf, float *gprev, float *gcoarse, int b, int y,
                                      int x, Add add) {
    const int R = a.R;
    const int64_t pl = (int64_t)R * R, o = (int64_t)b * 3 * pl + (int64_t)y * R + x;
    float f[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            float row = 0.f;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) row += a.kf[kx] * fs_head(a, b, ch, y + ky - 2, x + kx - 2);
            s += a.kf[ky] * row;
        }
        f[ch] = s;
    }
    float g0 = gacc[o], g1 = gacc[o + pl], g2 = gacc[o + 2 * pl];
    float acc2 = f[2], p2 = 0.f, a1 = 0.f;
    if (a.prev) {
        const float s0 = a.prev[o] + f[0], s1 = a.prev[o + pl] + f[1];
        if (!(s0 >= -a.scale && s0 <= a.scale)) g0 = 0.f;
        if (!(s1 >= -a.scale && s1 <= a.scale)) g1 = 0.f;
        p2 = a.prev[o + 2 * pl];
        a1 = f[2] * p2 + p2 * (1.f - p2);
        acc2 = fminf(fmaxf(a1, 0.f), 1.f);
    }
    if (a.coarse) {
        // bicubic, align_corners=True, clamped taps (ATen UpSampleBicubic2d): value and scatter of its gradient
        const int Rc = a.Rc;
        const float A = -0.75f;
        const float sc = (R > 1) ? (float)(Rc - 1) / (float)(R - 1) : 0.f;
        const float ry = sc * y, rx = sc * x;
        const int iy = (int)floorf(ry), ix = (int)floorf(rx);
        const float ty = ry - iy, tx = rx - ix;
        const float cx[4] = {fs_cubic2(tx + 1.f, A), fs_cubic1(tx, A), fs_cubic1(1.f - tx, A), fs_cubic2(2.f - tx, A)};
        const float cy[4] = {fs_cubic2(ty + 1.f, A), fs_cubic1(ty, A), fs_cubic1(1.f - ty, A), fs_cubic2(2.f - ty, A)};
        const float *src = a.coarse + ((int64_t)b * 3 + 2) * Rc * Rc;
        float u = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int yy = iy - 1 + j < 0 ? 0 : (iy - 1 + j > Rc - 1 ? Rc - 1 : iy - 1 + j);
            float row = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int xx = ix - 1 + i < 0 ? 0 : (ix - 1 + i > Rc - 1 ? Rc - 1 : ix - 1 + i);
                row += cx[i] * src[(int64_t)yy * Rc + xx];
            }
            u += cy[j] * row;
        }
        const float a2 = acc2 * u + u * (1.f - u);
        if (!(a2 >= 0.f && a2 <= 1.f)) g2 = 0.f;
        const float gu = g2 * (acc2 + 1.f - 2.f * u);
        g2 = g2 * u;                                             // gradient at acc2 (before the coarse PRM)
        if (gcoarse) {
            float *dst = gcoarse + ((int64_t)b * 3 + 2) * Rc * Rc;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int yy = iy - 1 + j < 0 ? 0 : (iy - 1 + j > Rc - 1 ? Rc - 1 : iy - 1 + j);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int xx = ix - 1 + i < 0 ? 0 : (ix - 1 + i > Rc - 1 ? Rc - 1 : ix - 1 + i);
                    add(dst + (int64_t)yy * Rc + xx, gu * cy[j] * cx[i]);
                }
            }
        }
    }
    float gp2 = 0.f;
    if (a.prev) {
        if (!(a1 >= 0.f && a1 <= 1.f)) g2 = 0.f;
        gp2 = g2 * (f[2] + 1.f - 2.f * p2);
        g2 = g2 * p2;
    }
    gf[o] = g0; gf[o + pl] = g1; gf[o + 2 * pl] = g2;
    if (gprev && a.prev) { gprev[o] = g0; gprev[o + pl] = g1; gprev[o + 2 * pl] = gp2; }
}

OOD_HD void field_step_bwd_pass2_item(const FieldBwdArgs &a, const float *gf, float *gz, int b, int y, int x) {
    const int R = a.R;
    const int64_t pl = (int64_t)R * R;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float *g = gf + ((int64_t)b * 3 + ch) * pl;
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const int yy = y - ky + 2;
            if (yy < 0 || yy >= R) continue;
            float row = 0.f;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int xx = x - kx + 2;
                if (xx >= 0 && xx < R) row += a.kf[kx] * g[(int64_t)yy * R + xx];
            }
            s += a.kf[ky] * row;
        }
        const int64_t o = ((int64_t)b * 3 + ch) * pl + (int64_t)y * R + x;
        const float t = a.z[o];
        float d;
        if (ch < 2) { const float th = tanhf(t); d = a.scale * (1.f - th * th); }
        else { const float sg = 1.f / (1.f + expf(-t)); d = sg * (1.f - sg); }
        gz[o] = s * d;
    }
}

}  // namespace ood_bwd
