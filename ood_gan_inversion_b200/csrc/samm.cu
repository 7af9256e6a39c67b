// Spatial-alignment kernels: alignment-field step, flow warp + alpha mix, invertibility-mask compose + blend.
// Reference: src/ops/SAMM/helpers.py:62-77 (new_PRM), :104-107 (tanh/sigmoid heads), :129-147 (add /
// upsample_add), :154 (field blur), :168-177 (grid, grid_sample, mix); src/archs/OOD_faceGAN_e4e_arch.py:315-347
// (blending_mask, blend).  All HBM-bound gathers / elementwise chains fused into one pass each.
#include "common.cuh"

namespace ood {

// ------------------------------------------------------------------ field step
constexpr int FT = 16;   // output tile edge

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// bicubic, align_corners=True, clamped taps (ATen UpSampleBicubic2d semantics)
__device__ float bicubic_ac(const float *__restrict__ src, int rc, int r, int y, int x) {
    const float A = -0.75f;
    const float sc = (r > 1) ? (float)(rc - 1) / (float)(r - 1) : 0.f;
    const float ry = sc * y, rx = sc * x;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    const float ty = ry - iy, tx = rx - ix;
    const float cx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
    const float cy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int yy = min(max(iy - 1 + j, 0), rc - 1);
        float row = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int xx = min(max(ix - 1 + i, 0), rc - 1);
            row += cx[i] * __ldg(src + (int64_t)yy * rc + xx);
        }
        acc += cy[j] * row;
    }
    return acc;
}

// z2 / coef (optional): the field is z*coef[b][ch][0] + z2*coef[b][ch][1] + coef[b][ch][2] -- the two instance norms and the
// residual sum that end the AlignNet (ood_alignnet_tail) folded into this kernel's load.
__global__ void __launch_bounds__(FT * FT) field_step_kernel(const float *__restrict__ z, const float *__restrict__ prev,
                                                              const float *__restrict__ coarse, float *__restrict__ acc,
                                                              float k0, float k1, float k2, float k3, float scale, int R,
                                                              int Rc, const float *__restrict__ z2, const float *__restrict__ coef) {
    __shared__ float sa[3][FT + 3][FT + 4];
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * FT, x0 = blockIdx.x * FT;
    const float *zb = z + (int64_t)b * 3 * R * R;
    const float *z2b = z2 ? z2 + (int64_t)b * 3 * R * R : nullptr;
    for (int i = threadIdx.x; i < 3 * (FT + 3) * (FT + 3); i += FT * FT) {
        const int ch = i / ((FT + 3) * (FT + 3));
        const int r = i % ((FT + 3) * (FT + 3));
        const int ty = r / (FT + 3), tx = r % (FT + 3);
        const int y = y0 + ty - 2, x = x0 + tx - 2;
        float v = 0.f;
        if (y >= 0 && y < R && x >= 0 && x < R) {
            float t = zb[((int64_t)ch * R + y) * R + x];
            if (z2b) {
                const float *cf = coef + ((int64_t)b * 3 + ch) * 3;
                t = fmaf(t, cf[0], fmaf(z2b[((int64_t)ch * R + y) * R + x], cf[1], cf[2]));
            }
            v = (ch < 2) ? tanhf(t) * scale : 1.f / (1.f + expf(-t));
        }
        sa[ch][ty][tx] = v;
    }
    __syncthreads();
    const int ty = threadIdx.x / FT, tx = threadIdx.x % FT;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= R || x >= R) return;
    const float kf[4] = {k0, k1, k2, k3};
    float f[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            float row = 0.f;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) row = fmaf(kf[kx], sa[ch][ty + ky][tx + kx], row);
            s = fmaf(kf[ky], row, s);
        }
        f[ch] = s;
    }
    const int64_t o = (int64_t)b * 3 * R * R + (int64_t)y * R + x;
    const int64_t pl = (int64_t)R * R;
    if (prev) {
        const float p0 = prev[o], p1 = prev[o + pl], p2 = prev[o + 2 * pl];
        f[0] = fminf(fmaxf(p0 + f[0], -scale), scale);
        f[1] = fminf(fmaxf(p1 + f[1], -scale), scale);
        f[2] = fminf(fmaxf(f[2] * p2 + p2 * (1.f - p2), 0.f), 1.f);
    }
    if (coarse) {
        const float u = bicubic_ac(coarse + ((int64_t)b * 3 + 2) * Rc * Rc, Rc, R, y, x);
        f[2] = fminf(fmaxf(f[2] * u + u * (1.f - u), 0.f), 1.f);
    }
    acc[o] = f[0];
    acc[o + pl] = f[1];
    acc[o + 2 * pl] = f[2];
}

// 16 bytes of T <-> N/2 fp32 pairs
template <typename T> __device__ __forceinline__ void wm_load(const T *p, float2 *dst);
template <> __device__ __forceinline__ void wm_load<float>(const float *p, float2 *dst) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
    dst[0] = make_float2(r.x, r.y); dst[1] = make_float2(r.z, r.w);
}
template <> __device__ __forceinline__ void wm_load<__nv_bfloat16>(const __nv_bfloat16 *p, float2 *dst) {
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    dst[0] = unpack_bf16x2(r.x); dst[1] = unpack_bf16x2(r.y); dst[2] = unpack_bf16x2(r.z); dst[3] = unpack_bf16x2(r.w);
}
template <typename T> __device__ __forceinline__ void wm_store(T *p, const float2 *v);
template <> __device__ __forceinline__ void wm_store<float>(float *p, const float2 *v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
}
template <> __device__ __forceinline__ void wm_store<__nv_bfloat16>(__nv_bfloat16 *p, const float2 *v) {
    uint4 r;
    r.x = pack_bf16x2(v[0].x, v[0].y); r.y = pack_bf16x2(v[1].x, v[1].y);
    r.z = pack_bf16x2(v[2].x, v[2].y); r.w = pack_bf16x2(v[3].x, v[3].y);
    *reinterpret_cast<uint4 *>(p) = r;
}

// ------------------------------------------------------------------ AlignNet tail (3-channel fp32)
// bottleneck_IR(2C -> 3) ends with  PReLU(3) -> conv3x3(3 -> 3) -> InstanceNorm(3)  on the residual branch and
// InstanceNorm(3) on the 1x1 shortcut (e4e/encoders/helpers.py:426-448 inside SAMM/helpers.py:97-101).  One pass computes
// r2 = conv(PReLU(res)) and the per-block partial moments of r2 and of the shortcut; the finalize kernel turns them into
// the affine coefficients that ood_field_step applies on load, so the two norms and the sum cost no pass of their own.
constexpr int TT = 16;
__global__ void __launch_bounds__(TT * TT) alignnet_tail_kernel(const float *__restrict__ res, const float *__restrict__ sc,
                                                                 const float *__restrict__ slope, const float *__restrict__ w,
                                                                 float *__restrict__ r2, float *__restrict__ partial, int R) {
    __shared__ float sa[3][TT + 2][TT + 3];
    __shared__ float sw[81];
    __shared__ float red[TT * TT / 32][12];
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * TT, x0 = blockIdx.x * TT;
    const int64_t pl = (int64_t)R * R;
    const float *rb = res + (int64_t)b * 3 * pl;
    if (threadIdx.x < 81) sw[threadIdx.x] = w[threadIdx.x];                    // [o][i][ky][kx]
    for (int i = threadIdx.x; i < 3 * (TT + 2) * (TT + 2); i += TT * TT) {
        const int ch = i / ((TT + 2) * (TT + 2)), r = i % ((TT + 2) * (TT + 2));
        const int ty = r / (TT + 2), tx = r % (TT + 2);
        const int y = y0 + ty - 1, x = x0 + tx - 1;
        float v = 0.f;
        if (y >= 0 && y < R && x >= 0 && x < R) {
            v = rb[ch * pl + (int64_t)y * R + x];
            v = v > 0.f ? v : v * slope[ch];
        }
        sa[ch][ty][tx] = v;
    }
    __syncthreads();
    const int ty = threadIdx.x / TT, tx = threadIdx.x % TT;
    const int y = y0 + ty, x = x0 + tx;
    const bool ok = y < R && x < R;
    float m[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = 0.f;
    if (ok) {
        const int64_t o = (int64_t)b * 3 * pl + (int64_t)y * R + x;
#pragma unroll
        for (int oc = 0; oc < 3; ++oc) {
            float a = 0.f;
#pragma unroll
            for (int ic = 0; ic < 3; ++ic)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) a = fmaf(sw[((oc * 3 + ic) * 3 + ky) * 3 + kx], sa[ic][ty + ky][tx + kx], a);
            r2[o + oc * pl] = a;
            const float s = sc[o + oc * pl];
            m[oc * 4 + 0] = a; m[oc * 4 + 1] = a * a; m[oc * 4 + 2] = s; m[oc * 4 + 3] = s * s;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        float v = m[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float v = 0.f;
        for (int wv = 0; wv < TT * TT / 32; ++wv) v += red[wv][threadIdx.x];
        const int blk = blockIdx.y * gridDim.x + blockIdx.x;
        partial[((int64_t)b * gridDim.x * gridDim.y + blk) * 12 + threadIdx.x] = v;
    }
}

// coef[b][ch] = {rstd_r*w_r, rstd_s*w_s, b_r + b_s - mu_r*rstd_r*w_r - mu_s*rstd_s*w_s}
// One warp per (b, ch): the lanes stride over the block partials (up to 256 of them at 256 px) and combine in a fixed
// shuffle order, so the result is deterministic; a single thread walking all partials took 21 us per launch.
__global__ void __launch_bounds__(128) alignnet_tail_finalize_kernel(const float *__restrict__ partial, int nblocks, double n, float eps,
                                                                      const float *__restrict__ wr, const float *__restrict__ br,
                                                                      const float *__restrict__ ws, const float *__restrict__ bs,
                                                                      float *__restrict__ coef, int total) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;        // b*3 + ch (warp-uniform)
    const int lane = threadIdx.x & 31;
    if (i >= total) return;
    const int b = i / 3, ch = i % 3;
    double s[4] = {0, 0, 0, 0};
    for (int k = lane; k < nblocks; k += 32) {
        const float4 p = __ldg(reinterpret_cast<const float4 *>(partial + ((int64_t)b * nblocks + k) * 12 + ch * 4));
        s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += p.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
    if (lane) return;
    const double mr = s[0] / n, vr = fmax(s[1] / n - mr * mr, 0.0), ms = s[2] / n, vs = fmax(s[3] / n - ms * ms, 0.0);
    const double gr = wr[ch] / sqrt(vr + eps), gs = ws[ch] / sqrt(vs + eps);
    coef[i * 3 + 0] = (float)gr;
    coef[i * 3 + 1] = (float)gs;
    coef[i * 3 + 2] = (float)(br[ch] + bs[ch] - mr * gr - ms * gs);
}

// ------------------------------------------------------------------ warp + alpha mix
// A thread owns VPT channel vectors of one pixel (slab-interleaved so a warp still reads contiguous runs): the grid /
// bilinear-weight arithmetic of the pixel is done once per VPT vectors instead of once per vector (ncu: the one-vector
// version was issue-bound at 70 % with 25 % of DRAM peak).  A block owns a 2-D tile of pixels -- `segw` columns by
// `rows_conc` rows in flight, walked down `WM_ITER` times -- so the four bilinear taps of neighbouring pixels (2x overlap
// in x, 2x in y) hit in L1 instead of being fetched from L2 four times: with the earlier 1-D pixel order the kernel sat
// at the L2 throughput limit (5 sector reads per algorithmic read) at 0.51 of HBM peak.

template <typename T, int VPT>
__global__ void __launch_bounds__(256, 4) warp_mix_kernel(const T *__restrict__ gen, const float *__restrict__ field,
                                                        T *__restrict__ out, int H, int W, int C, int segw, int tiles_x,
                                                        int tiles, int WM_ITER) {
    constexpr int N = Vec<T>::N, N2 = N / 2;
    const int cv = C / N;
    const int cvg = cv / VPT;                 // threads per pixel
    const int lanes_px = blockDim.x / cvg;    // pixels in flight per block
    const int rows_conc = lanes_px / segw;
    const int tile_h = rows_conc * WM_ITER;
    const int b = blockIdx.y;
    const int64_t P = (int64_t)H * W;
    const float *fb = field + (int64_t)b * 3 * P;
    const T *gb = gen + (int64_t)b * P * C;
    T *ob = out + (int64_t)b * P * C;
    const float stepx = W > 1 ? 2.f / (float)(W - 1) : 0.f, stepy = H > 1 ? 2.f / (float)(H - 1) : 0.f;
    const int pl = threadIdx.x / cvg, g = threadIdx.x - pl * cvg;
    const int px = pl % segw, py = pl / segw;
    if (pl >= lanes_px) return;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int x = tx * segw + px;
        if (x >= W) continue;
        // the field of the NEXT row of the walk is fetched before this row's gathers are issued: one memory latency per
        // row instead of two dependent ones
        float fn[3] = {0.f, 0.f, 0.f};
        {
            const int y = ty * tile_h + py;
            if (y < H) {
                const int64_t pix = (int64_t)y * W + x;
                fn[0] = __ldg(fb + pix); fn[1] = __ldg(fb + P + pix); fn[2] = __ldg(fb + 2 * P + pix);
            }
        }
#pragma unroll 1
        for (int it = 0; it < WM_ITER; ++it) {
            const int y = ty * tile_h + it * rows_conc + py;
            if (y >= H) break;
            const int64_t pix = (int64_t)y * W + x;
            const float f0 = fn[0], f1 = fn[1], alpha = fn[2];
            if (it + 1 < WM_ITER && y + rows_conc < H) {
                const int64_t pn = pix + (int64_t)rows_conc * W;
                fn[0] = __ldg(fb + pn); fn[1] = __ldg(fb + P + pn); fn[2] = __ldg(fb + 2 * P + pn);
            }
            // torch.linspace(-1, 1, n): start + i*step in the first half, end - (n-1-i)*step in the second
            const float lx = (x < W / 2) ? (-1.f + stepx * x) : (1.f - stepx * (W - 1 - x));
            const float ly = (y < H / 2) ? (-1.f + stepy * y) : (1.f - stepy * (H - 1 - y));
            const float gx = lx + f0, gy = ly + f1;
            const float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;   // align_corners=False
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const int x0i = (int)fx0, y0i = (int)fy0;
            const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
            float2 wq[4];
            int64_t oq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int xx = x0i + (q & 1), yy = y0i + (q >> 1);
                const bool ok = xx >= 0 && xx < W && yy >= 0 && yy < H;                            // zeros padding
                wq[q] = f2(ok ? ((q & 1) ? wx1 : wx0) * ((q >> 1) ? wy1 : wy0) : 0.f);
                oq[q] = ok ? ((int64_t)yy * W + xx) * C : 0;
            }
            const float2 al = f2(alpha), al1 = f2(1.f - alpha);
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
                const int c = (v * cvg + g) * N;
                float2 t[4][N2], gv[N2], o[N2];
#pragma unroll
                for (int q = 0; q < 4; ++q) wm_load<T>(gb + oq[q] + c, t[q]);
                wm_load<T>(gb + pix * C + c, gv);
#pragma unroll
                for (int j = 0; j < N2; ++j) {
                    // same association as the scalar form: ((w0 t0 + w1 t1) + w2 t2) + w3 t3, then smp*alpha + g*(1-alpha)
                    const float2 smp = fma2(wq[3], t[3][j], fma2(wq[2], t[2][j], fma2(wq[1], t[1][j], mul2(wq[0], t[0][j]))));
                    o[j] = fma2(smp, al, mul2(gv[j], al1));
                }
                wm_store<T>(ob + pix * C + c, o);
            }
        }
    }
}

// ------------------------------------------------------------------ mask compose + blend
struct MaskParams {
    const float *f[4];
    int r[4];
    float scale[4];     // ATen's area_pixel_compute_scale (align_corners=False): float(r) / float(S)
    int n;
};

__device__ __forceinline__ float bilinear_up(const float *__restrict__ a, int r, float scale, int y, int x) {
    // ATen upsample_bilinear2d, align_corners=False
    float sy = fmaxf(((float)y + 0.5f) * scale - 0.5f, 0.f), sx = fmaxf(((float)x + 0.5f) * scale - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < r - 1), x1 = x0 + (x0 < r - 1);
    const float ly1 = sy - y0, lx1 = sx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    return ly0 * (lx0 * __ldg(a + (int64_t)y0 * r + x0) + lx1 * __ldg(a + (int64_t)y0 * r + x1)) +
           ly1 * (lx0 * __ldg(a + (int64_t)y1 * r + x0) + lx1 * __ldg(a + (int64_t)y1 * r + x1));
}

// Four consecutive pixels of one row per thread, 2-D grid (no integer division per thread).  The six streaming 16-byte
// loads (x, gen: 3 planes each) are issued first so that they are in flight while the masks are composed.  When the image
// is at least 4x a mask level (always, in the reference configuration) the four pixels touch at most three mask columns and
// share the two mask rows, so a level costs 6 loads instead of 16.  The first form of this path was issue-bound (ncu: issue
// active 77 %, math-pipe throttle, DRAM at 0.66 of peak): ~15 int<->float conversions (quarter-rate pipe) and a float
// division per level per thread.  Now: the host passes the ATen scale (float(r) / float(S)), one F2I + I2F per axis per
// level (the other pixels' column offset is the exact float difference to the first column), clamped column loads make the
// right tap `column + 1` without a special case, and the three columns are interpolated vertically first
// (14 instead of 24 multiply-adds per level; differs from ATen's horizontal-first order by rounding only).
template <bool FAST>
__global__ void __launch_bounds__(256) mask_blend_kernel(const MaskParams mp, const float *__restrict__ xin,
                                                          const float *__restrict__ gen, float *__restrict__ out,
                                                          float *__restrict__ alpha_out, int S) {
    const int b = blockIdx.z;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= S || y >= S) return;
    const int64_t P = (int64_t)S * S;
    const int64_t pix = (int64_t)y * S + x;
    float4 xv[3], gv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int64_t o = ((int64_t)b * 3 + k) * P + pix;
        xv[k] = __ldg(reinterpret_cast<const float4 *>(xin + o));
        gv[k] = __ldg(reinterpret_cast<const float4 *>(gen + o));
    }
    float A[4] = {0.f, 0.f, 0.f, 0.f};
    const float xf05 = (float)x + 0.5f, yf05 = (float)y + 0.5f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= mp.n) break;
        const int r = mp.r[k];
        const float scale = mp.scale[k];
        const float *a = mp.f[k] + ((int64_t)b * 3 + 2) * r * r;
        float u[4];
        if (FAST) {
            const float sy = fmaxf(yf05 * scale - 0.5f, 0.f);
            const int y0 = (int)sy, y1 = min(y0 + 1, r - 1);
            const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
            const float sx0 = fmaxf(xf05 * scale - 0.5f, 0.f);
            const int base = (int)sx0;
            const float fbase = (float)base;
            const int c1 = min(base + 1, r - 1), c2 = min(base + 2, r - 1);
            const float *r0 = a + y0 * r, *r1 = a + y1 * r;
            const float t0 = __ldg(r0 + base), t1 = __ldg(r0 + c1), t2 = __ldg(r0 + c2);
            const float b0 = __ldg(r1 + base), b1 = __ldg(r1 + c1), b2 = __ldg(r1 + c2);
            const float v0 = ly0 * t0 + ly1 * b0, v1 = ly0 * t1 + ly1 * b1, v2 = ly0 * t2 + ly1 * b2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float sxj = j == 0 ? sx0 : fmaxf((xf05 + (float)j) * scale - 0.5f, 0.f);
                const float f = sxj - fbase;                       // exact; in [0, 2): the pixel's column is base or base + 1
                const bool ge = f >= 1.f;
                const float lx1 = ge ? f - 1.f : f, lx0 = 1.f - lx1;
                u[j] = lx0 * (ge ? v1 : v0) + lx1 * (ge ? v2 : v1);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = bilinear_up(a, r, scale, y, x + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) A[j] = (k == 0) ? u[j] : (u[j] * A[j] + A[j] * (1.f - A[j]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) A[j] = fminf(fmaxf(A[j], 0.f), 1.f);
    if (alpha_out) *reinterpret_cast<float4 *>(alpha_out + (int64_t)b * P + pix) = make_float4(A[0], A[1], A[2], A[3]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int64_t o = ((int64_t)b * 3 + k) * P + pix;
        float4 r;
        r.x = A[0] * xv[k].x + gv[k].x * (1.f - A[0]);
        r.y = A[1] * xv[k].y + gv[k].y * (1.f - A[1]);
        r.z = A[2] * xv[k].z + gv[k].z * (1.f - A[2]);
        r.w = A[3] * xv[k].w + gv[k].w * (1.f - A[3]);
        *reinterpret_cast<float4 *>(out + o) = r;
    }
}

// ------------------------------------------------------------------ bicubic upsample (align_corners=True) + add, NHWC
// FPN top-down merge of the encoder (src/ops/e4e/encoders/helpers.py:504-521): out = bicubic_up(x) + y.
template <typename T>
__global__ void __launch_bounds__(256) bicubic_up_add_kernel(const T *__restrict__ x, const T *__restrict__ y,
                                                              T *__restrict__ out, int h, int w, int H, int W, int C) {
    constexpr int N = Vec<T>::N;
    const int cv = C / N;
    const int b = blockIdx.y;
    const float A = -0.75f;
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const T *xb = x + (int64_t)b * h * w * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)H * W * cv; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = i / cv;
        const int c = (int)(i - pix * cv) * N;
        const int Y = (int)(pix / W), X = (int)(pix - (int64_t)Y * W);
        const float ry = sy * Y, rx = sx * X;
        const int iy = (int)floorf(ry), ix = (int)floorf(rx);
        const float ty = ry - iy, tx = rx - ix;
        const float cx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
        const float cy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
        float acc[N];
#pragma unroll
        for (int j = 0; j < N; ++j) acc[j] = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int yy = min(max(iy - 1 + r, 0), h - 1);
            float row[N];
#pragma unroll
            for (int j = 0; j < N; ++j) row[j] = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int xx = min(max(ix - 1 + q, 0), w - 1);
                const Vec<T> v = load_vec<T>(xb + ((int64_t)yy * w + xx) * C + c);
#pragma unroll
                for (int j = 0; j < N; ++j) row[j] = fmaf(cx[q], v.v[j], row[j]);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) acc[j] = fmaf(cy[r], row[j], acc[j]);
        }
        const int64_t off = ((int64_t)b * H * W + pix) * C + c;
        Vec<T> o;
        if (y) {
            const Vec<T> yv = load_vec<T>(y + off);
#pragma unroll
            for (int j = 0; j < N; ++j) o.v[j] = acc[j] + yv.v[j];
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) o.v[j] = acc[j];
        }
        store_vec<T>(out + off, o);
    }
}

}  // namespace ood

extern "C" int ood_bicubic_up_add(const void *x, const void *y, void *out, int batch, int h, int w, int H, int W,
                                  int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(x && out && batch > 0 && batch <= 65535 && h > 0 && w > 0 && H > 0 && W > 0, "bicubic_up_add: bad arguments");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16 || dtype == OOD_F16, "bicubic_up_add: bad dtype");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(channels % N == 0, "bicubic_up_add: channels (%d) must be a multiple of %d", channels, N);
    const int64_t work = (int64_t)H * W * (channels / N);
    dim3 grid((unsigned)std::min<int64_t>((work + 255) / 256, kNumSMs * 16), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) bicubic_up_add_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (const float *)y, (float *)out, h, w, H, W, channels);
    else if (dtype == OOD_F16) bicubic_up_add_kernel<__half><<<grid, 256, 0, st>>>((const __half *)x, (const __half *)y, (__half *)out, h, w, H, W, channels);
    else bicubic_up_add_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)y, (__nv_bfloat16 *)out, h, w, H, W, channels);
    return check_launch("bicubic_up_add");
}

extern "C" int64_t ood_alignnet_tail_workspace(int batch, int r) {
    const int nb = ood::ceil_div(r, ood::TT);
    return (int64_t)batch * nb * nb * 12 * (int64_t)sizeof(float);
}

extern "C" int ood_alignnet_tail(const float *res, const float *shortcut, const float *prelu_slope, const float *conv_w,
                                 const float *in_res_w, const float *in_res_b, const float *in_sc_w, const float *in_sc_b,
                                 float eps, float *r2, float *workspace, float *coef, int batch, int r, void *stream) {
    using namespace ood;
    OOD_REQUIRE(res && shortcut && prelu_slope && conv_w && in_res_w && in_res_b && in_sc_w && in_sc_b && r2 && workspace && coef,
                "alignnet_tail: null pointer");
    OOD_REQUIRE(batch > 0 && batch <= 65535 && r > 0, "alignnet_tail: bad sizes");
    const int nb = ceil_div(r, TT);
    dim3 grid(nb, nb, batch);
    cudaStream_t st = (cudaStream_t)stream;
    alignnet_tail_kernel<<<grid, TT * TT, 0, st>>>(res, shortcut, prelu_slope, conv_w, r2, workspace, r);
    alignnet_tail_finalize_kernel<<<ceil_div(batch * 3, 4), 128, 0, st>>>(workspace, nb * nb, (double)r * r, eps, in_res_w, in_res_b,
                                                                           in_sc_w, in_sc_b, coef, batch * 3);
    return check_launch("alignnet_tail", 2);
}

extern "C" int ood_field_step(const float *z, const float *prev, const float *coarse, float *acc, const float *taps_host,
                              float scale, int batch, int r, int rc, const float *z2, const float *coef, void *stream) {
    using namespace ood;
    OOD_REQUIRE(z && acc && taps_host && batch > 0 && batch <= 65535 && r > 0, "field_step: bad arguments");
    OOD_REQUIRE(!z2 == !coef, "field_step: z2 and coef come together");
    OOD_REQUIRE(!coarse || rc > 0, "field_step: coarse needs its size");
    dim3 grid(ceil_div(r, FT), ceil_div(r, FT), batch);
    // correlation with the flipped taps (upfirdn2d.py:179)
    field_step_kernel<<<grid, FT * FT, 0, (cudaStream_t)stream>>>(z, prev, coarse, acc, taps_host[3], taps_host[2],
                                                                   taps_host[1], taps_host[0], scale, r, rc, z2, coef);
    return check_launch("field_step");
}

extern "C" int ood_warp_mix(const void *gen, const float *field, void *out, int batch, int h, int w, int channels,
                            int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(gen && field && out && batch > 0 && batch <= 65535 && h > 0 && w > 0, "warp_mix: bad arguments");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "warp_mix: bad dtype");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(channels % N == 0, "warp_mix: channels (%d) must be a multiple of %d", channels, N);
    const int cv = channels / N;
    // measurement knobs, read per call (three getenv per launch; the step's launches are replayed from a CUDA graph)
    const char *e = getenv("OOD_WARP_ITER"), *e2 = getenv("OOD_WARP_SEG"), *e3 = getenv("OOD_WARP_VPT");
    int vpt = cv % 4 == 0 ? 4 : (cv % 2 == 0 ? 2 : 1);
    int WM_ITER = 4;
    // small levels: keep at least ~two blocks per SM in flight (a 32 px level with 4 vectors x 4 rows per thread is 256 blocks
    // of dependent gathers on 148 SMs)
    {
        const int64_t px = (int64_t)batch * h * w;
        auto blocks = [&](int v, int it) { return px * (cv / v) / (256 * (int64_t)it); };
        while (WM_ITER > 1 && blocks(vpt, WM_ITER) < kNumSMs * 4) WM_ITER >>= 1;
        while (vpt > 1 && blocks(vpt, WM_ITER) < kNumSMs * 4) vpt >>= 1;
    }
    if (e) WM_ITER = std::max(1, atoi(e));
    if (e3) { const int v = atoi(e3); if ((v == 1 || v == 2 || v == 4) && cv % v == 0) vpt = v; }
    const int seg_max = e2 ? std::max(1, atoi(e2)) : 16;
    const int cvg = cv / vpt;
    OOD_REQUIRE(cvg <= 256, "warp_mix: too many channels (%d)", channels);
    const int lanes_px = 256 / cvg;
    int segw = 1;
    while (segw * 2 <= lanes_px && segw < seg_max) segw *= 2;      // columns in flight (power of two); the rest are rows
    const int rows_conc = std::max(1, lanes_px / segw);
    const int tiles_x = ceil_div(w, segw), tiles_y = ceil_div(h, rows_conc * WM_ITER);
    const int64_t tiles = (int64_t)tiles_x * tiles_y;
    dim3 grid((unsigned)std::min<int64_t>(tiles, kNumSMs * 16), batch);
    cudaStream_t st = (cudaStream_t)stream;
#define OOD_WARP(T, V) warp_mix_kernel<T, V><<<grid, 256, 0, st>>>((const T *)gen, field, (T *)out, h, w, channels, segw, tiles_x, (int)tiles, WM_ITER)
    if (dtype == OOD_F32) { if (vpt == 4) OOD_WARP(float, 4); else if (vpt == 2) OOD_WARP(float, 2); else OOD_WARP(float, 1); }
    else { if (vpt == 4) OOD_WARP(__nv_bfloat16, 4); else if (vpt == 2) OOD_WARP(__nv_bfloat16, 2); else OOD_WARP(__nv_bfloat16, 1); }
#undef OOD_WARP
    return check_launch("warp_mix");
}

extern "C" int ood_mask_blend(const float *const *fields_host, const int *field_sizes_host, int n_fields, const float *x,
                              const float *gen, float *out, float *alpha_out, int batch, int size, void *stream) {
    using namespace ood;
    OOD_REQUIRE(fields_host && field_sizes_host && n_fields >= 1 && n_fields <= 4, "mask_blend: 1..4 fields supported");
    OOD_REQUIRE(x && gen && out && batch > 0 && batch <= 65535 && size > 0 && size % 4 == 0 && size <= 16384, "mask_blend: bad arguments");
    MaskParams mp{};
    mp.n = n_fields;
    for (int i = 0; i < n_fields; ++i) {
        OOD_REQUIRE(fields_host[i] && field_sizes_host[i] > 0, "mask_blend: bad field %d", i);
        mp.f[i] = fields_host[i];
        mp.r[i] = field_sizes_host[i];
        mp.scale[i] = (float)field_sizes_host[i] / (float)size;
    }
    int tx = 1;
    while (tx < 256 && tx * 4 < size) tx *= 2;            // threads along a row (4 pixels each); the rest of the block are rows
    const dim3 block(tx, 256 / tx);
    const dim3 grid(ceil_div(size / 4, tx), ceil_div(size, (int)block.y), batch);
    bool fast = true;                    // every level at most a quarter of the image: 4 pixels span <= 3 mask columns
    for (int i = 0; i < n_fields; ++i) fast = fast && (int64_t)mp.r[i] * 4 <= size;
    if (fast) mask_blend_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(mp, x, gen, out, alpha_out, size);
    else mask_blend_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(mp, x, gen, out, alpha_out, size);
    return check_launch("mask_blend");
}
