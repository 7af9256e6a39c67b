// NHWC bandwidth kernels of the StyledConv tail:
//   ood_blur_act : 4x4 separable FIR blur (pad (1,1)) of the transposed-conv output fused with
//                  demodulation, noise injection, bias, leaky-ReLU*sqrt2 and the NEXT layer's style scale;
//   ood_noise_act: the same epilogue without the blur (after the alignment callback replaced the image).
// Reference: Blur/upfirdn2d (src/ops/StyleGAN/model.py:71-88,257; src/ops/op/upfirdn2d.py:160-193),
// NoiseInjection (model.py:283-292), FusedLeakyReLU (src/ops/op/fused_act.py:96): five full-tensor passes there,
// one here.  HBM-bound: each thread owns a 16-byte channel vector for XPT adjacent pixels and slides down a strip
// of rows keeping the four most recent horizontally-filtered rows in registers, so every input element is read
// from DRAM once (horizontal neighbours hit L1) and every output is written once with 16-byte stores.
#include "common.cuh"

namespace ood {

constexpr int kStrip = 16;   // output rows per CTA strip

struct BlurParams {
    const void *in;
    void *out_img, *out_y, *out_ys;
    const float *d, *noise, *noise_w, *bias, *s_next;
    int64_t noise_bstride;
    float k[4];          // flipped 1-D taps
    int batch, ih, iw, oh, ow, C;
    int act;
    int pad0;            // leading pad: input pixel = output pixel - pad0 + tap
};

template <typename TIN, int N>
__device__ __forceinline__ void load_n(const TIN *p, float *dst);
template <> __device__ __forceinline__ void load_n<float, 4>(const float *p, float *dst) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
    dst[0] = r.x; dst[1] = r.y; dst[2] = r.z; dst[3] = r.w;
}
template <> __device__ __forceinline__ void load_n<float, 8>(const float *p, float *dst) {
    load_n<float, 4>(p, dst);
    load_n<float, 4>(p + 4, dst + 4);
}
template <> __device__ __forceinline__ void load_n<__nv_bfloat16, 8>(const __nv_bfloat16 *p, float *dst) {
    const Vec<__nv_bfloat16> v = load_vec<__nv_bfloat16>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = v.v[i];
}

template <typename T, int N>
__device__ __forceinline__ void epilogue_store(const BlurParams &p, float *v, int b, int oy, int ox, int c,
                                               const float *dreg, const float *breg, const float *sreg, float nw) {
    const int64_t off = (((int64_t)b * p.oh + oy) * p.ow + ox) * p.C + c;
    Vec<T> o;
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] *= dreg[j];
    if (p.out_img) {
#pragma unroll
        for (int j = 0; j < N; ++j) o.v[j] = v[j];
        store_vec<T>((T *)p.out_img + off, o);
    }
    if (!p.act) return;
    const float nz = p.noise ? nw * __ldg(p.noise + b * p.noise_bstride + (int64_t)oy * p.ow + ox) : 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = lrelu_sqrt2(v[j] + nz + breg[j]);
    if (p.out_y) {
#pragma unroll
        for (int j = 0; j < N; ++j) o.v[j] = v[j];
        store_vec<T>((T *)p.out_y + off, o);
    }
    if (p.out_ys) {
#pragma unroll
        for (int j = 0; j < N; ++j) o.v[j] = v[j] * sreg[j];
        store_vec<T>((T *)p.out_ys + off, o);
    }
}

template <typename T, typename TIN, int XPT>
__global__ void __launch_bounds__(128) blur_act_kernel(const BlurParams p) {
    constexpr int N = Vec<T>::N;
    const int cv = p.C / N;
    const int xgroups = (p.ow + XPT - 1) / XPT;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)xgroups * cv) return;
    const int c = (int)(idx % cv) * N;
    const int ox0 = (int)(idx / cv) * XPT;
    const int oy0 = blockIdx.y * kStrip;
    const int b = blockIdx.z;

    float dreg[N], breg[N], sreg[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        dreg[j] = p.d ? p.d[(int64_t)b * p.C + c + j] : 1.f;
        breg[j] = (p.act && p.bias) ? p.bias[c + j] : 0.f;
        sreg[j] = p.out_ys ? p.s_next[(int64_t)b * p.C + c + j] : 1.f;
    }
    const float nw = (p.noise && p.noise_w) ? *p.noise_w : 0.f;
    const TIN *src = reinterpret_cast<const TIN *>(p.in) + (int64_t)b * p.ih * p.iw * p.C + c;

    float hw[4][XPT][N];   // horizontally filtered rows, slot = input-row index & 3 (compile-time after unrolling)
#pragma unroll
    for (int i = 0; i < kStrip + 3; ++i) {
        const int iy = oy0 - p.pad0 + i;
        // ---- horizontal pass for input row iy ----
        float row[XPT + 3][N];
        const bool row_ok = iy >= 0 && iy < p.ih;
#pragma unroll
        for (int q = 0; q < XPT + 3; ++q) {
            const int ix = ox0 - p.pad0 + q;
            if (row_ok && ix >= 0 && ix < p.iw) {
                load_n<TIN, N>(src + ((int64_t)iy * p.iw + ix) * p.C, row[q]);
            } else {
#pragma unroll
                for (int j = 0; j < N; ++j) row[q][j] = 0.f;
            }
        }
#pragma unroll
        for (int xp = 0; xp < XPT; ++xp)
#pragma unroll
            for (int j = 0; j < N; ++j)
                hw[i & 3][xp][j] = p.k[0] * row[xp][j] + p.k[1] * row[xp + 1][j] + p.k[2] * row[xp + 2][j] + p.k[3] * row[xp + 3][j];
        // ---- vertical pass: output row oy uses input rows oy-1 .. oy+2  (loop steps i-3 .. i) ----
        if (i >= 3) {
            const int oy = oy0 + i - 3;
            if (oy < p.oh) {
#pragma unroll
                for (int xp = 0; xp < XPT; ++xp) {
                    if (ox0 + xp < p.ow) {
                        float v[N];
#pragma unroll
                        for (int j = 0; j < N; ++j)
                            v[j] = p.k[0] * hw[(i - 3) & 3][xp][j] + p.k[1] * hw[(i - 2) & 3][xp][j] +
                                   p.k[2] * hw[(i - 1) & 3][xp][j] + p.k[3] * hw[i & 3][xp][j];
                        epilogue_store<T, N>(p, v, b, oy, ox0 + xp, c, dreg, breg, sreg, nw);
                    }
                }
            }
        }
    }
}

// thread = one 16-byte channel vector walking down a chunk of pixels: bias / next-style live in registers
template <typename T>
__global__ void __launch_bounds__(256) noise_act_kernel(const T *__restrict__ img, T *__restrict__ out_y,
                                                         T *__restrict__ out_ys, const float *__restrict__ noise,
                                                         int64_t noise_bstride, const float *__restrict__ noise_w,
                                                         const float *__restrict__ bias, const float *__restrict__ s_next,
                                                         int64_t pixels, int C, int64_t chunk) {
    constexpr int N = Vec<T>::N;
    const int cv = C / N;
    const int lanes = blockDim.x / cv;
    const int vec = threadIdx.x % cv, lane = threadIdx.x / cv;
    if (lane >= lanes) return;
    const int b = blockIdx.y, c = vec * N;
    const float nw = (noise && noise_w) ? *noise_w : 0.f;
    float breg[N], sreg[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        breg[j] = bias ? bias[c + j] : 0.f;
        sreg[j] = out_ys ? s_next[(int64_t)b * C + c + j] : 1.f;
    }
    const int64_t p0 = (int64_t)blockIdx.x * chunk, p1 = min(p0 + chunk, pixels);
#pragma unroll 4
    for (int64_t pix = p0 + lane; pix < p1; pix += lanes) {
        const int64_t off = ((int64_t)b * pixels + pix) * C + c;
        Vec<T> x = load_vec<T>(img + off);
        const float nz = noise ? nw * __ldg(noise + b * noise_bstride + pix) : 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) x.v[j] = lrelu_sqrt2(x.v[j] + nz + breg[j]);
        if (out_y) store_vec<T>(out_y + off, x);
        if (out_ys) {
#pragma unroll
            for (int j = 0; j < N; ++j) x.v[j] *= sreg[j];
            store_vec<T>(out_ys + off, x);
        }
    }
}

int blur_act_tma(const ood_blur_act_args *a, cudaStream_t st, int *handled);    // blur_tma.cu
int blur_act_rows(const ood_blur_act_args *a, cudaStream_t st, int *handled);   // blur_rows.cu

}  // namespace ood

extern "C" int ood_blur_act(const ood_blur_act_args *a, void *stream) {
    using namespace ood;
    OOD_REQUIRE(a && a->in, "blur_act: null input");
    OOD_REQUIRE(a->batch > 0 && a->batch <= 65535 && a->ih >= 2 && a->iw >= 2, "blur_act: bad sizes");
    OOD_REQUIRE(a->dtype == OOD_F32 || a->dtype == OOD_BF16, "blur_act: bad dtype");
    OOD_REQUIRE(a->out_img || a->out_y || a->out_ys, "blur_act: no output requested");
    OOD_REQUIRE(a->act || (!a->out_y && !a->out_ys), "blur_act: out_y/out_ys need act=1");
    OOD_REQUIRE(!a->out_ys || a->s_next, "blur_act: out_ys needs s_next");
    const int N = a->dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(a->channels % N == 0, "blur_act: channels (%d) must be a multiple of %d", a->channels, N);
    {   // TMA-staged paths (channels % 32 == 0 on sm_100): row-streaming strips for images >= one strip wide, [16 x 32]
        // tiles below that; otherwise the register-window kernel below
        int handled = 0;
        int rc = blur_act_rows(a, (cudaStream_t)stream, &handled);
        if (handled) return rc;
        rc = blur_act_tma(a, (cudaStream_t)stream, &handled);
        if (handled) return rc;
    }
    BlurParams p;
    p.in = a->in; p.out_img = a->out_img; p.out_y = a->out_y; p.out_ys = a->out_ys;
    p.d = a->d; p.noise = a->noise; p.noise_w = a->noise_w; p.bias = a->bias; p.s_next = a->s_next;
    p.noise_bstride = a->noise_bstride;
    for (int i = 0; i < 4; ++i) p.k[i] = a->taps[3 - i];   // correlation with the flipped FIR (upfirdn2d.py:179)
    p.batch = a->batch; p.ih = a->ih; p.iw = a->iw; p.pad0 = a->pad0 > 0 ? a->pad0 : 1;
    const int pad1 = a->pad0 > 0 ? a->pad1 : 1;
    p.oh = a->ih + p.pad0 + pad1 - 3; p.ow = a->iw + p.pad0 + pad1 - 3; p.C = a->channels;
    p.act = a->act;
    constexpr int XPT = 2;
    const int cv = p.C / N;
    const int64_t threads = (int64_t)((p.ow + XPT - 1) / XPT) * cv;
    dim3 grid(ceil_div(threads, 128), ceil_div(p.oh, kStrip), p.batch);
    OOD_REQUIRE(grid.y <= 65535, "blur_act: image too tall");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->dtype == OOD_F32) blur_act_kernel<float, float, XPT><<<grid, 128, 0, st>>>(p);
    else if (a->in_f32) blur_act_kernel<__nv_bfloat16, float, XPT><<<grid, 128, 0, st>>>(p);
    else blur_act_kernel<__nv_bfloat16, __nv_bfloat16, XPT><<<grid, 128, 0, st>>>(p);
    return check_launch("blur_act");
}

extern "C" int ood_noise_act(const void *img, void *out_y, void *out_ys, const float *noise, int64_t noise_bstride,
                             const float *noise_w, const float *bias, const float *s_next, int batch, int64_t pixels,
                             int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(img && (out_y || out_ys) && batch > 0 && batch <= 65535 && pixels > 0, "noise_act: bad arguments");
    OOD_REQUIRE(!out_ys || s_next, "noise_act: out_ys needs s_next");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "noise_act: bad dtype");
    OOD_REQUIRE(channels % N == 0, "noise_act: channels (%d) must be a multiple of %d", channels, N);
    OOD_REQUIRE(channels / N <= 256, "noise_act: too many channels (%d)", channels);
    const int lanes = std::max(1, 256 / (channels / N));
    const int64_t want = std::max<int64_t>(1, (int64_t)kNumSMs * pixwalk_blocks_per_sm(8) / batch);
    int64_t chunk = std::max<int64_t>((pixels + want - 1) / want, (int64_t)lanes * 4);
    chunk = (chunk + lanes - 1) / lanes * lanes;
    dim3 grid((unsigned)((pixels + chunk - 1) / chunk), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32)
        noise_act_kernel<float><<<grid, 256, 0, st>>>((const float *)img, (float *)out_y, (float *)out_ys, noise, noise_bstride,
                                                      noise_w, bias, s_next, pixels, channels, chunk);
    else
        noise_act_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)img, (__nv_bfloat16 *)out_y,
                                                              (__nv_bfloat16 *)out_ys, noise, noise_bstride, noise_w, bias,
                                                              s_next, pixels, channels, chunk);
    return check_launch("noise_act");
}
