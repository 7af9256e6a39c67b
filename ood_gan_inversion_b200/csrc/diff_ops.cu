// Elementwise building blocks of the differentiable alignment path: the backward of
// InstanceNorm2d, PReLU and of the AlignNet's 2C -> 3 head, on NHWC activations in the pipeline's storage type.
// Reference: autograd through src/ops/SAMM/helpers.py:85-109,149-179 and e4e/encoders/helpers.py:426-448 (bottleneck_IR).
//
//   ood_nhwc_affine2: out[b,p,c] = A[b,c]*x1[b,p,c] + B[b,c]*x2[b,p,c] + C[b,c] on channel SLICES of wider tensors (pitch / offset
//                     per operand).  One kernel covers: the InstanceNorm backward's combine step
//                         g_x = (r*gamma) * g - (r*gamma*r*m2) * x + (r*gamma*(r*m2*mu - m1))
//                     (m1 = mean(g), m2 = mean(g * xhat), from ood_in_stats / ood_dot_reduce), residual sums and differences,
//                     channel concatenation and its adjoint (the AlignNet input cat[IN(cur) - IN(enc), IN(enc)]).
//   ood_prelu / ood_prelu_bwd: y = x > 0 ? x : slope[c]*x;  g_x = g * (x > 0 ? 1 : slope[c]).
//   ood_tap_gather:   adjoint of ood_tap_sum: G[b,y,x,3t+k] = g[b,k,y-dy_t,x-dx_t] (zero outside), channels 27..cp-1 zero: the
//                     gradient of the nine shifted partial sums w.r.t. the per-pixel projections, as the operand of a 1x1
//                     convolution with the transposed projection weights (data gradient of the 2C -> 3 3x3 convolution).
#include "common.cuh"

namespace ood {

template <typename T>
__global__ void __launch_bounds__(256) affine2_kernel(const T *__restrict__ x1, int p1, int o1, const T *__restrict__ x2, int p2, int o2,
                                                       const float *__restrict__ A, const float *__restrict__ B, const float *__restrict__ Cc,
                                                       T *__restrict__ out, int po, int oo, int64_t P, int C, int64_t nvec) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const int cv = C / N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * N;
        const int64_t pix = (int64_t)b * P + i / cv;
        Vec<T> a = load_vec<T>(x1 + pix * p1 + o1 + c);
        Vec<T> o;
        if (x2) {
            const Vec<T> bb = load_vec<T>(x2 + pix * p2 + o2 + c);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const float ca = A ? A[(int64_t)b * C + c + j] : 1.f, cb = B ? B[(int64_t)b * C + c + j] : 1.f;
                o.v[j] = fmaf(ca, a.v[j], fmaf(cb, bb.v[j], Cc ? Cc[(int64_t)b * C + c + j] : 0.f));
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) o.v[j] = fmaf(A ? A[(int64_t)b * C + c + j] : 1.f, a.v[j], Cc ? Cc[(int64_t)b * C + c + j] : 0.f);
        }
        store_vec<T>(out + pix * po + oo + c, o);
    }
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(256) prelu_kernel(const T *__restrict__ x, const T *__restrict__ g, const float *__restrict__ slope,
                                                     T *__restrict__ out, int C, int64_t nvec) {
    constexpr int N = Vec<T>::N;
    const int cv = C / N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * N;
        const Vec<T> xv = load_vec<T>(x + i * N);
        Vec<T> o;
        if (BWD) {
            const Vec<T> gv = load_vec<T>(g + i * N);
#pragma unroll
            for (int j = 0; j < N; ++j) o.v[j] = xv.v[j] > 0.f ? gv.v[j] : gv.v[j] * slope[c + j];
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) o.v[j] = xv.v[j] > 0.f ? xv.v[j] : xv.v[j] * slope[c + j];
        }
        store_vec<T>(out + i * N, o);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) tap_gather_kernel(const float *__restrict__ g, T *__restrict__ out, int H, int W, int CP) {
    const int b = blockIdx.y;
    const int64_t P = (int64_t)H * W;
    for (int64_t pix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pix < P; pix += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(pix / W), x = (int)(pix - (int64_t)y * W);
        T *o = out + ((int64_t)b * P + pix) * CP;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);       // res[yy,xx] read proj[y,x] through tap t
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
#pragma unroll
            for (int k = 0; k < 3; ++k) v[3 * t + k] = __ldg(g + ((int64_t)b * 3 + k) * P + (int64_t)yy * W + xx);
        }
        constexpr int N = Vec<T>::N;
        for (int q = 0; q < CP / N; ++q) {
            Vec<T> t;
#pragma unroll
            for (int j = 0; j < N; ++j) t.v[j] = q * N + j < 32 ? v[q * N + j] : 0.f;
            store_vec<T>(o + q * N, t);
        }
    }
}

template <typename T> static bool aligned16(const void *p) { return ((uintptr_t)p % 16) == 0; }

}  // namespace ood

extern "C" int ood_nhwc_affine2(const void *x1, int pitch1, int off1, const void *x2, int pitch2, int off2, const float *a, const float *b,
                                const float *c, void *out, int pitch_out, int off_out, int batch, int64_t pixels, int channels, int dtype,
                                void *stream) {
    using namespace ood;
    OOD_REQUIRE(x1 && out && batch > 0 && batch <= 65535 && pixels > 0 && channels > 0, "nhwc_affine2: bad arguments");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "nhwc_affine2: bad dtype");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(channels % N == 0 && pitch1 % N == 0 && off1 % N == 0 && pitch_out % N == 0 && off_out % N == 0 && (!x2 || (pitch2 % N == 0 && off2 % N == 0)),
                "nhwc_affine2: channels, pitches and offsets must be multiples of %d", N);
    OOD_REQUIRE(off1 + channels <= pitch1 && off_out + channels <= pitch_out && (!x2 || off2 + channels <= pitch2), "nhwc_affine2: slice outside its tensor");
    OOD_REQUIRE(((uintptr_t)x1 % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)x2 % 16) == 0, "nhwc_affine2: tensors must be 16-byte aligned");
    const int64_t nvec = pixels * (channels / N);
    dim3 grid((unsigned)std::min<int64_t>((nvec + 255) / 256, kNumSMs * 16), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32)
        affine2_kernel<float><<<grid, 256, 0, st>>>((const float *)x1, pitch1, off1, (const float *)x2, pitch2, off2, a, b, c, (float *)out, pitch_out, off_out,
                                                    pixels, channels, nvec);
    else
        affine2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x1, pitch1, off1, (const __nv_bfloat16 *)x2, pitch2, off2, a, b, c,
                                                            (__nv_bfloat16 *)out, pitch_out, off_out, pixels, channels, nvec);
    return check_launch("nhwc_affine2");
}

extern "C" int ood_prelu(const void *x, const void *g, const float *slope, void *out, int64_t pixels_total, int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(x && slope && out && pixels_total > 0 && channels > 0, "prelu: bad arguments");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "prelu: bad dtype");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(channels % N == 0, "prelu: channels (%d) must be a multiple of %d", channels, N);
    const int64_t nvec = pixels_total * (channels / N);
    const unsigned grid = (unsigned)std::min<int64_t>((nvec + 255) / 256, kNumSMs * 16);
    cudaStream_t st = (cudaStream_t)stream;
#define OOD_PRELU(T, BWD) prelu_kernel<T, BWD><<<grid, 256, 0, st>>>((const T *)x, (const T *)g, slope, (T *)out, channels, nvec)
    if (dtype == OOD_F32) { if (g) OOD_PRELU(float, true); else OOD_PRELU(float, false); }
    else { if (g) OOD_PRELU(__nv_bfloat16, true); else OOD_PRELU(__nv_bfloat16, false); }
#undef OOD_PRELU
    return check_launch("prelu");
}

extern "C" int ood_tap_gather(const float *g, void *out, int batch, int h, int w, int cp, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(g && out && batch > 0 && batch <= 65535 && h > 0 && w > 0 && cp >= 32, "tap_gather: bad arguments");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "tap_gather: bad dtype");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(cp % N == 0, "tap_gather: cp (%d) must be a multiple of %d", cp, N);
    const int64_t P = (int64_t)h * w;
    dim3 grid((unsigned)std::min<int64_t>((P + 255) / 256, kNumSMs * 8), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) tap_gather_kernel<float><<<grid, 256, 0, st>>>(g, (float *)out, h, w, cp);
    else tap_gather_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g, (__nv_bfloat16 *)out, h, w, cp);
    return check_launch("tap_gather");
}
