// NHWC instance-norm kernels of the alignment network (SAMM AlignNet).
// Reference: src/ops/SAMM/helpers.py:85-109 (AlignNet.forward) + bottleneck_IR / BN('InstanceNorm')
// (src/ops/e4e/encoders/helpers.py:93-99,426-448).  In the reference every InstanceNorm2d is an ATen batch-norm call on a
// [1, B*C, H, W] reshape (a layout flip on channels-last data), followed by separate cat / sub / add / PReLU passes.
// Here: one statistics pass over (cur, enc) yields the moments of IN(cur), IN(enc) AND of z0 = cat[IN(cur)-IN(enc),
// IN(enc)] analytically; one elementwise pass builds the first conv's input; the residual add re-derives z0 instead of
// storing it.  All HBM-bound, 16-byte channel vectors, deterministic two-stage reductions (no float atomics).
#include "common.cuh"

namespace ood {

// pixels per partial-sum block: small tensors get small chunks so that the grid still fills the 148 SMs
static inline int stat_chunk(int64_t P, int batch) {
    int c = 512;
    while (c > 32 && (int64_t)batch * ((P + c - 1) / c) < 1024) c >>= 1;
    return c;
}

// ---------------------------------------------------------------------------------------------- statistics
// partial[b][chunk][c][K]: K = 2 (sum x, sum x^2) or 5 (+ sum y, sum y^2, sum xy)
template <typename T, int K>
__global__ void __launch_bounds__(256) in_partial_kernel(const T *__restrict__ x, const T *__restrict__ y,
                                                          float *__restrict__ partial, int64_t P, int C, int nchunks,
                                                          int chunk_px) {
    constexpr int N = Vec<T>::N;
    extern __shared__ float red[];                      // [lanes][cv*N*K]
    const int cv = C / N;
    const int lanes = 256 / cv > 0 ? 256 / cv : 1;      // pixel lanes per block (cv <= 256 enforced by the host)
    const int vec = threadIdx.x % cv, lane = threadIdx.x / cv;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int64_t p0 = (int64_t)chunk * chunk_px, p1 = min(p0 + chunk_px, P);
    float acc[N][K];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[j][k] = 0.f;
    if (lane < lanes) {
#pragma unroll 4
        for (int64_t p = p0 + lane; p < p1; p += lanes) {
            const int64_t off = ((int64_t)b * P + p) * C + vec * N;
            const Vec<T> xv = load_vec<T>(x + off);
            if constexpr (K == 5) {
                const Vec<T> yv = load_vec<T>(y + off);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    acc[j][0] += xv.v[j]; acc[j][1] = fmaf(xv.v[j], xv.v[j], acc[j][1]);
                    acc[j][2] += yv.v[j]; acc[j][3] = fmaf(yv.v[j], yv.v[j], acc[j][3]);
                    acc[j][4] = fmaf(xv.v[j], yv.v[j], acc[j][4]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < N; ++j) { acc[j][0] += xv.v[j]; acc[j][1] = fmaf(xv.v[j], xv.v[j], acc[j][1]); }
            }
        }
        float *r = red + ((size_t)lane * cv + vec) * N * K;
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k) r[j * K + k] = acc[j][k];
    }
    __syncthreads();
    // fixed-order reduction over the pixel lanes
    for (int i = threadIdx.x; i < C * K; i += 256) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[(size_t)l * C * K + i];
        partial[(((int64_t)b * nchunks + chunk) * C) * K + i] = s;
    }
}

// pair statistics -> st6[b][c] = {mu_x, rstd_x, mu_y, rstd_y, rstd_d, rstd_e2}
//   a = IN(x), e = IN(y):  var(a-e) = var(a) + var(e) - 2 cov(a,e), mean 0;  var(e) = s_y^2/(s_y^2+eps), mean 0
__global__ void in_finalize_pair_kernel(const float *__restrict__ partial, float *__restrict__ st6, int64_t P, int C,
                                        int nchunks, float eps, int64_t total) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // b*C + c
    if (i >= total) return;
    const int64_t b = i / C;
    const int c = (int)(i % C);
    double s[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < nchunks; ++k) {
        const float *p = partial + (((b * nchunks + k) * C) + c) * 5;
        for (int j = 0; j < 5; ++j) s[j] += p[j];
    }
    const double n = (double)P;
    const double mx = s[0] / n, my = s[2] / n;
    const double vx = fmax(s[1] / n - mx * mx, 0.0), vy = fmax(s[3] / n - my * my, 0.0);
    const double cxy = s[4] / n - mx * my;
    const double rx = 1.0 / sqrt(vx + eps), ry = 1.0 / sqrt(vy + eps);
    const double va = vx * rx * rx, ve = vy * ry * ry, cae = cxy * rx * ry;
    const double vd = fmax(va + ve - 2.0 * cae, 0.0);
    float *o = st6 + i * 6;
    o[0] = (float)mx; o[1] = (float)rx; o[2] = (float)my; o[3] = (float)ry;
    o[4] = (float)(1.0 / sqrt(vd + eps));
    o[5] = (float)(1.0 / sqrt(ve + eps));
}

// single statistics -> st2[b][c] = {mu, rstd}
__global__ void in_finalize_kernel(const float *__restrict__ partial, float *__restrict__ st2, int64_t P, int C, int nchunks,
                                   float eps, int64_t total) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t b = i / C;
    const int c = (int)(i % C);
    double s0 = 0, s1 = 0;
    for (int k = 0; k < nchunks; ++k) {
        const float *p = partial + (((b * nchunks + k) * C) + c) * 2;
        s0 += p[0]; s1 += p[1];
    }
    const double m = s0 / (double)P, v = fmax(s1 / (double)P - m * m, 0.0);
    st2[i * 2] = (float)m;
    st2[i * 2 + 1] = (float)(1.0 / sqrt(v + eps));
}

// ---------------------------------------------------------------------------------------------- elementwise passes
// Skeleton shared by the passes below: a thread owns ONE 16-byte channel vector and walks down a chunk of pixels, so the
// per-(b,c) coefficients are folded once into registers and the loop body is pure vector load / FMA / vector store.
struct PixSpan {
    int c;              // first channel of this thread's vector
    int64_t p, p_end;   // pixel range of this thread
    int step;
    bool active;
};
template <int N>
__device__ __forceinline__ PixSpan pix_span(int C, int64_t P, int64_t chunk) {
    const int cv = C / N;
    const int lanes = blockDim.x / cv;
    const int vec = threadIdx.x % cv, lane = threadIdx.x / cv;
    PixSpan s;
    s.c = vec * N;
    s.active = lane < lanes;
    const int64_t p0 = (int64_t)blockIdx.x * chunk;
    s.p = p0 + lane;
    s.p_end = min(p0 + chunk, P);
    s.step = lanes;
    return s;
}
static inline void pix_grid(int C, int N, int64_t P, int batch, dim3 &grid, int64_t &chunk) {
    const int lanes = std::max(1, 256 / (C / N));
    const int64_t want_blocks = std::max<int64_t>(1, (int64_t)kNumSMs * 8 / batch);
    chunk = std::max<int64_t>((P + want_blocks - 1) / want_blocks, (int64_t)lanes * 4);
    chunk = (chunk + lanes - 1) / lanes * lanes;
    grid = dim3((unsigned)((P + chunk - 1) / chunk), batch);
}

// mode 0 (front): out[b,p,0:C]  = (IN(cur)-IN(enc)) * rstd_d * w[c]   + bias[c]
//                 out[b,p,C:2C] =  IN(enc)          * rstd_e2 * w[C+c] + bias[C+c]          (conv input of block 0)
// mode 1 (res0):  out = (t - mu_t) * rstd_t * w + bias + z0,  z0 = cat[IN(cur)-IN(enc), IN(enc)]   (block-0 output)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) alignnet_ew_kernel(const T *__restrict__ cur, const T *__restrict__ enc,
                                                           const float *__restrict__ st6, const T *__restrict__ t,
                                                           const float *__restrict__ st2, const float *__restrict__ w,
                                                           const float *__restrict__ bias, T *__restrict__ out, int64_t P,
                                                           int C, int64_t chunk) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const PixSpan sp = pix_span<N>(C, P, chunk);
    if (!sp.active) return;
    // lo = tl*ct_lo + cu*a1 + en*a2 + a3 ; hi = th*ct_hi + en*b1 + b2
    float a1[N], a2[N], a3[N], b1[N], b2[N], ctl[N], cth[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const int c = sp.c + j;
        const float *s = st6 + ((int64_t)b * C + c) * 6;
        const float mc = s[0], rc = s[1], me = s[2], re = s[3];
        if constexpr (MODE == 0) {
            const float gl = s[4] * w[c], gh = s[5] * w[C + c];
            a1[j] = rc * gl; a2[j] = -re * gl; a3[j] = (me * re - mc * rc) * gl + bias[c];
            b1[j] = re * gh; b2[j] = -me * re * gh + bias[C + c];
            ctl[j] = cth[j] = 0.f;
        } else {
            const float *q0 = st2 + ((int64_t)b * 2 * C + c) * 2, *q1 = st2 + ((int64_t)b * 2 * C + C + c) * 2;
            ctl[j] = q0[1] * w[c]; cth[j] = q1[1] * w[C + c];
            a1[j] = rc; a2[j] = -re; a3[j] = (me * re - mc * rc) - q0[0] * ctl[j] + bias[c];
            b1[j] = re; b2[j] = -me * re - q1[0] * cth[j] + bias[C + c];
        }
    }
#pragma unroll 2
    for (int64_t p = sp.p; p < sp.p_end; p += sp.step) {
        const int64_t off1 = ((int64_t)b * P + p) * C + sp.c;
        const int64_t off2 = ((int64_t)b * P + p) * 2 * C + sp.c;
        const Vec<T> cu = load_vec<T>(cur + off1), en = load_vec<T>(enc + off1);
        Vec<T> lo, hi;
        if constexpr (MODE == 1) {
            const Vec<T> tl = load_vec<T>(t + off2), th = load_vec<T>(t + off2 + C);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                lo.v[j] = fmaf(tl.v[j], ctl[j], fmaf(cu.v[j], a1[j], fmaf(en.v[j], a2[j], a3[j])));
                hi.v[j] = fmaf(th.v[j], cth[j], fmaf(en.v[j], b1[j], b2[j]));
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                lo.v[j] = fmaf(cu.v[j], a1[j], fmaf(en.v[j], a2[j], a3[j]));
                hi.v[j] = fmaf(en.v[j], b1[j], b2[j]);
            }
        }
        store_vec<T>(out + off2, lo);
        store_vec<T>(out + off2 + C, hi);
    }
}

// y = (x - mu) * rstd * w + bias on NHWC
template <typename T>
__global__ void __launch_bounds__(256) in_apply_kernel(const T *__restrict__ x, const float *__restrict__ st2,
                                                        const float *__restrict__ w, const float *__restrict__ bias,
                                                        T *__restrict__ out, int64_t P, int C, int64_t chunk) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const PixSpan sp = pix_span<N>(C, P, chunk);
    if (!sp.active) return;
    float g[N], h[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const float *q = st2 + ((int64_t)b * C + sp.c + j) * 2;
        g[j] = q[1] * (w ? w[sp.c + j] : 1.f);
        h[j] = (bias ? bias[sp.c + j] : 0.f) - q[0] * g[j];
    }
#pragma unroll 4
    for (int64_t p = sp.p; p < sp.p_end; p += sp.step) {
        const int64_t off = ((int64_t)b * P + p) * C + sp.c;
        Vec<T> v = load_vec<T>(x + off);
#pragma unroll
        for (int j = 0; j < N; ++j) v.v[j] = fmaf(v.v[j], g[j], h[j]);
        store_vec<T>(out + off, v);
    }
}

template <typename T>
static int launch_stats(const void *x, const void *y, float *partial, float *st, int batch, int64_t P, int C, float eps,
                        cudaStream_t s) {
    constexpr int N = Vec<T>::N;
    OOD_REQUIRE(C % N == 0 && C / N <= 256, "in_stats: channels (%d) must be a multiple of %d and at most %d", C, N, 256 * N);
    const int chunk_px = stat_chunk(P, batch);
    const int nchunks = ceil_div(P, chunk_px);
    const int cv = C / N, lanes = 256 / cv;
    const int K = y ? 5 : 2;
    const size_t smem = (size_t)lanes * C * K * sizeof(float);
    OOD_REQUIRE(smem <= 160 * 1024, "in_stats: reduction buffer too large");
    dim3 grid(nchunks, batch);
    if (y) {
        auto kern = in_partial_kernel<T, 5>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, 256, smem, s>>>((const T *)x, (const T *)y, partial, P, C, nchunks, chunk_px);
        const int64_t total = (int64_t)batch * C;
        in_finalize_pair_kernel<<<ceil_div(total, 256), 256, 0, s>>>(partial, st, P, C, nchunks, eps, total);
    } else {
        auto kern = in_partial_kernel<T, 2>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, 256, smem, s>>>((const T *)x, nullptr, partial, P, C, nchunks, chunk_px);
        const int64_t total = (int64_t)batch * C;
        in_finalize_kernel<<<ceil_div(total, 256), 256, 0, s>>>(partial, st, P, C, nchunks, eps, total);
    }
    return check_launch("in_stats", 2);
}

}  // namespace ood

extern "C" int64_t ood_in_stats_workspace(int batch, int64_t pixels, int channels, int pair) {
    return (int64_t)batch * ood::ceil_div(pixels, ood::stat_chunk(pixels, batch)) * channels * (pair ? 5 : 2) * (int64_t)sizeof(float);
}

extern "C" int ood_in_stats(const void *x, const void *y, float *workspace, float *stats, int batch, int64_t pixels,
                            int channels, float eps, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(x && workspace && stats && batch > 0 && batch <= 65535 && pixels > 0 && channels > 0, "in_stats: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == OOD_F32) return launch_stats<float>(x, y, workspace, stats, batch, pixels, channels, eps, s);
    if (dtype == OOD_BF16) return launch_stats<__nv_bfloat16>(x, y, workspace, stats, batch, pixels, channels, eps, s);
    OOD_REQUIRE(false, "in_stats: bad dtype");
}

extern "C" int ood_alignnet_front(const void *cur, const void *enc, const float *st6, const float *w, const float *bias,
                                  void *out, int batch, int64_t pixels, int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(cur && enc && st6 && w && bias && out && batch > 0 && batch <= 65535 && pixels > 0, "alignnet_front: bad arguments");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "alignnet_front: bad dtype");
    OOD_REQUIRE(channels % N == 0, "alignnet_front: channels (%d) must be a multiple of %d", channels, N);
    OOD_REQUIRE(channels / N <= 256, "alignnet_front: too many channels (%d)", channels);
    dim3 grid; int64_t chunk;
    pix_grid(channels, N, pixels, batch, grid, chunk);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == OOD_F32)
        alignnet_ew_kernel<float, 0><<<grid, 256, 0, s>>>((const float *)cur, (const float *)enc, st6, nullptr, nullptr, w, bias, (float *)out, pixels, channels, chunk);
    else
        alignnet_ew_kernel<__nv_bfloat16, 0><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)cur, (const __nv_bfloat16 *)enc, st6, nullptr, nullptr, w, bias, (__nv_bfloat16 *)out, pixels, channels, chunk);
    return check_launch("alignnet_front");
}

extern "C" int ood_alignnet_res0(const void *t, const float *st2, const float *w, const float *bias, const void *cur,
                                 const void *enc, const float *st6, void *out, int batch, int64_t pixels, int channels,
                                 int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(t && st2 && w && bias && cur && enc && st6 && out && batch > 0 && batch <= 65535 && pixels > 0, "alignnet_res0: bad arguments");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "alignnet_res0: bad dtype");
    OOD_REQUIRE(channels % N == 0, "alignnet_res0: channels (%d) must be a multiple of %d", channels, N);
    OOD_REQUIRE(channels / N <= 256, "alignnet_res0: too many channels (%d)", channels);
    dim3 grid; int64_t chunk;
    pix_grid(channels, N, pixels, batch, grid, chunk);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == OOD_F32)
        alignnet_ew_kernel<float, 1><<<grid, 256, 0, s>>>((const float *)cur, (const float *)enc, st6, (const float *)t, st2, w, bias, (float *)out, pixels, channels, chunk);
    else
        alignnet_ew_kernel<__nv_bfloat16, 1><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)cur, (const __nv_bfloat16 *)enc, st6, (const __nv_bfloat16 *)t, st2, w, bias, (__nv_bfloat16 *)out, pixels, channels, chunk);
    return check_launch("alignnet_res0");
}

extern "C" int ood_in_apply(const void *x, const float *st2, const float *w, const float *bias, void *out, int batch,
                            int64_t pixels, int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(x && st2 && out && batch > 0 && batch <= 65535 && pixels > 0, "in_apply: bad arguments");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "in_apply: bad dtype");
    OOD_REQUIRE(channels % N == 0, "in_apply: channels (%d) must be a multiple of %d", channels, N);
    OOD_REQUIRE(channels / N <= 256, "in_apply: too many channels (%d)", channels);
    dim3 grid; int64_t chunk;
    pix_grid(channels, N, pixels, batch, grid, chunk);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == OOD_F32) in_apply_kernel<float><<<grid, 256, 0, s>>>((const float *)x, st2, w, bias, (float *)out, pixels, channels, chunk);
    else in_apply_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)x, st2, w, bias, (__nv_bfloat16 *)out, pixels, channels, chunk);
    return check_launch("in_apply");
}
