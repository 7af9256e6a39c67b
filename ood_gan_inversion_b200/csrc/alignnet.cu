// NHWC instance-norm kernels of the alignment network (SAMM AlignNet).
// Reference: src/ops/SAMM/helpers.py:85-109 (AlignNet.forward) + bottleneck_IR / BN('InstanceNorm')
// (src/ops/e4e/encoders/helpers.py:93-99,426-448).  In the reference every InstanceNorm2d is an ATen batch-norm call on a
// [1, B*C, H, W] reshape (a layout flip on channels-last data), followed by separate cat / sub / add / PReLU passes.
// Here: one statistics pass over (cur, enc) yields the moments of IN(cur), IN(enc) AND of z0 = cat[IN(cur)-IN(enc),
// IN(enc)] analytically; one elementwise pass builds the first conv's input; the residual add re-derives z0 instead of
// storing it.  All HBM-bound, 16-byte channel vectors, deterministic two-stage reductions (no float atomics).
#include "common.cuh"

namespace ood {

// pixels per partial-sum block: ONE resident wave (two CTAs per SM) of long streaming blocks.  The earlier 512-pixel
// chunks ran ~7 short waves whose start-up (first-load latency) and tail (shared-memory reduction, partial store) cost
// as much as the streaming itself (0.44 of HBM peak on a pure read).
static inline int stat_chunk(int64_t P, int batch) {
    const int64_t per_image = std::max<int64_t>(1, (int64_t)kNumSMs * 2 / batch);
    return (int)std::max<int64_t>((P + per_image - 1) / per_image, 32);
}

// 16 bytes of T -> N/2 fp32 pairs
template <typename T> __device__ __forceinline__ void load_pairs(const T *p, float2 *dst);
template <> __device__ __forceinline__ void load_pairs<float>(const float *p, float2 *dst) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
    dst[0] = make_float2(r.x, r.y); dst[1] = make_float2(r.z, r.w);
}
template <> __device__ __forceinline__ void load_pairs<__nv_bfloat16>(const __nv_bfloat16 *p, float2 *dst) {
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    dst[0] = unpack_bf16x2(r.x); dst[1] = unpack_bf16x2(r.y); dst[2] = unpack_bf16x2(r.z); dst[3] = unpack_bf16x2(r.w);
}
template <> __device__ __forceinline__ void load_pairs<__half>(const __half *p, float2 *dst) {
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    dst[0] = unpack_f16x2(r.x); dst[1] = unpack_f16x2(r.y); dst[2] = unpack_f16x2(r.z); dst[3] = unpack_f16x2(r.w);
}
template <typename T> __device__ __forceinline__ void store_pairs(T *p, const float2 *v);
template <> __device__ __forceinline__ void store_pairs<__half>(__half *p, const float2 *v) {
    uint4 r;
    r.x = pack_f16x2(v[0].x, v[0].y); r.y = pack_f16x2(v[1].x, v[1].y);
    r.z = pack_f16x2(v[2].x, v[2].y); r.w = pack_f16x2(v[3].x, v[3].y);
    *reinterpret_cast<uint4 *>(p) = r;
}
template <> __device__ __forceinline__ void store_pairs<float>(float *p, const float2 *v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
}
template <> __device__ __forceinline__ void store_pairs<__nv_bfloat16>(__nv_bfloat16 *p, const float2 *v) {
    uint4 r;
    r.x = pack_bf16x2(v[0].x, v[0].y); r.y = pack_bf16x2(v[1].x, v[1].y);
    r.z = pack_bf16x2(v[2].x, v[2].y); r.w = pack_bf16x2(v[3].x, v[3].y);
    *reinterpret_cast<uint4 *>(p) = r;
}

// values as they will read back from storage type T
template <typename T> __device__ __forceinline__ void round_pairs(float2 *v);
template <> __device__ __forceinline__ void round_pairs<float>(float2 *) {}
template <> __device__ __forceinline__ void round_pairs<__nv_bfloat16>(float2 *v) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = unpack_bf16x2(pack_bf16x2(v[j].x, v[j].y));
}

template <> __device__ __forceinline__ void round_pairs<__half>(float2 *v) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = unpack_f16x2(pack_f16x2(v[j].x, v[j].y));
}

// ---------------------------------------------------------------------------------------------- statistics
// partial[b][chunk][c][K]: K = 2 (sum x, sum x^2) or 5 (+ sum y, sum y^2, sum xy); packed fp32x2 accumulation
template <typename T, int K>
__global__ void __launch_bounds__(256, 2) in_partial_kernel(const T *__restrict__ x, const T *__restrict__ y,
                                                             float *__restrict__ partial, int64_t P, int C, int nchunks,
                                                             int chunk_px) {
    constexpr int N = Vec<T>::N, N2 = N / 2;
    constexpr int UNROLL = K == 5 ? 4 : 8;
    extern __shared__ float red[];                      // [lanes][cv*N*K]
    const int cv = C / N;
    const int lanes = 256 / cv > 0 ? 256 / cv : 1;      // pixel lanes per block (cv <= 256 enforced by the host)
    const int vec = threadIdx.x % cv, lane = threadIdx.x / cv;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int64_t p0 = (int64_t)chunk * chunk_px, p1 = min(p0 + chunk_px, P);
    float2 acc[N2][K];
#pragma unroll
    for (int j = 0; j < N2; ++j)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[j][k] = f2(0.f);
    if (lane < lanes) {
        const int64_t stride = (int64_t)lanes * C;
        const T *xp = x + ((int64_t)b * P + p0 + lane) * C + vec * N;
        const T *yp = K == 5 ? y + ((int64_t)b * P + p0 + lane) * C + vec * N : nullptr;
        int64_t n = p0 + lane < p1 ? (p1 - p0 - lane + lanes - 1) / lanes : 0;      // pixels of this thread
        for (; n >= UNROLL; n -= UNROLL) {
            float2 xv[UNROLL][N2], yv[K == 5 ? UNROLL : 1][N2];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                load_pairs<T>(xp + u * stride, xv[u]);
                if constexpr (K == 5) load_pairs<T>(yp + u * stride, yv[u]);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int j = 0; j < N2; ++j) {
                    acc[j][0] = add2(acc[j][0], xv[u][j]);
                    acc[j][1] = fma2(xv[u][j], xv[u][j], acc[j][1]);
                    if constexpr (K == 5) {
                        acc[j][2] = add2(acc[j][2], yv[u][j]);
                        acc[j][3] = fma2(yv[u][j], yv[u][j], acc[j][3]);
                        acc[j][4] = fma2(xv[u][j], yv[u][j], acc[j][4]);
                    }
                }
            xp += UNROLL * stride;
            if constexpr (K == 5) yp += UNROLL * stride;
        }
        for (; n > 0; --n) {
            float2 xv[N2], yv[N2];
            load_pairs<T>(xp, xv);
            if constexpr (K == 5) load_pairs<T>(yp, yv);
#pragma unroll
            for (int j = 0; j < N2; ++j) {
                acc[j][0] = add2(acc[j][0], xv[j]);
                acc[j][1] = fma2(xv[j], xv[j], acc[j][1]);
                if constexpr (K == 5) {
                    acc[j][2] = add2(acc[j][2], yv[j]);
                    acc[j][3] = fma2(yv[j], yv[j], acc[j][3]);
                    acc[j][4] = fma2(xv[j], yv[j], acc[j][4]);
                }
            }
            xp += stride;
            if constexpr (K == 5) yp += stride;
        }
        float *r = red + ((size_t)lane * cv + vec) * N * K;
#pragma unroll
        for (int j = 0; j < N2; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k) { r[(2 * j) * K + k] = acc[j][k].x; r[(2 * j + 1) * K + k] = acc[j][k].y; }
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
    // ONE resident wave (two 256-thread blocks per SM at 82-128 registers), like the statistics pass: a thread folds ~100 scalar
    // coefficients (statistics, affine weights) before its first pixel, so blocks should be few and long, and a second,
    // partial wave costs a whole block time.  The earlier fixed 8 blocks per SM ran 4-14 pixels per thread at 32 / 64 px:
    // 75 / 130 us for passes whose HBM time is 15 / 60 us (scripts/seed_bench.py; OOD_EW_WAVES overrides the wave count).
    static int waves = -1;
    if (waves < 0) { const char *e = getenv("OOD_EW_WAVES"); waves = e ? atoi(e) : 1; if (waves <= 0) waves = 1; }
    const int lanes = std::max(1, 256 / (C / N));
    const int64_t per_image = std::max<int64_t>(1, (int64_t)kNumSMs * 2 * waves / batch);
    chunk = std::max<int64_t>((P + per_image - 1) / per_image, (int64_t)lanes * 4);
    chunk = (chunk + lanes - 1) / lanes * lanes;
    grid = dim3((unsigned)((P + chunk - 1) / chunk), batch);
}

// mode 0 (front): lo[b,p,0:C] = (IN(cur)-IN(enc)) * rstd_d * w[c]   + bias[c]
//                 hi[b,p,0:C] =  IN(enc)          * rstd_e2 * w[C+c] + bias[C+c]            (conv input of block 0)
//                 lo = out, hi = out_hi, both with a pixel pitch of `pitch` channels: one [.,2C] tensor (out_hi = out + C,
//                 pitch 2C) or two [.,C] tensors; out_hi == nullptr skips the hi half (it does not depend on `cur`, so the
//                 second alignment cycle does not rebuild it)
// mode 1 (res0):  out = (t - mu_t) * rstd_t * w + bias + z0,  z0 = cat[IN(cur)-IN(enc), IN(enc)]   (block-0 output)
// mode 2:         mode 1 + per-block partial moments (sum, sum of squares) of `out` AS STORED (after the rounding to T):
//                 partial[b][block][2C][2], finished by in_finalize_kernel -- the next InstanceNorm's statistics without a
//                 pass over `out`
template <typename T, int MODE>
__global__ void __launch_bounds__(256) alignnet_ew_kernel(const T *__restrict__ cur, const T *__restrict__ enc,
                                                           const float *__restrict__ st6, const T *__restrict__ t,
                                                           const float *__restrict__ st2, const float *__restrict__ w,
                                                           const float *__restrict__ bias, T *__restrict__ out,
                                                           T *__restrict__ out_hi, int pitch, float *__restrict__ partial,
                                                           int64_t P, int C, int64_t chunk) {
    constexpr int N = Vec<T>::N;
    constexpr bool RES = MODE != 0;
    const int b = blockIdx.y;
    const PixSpan sp = pix_span<N>(C, P, chunk);
    if (MODE != 2 && !sp.active) return;
    if (MODE == 2 && !sp.active) { __syncthreads(); return; }       // cv divides 256 on this path (host check): never taken
    // lo = tl*ct_lo + cu*a1 + en*a2 + a3 ; hi = th*ct_hi + en*b1 + b2      (channel pairs: FFMA2)
    constexpr int N2 = N / 2;
    float2 a1[N2], a2[N2], a3[N2], b1[N2], b2[N2], ctl[N2], cth[N2];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const int c = sp.c + j;
        const float *s = st6 + ((int64_t)b * C + c) * 6;
        const float mc = s[0], rc = s[1], me = s[2], re = s[3];
        float va1, va2, va3, vb1, vb2, vctl = 0.f, vcth = 0.f;
        if constexpr (MODE == 0) {
            const float gl = s[4] * w[c], gh = s[5] * w[C + c];
            va1 = rc * gl; va2 = -re * gl; va3 = (me * re - mc * rc) * gl + bias[c];
            vb1 = re * gh; vb2 = -me * re * gh + bias[C + c];
        } else {
            const float *q0 = st2 + ((int64_t)b * 2 * C + c) * 2, *q1 = st2 + ((int64_t)b * 2 * C + C + c) * 2;
            vctl = q0[1] * w[c]; vcth = q1[1] * w[C + c];
            va1 = rc; va2 = -re; va3 = (me * re - mc * rc) - q0[0] * vctl + bias[c];
            vb1 = re; vb2 = -me * re - q1[0] * vcth + bias[C + c];
        }
        if (j & 1) { a1[j / 2].y = va1; a2[j / 2].y = va2; a3[j / 2].y = va3; b1[j / 2].y = vb1; b2[j / 2].y = vb2; ctl[j / 2].y = vctl; cth[j / 2].y = vcth; }
        else       { a1[j / 2].x = va1; a2[j / 2].x = va2; a3[j / 2].x = va3; b1[j / 2].x = vb1; b2[j / 2].x = vb2; ctl[j / 2].x = vctl; cth[j / 2].x = vcth; }
    }
    const int64_t s1 = (int64_t)sp.step * C, s2 = 2 * s1, so = (int64_t)sp.step * pitch;
    const T *cup = cur + ((int64_t)b * P + sp.p) * C + sp.c, *enp = enc + ((int64_t)b * P + sp.p) * C + sp.c;
    const T *tp = RES ? t + ((int64_t)b * P + sp.p) * 2 * C + sp.c : nullptr;
    T *op = out + ((int64_t)b * P + sp.p) * pitch + sp.c;
    T *oph = out_hi ? out_hi + ((int64_t)b * P + sp.p) * pitch + sp.c : nullptr;
    float2 m_lo[MODE == 2 ? N2 : 1][2], m_hi[MODE == 2 ? N2 : 1][2];
    if constexpr (MODE == 2) {
#pragma unroll
        for (int j = 0; j < N2; ++j) { m_lo[j][0] = m_lo[j][1] = m_hi[j][0] = m_hi[j][1] = f2(0.f); }
    }
#pragma unroll 2
    for (int64_t p = sp.p; p < sp.p_end; p += sp.step) {
        float2 cu[N2], en[N2], lo[N2], hi[N2];
        load_pairs<T>(cup, cu);
        load_pairs<T>(enp, en);
        if constexpr (RES) {
            float2 tl[N2], th[N2];
            load_pairs<T>(tp, tl);
            load_pairs<T>(tp + C, th);
#pragma unroll
            for (int j = 0; j < N2; ++j) {
                lo[j] = fma2(tl[j], ctl[j], fma2(cu[j], a1[j], fma2(en[j], a2[j], a3[j])));
                hi[j] = fma2(th[j], cth[j], fma2(en[j], b1[j], b2[j]));
            }
            tp += s2;
        } else {
#pragma unroll
            for (int j = 0; j < N2; ++j) {
                lo[j] = fma2(cu[j], a1[j], fma2(en[j], a2[j], a3[j]));
                hi[j] = fma2(en[j], b1[j], b2[j]);
            }
        }
        if constexpr (MODE == 2) {
            round_pairs<T>(lo);
            round_pairs<T>(hi);
#pragma unroll
            for (int j = 0; j < N2; ++j) {
                m_lo[j][0] = add2(m_lo[j][0], lo[j]); m_lo[j][1] = fma2(lo[j], lo[j], m_lo[j][1]);
                m_hi[j][0] = add2(m_hi[j][0], hi[j]); m_hi[j][1] = fma2(hi[j], hi[j], m_hi[j][1]);
            }
        }
        store_pairs<T>(op, lo);
        if (MODE != 0 || oph) store_pairs<T>(oph, hi);
        cup += s1; enp += s1; op += so;
        if (MODE != 0 || oph) oph += so;
    }
    if constexpr (MODE == 2) {
        // fixed-order reduction over the block's pixel lanes: red[lane][channel of 2C][2]
        __shared__ float red[256 * N * 4];
        const int lane = threadIdx.x / (C / N);
        float *r = red + (size_t)lane * 2 * C * 2;
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            float *q = r + (sp.c + 2 * j) * 2;
            q[0] = m_lo[j][0].x; q[1] = m_lo[j][1].x; q[2] = m_lo[j][0].y; q[3] = m_lo[j][1].y;
            q += 2 * C;
            q[0] = m_hi[j][0].x; q[1] = m_hi[j][1].x; q[2] = m_hi[j][0].y; q[3] = m_hi[j][1].y;
        }
        __syncthreads();
        const int lanes = 256 / (C / N);
        for (int i = threadIdx.x; i < 4 * C; i += 256) {
            float sum = 0.f;
            for (int l = 0; l < lanes; ++l) sum += red[(size_t)l * 4 * C + i];
            partial[((int64_t)b * gridDim.x + blockIdx.x) * 4 * C + i] = sum;
        }
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
    constexpr int N2 = N / 2;
    float2 g[N2], h[N2];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const float *q = st2 + ((int64_t)b * C + sp.c + j) * 2;
        const float gj = q[1] * (w ? w[sp.c + j] : 1.f);
        const float hj = (bias ? bias[sp.c + j] : 0.f) - q[0] * gj;
        if (j & 1) { g[j / 2].y = gj; h[j / 2].y = hj; } else { g[j / 2].x = gj; h[j / 2].x = hj; }
    }
    const int64_t s1 = (int64_t)sp.step * C;
    const T *xp = x + ((int64_t)b * P + sp.p) * C + sp.c;
    T *op = out + ((int64_t)b * P + sp.p) * C + sp.c;
#pragma unroll 4
    for (int64_t p = sp.p; p < sp.p_end; p += sp.step) {
        float2 v[N2];
        load_pairs<T>(xp, v);
#pragma unroll
        for (int j = 0; j < N2; ++j) v[j] = fma2(v[j], g[j], h[j]);
        store_pairs<T>(op, v);
        xp += s1; op += s1;
    }
}

// The same for C % 32 == 0 with the chunk loop spread over eight thread rows: a block owns (image, 32 channels), row r sums
// chunks r, r+8, ..., and the eight row sums are combined in a fixed order.  With the convolution's fused statistics there
// are up to 512 partials per channel (one per 128-pixel tile of a 256 px image); one thread walking them took ~35 us.
__global__ void __launch_bounds__(256) in_finalize_wide_kernel(const float *__restrict__ partial, float *__restrict__ st2, int64_t P,
                                                                int C, int nchunks, float eps) {
    __shared__ double s0[8][32], s1[8][32];
    const int cg = C / 32;
    const int b = blockIdx.x / cg, c = (blockIdx.x - b * cg) * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
    const float2 *p = reinterpret_cast<const float2 *>(partial) + ((int64_t)b * nchunks + r) * C + c;
    double a0 = 0, a1 = 0;
#pragma unroll 4
    for (int k = r; k < nchunks; k += 8, p += (int64_t)8 * C) {
        const float2 v = __ldg(p);
        a0 += v.x; a1 += v.y;
    }
    s0[r][threadIdx.x & 31] = a0; s1[r][threadIdx.x & 31] = a1;
    __syncthreads();
    if (r == 0) {
#pragma unroll
        for (int k = 1; k < 8; ++k) { a0 += s0[k][threadIdx.x]; a1 += s1[k][threadIdx.x]; }
        const double m = a0 / (double)P, v = fmax(a1 / (double)P - m * m, 0.0);
        *reinterpret_cast<float2 *>(st2 + ((int64_t)b * C + c) * 2) = make_float2((float)m, (float)(1.0 / sqrt(v + eps)));
    }
}

// finalize [B][nchunks][C][2] partial moments (shared with the convolution's fused statistics, conv_tc.cu)
void in_finalize_launch(const float *partial, float *st2, int64_t P, int C, int nchunks, float eps, int batch, cudaStream_t s) {
    if (C % 32 == 0 && nchunks > 8) {
        in_finalize_wide_kernel<<<batch * (C / 32), 256, 0, s>>>(partial, st2, P, C, nchunks, eps);
        return;
    }
    const int64_t total = (int64_t)batch * C;
    in_finalize_kernel<<<ceil_div(total, 256), 256, 0, s>>>(partial, st2, P, C, nchunks, eps, total);
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
        in_finalize_launch(partial, st, P, C, nchunks, eps, batch, s);
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
    if (dtype == OOD_F16) return launch_stats<__half>(x, y, workspace, stats, batch, pixels, channels, eps, s);
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
        alignnet_ew_kernel<float, 0><<<grid, 256, 0, s>>>((const float *)cur, (const float *)enc, st6, nullptr, nullptr, w, bias, (float *)out, (float *)out + channels, 2 * channels, nullptr, pixels, channels, chunk);
    else
        alignnet_ew_kernel<__nv_bfloat16, 0><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)cur, (const __nv_bfloat16 *)enc, st6, nullptr, nullptr, w, bias, (__nv_bfloat16 *)out, (__nv_bfloat16 *)out + channels, 2 * channels, nullptr, pixels, channels, chunk);
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
        alignnet_ew_kernel<float, 1><<<grid, 256, 0, s>>>((const float *)cur, (const float *)enc, st6, (const float *)t, st2, w, bias, (float *)out, (float *)out + channels, 2 * channels, nullptr, pixels, channels, chunk);
    else
        alignnet_ew_kernel<__nv_bfloat16, 1><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)cur, (const __nv_bfloat16 *)enc, st6, (const __nv_bfloat16 *)t, st2, w, bias, (__nv_bfloat16 *)out, (__nv_bfloat16 *)out + channels, 2 * channels, nullptr, pixels, channels, chunk);
    return check_launch("alignnet_res0");
}

extern "C" int ood_alignnet_front_split(const void *cur, const void *enc, const float *st6, const float *w, const float *bias,
                                        void *out_lo, void *out_hi, int batch, int64_t pixels, int channels, int dtype,
                                        void *stream) {
    using namespace ood;
    OOD_REQUIRE(cur && enc && st6 && w && bias && out_lo && batch > 0 && batch <= 65535 && pixels > 0, "alignnet_front_split: bad arguments");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "alignnet_front_split: bad dtype");
    OOD_REQUIRE(channels % N == 0, "alignnet_front_split: channels (%d) must be a multiple of %d", channels, N);
    OOD_REQUIRE(channels / N <= 256, "alignnet_front_split: too many channels (%d)", channels);
    dim3 grid; int64_t chunk;
    pix_grid(channels, N, pixels, batch, grid, chunk);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == OOD_F32)
        alignnet_ew_kernel<float, 0><<<grid, 256, 0, s>>>((const float *)cur, (const float *)enc, st6, nullptr, nullptr, w, bias, (float *)out_lo, (float *)out_hi, channels, nullptr, pixels, channels, chunk);
    else
        alignnet_ew_kernel<__nv_bfloat16, 0><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)cur, (const __nv_bfloat16 *)enc, st6, nullptr, nullptr, w, bias, (__nv_bfloat16 *)out_lo, (__nv_bfloat16 *)out_hi, channels, nullptr, pixels, channels, chunk);
    return check_launch("alignnet_front_split");
}

extern "C" int64_t ood_alignnet_res0_workspace(int batch, int64_t pixels, int channels, int dtype) {
    using namespace ood;
    dim3 grid; int64_t chunk;
    pix_grid(channels, dtype == OOD_F32 ? 4 : 8, pixels, batch > 0 ? batch : 1, grid, chunk);
    return (int64_t)batch * grid.x * 4 * channels * (int64_t)sizeof(float);
}

extern "C" int ood_alignnet_res0_stats(const void *t, const float *st2, const float *w, const float *bias, const void *cur,
                                       const void *enc, const float *st6, void *out, float *workspace, float *stats_out,
                                       float eps, int batch, int64_t pixels, int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(t && st2 && w && bias && cur && enc && st6 && out && workspace && stats_out && batch > 0 && batch <= 65535 && pixels > 0,
                "alignnet_res0_stats: bad arguments");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "alignnet_res0_stats: bad dtype");
    OOD_REQUIRE(channels % N == 0 && channels / N <= 256 && 256 % (channels / N) == 0,
                "alignnet_res0_stats: channels / %d (%d) must divide 256", N, channels / N);
    dim3 grid; int64_t chunk;
    pix_grid(channels, N, pixels, batch, grid, chunk);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == OOD_F32)
        alignnet_ew_kernel<float, 2><<<grid, 256, 0, s>>>((const float *)cur, (const float *)enc, st6, (const float *)t, st2, w, bias, (float *)out, (float *)out + channels, 2 * channels, workspace, pixels, channels, chunk);
    else
        alignnet_ew_kernel<__nv_bfloat16, 2><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)cur, (const __nv_bfloat16 *)enc, st6, (const __nv_bfloat16 *)t, st2, w, bias, (__nv_bfloat16 *)out, (__nv_bfloat16 *)out + channels, 2 * channels, workspace, pixels, channels, chunk);
    in_finalize_launch(workspace, stats_out, pixels, 2 * channels, (int)grid.x, eps, batch, s);
    return check_launch("alignnet_res0_stats", 2);
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

// Per-sample weights of the AlignNet's 2C -> 3 head as a 1x1 projection (SAMM/helpers.py:85-109; bottleneck_IR res_layer[0..1] of
// the second block, e4e/encoders/helpers.py:436-441): the affine InstanceNorm in front of the 3x3 convolution is folded into it,
//   W27 . (g*x + h) = (W27 diag(g_b)) . x + W27 . h_b,   g_b = rstd_b * in_w,  h_b = in_b - mean_b * g_b     (st = {mean, rstd}),
// so the normalised copy of x is never written.  Rows 27..29 optionally carry the bottleneck's 1x1 shortcut convolution on the
// UN-normalised x (w1, no g, no bias), rows 30..31 are zero.  One block per image; replaces seven ATen launches per cycle.
namespace ood {
template <typename T>
__global__ void __launch_bounds__(256) head_weights_kernel(const float *__restrict__ st, const float *__restrict__ in_w,
                                                            const float *__restrict__ in_b, const float *__restrict__ w27,
                                                            const float *__restrict__ w1, T *__restrict__ wps, float *__restrict__ bias, int C) {
    extern __shared__ float hw_s[];          // g[C], h[C]
    float *g = hw_s, *h = hw_s + C;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < C; c += blockDim.x) {
        const float mean = st[((int64_t)b * C + c) * 2], rstd = st[((int64_t)b * C + c) * 2 + 1];
        const float gg = rstd * in_w[c];
        g[c] = gg;
        h[c] = in_b[c] - mean * gg;
    }
    __syncthreads();
    // blockIdx.y owns rows [r0, r0 + rows): 8 row groups per image, so that 16 images fill 128 SMs instead of 16
    const int rows = 32 / gridDim.y, r0 = blockIdx.y * rows;
    for (int i = r0 * C + tid; i < (r0 + rows) * C; i += blockDim.x) {
        const int r = i / C, c = i - r * C;
        float v = w27[i] * g[c];
        if (w1 && r >= 27 && r < 30) v = w1[(r - 27) * C + c];
        wps[(int64_t)b * 32 * C + i] = from_f32<T>(v);
    }
    for (int r = r0 + warp; r < r0 + rows; r += blockDim.x / 32) {
        float sacc = 0.f;
        for (int c = lane; c < C; c += 32) sacc = fmaf(h[c], w27[r * C + c], sacc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if (lane == 0) bias[b * 32 + r] = sacc;
    }
}
}  // namespace ood

extern "C" int ood_alignnet_head_weights(const float *stats, const float *in_w, const float *in_b, const float *w27, const float *w1,
                                         void *wps, float *bias, int batch, int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(stats && in_w && in_b && w27 && wps && bias && batch > 0 && channels > 0, "alignnet_head_weights: bad arguments");
    const size_t smem = (size_t)channels * 2 * sizeof(float);
    OOD_REQUIRE(smem <= 48 * 1024, "alignnet_head_weights: too many channels (%d)", channels);
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 hgrid(batch, 8);
    if (dtype == OOD_BF16) head_weights_kernel<__nv_bfloat16><<<hgrid, 256, smem, s>>>(stats, in_w, in_b, w27, w1, (__nv_bfloat16 *)wps, bias, channels);
    else if (dtype == OOD_F32) head_weights_kernel<float><<<hgrid, 256, smem, s>>>(stats, in_w, in_b, w27, w1, (float *)wps, bias, channels);
    else OOD_REQUIRE(false, "alignnet_head_weights: bad dtype");
    return check_launch("alignnet_head_weights");
}
