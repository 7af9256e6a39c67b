// ToRGB: 1x1 modulated convolution to 3 channels (no demodulation) + bias + FIR-upsampled skip accumulate.
// Reference: src/ops/StyleGAN/model.py:353-372 (ToRGB), :30-47 (Upsample), upfirdn2d up=2 pad (2,1).
// HBM-bound (reads C channels per pixel to produce 3): a group of G lanes owns one pixel, each lane a 16-byte
// channel vector (KCH of them when C > 32 vectors), so a warp reads contiguous 512-byte runs; the lane's slice of the
// per-sample RGB weights lives in registers for the whole kernel; RGB partials are butterfly-reduced and lanes 0..2 of
// the group each finish one colour plane (bias + 2x2-tap upsampled skip + one fp32 store).
#include "common.cuh"

namespace ood {

// KCH > 0: channel chunks per lane, weights register-resident.  KCH == 0: generic (weights re-read through L1).
template <typename T, int KCH>
__global__ void __launch_bounds__(256) torgb_kernel(const T *__restrict__ y, const float *__restrict__ wrgb,
                                                     const float *__restrict__ bias, const float *__restrict__ skip,
                                                     float *__restrict__ out, int H, int W, int C, int G, float kf0,
                                                     float kf1, float kf2, float kf3) {
    constexpr int N = Vec<T>::N;
    constexpr int KR = KCH > 0 ? KCH : 1;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;                       // lane within the pixel group
    const int groups_per_block = blockDim.x / G;
    const int64_t P = (int64_t)H * W;
    const float *wb = wrgb + (int64_t)b * 3 * C;
    float wreg[KR][3][N];
    if constexpr (KCH > 0) {
#pragma unroll
        for (int k = 0; k < KCH; ++k)
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int j = 0; j < N; ++j) wreg[k][q][j] = wb[q * C + (gl + k * G) * N + j];
    }
    const int h2 = H >> 1, w2 = W >> 1;
    constexpr int PPI = KCH > 0 ? (KCH == 1 ? 4 : 2) : 1;      // pixels per group per iteration (independent loads in flight)
    // block-uniform trip count: every lane takes part in the shuffles, out-of-range groups contribute nothing
    for (int64_t base = (int64_t)blockIdx.x * groups_per_block * PPI; base < P;
         base += (int64_t)gridDim.x * groups_per_block * PPI) {
        float r[PPI][3];
        int64_t pixs[PPI];
        if constexpr (KCH > 0) {
            Vec<T> x[PPI][KCH];
#pragma unroll
            for (int q = 0; q < PPI; ++q) {
                pixs[q] = base + (int64_t)q * groups_per_block + threadIdx.x / G;
                const T *src = y + ((int64_t)b * P + (pixs[q] < P ? pixs[q] : 0)) * C;
#pragma unroll
                for (int k = 0; k < KCH; ++k) x[q][k] = load_vec<T>(src + (gl + k * G) * N);
            }
#pragma unroll
            for (int q = 0; q < PPI; ++q) {
                r[q][0] = r[q][1] = r[q][2] = 0.f;
#pragma unroll
                for (int k = 0; k < KCH; ++k)
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        r[q][0] = fmaf(x[q][k].v[j], wreg[k][0][j], r[q][0]);
                        r[q][1] = fmaf(x[q][k].v[j], wreg[k][1][j], r[q][1]);
                        r[q][2] = fmaf(x[q][k].v[j], wreg[k][2][j], r[q][2]);
                    }
            }
        } else {
            pixs[0] = base + threadIdx.x / G;
            r[0][0] = r[0][1] = r[0][2] = 0.f;
            if (pixs[0] < P) {
                const T *src = y + ((int64_t)b * P + pixs[0]) * C;
                for (int c = gl * N; c < C; c += G * N) {
                    const Vec<T> x = load_vec<T>(src + c);
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        r[0][0] = fmaf(x.v[j], __ldg(wb + c + j), r[0][0]);
                        r[0][1] = fmaf(x.v[j], __ldg(wb + C + c + j), r[0][1]);
                        r[0][2] = fmaf(x.v[j], __ldg(wb + 2 * C + c + j), r[0][2]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PPI; ++q) {
            for (int o = G >> 1; o > 0; o >>= 1) {
                r[q][0] += __shfl_xor_sync(0xffffffffu, r[q][0], o);
                r[q][1] += __shfl_xor_sync(0xffffffffu, r[q][1], o);
                r[q][2] += __shfl_xor_sync(0xffffffffu, r[q][2], o);
            }
        }
#pragma unroll
        for (int q = 0; q < PPI; ++q) {
            const int64_t pix = pixs[q];
            if (pix >= P) continue;
            const int Y = (int)(pix / W), X = (int)(pix - (int64_t)Y * W);
            // colour k is finished by lane k of the group (all three by lane 0 when the group is smaller than 4 lanes)
            const int k_lo = G >= 4 ? gl : 0, k_hi = G >= 4 ? gl + 1 : 3;
            if (gl >= (G >= 4 ? 3 : 1)) continue;
            for (int k = k_lo; k < k_hi; ++k) {
                float v = (k == 0 ? r[q][0] : (k == 1 ? r[q][1] : r[q][2])) + bias[k];
                if (skip) {
                    // up-2 polyphase: output row Y sees skip rows (Y + ky0 - 2)/2 and the next one, taps ky0, ky0+2 (ky0 = Y & 1)
                    const float *sp = skip + ((int64_t)b * 3 + k) * h2 * w2;
                    const int ky0 = Y & 1, kx0 = X & 1;
                    const int ra = (Y + ky0 - 2) >> 1, ca = (X + kx0 - 2) >> 1;
                    const float wy[2] = {ky0 ? kf1 : kf0, ky0 ? kf3 : kf2}, wx[2] = {kx0 ? kf1 : kf0, kx0 ? kf3 : kf2};
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
                out[((int64_t)b * 3 + k) * P + pix] = v;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// bf16 path: the [pixels x C] x [C x 3] contraction on warp-level MMA (mma.sync m16n8k16, fp32 accumulate).  The SIMT
// kernel above spends ~3 FMAs per loaded element plus a butterfly reduction per pixel and reached 0.23 of HBM peak;
// here a warp owns 32 consecutive pixels, a lane's two 16-byte loads per 32-channel block ARE its A fragments (the K
// order inside a block is permuted so that lane t of a quad holds channels 8t..8t+7 -- the B fragments are permuted the
// same way), the RGB weights are split into bf16 hi + lo parts so no precision is lost against the fp32-weight FMA
// form, and after one shuffle exchange lane L finishes pixel L: bias + polyphase up-FIR of the skip + three fully
// coalesced 128-byte plane stores.  Tensor throughput is irrelevant here (N = 3); the point is ~1 instruction per
// loaded 16 bytes.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// hi / lo bf16 split of eight consecutive RGB weights -> B fragments of the two MMAs of a 32-channel block
__device__ __forceinline__ void torgb_bfrag(const float *w8, uint32_t *hi, uint32_t *lo) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = w8[2 * i], b = w8[2 * i + 1];
        const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
        hi[i] = pack_bf16x2(__bfloat162float(ah), __bfloat162float(bh));
        lo[i] = pack_bf16x2(a - __bfloat162float(ah), b - __bfloat162float(bh));
    }
}

// KCB > 0: 32-channel blocks per pixel, B fragments register-resident.  KCB == 0: any C % 32 == 0, B fragments staged in
// shared memory ([C/8][3][8] words).
template <int KCB>
__global__ void __launch_bounds__(256) torgb_mma_kernel(const __nv_bfloat16 *__restrict__ y, const float *__restrict__ wrgb,
                                                         const float *__restrict__ bias, const float *__restrict__ skip,
                                                         float *__restrict__ out, int H, int W, int C, float kf0, float kf1,
                                                         float kf2, float kf3) {
    extern __shared__ uint32_t tg_wsm[];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t P = (int64_t)H * W;
    const float *wb = wrgb + (int64_t)b * 3 * C;
    const int ncb = C >> 5;
    constexpr int KR = KCB > 0 ? KCB : 1;
    uint32_t bh[KR][4], bl[KR][4];
    if constexpr (KCB > 0) {
#pragma unroll
        for (int cb = 0; cb < KCB; ++cb) {
            float w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = g < 3 ? wb[g * C + cb * 32 + t * 8 + j] : 0.f;
            torgb_bfrag(w8, bh[cb], bl[cb]);
        }
    } else {
        for (int i = threadIdx.x; i < (C >> 3) * 3; i += blockDim.x) {
            const int c8 = i / 3, n = i - c8 * 3;
            float w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = wb[n * C + c8 * 8 + j];
            uint32_t hi[4], lo[4];
            torgb_bfrag(w8, hi, lo);
            uint32_t *dst = tg_wsm + (size_t)i * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) { dst[j] = hi[j]; dst[4 + j] = lo[j]; }
        }
        __syncthreads();
    }
    const float bias_r = bias[0], bias_g = bias[1], bias_b = bias[2];
    const int h2 = H >> 1, w2 = W >> 1;
    const __nv_bfloat16 *yb = y + (int64_t)b * P * C + t * 8;

    for (int64_t p0 = ((int64_t)blockIdx.x * 8 + warp) * 32; p0 < P; p0 += (int64_t)gridDim.x * 8 * 32) {
        float acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        // rows of the two 16-pixel tiles: tile*16 + g and + 8
        const __nv_bfloat16 *rp[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t pix = p0 + (i >> 1) * 16 + (i & 1) * 8 + g;
            ok[i] = pix < P;
            rp[i] = yb + (ok[i] ? pix : 0) * C;
        }
        auto body = [&](int cb, const uint32_t *fh, const uint32_t *fl) {
            uint4 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = ok[i] ? __ldg(reinterpret_cast<const uint4 *>(rp[i] + cb * 32)) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int tl = 0; tl < 2; ++tl) {
                const uint4 &xa = x[2 * tl], &xb = x[2 * tl + 1];
                mma_bf16_16816(acc[tl], xa.x, xb.x, xa.y, xb.y, fh[0], fh[1]);
                mma_bf16_16816(acc[tl], xa.z, xb.z, xa.w, xb.w, fh[2], fh[3]);
                mma_bf16_16816(acc[tl], xa.x, xb.x, xa.y, xb.y, fl[0], fl[1]);
                mma_bf16_16816(acc[tl], xa.z, xb.z, xa.w, xb.w, fl[2], fl[3]);
            }
        };
        if constexpr (KCB > 0) {
#pragma unroll
            for (int cb = 0; cb < KCB; ++cb) body(cb, bh[cb], bl[cb]);
        } else {
#pragma unroll 2
            for (int cb = 0; cb < ncb; ++cb) {
                uint32_t fh[4] = {0, 0, 0, 0}, fl[4] = {0, 0, 0, 0};
                if (g < 3) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(tg_wsm + ((size_t)(cb * 4 + t) * 3 + g) * 8);
                    const uint4 h4 = src[0], l4 = src[1];
                    fh[0] = h4.x; fh[1] = h4.y; fh[2] = h4.z; fh[3] = h4.w;
                    fl[0] = l4.x; fl[1] = l4.y; fl[2] = l4.z; fl[3] = l4.w;
                }
                body(cb, fh, fl);
            }
        }
        // D fragment: lane (g, t) holds columns 2t, 2t+1 of rows g (d0, d1) and g + 8 (d2, d3).  Lane L finishes pixel
        // p0 + L: tile L / 16, row L % 16 -> source quad (L % 8), R and G in its lane 0, B in its lane 1
        const int src = (lane & 7) * 4;
        const bool hi_row = (lane >> 3) & 1, tile1 = lane >> 4;
        float rgb[3];
        {
            float v[2][3];
#pragma unroll
            for (int tl = 0; tl < 2; ++tl) {
                const float r0 = __shfl_sync(0xffffffffu, acc[tl][0], src), r2 = __shfl_sync(0xffffffffu, acc[tl][2], src);
                const float g0 = __shfl_sync(0xffffffffu, acc[tl][1], src), g2 = __shfl_sync(0xffffffffu, acc[tl][3], src);
                const float b0 = __shfl_sync(0xffffffffu, acc[tl][0], src + 1), b2 = __shfl_sync(0xffffffffu, acc[tl][2], src + 1);
                v[tl][0] = hi_row ? r2 : r0; v[tl][1] = hi_row ? g2 : g0; v[tl][2] = hi_row ? b2 : b0;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) rgb[k] = tile1 ? v[1][k] : v[0][k];
        }
        const int64_t pix = p0 + lane;
        if (pix >= P) continue;
        rgb[0] += bias_r; rgb[1] += bias_g; rgb[2] += bias_b;
        if (skip) {
            // up-2 polyphase: output row Y sees skip rows (Y + ky0 - 2)/2 and the next one, taps ky0, ky0+2 (ky0 = Y & 1).
            // 32-bit pixel arithmetic (P < 2^31 is checked by the host): the 64-bit division of the first version was a
            // third of the kernel's instructions (ncu: 381 warp-instructions per 32-pixel chunk, issue-bound).
            const int pi = (int)pix;
            const int Y = pi / W, X = pi - Y * W;
            const int ky0 = Y & 1, kx0 = X & 1;
            const int ra = (Y + ky0 - 2) >> 1, ca = (X + kx0 - 2) >> 1;
            const float wy[2] = {ky0 ? kf1 : kf0, ky0 ? kf3 : kf2}, wx[2] = {kx0 ? kf1 : kf0, kx0 ? kf3 : kf2};
            const int plane = h2 * w2;
            const float *sp = skip + (int64_t)b * 3 * plane;
            const bool cok[2] = {ca >= 0, ca + 1 < w2};
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int rr = ra + dy;
                if (rr < 0 || rr >= h2) continue;
                const float *srow = sp + rr * w2 + ca;
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    if (!cok[dx]) continue;
                    const float wgt = wy[dy] * wx[dx];
#pragma unroll
                    for (int k = 0; k < 3; ++k) rgb[k] = fmaf(wgt, __ldg(srow + dx + k * plane), rgb[k]);
                }
            }
        }
        float *op = out + (int64_t)b * 3 * P + pix;
#pragma unroll
        for (int k = 0; k < 3; ++k) op[(int64_t)k * P] = rgb[k];
    }
}

static int launch_torgb_mma(const void *y, const float *wrgb, const float *bias, const float *skip, float *out, const float *kf,
                            int batch, int h, int w, int C, cudaStream_t st) {
    const int64_t P = (int64_t)h * w;
    OOD_REQUIRE(P < (1LL << 31) / 4, "torgb: image too large");
    const int64_t chunks = (P + 255) / 256;              // 8 warps x 32 pixels per block iteration
    dim3 grid((unsigned)std::min<int64_t>(chunks, std::max(1, kNumSMs * 8 / batch)), batch);
    const __nv_bfloat16 *yy = (const __nv_bfloat16 *)y;
    if (C == 32) torgb_mma_kernel<1><<<grid, 256, 0, st>>>(yy, wrgb, bias, skip, out, h, w, C, kf[0], kf[1], kf[2], kf[3]);
    else if (C == 64) torgb_mma_kernel<2><<<grid, 256, 0, st>>>(yy, wrgb, bias, skip, out, h, w, C, kf[0], kf[1], kf[2], kf[3]);
    else {
        const size_t smem = (size_t)(C / 8) * 3 * 8 * sizeof(uint32_t);
        OOD_REQUIRE(smem <= 48 * 1024, "torgb: too many channels (%d)", C);
        torgb_mma_kernel<0><<<grid, 256, smem, st>>>(yy, wrgb, bias, skip, out, h, w, C, kf[0], kf[1], kf[2], kf[3]);
    }
    return check_launch("torgb");
}

template <typename T>
static int launch_torgb(const void *y, const float *wrgb, const float *bias, const float *skip, float *out, const float *kf,
                        int batch, int h, int w, int C, cudaStream_t st) {
    constexpr int N = Vec<T>::N;
    int G = 1;
    while (G < 32 && G * 2 * N <= C) G *= 2;        // power of two, <= 32, G*N <= C
    OOD_REQUIRE(C % (G * N) == 0, "torgb: channels (%d) must be a multiple of %d", C, G * N);
    const int kch = C / (G * N);
    const int64_t P = (int64_t)h * w;
    const int gpb = 256 / G;
    dim3 grid((unsigned)std::min<int64_t>((P + gpb - 1) / gpb, std::max(1, kNumSMs * 16 / batch)), batch);
#define OOD_TORGB(K) torgb_kernel<T, K><<<grid, 256, 0, st>>>((const T *)y, wrgb, bias, skip, out, h, w, C, G, kf[0], kf[1], kf[2], kf[3])
    if (kch == 1) OOD_TORGB(1);
    else if (kch == 2) OOD_TORGB(2);
    else if (kch == 4 && N == 8) OOD_TORGB(4);
    else OOD_TORGB(0);
#undef OOD_TORGB
    return check_launch("torgb");
}

}  // namespace ood

extern "C" int ood_torgb(const void *y, const float *wrgb, const float *bias, const float *skip, float *out,
                         const float *taps_up_host, int batch, int h, int w, int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(y && wrgb && bias && out && batch > 0 && batch <= 65535 && h > 0 && w > 0, "torgb: bad arguments");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "torgb: bad dtype");
    OOD_REQUIRE(!skip || (taps_up_host && h % 2 == 0 && w % 2 == 0), "torgb: skip needs taps and even size");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(channels % N == 0, "torgb: channels (%d) must be a multiple of %d", channels, N);
    float kf[4] = {0, 0, 0, 0};
    if (skip) for (int i = 0; i < 4; ++i) kf[i] = taps_up_host[3 - i];
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) return launch_torgb<float>(y, wrgb, bias, skip, out, kf, batch, h, w, channels, st);
    if (channels % 32 == 0 && channels <= 4096 && ((uintptr_t)y % 16) == 0)
        return launch_torgb_mma(y, wrgb, bias, skip, out, kf, batch, h, w, channels, st);
    return launch_torgb<__nv_bfloat16>(y, wrgb, bias, skip, out, kf, batch, h, w, channels, st);
}
