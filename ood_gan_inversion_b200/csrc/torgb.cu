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
    return dtype == OOD_F32 ? launch_torgb<float>(y, wrgb, bias, skip, out, kf, batch, h, w, channels, st)
                            : launch_torgb<__nv_bfloat16>(y, wrgb, bias, skip, out, kf, batch, h, w, channels, st);
}
