// ToRGB: 1x1 modulated convolution to 3 channels (no demodulation) + bias + FIR-upsampled skip accumulate.
// Reference: src/ops/StyleGAN/model.py:353-372 (ToRGB), :30-47 (Upsample), upfirdn2d up=2 pad (2,1).
// HBM-bound (reads C channels per pixel to produce 3): a group of G lanes owns one pixel, each lane a 16-byte
// channel vector, so a warp reads one contiguous 512-byte run per load; RGB partials are shuffle-reduced and the
// group leader adds bias and the 2x2-tap upsampled skip before one fp32 store per colour plane.
#include "common.cuh"

namespace ood {

template <typename T>
__global__ void __launch_bounds__(256) torgb_kernel(const T *__restrict__ y, const float *__restrict__ wrgb,
                                                     const float *__restrict__ bias, const float *__restrict__ skip,
                                                     float *__restrict__ out, int H, int W, int C, int G, float kf0,
                                                     float kf1, float kf2, float kf3) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;                       // lane within the pixel group
    const int groups_per_block = blockDim.x / G;
    const int64_t P = (int64_t)H * W;
    const float kf[4] = {kf0, kf1, kf2, kf3};
    const float *wb = wrgb + (int64_t)b * 3 * C;
    // block-uniform trip count: every lane takes part in the shuffles, out-of-range groups contribute nothing
    for (int64_t base = (int64_t)blockIdx.x * groups_per_block; base < P; base += (int64_t)gridDim.x * groups_per_block) {
        const int64_t pix = base + threadIdx.x / G;
        const bool valid = pix < P;
        const T *src = y + ((int64_t)b * P + (valid ? pix : 0)) * C;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        for (int c = gl * N; valid && c < C; c += G * N) {
            const Vec<T> x = load_vec<T>(src + c);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                r0 = fmaf(x.v[j], __ldg(wb + c + j), r0);
                r1 = fmaf(x.v[j], __ldg(wb + C + c + j), r1);
                r2 = fmaf(x.v[j], __ldg(wb + 2 * C + c + j), r2);
            }
        }
        for (int o = G >> 1; o > 0; o >>= 1) {
            r0 += __shfl_xor_sync(0xffffffffu, r0, o);
            r1 += __shfl_xor_sync(0xffffffffu, r1, o);
            r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        }
        if (valid && gl == 0) {
            const int Y = (int)(pix / W), X = (int)(pix - (int64_t)Y * W);
            float rgb[3] = {r0 + bias[0], r1 + bias[1], r2 + bias[2]};
            if (skip) {
                const int h2 = H >> 1, w2 = W >> 1;
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const int u = Y + ky - 2;
                    if (u < 0 || (u & 1) || (u >> 1) >= h2) continue;
#pragma unroll
                    for (int kx = 0; kx < 4; ++kx) {
                        const int v = X + kx - 2;
                        if (v < 0 || (v & 1) || (v >> 1) >= w2) continue;
                        const float wgt = kf[ky] * kf[kx];
                        const float *sp = skip + ((int64_t)b * 3 * h2 + (u >> 1)) * w2 + (v >> 1);
#pragma unroll
                        for (int k = 0; k < 3; ++k) rgb[k] = fmaf(wgt, __ldg(sp + (int64_t)k * h2 * w2), rgb[k]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) out[((int64_t)b * 3 + k) * P + pix] = rgb[k];
        }
    }
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
    int G = 1;
    while (G < 32 && G * 2 * N <= channels) G *= 2;        // power of two, <= 32, G*N <= C
    OOD_REQUIRE(channels % (G * N) == 0, "torgb: channels (%d) must be a multiple of %d", channels, G * N);
    float kf[4] = {0, 0, 0, 0};
    if (skip) for (int i = 0; i < 4; ++i) kf[i] = taps_up_host[3 - i];
    const int64_t P = (int64_t)h * w;
    const int gpb = 256 / G;
    dim3 grid((unsigned)std::min<int64_t>((P + gpb - 1) / gpb, kNumSMs * 16), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32)
        torgb_kernel<float><<<grid, 256, 0, st>>>((const float *)y, wrgb, bias, skip, out, h, w, channels, G, kf[0], kf[1], kf[2], kf[3]);
    else
        torgb_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)y, wrgb, bias, skip, out, h, w, channels, G, kf[0], kf[1], kf[2], kf[3]);
    return check_launch("torgb");
}
