// Host I/O either side of the path: the byte formats of the reference's inference script.
//
//   in : cv2.imread(file) / 255.0 -> img2tensor(bgr2rgb=True) -> (t - 0.5) * 2      run_ood_faceGAN_inversion.py:158-159,
//        BasicSR/basicsr/utils/img_util.py:10-36:  uint8 BGR [H][W][3] -> fp32 RGB planes [3][H][W] in [-1, 1]
//   out: tensor2img(t, rgb2bgr=True, min_max=(-1, 1))                               run_ood_faceGAN_inversion.py:64-72,
//        img_util.py:38-94:  fp32 RGB planes -> clamp -> (t - min) / (max - min) -> (* 255.0).round() -> uint8 BGR [H][W][3]
//
// With these two kernels at the ends of the captured step, a batch crosses PCIe as 3 bytes per pixel each way instead of
// 12.  Byte work: results are bit-exact against the reference's arithmetic --
//   * the input value is float32(double(v) / 255.0) (numpy divides in float64, img2tensor then casts to float32): a
//     256-entry table built per block with IEEE double division, then the fp32 subtract and multiply as separate roundings;
//   * the output value is rint(((clamp(x) - min) / (max - min)) * 255.0f) with round-half-to-even (numpy.round), every
//     step a separately rounded fp32 operation (no FMA contraction).
// HBM-bound: 15 bytes per pixel.  A thread owns four consecutive pixels of one image: three 4-byte words of interleaved
// bytes on the uint8 side, one 16-byte vector per colour plane on the fp32 side.
#include <cmath>

#include "common.cuh"

namespace ood {

__device__ __forceinline__ float u8_unit(int v) { return (float)((double)v / 255.0); }

template <bool VEC>
__global__ void __launch_bounds__(256) img2tensor_u8_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, int P,
                                                             int swap, float sub, float mul) {
    __shared__ float lut[256];
    lut[threadIdx.x] = __fmul_rn(__fsub_rn(u8_unit(threadIdx.x), sub), mul);
    __syncthreads();
    const int b = blockIdx.y;
    const uint8_t *ib = in + (int64_t)b * P * 3;
    float *ob = out + (int64_t)b * P * 3;
    if constexpr (VEC) {            // P % 4 == 0, 4-byte aligned input, 16-byte aligned output
        const int nq = P / 4;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(ib) + (int64_t)q * 3;
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
            // bytes: p0c0 p0c1 p0c2 p1c0 | p1c1 p1c2 p2c0 p2c1 | p2c2 p3c0 p3c1 p3c2
            const float c0[4] = {lut[w0 & 255], lut[w0 >> 24], lut[(w1 >> 16) & 255], lut[(w2 >> 8) & 255]};
            const float c1[4] = {lut[(w0 >> 8) & 255], lut[w1 & 255], lut[w1 >> 24], lut[(w2 >> 16) & 255]};
            const float c2[4] = {lut[(w0 >> 16) & 255], lut[(w1 >> 8) & 255], lut[w2 & 255], lut[w2 >> 24]};
            float4 *o = reinterpret_cast<float4 *>(ob) + q;
            const int64_t plane = P / 4;
            o[swap ? 2 * plane : 0] = make_float4(c0[0], c0[1], c0[2], c0[3]);
            o[plane] = make_float4(c1[0], c1[1], c1[2], c1[3]);
            o[swap ? 0 : 2 * plane] = make_float4(c2[0], c2[1], c2[2], c2[3]);
        }
    } else {
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
            const uint8_t *px = ib + (int64_t)p * 3;
            ob[(swap ? 2 : 0) * (int64_t)P + p] = lut[px[0]];
            ob[(int64_t)P + p] = lut[px[1]];
            ob[(swap ? 0 : 2) * (int64_t)P + p] = lut[px[2]];
        }
    }
}

// `range` > 0: divide; `range` < 0: -range is the exact reciprocal of a power-of-two range (same result as the division)
__device__ __forceinline__ uint32_t quant_u8(float x, float lo, float hi, float range) {
    const float v = fminf(fmaxf(x, lo), hi);                                  // clamp_(min, max)
    const float d = __fsub_rn(v, lo);
    const float u = range < 0.f ? __fmul_rn(d, -range) : __fdiv_rn(d, range);  // (t - min) / (max - min)
    return (uint32_t)__float2int_rn(__fmul_rn(u, 255.0f)) & 255u;             // (img * 255.0).round().astype(uint8)
}

template <bool VEC>
__global__ void __launch_bounds__(256) tensor2img_u8_kernel(const float *__restrict__ in, uint8_t *__restrict__ out, int P,
                                                             int swap, float lo, float hi, float range) {
    const int b = blockIdx.y;
    const float *ib = in + (int64_t)b * P * 3;
    uint8_t *ob = out + (int64_t)b * P * 3;
    if constexpr (VEC) {
        const int nq = P / 4;
        const int64_t plane = P / 4;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
            const float4 *i4 = reinterpret_cast<const float4 *>(ib) + q;
            const float4 a = __ldg(i4 + (swap ? 2 * plane : 0)), g = __ldg(i4 + plane), c = __ldg(i4 + (swap ? 0 : 2 * plane));
            const float f0[4] = {a.x, a.y, a.z, a.w}, f1[4] = {g.x, g.y, g.z, g.w}, f2[4] = {c.x, c.y, c.z, c.w};
            uint32_t c0[4], c1[4], c2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c0[j] = quant_u8(f0[j], lo, hi, range);
                c1[j] = quant_u8(f1[j], lo, hi, range);
                c2[j] = quant_u8(f2[j], lo, hi, range);
            }
            uint32_t *w = reinterpret_cast<uint32_t *>(ob) + (int64_t)q * 3;
            w[0] = c0[0] | (c1[0] << 8) | (c2[0] << 16) | (c0[1] << 24);
            w[1] = c1[1] | (c2[1] << 8) | (c0[2] << 16) | (c1[2] << 24);
            w[2] = c2[2] | (c0[3] << 8) | (c1[3] << 16) | (c2[3] << 24);
        }
    } else {
        for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
            uint8_t *px = ob + (int64_t)p * 3;
            px[0] = (uint8_t)quant_u8(ib[(swap ? 2 : 0) * (int64_t)P + p], lo, hi, range);
            px[1] = (uint8_t)quant_u8(ib[(int64_t)P + p], lo, hi, range);
            px[2] = (uint8_t)quant_u8(ib[(swap ? 0 : 2) * (int64_t)P + p], lo, hi, range);
        }
    }
}

static inline bool aligned_to(const void *p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

}  // namespace ood

extern "C" int ood_img2tensor_u8(const uint8_t *in, float *out, int batch, int h, int w, int swap_rb, float sub, float mul,
                                 void *stream) {
    using namespace ood;
    OOD_REQUIRE(in && out && batch > 0 && batch <= 65535 && h > 0 && w > 0 && (int64_t)h * w < (1ll << 30),
                "img2tensor_u8: bad arguments");
    const int P = h * w;
    const bool vec = P % 4 == 0 && aligned_to(in, 4) && aligned_to(out, 16);
    const int units = vec ? P / 4 : P;
    // about eight blocks per SM over the batch: a block builds its 256-entry table once and then streams
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(units, 256), std::max(1, ceil_div(kNumSMs * 8, batch))), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) img2tensor_u8_kernel<true><<<grid, 256, 0, st>>>(in, out, P, swap_rb ? 1 : 0, sub, mul);
    else img2tensor_u8_kernel<false><<<grid, 256, 0, st>>>(in, out, P, swap_rb ? 1 : 0, sub, mul);
    return check_launch("img2tensor_u8");
}

extern "C" int ood_tensor2img_u8(const float *in, uint8_t *out, int batch, int h, int w, int swap_rb, float lo, float hi,
                                 void *stream) {
    using namespace ood;
    OOD_REQUIRE(in && out && batch > 0 && batch <= 65535 && h > 0 && w > 0 && (int64_t)h * w < (1ll << 30),
                "tensor2img_u8: bad arguments");
    OOD_REQUIRE(hi > lo, "tensor2img_u8: min_max must be increasing (%g, %g)", (double)lo, (double)hi);
    const int P = h * w;
    const bool vec = P % 4 == 0 && aligned_to(in, 16) && aligned_to(out, 4);
    const int units = vec ? P / 4 : P;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(units, 256), std::max(1, ceil_div(kNumSMs * 16, batch))), batch);
    cudaStream_t st = (cudaStream_t)stream;
    float range = (float)((double)hi - (double)lo);
    int ex = 0;
    if (std::frexp(range, &ex) == 0.5f && ex > -100 && ex < 100) range = -(1.0f / range);   // power of two: exact reciprocal
    if (vec) tensor2img_u8_kernel<true><<<grid, 256, 0, st>>>(in, out, P, swap_rb ? 1 : 0, lo, hi, range);
    else tensor2img_u8_kernel<false><<<grid, 256, 0, st>>>(in, out, P, swap_rb ? 1 : 0, lo, hi, range);
    return check_launch("tensor2img_u8");
}
