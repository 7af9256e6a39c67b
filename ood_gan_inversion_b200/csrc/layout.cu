// Layout changes between the module API (NCHW fp32) and the internal NHWC storage type, with the
// style modulation of the consuming convolution folded in (x*s is what the implicit GEMM reads).
#include "common.cuh"

namespace ood {

// in [B][C][P] fp32 (batch stride given) -> out [B][P][C] T, times scale[b][c].
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float *__restrict__ in, int64_t in_bstride,
                                                            const float *__restrict__ scale, T *__restrict__ out, int C,
                                                            int64_t P) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float *src = in + b * in_bstride;
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t p = p0 + tx;
        float v = 0.f;
        if (c < C && p < P) {
            v = src[(int64_t)c * P + p];
            if (scale) v *= scale[(int64_t)b * C + c];
        }
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int64_t p = p0 + j;
        const int c = c0 + tx;
        if (c < C && p < P) out[((int64_t)b * P + p) * C + c] = from_f32<T>(tile[tx][j]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T *__restrict__ in, float *__restrict__ out, int C,
                                                            int64_t P) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int64_t p = p0 + j;
        const int c = c0 + tx;
        tile[j][tx] = (c < C && p < P) ? to_f32(in[((int64_t)b * P + p) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t p = p0 + tx;
        if (c < C && p < P) out[((int64_t)b * C + c) * P + p] = tile[tx][j];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_scale_kernel(const T *__restrict__ in, const float *__restrict__ scale,
                                                          T *__restrict__ out, int C, int64_t P, int64_t nvec_per_b) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const int cv = C / N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec_per_b; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * N;
        const int64_t off = (int64_t)b * P * C + i * N;
        Vec<T> x = load_vec<T>(in + off);
#pragma unroll
        for (int j = 0; j < N; ++j) x.v[j] *= scale[(int64_t)b * C + c + j];
        store_vec<T>(out + off, x);
    }
}

}  // namespace ood

extern "C" int ood_nchw_to_nhwc(const float *in, int64_t in_batch_stride, const float *scale_bc, void *out, int batch,
                                int channels, int h, int w, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(in && out && batch > 0 && channels > 0 && h > 0 && w > 0, "nchw_to_nhwc: bad arguments");
    OOD_REQUIRE(batch <= 65535, "nchw_to_nhwc: batch too large");
    const int64_t P = (int64_t)h * w;
    dim3 grid(ceil_div(P, 32), ceil_div(channels, 32), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) nchw_to_nhwc_kernel<float><<<grid, 256, 0, st>>>(in, in_batch_stride, scale_bc, (float *)out, channels, P);
    else if (dtype == OOD_BF16) nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(in, in_batch_stride, scale_bc, (__nv_bfloat16 *)out, channels, P);
    else OOD_REQUIRE(false, "nchw_to_nhwc: bad dtype");
    return check_launch("nchw_to_nhwc");
}

extern "C" int ood_nhwc_to_nchw(const void *in, float *out, int batch, int channels, int h, int w, int dtype,
                                void *stream) {
    using namespace ood;
    OOD_REQUIRE(in && out && batch > 0 && channels > 0 && h > 0 && w > 0, "nhwc_to_nchw: bad arguments");
    OOD_REQUIRE(batch <= 65535, "nhwc_to_nchw: batch too large");
    const int64_t P = (int64_t)h * w;
    dim3 grid(ceil_div(P, 32), ceil_div(channels, 32), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, st>>>((const float *)in, out, channels, P);
    else if (dtype == OOD_BF16) nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)in, out, channels, P);
    else OOD_REQUIRE(false, "nhwc_to_nchw: bad dtype");
    return check_launch("nhwc_to_nchw");
}

extern "C" int ood_nhwc_scale(const void *in, const float *scale_bc, void *out, int batch, int channels, int64_t pixels,
                              int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(in && out && scale_bc && batch > 0 && channels > 0 && pixels > 0, "nhwc_scale: bad arguments");
    OOD_REQUIRE(batch <= 65535, "nhwc_scale: batch too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(channels % N == 0, "nhwc_scale: channels (%d) must be a multiple of %d", channels, N);
    const int64_t nvec = pixels * channels / N;
    dim3 grid((unsigned)std::min<int64_t>((nvec + 255) / 256, kNumSMs * 8), batch);
    if (dtype == OOD_F32) nhwc_scale_kernel<float><<<grid, 256, 0, st>>>((const float *)in, scale_bc, (float *)out, channels, pixels, nvec);
    else if (dtype == OOD_BF16) nhwc_scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)in, scale_bc, (__nv_bfloat16 *)out, channels, pixels, nvec);
    else OOD_REQUIRE(false, "nhwc_scale: bad dtype");
    return check_launch("nhwc_scale");
}
