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

// Bilinear thumbnail of the network input for the encoder (OOD_faceGAN_e4e_arch.py:256: F.interpolate(x, (256, 256), mode='bilinear'),
// align_corners=False, no antialias) written directly as the encoder's first-convolution operand: NHWC storage type with the channel
// dimension zero-padded to Cp (the tcgen05 kernel's K granule).  Same arithmetic as ATen's upsample_bilinear2d: source index
// scale*(dst+0.5)-0.5 clamped at 0, val = h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11).  At 1024 -> 256 this is the mean of the 2x2 centre
// pixels of every 4x4 block.  One thread per output pixel; a warp writes 32 consecutive pixels (Cp*sizeof(T) bytes each).
template <typename T, int C, int CP>
__global__ void __launch_bounds__(256) thumbnail_nhwc_kernel(const float *__restrict__ in, T *__restrict__ out, int H, int W, int OH, int OW,
                                                              float sy, float sx) {
    const int b = blockIdx.y;
    const int64_t P = (int64_t)OH * OW;
    for (int64_t pix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pix < P; pix += (int64_t)gridDim.x * blockDim.x) {
        const int oy = (int)(pix / OW), ox = (int)(pix - (int64_t)oy * OW);
        const float fy = fmaxf(sy * (oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * (ox + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
        float v[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) v[c] = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float *pl = in + ((int64_t)b * C + c) * H * W;
            const float a = __ldg(pl + (int64_t)y0 * W + x0), bb = __ldg(pl + (int64_t)y0 * W + x1);
            const float cc = __ldg(pl + (int64_t)y1 * W + x0), dd = __ldg(pl + (int64_t)y1 * W + x1);
            v[c] = hy * (hx * a + lx * bb) + ly * (hx * cc + lx * dd);
        }
        T *o = out + ((int64_t)b * P + pix) * CP;
        constexpr int N = Vec<T>::N;
#pragma unroll
        for (int q = 0; q < CP / N; ++q) {
            Vec<T> t;
#pragma unroll
            for (int j = 0; j < N; ++j) t.v[j] = v[q * N + j];
            store_vec<T>(o + q * N, t);
        }
    }
}

}  // namespace ood

extern "C" int ood_thumbnail_nhwc(const float *in, void *out, int batch, int channels, int h, int w, int oh, int ow, int cp, int dtype,
                                  void *stream) {
    using namespace ood;
    OOD_REQUIRE(in && out && batch > 0 && batch <= 65535 && h > 0 && w > 0 && oh > 0 && ow > 0, "thumbnail_nhwc: bad arguments");
    OOD_REQUIRE(channels == 3 && cp == 32, "thumbnail_nhwc: built for 3 input channels padded to 32 (got %d -> %d)", channels, cp);
    OOD_REQUIRE((uintptr_t)out % 16 == 0, "thumbnail_nhwc: output must be 16-byte aligned");
    const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;
    const int64_t P = (int64_t)oh * ow;
    dim3 grid((unsigned)std::min<int64_t>((P + 255) / 256, kNumSMs * 8), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) thumbnail_nhwc_kernel<float, 3, 32><<<grid, 256, 0, st>>>(in, (float *)out, h, w, oh, ow, sy, sx);
    else if (dtype == OOD_BF16) thumbnail_nhwc_kernel<__nv_bfloat16, 3, 32><<<grid, 256, 0, st>>>(in, (__nv_bfloat16 *)out, h, w, oh, ow, sy, sx);
    else if (dtype == OOD_F16) thumbnail_nhwc_kernel<__half, 3, 32><<<grid, 256, 0, st>>>(in, (__half *)out, h, w, oh, ow, sy, sx);
    else OOD_REQUIRE(false, "thumbnail_nhwc: bad dtype");
    return check_launch("thumbnail_nhwc");
}

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
