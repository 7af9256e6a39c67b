// Fused bias + leaky-ReLU*scale and its gradient forms on contiguous [B][C][inner] data.
// Replaces fused.fused_bias_act (reference src/ops/op/fused_bias_act.cpp:11-21,
// fused_bias_act_kernel.cu:18-99) and the native branch src/ops/op/fused_act.py:96.  HBM-bound:
// 16-byte vector loads/stores, channel index derived per vector (inner % vec == 0) or per element.
#include "common.cuh"

namespace ood {

template <typename T, int GRAD, bool VEC>
__global__ void __launch_bounds__(256) bias_act_kernel(const T *__restrict__ in, const float *__restrict__ bias,
                                                        const T *__restrict__ refer, T *__restrict__ out, int64_t numel,
                                                        int channels, int64_t inner, float alpha, float scale) {
    constexpr int N = VEC ? Vec<T>::N : 1;
    const int64_t nvec = numel / N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * N;
        if constexpr (VEC) {
            Vec<T> x = load_vec<T>(in + e), r;
            if constexpr (GRAD == 1) r = load_vec<T>(refer + e);
            const float b = (GRAD == 0 && bias) ? bias[(e / inner) % channels] : 0.f;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if constexpr (GRAD == 0) {
                    const float v = x.v[j] + b;
                    x.v[j] = (v > 0.f ? v : v * alpha) * scale;
                } else {
                    x.v[j] = x.v[j] * (r.v[j] > 0.f ? 1.f : alpha) * scale;
                }
            }
            store_vec<T>(out + e, x);
        } else {
            float v = to_f32(in[e]);
            if constexpr (GRAD == 0) {
                v += bias ? bias[(e / inner) % channels] : 0.f;
                v = (v > 0.f ? v : v * alpha) * scale;
            } else {
                v = v * (to_f32(refer[e]) > 0.f ? 1.f : alpha) * scale;
            }
            out[e] = from_f32<T>(v);
        }
    }
}

// grad_bias[c] = sum_{b,inner} g[b][c][inner]: one block per channel slice, deterministic tree reduce.
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const T *__restrict__ g, float *__restrict__ gb, int64_t batch,
                                                         int channels, int64_t inner) {
    const int c = blockIdx.x;
    float acc = 0.f;
    for (int64_t b = 0; b < batch; ++b) {
        const T *row = g + (b * channels + c) * inner;
        for (int64_t i = threadIdx.x; i < inner; i += blockDim.x) acc += to_f32(row[i]);
    }
    __shared__ float red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) gb[c] = red[0];
}

template <typename T>
static int launch_bias_act(const void *in, const void *bias, const void *refer, void *out, int64_t numel, int channels,
                           int64_t inner, int grad, float alpha, float scale, cudaStream_t st) {
    constexpr int N = Vec<T>::N;
    const bool vec = (inner % N == 0) && (numel % N == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     (grad != 1 || (uintptr_t)refer % 16 == 0);
    const int64_t work = vec ? numel / N : numel;
    const int blocks = (int)std::min<int64_t>((work + 255) / 256, (int64_t)kNumSMs * 16);
    const T *i_ = (const T *)in, *r_ = (const T *)refer;
    const float *b_ = (const float *)bias;
    T *o_ = (T *)out;
    if (grad == 0) {
        if (vec) bias_act_kernel<T, 0, true><<<blocks, 256, 0, st>>>(i_, b_, r_, o_, numel, channels, inner, alpha, scale);
        else bias_act_kernel<T, 0, false><<<blocks, 256, 0, st>>>(i_, b_, r_, o_, numel, channels, inner, alpha, scale);
    } else {
        if (vec) bias_act_kernel<T, 1, true><<<blocks, 256, 0, st>>>(i_, b_, r_, o_, numel, channels, inner, alpha, scale);
        else bias_act_kernel<T, 1, false><<<blocks, 256, 0, st>>>(i_, b_, r_, o_, numel, channels, inner, alpha, scale);
    }
    return check_launch("fused_bias_act");
}

}  // namespace ood

extern "C" int ood_fused_bias_act(const void *in, const void *bias, const void *refer, void *out, int64_t numel,
                                  int channels, int64_t inner, int act, int grad, float alpha, float scale, int dtype,
                                  void *stream) {
    using namespace ood;
    OOD_REQUIRE(act == 3, "fused_bias_act: only act=3 (leaky relu) is implemented, got %d", act);
    OOD_REQUIRE(grad >= 0 && grad <= 2, "fused_bias_act: grad must be 0, 1 or 2");
    OOD_REQUIRE(in && out, "fused_bias_act: null pointer");
    OOD_REQUIRE(grad != 1 || refer, "fused_bias_act: grad=1 needs refer");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "fused_bias_act: bad dtype");
    OOD_REQUIRE(channels >= 1 && inner >= 1, "fused_bias_act: bad channels/inner");
    if (numel == 0) return OOD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (grad == 2) {  // second derivative of a piecewise-linear function
        cudaError_t e = cudaMemsetAsync(out, 0, numel * (dtype == OOD_F32 ? 4 : 2), st);
        if (e != cudaSuccess) { set_error("fused_bias_act: %s", cudaGetErrorString(e)); return OOD_ERR_CUDA; }
        return OOD_OK;
    }
    return dtype == OOD_F32 ? launch_bias_act<float>(in, bias, refer, out, numel, channels, inner, grad, alpha, scale, st)
                            : launch_bias_act<__nv_bfloat16>(in, bias, refer, out, numel, channels, inner, grad, alpha, scale, st);
}

extern "C" int ood_bias_grad(const void *g, float *grad_bias, int64_t batch, int channels, int64_t inner, int dtype,
                             void *stream) {
    using namespace ood;
    OOD_REQUIRE(g && grad_bias && channels >= 1, "bias_grad: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) bias_grad_kernel<float><<<channels, 256, 0, st>>>((const float *)g, grad_bias, batch, channels, inner);
    else bias_grad_kernel<__nv_bfloat16><<<channels, 256, 0, st>>>((const __nv_bfloat16 *)g, grad_bias, batch, channels, inner);
    return check_launch("bias_grad");
}
