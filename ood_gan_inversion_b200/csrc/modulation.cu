// Style modulation s = EqualLinear(latent), demodulation coefficients d, and weight packing.
// Reference: src/ops/StyleGAN/model.py:129-163 (EqualLinear), :236-241 (modulate / demodulate).
// The reference materialises B x Co x Ci x k x k modulated weights per call; here the weights stay
// shared and only s[B,Ci] and d[B,Co] are computed:  d = scale * rsqrt(scale^2 * sum_i s_i^2 * Wsq[o,i] + eps).
#include "common.cuh"

namespace ood {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per (b, i): grid (ceil(cin/8), batch)
__global__ void __launch_bounds__(256) style_s_kernel(const float *__restrict__ latent, int64_t lstride,
                                                       const float *__restrict__ mod_w, const float *__restrict__ mod_b,
                                                       float *__restrict__ s, int batch, int D, int cin, float lin_scale) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (i >= cin) return;
    const float *wrow = mod_w + (int64_t)i * D;
    const float *l = latent + b * lstride;
    float acc = 0.f;
    for (int j = lane; j < D; j += 32) acc = fmaf(__ldg(l + j), __ldg(wrow + j), acc);
    acc = warp_sum(acc);
    if (lane == 0) s[(int64_t)b * cin + i] = acc * lin_scale + (mod_b ? mod_b[i] : 0.f);
}

// one warp per (b, o)
__global__ void __launch_bounds__(256) demod_kernel(const float *__restrict__ s, const float *__restrict__ wsq,
                                                     float *__restrict__ d, int batch, int cin, int cout, float cs) {
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= (int64_t)batch * cout) return;
    const int b = (int)(wid / cout), o = (int)(wid % cout);
    const float *sr = s + (int64_t)b * cin;
    const float *wr = wsq + (int64_t)o * cin;
    float acc = 0.f;
    for (int i = lane; i < cin; i += 32) acc = fmaf(sr[i] * sr[i], wr[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) d[wid] = cs * rsqrtf(cs * cs * acc + 1e-8f);
}

__global__ void fill_kernel(float *p, int64_t n, float v) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void sumsq_kernel(const float *__restrict__ w, float *__restrict__ wsq, int64_t n, int taps) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int k = 0; k < taps; ++k) { const float v = w[i * taps + k]; acc = fmaf(v, v, acc); }
    wsq[i] = acc;
}

template <typename T>
__global__ void pack_weight_kernel(const float *__restrict__ w, T *__restrict__ out, int cout, int cin, int taps,
                                   int ci_major) {
    const int64_t n = (int64_t)cout * cin * taps;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // index into out
    if (i >= n) return;
    int t, co, ci;
    if (ci_major) { co = (int)(i % cout); ci = (int)((i / cout) % cin); t = (int)(i / ((int64_t)cout * cin)); }
    else          { ci = (int)(i % cin); co = (int)((i / cin) % cout); t = (int)(i / ((int64_t)cout * cin)); }
    out[i] = from_f32<T>(w[((int64_t)co * cin + ci) * taps + t]);
}

__global__ void torgb_weight_kernel(const float *__restrict__ w, const float *__restrict__ s, float *__restrict__ o,
                                    int batch, int C, float scale) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)batch * 3 * C) return;
    const int c = (int)(i % C), k = (int)((i / C) % 3), b = (int)(i / (3 * (int64_t)C));
    o[i] = w[(int64_t)k * C + c] * s[(int64_t)b * C + c] * scale;
}

}  // namespace ood

extern "C" int ood_modulation(const float *latent, int64_t latent_stride, const float *mod_w, const float *mod_b,
                              const float *wsq, float conv_scale, float *s_out, float *d_out, int batch, int style_dim,
                              int cin, int cout, void *stream) {
    using namespace ood;
    OOD_REQUIRE(latent && mod_w && s_out && batch > 0 && style_dim > 0 && cin > 0, "modulation: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    OOD_REQUIRE(batch <= 65535, "modulation: batch too large");
    style_s_kernel<<<dim3(ceil_div(cin, 8), batch), 256, 0, st>>>(latent, latent_stride, mod_w, mod_b, s_out, batch, style_dim, cin,
                                                      1.0f / sqrtf((float)style_dim));
    if (d_out) {
        OOD_REQUIRE(cout > 0, "modulation: cout");
        const int64_t n = (int64_t)batch * cout;
        if (wsq) demod_kernel<<<ceil_div(n, 8), 256, 0, st>>>(s_out, wsq, d_out, batch, cin, cout, conv_scale);
        else fill_kernel<<<ceil_div(n, 256), 256, 0, st>>>(d_out, n, conv_scale);
    }
    return check_launch("modulation", d_out ? 2 : 1);
}

extern "C" int ood_weight_sumsq(const float *w, float *wsq, int cout, int cin, int taps, void *stream) {
    using namespace ood;
    OOD_REQUIRE(w && wsq && cout > 0 && cin > 0 && taps > 0, "weight_sumsq: bad arguments");
    const int64_t n = (int64_t)cout * cin;
    sumsq_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w, wsq, n, taps);
    return check_launch("weight_sumsq");
}

extern "C" int ood_pack_conv_weight(const float *w, void *out, int cout, int cin, int taps, int ci_major_out, int dtype,
                                    void *stream) {
    using namespace ood;
    OOD_REQUIRE(w && out && cout > 0 && cin > 0 && taps > 0, "pack_conv_weight: bad arguments");
    const int64_t n = (int64_t)cout * cin * taps;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) pack_weight_kernel<float><<<ceil_div(n, 256), 256, 0, st>>>(w, (float *)out, cout, cin, taps, ci_major_out);
    else if (dtype == OOD_BF16) pack_weight_kernel<__nv_bfloat16><<<ceil_div(n, 256), 256, 0, st>>>(w, (__nv_bfloat16 *)out, cout, cin, taps, ci_major_out);
    else if (dtype == OOD_F16) pack_weight_kernel<__half><<<ceil_div(n, 256), 256, 0, st>>>(w, (__half *)out, cout, cin, taps, ci_major_out);
    else OOD_REQUIRE(false, "pack_conv_weight: bad dtype");
    return check_launch("pack_conv_weight");
}

extern "C" int ood_torgb_weight(const float *w, const float *s, float *wrgb, float scale, int batch, int channels,
                                void *stream) {
    using namespace ood;
    OOD_REQUIRE(w && s && wrgb && batch > 0 && channels > 0, "torgb_weight: bad arguments");
    const int64_t n = (int64_t)batch * 3 * channels;
    torgb_weight_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w, s, wrgb, batch, channels, scale);
    return check_launch("torgb_weight");
}
