// Shared helpers for libood_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>

#include "../../include/ood_b200.h"

namespace ood {

void set_error(const char *fmt, ...);
int check_launch(const char *what, int kernels = 1);
// blocks per SM for the pixel-walk elementwise kernels (a thread owns a channel vector and walks a chunk of pixels): few long
// blocks -- one resident wave -- beat many short ones, whose per-thread coefficient set-up and partial last wave dominate on
// small tensors (alignnet.cu: pix_grid).  OOD_PW_BPS overrides for experiments.
static inline int pixwalk_blocks_per_sm(int dflt) {
    static int v = -1;
    if (v < 0) { const char *e = getenv("OOD_PW_BPS"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
    return v > 0 ? v : dflt;
}

// One-time per-DEVICE set-up (cudaFuncSetAttribute is per context): `static DeviceOnce once; if (once.first()) {...}`.
// A process that drives several GPUs (the reference's ops take the current device per call, upfirdn2d_kernel.cu:143-145)
// gets the dynamic-shared-memory opt-in on each of them.
struct DeviceOnce {
    bool done[64] = {};
    bool first() {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

#define OOD_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            ood::set_error(__VA_ARGS__);    \
            return OOD_ERR_ARG;             \
        }                                   \
    } while (0)

constexpr int kNumSMs = 148;
extern thread_local int g_conv_route;      // which kernel family the last ood_conv3x3 of this thread ran on (ood_last_conv_route)
constexpr float kSqrt2 = 1.4142135623730951f;

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- storage-type traits: T in {float, __nv_bfloat16}; arithmetic is fp32 ----
template <typename T> struct Vec;   // 16-byte vector of T
template <> struct Vec<float> {
    static constexpr int N = 4;
    float v[4];
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    float v[8];
};

template <> struct Vec<__half> {
    static constexpr int N = 8;
    float v[8];
};

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__half x) { return __half2float(x); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
// fp16 storage (OOD_F16: the encoder, whose activations are normalised and O(1); 11 significant bits instead of bf16's 8).
// Conversions saturate to the largest finite half instead of producing inf.
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2 *>(&w)); }
template <> __device__ __forceinline__ __half from_f32<__half>(float x) {
    const uint32_t r = pack_f16x2(x, 0.f);
    return __ushort_as_half((unsigned short)(r & 0xffffu));
}

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100): two lanes per issue slot.  The bandwidth kernels are
// issue-bound long before they are FMA-pipe-bound (ncu: blur 27 instructions per element at 69 % issue utilisation), so
// halving the arithmetic instruction count is what moves them towards the HBM roofline. ----
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 max2(float2 a, float2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
// bf16x2 word -> (low element, high element)
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// 16-byte global load of T, unpacked to fp32 lanes.
template <typename T> __device__ __forceinline__ Vec<T> load_vec(const T *p);
template <> __device__ __forceinline__ Vec<float> load_vec<float>(const float *p) {
    float4 r = __ldg(reinterpret_cast<const float4 *>(p));
    Vec<float> o;
    o.v[0] = r.x; o.v[1] = r.y; o.v[2] = r.z; o.v[3] = r.w;
    return o;
}
template <> __device__ __forceinline__ Vec<__nv_bfloat16> load_vec<__nv_bfloat16>(const __nv_bfloat16 *p) {
    uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    Vec<__nv_bfloat16> o;
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o.v[2 * i] = __uint_as_float(w[i] << 16);
        o.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return o;
}
template <> __device__ __forceinline__ Vec<__half> load_vec<__half>(const __half *p) {
    uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    Vec<__half> o;
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = unpack_f16x2(w[i]);
        o.v[2 * i] = t.x;
        o.v[2 * i + 1] = t.y;
    }
    return o;
}
template <typename T> __device__ __forceinline__ void store_vec(T *p, const Vec<T> &x);
template <> __device__ __forceinline__ void store_vec<float>(float *p, const Vec<float> &x) {
    *reinterpret_cast<float4 *>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
// 32-byte global store (sm_100: STG.256).  A thread that owns a 64-byte run and writes it as four 16-byte stores sends four
// half-filled 32-byte sectors to L2 (ncu on the transposed convolution: 2.15 GB of L1->L2 writes for 1.08 GB of output); two
// 32-byte stores send whole sectors.  `p` must be 32-byte aligned.
__device__ __forceinline__ void st_global_256(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                              uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h)
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
template <> __device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16 *p, const Vec<__nv_bfloat16> &x) {
    uint4 r;
    r.x = pack_bf16x2(x.v[0], x.v[1]);
    r.y = pack_bf16x2(x.v[2], x.v[3]);
    r.z = pack_bf16x2(x.v[4], x.v[5]);
    r.w = pack_bf16x2(x.v[6], x.v[7]);
    *reinterpret_cast<uint4 *>(p) = r;
}

template <> __device__ __forceinline__ void store_vec<__half>(__half *p, const Vec<__half> &x) {
    uint4 r;
    r.x = pack_f16x2(x.v[0], x.v[1]);
    r.y = pack_f16x2(x.v[2], x.v[3]);
    r.z = pack_f16x2(x.v[4], x.v[5]);
    r.w = pack_f16x2(x.v[6], x.v[7]);
    *reinterpret_cast<uint4 *>(p) = r;
}

__device__ __forceinline__ float lrelu_sqrt2(float v) { return fmaxf(v, 0.2f * v) * kSqrt2; }

}  // namespace ood
