// upfirdn2d on NCHW planes: zero-insert up, pad/crop, correlate with the flipped FIR, decimate.
// Replaces upfirdn2d_op.upfirdn2d (reference src/ops/op/upfirdn2d.cpp:12-23, upfirdn2d_kernel.cu:52-272)
// for the minor==1 form the Python wrapper issues; semantics follow upfirdn2d_native
// (src/ops/op/upfirdn2d.py:160-193).  HBM-bound: every input element is staged once per tile in shared
// memory (halo included), only taps that hit a real sample are visited, every parameter is generic.
#include "upfirdn_common.cuh"

namespace ood {

int plane_fir(const UpfirdnParams &p, int pad_x1, int pad_y1, cudaStream_t st, int *handled);   // plane_fir.cu

constexpr int kTOH = 32, kTOW = 64, kThreads = 256;

__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__device__ __forceinline__ int pos_mod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

// UP/DOWN/K > 0: compile-time specialisation (square, same on both axes); 0: runtime values.
template <typename T, int UP, int DOWN, int K>
__global__ void __launch_bounds__(kThreads) upfirdn2d_kernel(const UpfirdnParams p) {
    extern __shared__ float smem[];
    const int kh = K ? K : p.kh, kw = K ? K : p.kw;
    const int up_x = UP ? UP : p.up_x, up_y = UP ? UP : p.up_y;
    const int down_x = DOWN ? DOWN : p.down_x, down_y = DOWN ? DOWN : p.down_y;
    float *s_k = smem;                 // flipped taps [kh][kw]
    float *s_in = smem + kh * kw;      // [sih][siw]

    const int64_t tiles_per_plane = (int64_t)p.tiles_x * p.tiles_y;
    const int64_t plane = blockIdx.x / tiles_per_plane;
    const int t = (int)(blockIdx.x % tiles_per_plane);
    const int oy0 = (t / p.tiles_x) * kTOH, ox0 = (t % p.tiles_x) * kTOW;

    for (int i = threadIdx.x; i < kh * kw; i += kThreads) {
        const int ky = i / kw, kx = i % kw;
        s_k[i] = p.kernel[(kh - 1 - ky) * kw + (kw - 1 - kx)];
    }
    const int iy_lo = floor_div(oy0 * down_y - p.pad_y0 + up_y - 1, up_y);
    const int ix_lo = floor_div(ox0 * down_x - p.pad_x0 + up_x - 1, up_x);
    const T *src = reinterpret_cast<const T *>(p.in) + plane * (int64_t)p.in_h * p.in_w;
    // batches of 8 independent loads per thread keep ~8x more bytes in flight than a load/store loop
    const int n_in = p.sih * p.siw;
    for (int i0 = threadIdx.x; i0 < n_in; i0 += kThreads * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * kThreads;
            v[u] = 0.f;
            if (i < n_in) {
                const int r = i / p.siw, c = i - r * p.siw;
                const int iy = iy_lo + r, ix = ix_lo + c;
                if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v[u] = to_f32(src[(int64_t)iy * p.in_w + ix]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * kThreads;
            if (i < n_in) s_in[i] = v[u];
        }
    }
    __syncthreads();

    T *dst = reinterpret_cast<T *>(p.out) + plane * (int64_t)p.out_h * p.out_w;
    for (int i = threadIdx.x; i < kTOH * kTOW; i += kThreads) {
        const int oy = oy0 + i / kTOW, ox = ox0 + i % kTOW;
        if (oy >= p.out_h || ox >= p.out_w) continue;
        const int by = oy * down_y - p.pad_y0, bx = ox * down_x - p.pad_x0;
        const int ky0 = pos_mod(-by, up_y), kx0 = pos_mod(-bx, up_x);
        float acc = 0.f;
#pragma unroll
        for (int ky = ky0; ky < kh; ky += up_y) {
            const float *row = s_in + ((by + ky) / up_y - iy_lo) * p.siw - ix_lo;
            const float *krow = s_k + ky * kw;
#pragma unroll
            for (int kx = kx0; kx < kw; kx += up_x) acc = fmaf(row[(bx + kx) / up_x], krow[kx], acc);
        }
        dst[(int64_t)oy * p.out_w + ox] = from_f32<T>(acc);
    }
}

// up == 1 fast path (blur, down-sampling): the generic kernel above is issue-bound (ncu: 86 % issue slots busy, 144
// thread-instructions per output: two LDS + address arithmetic per tap).  Here the K x K taps live in registers and a
// thread produces MR vertically adjacent outputs from one ((MR-1)*DOWN + K) x K register window, so each input value
// is read from shared memory once per MR outputs (blur: 7 LDS + 16 FFMA per output instead of 32 LDS + 16 FFMA).
template <typename T, int DOWN, int K, int MR>
__global__ void __launch_bounds__(kThreads) upfirdn2d_up1_kernel(const UpfirdnParams p) {
    constexpr int SIH = (kTOH - 1) * DOWN + K, SIW = (kTOW - 1) * DOWN + K;
    constexpr int WR = (MR - 1) * DOWN + K;
    __shared__ float s_in[SIH * SIW];
    const int64_t tiles_per_plane = (int64_t)p.tiles_x * p.tiles_y;
    const int64_t plane = blockIdx.x / tiles_per_plane;
    const int t = (int)(blockIdx.x % tiles_per_plane);
    const int oy0 = (t / p.tiles_x) * kTOH, ox0 = (t % p.tiles_x) * kTOW;
    float w[K][K];
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) w[ky][kx] = __ldg(p.kernel + (K - 1 - ky) * K + (K - 1 - kx));   // flipped FIR
    const int iy_lo = oy0 * DOWN - p.pad_y0, ix_lo = ox0 * DOWN - p.pad_x0;
    const T *src = reinterpret_cast<const T *>(p.in) + plane * (int64_t)p.in_h * p.in_w;
    {   // tile load: (row, col) advance incrementally (no per-element division), 8 independent loads in flight per thread
        constexpr int DR = kThreads / SIW, DC = kThreads % SIW;
        int r = threadIdx.x / SIW, c = threadIdx.x % SIW;
        for (int i0 = threadIdx.x; i0 < SIH * SIW; i0 += kThreads * 8) {
            float v[8];
            int si[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int iy = iy_lo + r, ix = ix_lo + c;
                si[u] = (r < SIH) ? r * SIW + c : -1;
                v[u] = 0.f;
                if (r < SIH && iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v[u] = to_f32(src[(int64_t)iy * p.in_w + ix]);
                r += DR; c += DC;
                if (c >= SIW) { c -= SIW; ++r; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (si[u] >= 0) s_in[si[u]] = v[u];
        }
    }
    __syncthreads();
    T *dst = reinterpret_cast<T *>(p.out) + plane * (int64_t)p.out_h * p.out_w;
    for (int m = threadIdx.x; m < (kTOH / MR) * kTOW; m += kThreads) {
        const int tx = m % kTOW, tr = (m / kTOW) * MR;          // consecutive lanes -> consecutive columns
        float acc[MR];
#pragma unroll
        for (int j = 0; j < MR; ++j) acc[j] = 0.f;
        const float *base = s_in + (tr * DOWN) * SIW + tx * DOWN;
#pragma unroll
        for (int r = 0; r < WR; ++r) {
            float win[K];
#pragma unroll
            for (int kx = 0; kx < K; ++kx) win[kx] = base[r * SIW + kx];
#pragma unroll
            for (int j = 0; j < MR; ++j) {
                const int ky = r - j * DOWN;                   // compile-time after unrolling
                if (ky >= 0 && ky < K) {
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) acc[j] = fmaf(win[kx], w[ky][kx], acc[j]);
                }
            }
        }
        const int ox = ox0 + tx;
        if (ox < p.out_w) {
            T *o = dst + (int64_t)(oy0 + tr) * p.out_w + ox;
#pragma unroll
            for (int j = 0; j < MR; ++j)
                if (oy0 + tr + j < p.out_h) o[(int64_t)j * p.out_w] = from_f32<T>(acc[j]);
        }
    }
}

// up == 1, down == 1 (the blur after the transposed conv, the field blur): no shared memory at all.  A lane owns one
// input column and walks down a strip of rows; the K-1 right-hand neighbours come from warp shuffles, the last K input
// rows stay in registers, so per output there is 1 coalesced LDG, K-1 SHFL, K*K FFMA and 1 coalesced STG.
// (A warp produces 32-(K-1) output columns per row.)
template <typename T, int K>
__global__ void __launch_bounds__(kThreads) upfirdn2d_blur_kernel(const UpfirdnParams p) {
    constexpr int S = 32;                                  // output rows per strip
    constexpr int OW = 32 - (K - 1);                       // output columns per warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t plane = blockIdx.z;
    const int oy0 = blockIdx.y * S;
    const int ox = (blockIdx.x * (kThreads / 32) + warp) * OW + lane;      // output column of this lane (lanes >= OW: loaders only)
    const int ix = ox - p.pad_x0;
    float w[K][K];
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) w[ky][kx] = __ldg(p.kernel + (K - 1 - ky) * K + (K - 1 - kx));
    const T *src = reinterpret_cast<const T *>(p.in) + plane * (int64_t)p.in_h * p.in_w;
    T *dst = reinterpret_cast<T *>(p.out) + plane * (int64_t)p.out_h * p.out_w;
    const bool col_ok = ix >= 0 && ix < p.in_w;
    const bool out_ok = lane < OW && ox < p.out_w;
    float win[K][K];
#pragma unroll
    for (int i = 0; i < S + K - 1; ++i) {
        const int iy = oy0 - p.pad_y0 + i;
        float v = 0.f;
        if (col_ok && iy >= 0 && iy < p.in_h) v = to_f32(src[(int64_t)iy * p.in_w + ix]);
        win[i % K][0] = v;
#pragma unroll
        for (int kx = 1; kx < K; ++kx) win[i % K][kx] = __shfl_down_sync(0xffffffffu, v, kx);
        if (i >= K - 1) {
            const int oy = oy0 + i - (K - 1);
            float acc = 0.f;
#pragma unroll
            for (int ky = 0; ky < K; ++ky)
#pragma unroll
                for (int kx = 0; kx < K; ++kx) acc = fmaf(win[(i - (K - 1) + ky) % K][kx], w[ky][kx], acc);
            if (out_ok && oy < p.out_h) dst[(int64_t)oy * p.out_w + ox] = from_f32<T>(acc);
        }
    }
}

template <typename T>
static int launch_upfirdn(const UpfirdnParams &p, cudaStream_t st) {
    if (p.up_x == 1 && p.up_y == 1 && p.kh == 4 && p.kw == 4 && p.down_x == 1 && p.down_y == 1 && p.planes <= 65535 &&
        ceil_div(p.out_h, 32) <= 65535) {
        dim3 grid(ceil_div(p.out_w, (kThreads / 32) * 29), ceil_div(p.out_h, 32), (unsigned)p.planes);
        upfirdn2d_blur_kernel<T, 4><<<grid, kThreads, 0, st>>>(p);
        return check_launch("upfirdn2d");
    }
    if (p.up_x == 1 && p.up_y == 1 && p.kh == 4 && p.kw == 4 && p.down_x == p.down_y && (p.down_x == 1 || p.down_x == 2)) {
        const int64_t nblocks = p.planes * p.tiles_x * p.tiles_y;
        OOD_REQUIRE(nblocks < (1LL << 31), "upfirdn2d: grid too large");
        if (p.down_x == 1) upfirdn2d_up1_kernel<T, 1, 4, 4><<<(unsigned)nblocks, kThreads, 0, st>>>(p);
        else upfirdn2d_up1_kernel<T, 2, 4, 2><<<(unsigned)nblocks, kThreads, 0, st>>>(p);
        return check_launch("upfirdn2d");
    }
    const size_t smem = sizeof(float) * ((size_t)p.kh * p.kw + (size_t)p.sih * p.siw);
    OOD_REQUIRE(smem <= 200 * 1024, "upfirdn2d: FIR %dx%d with up %d/%d down %d/%d needs %zu B of shared memory",
                p.kh, p.kw, p.up_x, p.up_y, p.down_x, p.down_y, smem);
    const int64_t blocks = p.planes * p.tiles_x * p.tiles_y;
    OOD_REQUIRE(blocks < (1LL << 31), "upfirdn2d: grid too large");
    const bool sq = p.kh == p.kw && p.up_x == p.up_y && p.down_x == p.down_y;
#define OOD_UPF_LAUNCH(UP, DOWN, K)                                                                        \
    do {                                                                                                   \
        auto kern = upfirdn2d_kernel<T, UP, DOWN, K>;                                                      \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        kern<<<(unsigned)blocks, kThreads, smem, st>>>(p);                                                 \
    } while (0)
    if (sq && p.kh == 4 && p.up_x == 1 && p.down_x == 1) OOD_UPF_LAUNCH(1, 1, 4);
    else if (sq && p.kh == 4 && p.up_x == 2 && p.down_x == 1) OOD_UPF_LAUNCH(2, 1, 4);
    else if (sq && p.kh == 4 && p.up_x == 1 && p.down_x == 2) OOD_UPF_LAUNCH(1, 2, 4);
    else OOD_UPF_LAUNCH(0, 0, 0);
#undef OOD_UPF_LAUNCH
    return check_launch("upfirdn2d");
}

}  // namespace ood

extern "C" int ood_upfirdn2d(const void *in, void *out, const float *kernel, int64_t planes, int in_h, int in_w,
                             int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                             int pad_y0, int pad_y1, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(in && out && kernel, "upfirdn2d: null pointer");
    OOD_REQUIRE(planes >= 0 && in_h > 0 && in_w > 0 && kh > 0 && kw > 0, "upfirdn2d: bad sizes");
    OOD_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "upfirdn2d: bad dtype %d", dtype);
    UpfirdnParams p;
    p.in = in; p.out = out; p.kernel = kernel; p.planes = planes;
    p.in_h = in_h; p.in_w = in_w;
    p.kh = kh; p.kw = kw; p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
    p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
    const int full_h = in_h * up_y + pad_y0 + pad_y1 - kh, full_w = in_w * up_x + pad_x0 + pad_x1 - kw;
    OOD_REQUIRE(full_h >= 0 && full_w >= 0, "upfirdn2d: empty output (in %dx%d, pads %d,%d,%d,%d, fir %dx%d)", in_h,
                in_w, pad_x0, pad_x1, pad_y0, pad_y1, kh, kw);
    p.out_h = full_h / down_y + 1;
    p.out_w = full_w / down_x + 1;
    if (planes == 0) return OOD_OK;
    p.tiles_x = ceil_div(p.out_w, kTOW);
    p.tiles_y = ceil_div(p.out_h, kTOH);
    p.sih = ((kTOH - 1) * down_y + kh - 1) / up_y + 2;
    p.siw = ((kTOW - 1) * down_x + kw - 1) / up_x + 2;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) {      // wide fp32 planes: row-streaming cp.async kernel (plane_fir.cu)
        int handled = 0;
        const int rc = plane_fir(p, pad_x1, pad_y1, st, &handled);
        if (handled) return rc;
    }
    return dtype == OOD_F32 ? launch_upfirdn<float>(p, st) : launch_upfirdn<__nv_bfloat16>(p, st);
}
