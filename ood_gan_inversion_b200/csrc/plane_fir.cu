// Row-streaming 4x4 FIR resampling of wide fp32 NCHW planes: blur (up 1, down 1), down-sampling by 2 and up-sampling by 2.
// Same op as upfirdn2d.cu (reference src/ops/op/upfirdn2d.py:160-193, upfirdn2d_kernel.cu:52-137) for the three
// configurations the reference actually instantiates (model.py:30-88: Upsample pad (2,1), Downsample pad (1,1), Blur).
//
// The tile kernels in upfirdn2d.cu stage a [32 x 64] output tile per CTA with scalar loads that wait in registers, and
// sat at 0.14-0.45 of HBM peak (sweep_r01.json).  Here a CTA owns a 64 / 128 / 256 / 512-column strip of one plane and streams its input
// rows ONCE through a shared-memory ring with cp.async (4-byte copies: plane rows of odd width, e.g. the 1025-wide
// transposed-conv output, are only 4-byte aligned, which rules out TMA and 16-byte vectors), so 12-16 KB per CTA are in
// flight without holding registers.  A thread owns the output columns t and t + SW/2 as one fp32x2 pair: the vertical
// taps slide through a 4-row register window, the 16 (4 for up-sampling) multiply-adds per output pair are FFMA2 with a
// broadcast weight, and every load / store instruction of a warp touches one contiguous 128-byte (256 for up 2) run.
#include "upfirdn_common.cuh"

namespace ood {

constexpr int PF_G = 4;      // input rows per ring slot
// SW: strip width (MODE 0 / 1: output columns, MODE 2: input columns); SW / 2 threads, thread t owns columns t and t + SW / 2

struct PlaneFirParams {
    const float *in;
    float *out;
    const float *kernel;
    int in_h, in_w, out_h, out_w, pad_x0, pad_y0;
    int chunk_rows;            // MODE 0 / 1: output rows per unit; MODE 2: input rows per unit
};

template <int MODE, int SW> struct PfCfg {
    static constexpr int SEG = MODE == 0 ? SW + 3 : (MODE == 1 ? 2 * SW + 2 : SW + 2);            // input elements per row of a strip
    static constexpr int SEGP = (SEG + 3) & ~3;
    static constexpr int NG = MODE == 1 ? 3 : 4;                                                   // ring depth in 4-row groups
    static constexpr int SMEM = NG * PF_G * SEGP * (int)sizeof(float);
};

__device__ __forceinline__ void pf_cp_async4(uint32_t dst_saddr, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void pf_st_zero(uint32_t dst_saddr) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst_saddr), "r"(0) : "memory"); }
__device__ __forceinline__ void pf_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void pf_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MODE, int SW>
__global__ void __launch_bounds__(SW / 2) plane_fir_kernel(const PlaneFirParams p) {
    using Cfg = PfCfg<MODE, SW>;
    constexpr int PF_T = SW / 2, PF_SW = SW, HALF = SW / 2;
    constexpr int SEG = Cfg::SEG, SEGP = Cfg::SEGP, NG = Cfg::NG;
    extern __shared__ float pf_ring[];                      // [NG][PF_G][SEGP]
    const int t = threadIdx.x;
    const int strip = blockIdx.x, chunk = blockIdx.y;
    const int64_t plane = blockIdx.z;
    const float *src = p.in + plane * (int64_t)p.in_h * p.in_w;
    float *dst = p.out + plane * (int64_t)p.out_h * p.out_w;

    // flipped FIR (correlation form, upfirdn2d.py:179)
    float w[4][4];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) w[ky][kx] = __ldg(p.kernel + (3 - ky) * 4 + (3 - kx));

    // unit geometry: local input row r <-> global row iy0 + r, segment element e <-> global column ix0 + e
    int u0, u1, iy0, ix0, nrows;
    if (MODE == 0) {
        u0 = chunk * p.chunk_rows; u1 = min(p.out_h, u0 + p.chunk_rows);
        iy0 = u0 - p.pad_y0; ix0 = strip * PF_SW - p.pad_x0; nrows = (u1 - u0) + 3;
    } else if (MODE == 1) {
        u0 = chunk * p.chunk_rows; u1 = min(p.out_h, u0 + p.chunk_rows);
        iy0 = 2 * u0 - p.pad_y0; ix0 = 2 * strip * PF_SW - p.pad_x0; nrows = 2 * (u1 - u0) + 2;
    } else {
        u0 = chunk * p.chunk_rows; u1 = min(p.in_h, u0 + p.chunk_rows);
        iy0 = u0 - 1; ix0 = strip * PF_SW - 1; nrows = (u1 - u0) + 2;
    }
    const int ngroups = (nrows + PF_G - 1) / PF_G;

    // Copy plan of this thread, fixed for the whole unit: elements t, t + T, t + 2T, ... of every row segment.  The first
    // version recomputed bounds and 64-bit addresses per element and spent half of its 46 instructions per output there
    // (ncu: issue-bound at 70 %, 0.66 of HBM peak).  Columns outside the image are zeroed once: no copy ever lands there.
    constexpr int NE = (SEG + PF_T - 1) / PF_T;
    bool cok[NE];
#pragma unroll
    for (int j = 0; j < NE; ++j) {
        const int e = t + j * PF_T, ix = ix0 + e;
        cok[j] = e < SEG && ix >= 0 && ix < p.in_w;
    }
    for (int i = t; i < NG * PF_G * SEGP; i += PF_T) pf_ring[i] = 0.f;
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(pf_ring) + t * 4;
    const float *gcol = src + ix0 + t;                      // dereferenced only where cok[] holds

    auto issue = [&](int g) {
        if (g < ngroups) {
#pragma unroll
            for (int rr = 0; rr < PF_G; ++rr) {
                const int iy = iy0 + g * PF_G + rr;
                const uint32_t srow = sbase + (uint32_t)(((g % NG) * PF_G + rr) * SEGP) * 4u;
                if (iy >= 0 && iy < p.in_h) {
                    const float *grow = gcol + (int64_t)iy * p.in_w;
#pragma unroll
                    for (int j = 0; j < NE; ++j)
                        if (cok[j]) pf_cp_async4(srow + j * PF_T * 4, grow + j * PF_T);
                } else {                                    // rows above / below the image: zero padding
#pragma unroll
                    for (int j = 0; j < NE; ++j)
                        if (cok[j]) pf_st_zero(srow + j * PF_T * 4);
                }
            }
        }
        pf_commit();
    };

#pragma unroll
    for (int g = 0; g < NG - 1; ++g) issue(g);

    float2 win[4][4];                                       // [row & 3][tap]: columns (t, t + 256) of the strip
    const int oxa = strip * PF_SW + t, oxb = oxa + HALF;     // MODE 0 / 1: output columns; MODE 2: input columns
    const bool oka = oxa < (MODE == 2 ? p.in_w : p.out_w), okb = oxb < (MODE == 2 ? p.in_w : p.out_w);
    // first output element of this thread; advanced by one (MODE 2: two) output rows per result
    float *optr = dst + (int64_t)(MODE == 2 ? 2 * u0 : u0) * p.out_w + (MODE == 2 ? 2 * oxa : oxa);
    for (int g = 0; g < ngroups; ++g) {
        issue(g + NG - 1);                                  // its slot was released by the barrier that ended iteration g - 1
        pf_wait<NG - 1>();
        __syncthreads();
        const float *slot = pf_ring + (size_t)(g % NG) * PF_G * SEGP;
#pragma unroll
        for (int rr = 0; rr < PF_G; ++rr) {
            const int r = g * PF_G + rr;                    // local input row; r & 3 == rr
            const float *srow = slot + rr * SEGP;
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) win[rr][k] = make_float2(srow[t + k], srow[t + HALF + k]);
                const int oy = u0 + r - 3;
                if (r >= 3 && oy < u1) {
                    float2 acc = f2(0.f);
#pragma unroll
                    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 4; ++kx) acc = fma2(f2(w[ky][kx]), win[(rr + 1 + ky) & 3][kx], acc);
                    if (oka) optr[0] = acc.x;
                    if (okb) optr[HALF] = acc.y;
                    optr += p.out_w;
                }
            } else if (MODE == 1) {
                const float2 a01 = *reinterpret_cast<const float2 *>(srow + 2 * t), a23 = *reinterpret_cast<const float2 *>(srow + 2 * t + 2);
                const float2 b01 = *reinterpret_cast<const float2 *>(srow + 2 * t + SW), b23 = *reinterpret_cast<const float2 *>(srow + 2 * t + SW + 2);
                win[rr][0] = make_float2(a01.x, b01.x); win[rr][1] = make_float2(a01.y, b01.y);
                win[rr][2] = make_float2(a23.x, b23.x); win[rr][3] = make_float2(a23.y, b23.y);
                const int oy = u0 + (r - 3) / 2;
                if ((rr & 1) && r >= 3 && oy < u1) {         // output row j uses local rows 2j .. 2j + 3
                    float2 acc = f2(0.f);
#pragma unroll
                    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 4; ++kx) acc = fma2(f2(w[ky][kx]), win[(rr + 1 + ky) & 3][kx], acc);
                    if (oka) optr[0] = acc.x;
                    if (okb) optr[HALF] = acc.y;
                    optr += p.out_w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) win[rr][k] = make_float2(srow[t + k], srow[t + HALF + k]);
                const int y = u0 + r - 2;                    // window rows r-2, r-1, r = input rows y-1, y, y+1
                if (r >= 2 && y < u1) {
                    const float2 *R0 = win[(rr + 2) & 3], *R1 = win[(rr + 3) & 3], *R2 = win[rr];
                    // polyphase (pad (2,1)): even outputs use taps 0,2 on inputs (i-1, i); odd outputs taps 1,3 on (i, i+1)
                    const float2 e0 = fma2(f2(w[2][2]), R1[1], fma2(f2(w[2][0]), R1[0], fma2(f2(w[0][2]), R0[1], mul2(f2(w[0][0]), R0[0]))));
                    const float2 e1 = fma2(f2(w[2][3]), R1[2], fma2(f2(w[2][1]), R1[1], fma2(f2(w[0][3]), R0[2], mul2(f2(w[0][1]), R0[1]))));
                    const float2 o0 = fma2(f2(w[3][2]), R2[1], fma2(f2(w[3][0]), R2[0], fma2(f2(w[1][2]), R1[1], mul2(f2(w[1][0]), R1[0]))));
                    const float2 o1 = fma2(f2(w[3][3]), R2[2], fma2(f2(w[3][1]), R2[1], fma2(f2(w[1][3]), R1[2], mul2(f2(w[1][1]), R1[1]))));
                    float *row1 = optr + p.out_w;
                    if (oka) {
                        *reinterpret_cast<float2 *>(optr) = make_float2(e0.x, e1.x);
                        *reinterpret_cast<float2 *>(row1) = make_float2(o0.x, o1.x);
                    }
                    if (okb) {
                        *reinterpret_cast<float2 *>(optr + SW) = make_float2(e0.y, e1.y);
                        *reinterpret_cast<float2 *>(row1 + SW) = make_float2(o0.y, o1.y);
                    }
                    optr += 2 * p.out_w;
                }
            }
        }
        __syncthreads();
    }
}

template <int MODE, int SW>
static int launch_plane_fir(const PlaneFirParams &p0, int64_t planes, cudaStream_t st) {
    using Cfg = PfCfg<MODE, SW>;
    constexpr int PF_T = SW / 2, PF_SW = SW;
    PlaneFirParams p = p0;
    auto kern = plane_fir_kernel<MODE, SW>;
    static DeviceOnce attr;
    if (attr.first()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("upfirdn2d plane: smem attribute: %s", cudaGetErrorString(e)); return OOD_ERR_CUDA; }
    }
    const int unit_h = MODE == 2 ? p.in_h : p.out_h, unit_w = MODE == 2 ? p.in_w : p.out_w;
    const int strips = ceil_div(unit_w, PF_SW);
    // enough units for ~8 CTAs per SM, chunks of at least 32 rows (3 halo rows are re-read per chunk)
    int chunks = (int)std::min<int64_t>(std::max<int64_t>(1, (int64_t)kNumSMs * 8 / std::max<int64_t>(1, planes * strips)), std::max(1, unit_h / 32));
    p.chunk_rows = ceil_div(unit_h, chunks);
    if (MODE == 1) p.chunk_rows = (p.chunk_rows + 1) & ~1;
    chunks = ceil_div(unit_h, p.chunk_rows);
    dim3 grid(strips, chunks, (unsigned)planes);
    kern<<<grid, PF_T, Cfg::SMEM, st>>>(p);
    return check_launch("upfirdn2d");
}

int plane_fir(const UpfirdnParams &u, int pad_x1, int pad_y1, cudaStream_t st, int *handled) {
    *handled = 0;
    if (u.kh != 4 || u.kw != 4 || u.planes > 65535 || u.planes <= 0) return OOD_OK;
    if (u.up_x != u.up_y || u.down_x != u.down_y) return OOD_OK;
    if (((uintptr_t)u.in % 4) != 0 || ((uintptr_t)u.out % 8) != 0) return OOD_OK;
    PlaneFirParams p;
    p.in = (const float *)u.in; p.out = (float *)u.out; p.kernel = u.kernel;
    p.in_h = u.in_h; p.in_w = u.in_w; p.out_h = u.out_h; p.out_w = u.out_w; p.pad_x0 = u.pad_x0; p.pad_y0 = u.pad_y0;
    p.chunk_rows = 0;
    int rc;
    const bool pads_ok = u.pad_x0 >= 0 && u.pad_y0 >= 0;
    // strip width: the widest of 512 / 256 / 128 / 64 that the plane fills (a thread owns columns t and t + SW/2)
#define OOD_PF(MODE, W) ((W) >= 384 ? launch_plane_fir<MODE, 512>(p, u.planes, st) : (W) >= 192 ? launch_plane_fir<MODE, 256>(p, u.planes, st) \
                        : (W) >= 96 ? launch_plane_fir<MODE, 128>(p, u.planes, st) : launch_plane_fir<MODE, 64>(p, u.planes, st))
    if (u.up_x == 1 && u.down_x == 1 && u.out_w >= 48 && pads_ok) rc = OOD_PF(0, u.out_w);
    else if (u.up_x == 1 && u.down_x == 2 && u.out_w >= 48 && pads_ok) rc = OOD_PF(1, u.out_w);
    else if (u.up_x == 2 && u.down_x == 1 && u.in_w >= 48 && u.pad_x0 == 2 && u.pad_y0 == 2 && pad_x1 == 1 && pad_y1 == 1 &&
             u.out_w % 2 == 0) rc = OOD_PF(2, u.in_w);
    else return OOD_OK;
#undef OOD_PF
    *handled = 1;
    return rc;
}

}  // namespace ood
