// Row-streaming version of the fused FIR-blur + StyledConv tail (see blur_act.cu for the op and its reference lines:
// upfirdn2d.py:160-193 with up=down=1, model.py:286-292 noise, fused_act.py:84-96 bias + leaky-ReLU*sqrt2).
//
// blur_tma.cu filters independent [16 x 32] tiles: every tile re-reads a 3-row halo (19/16 of the rows through shared
// memory and the LSU), recomputes the horizontal pass for it, and the CTA drains at every tile boundary; ncu showed it
// issue-bound (33 instructions per element, 58 % issue utilisation, 41 % of DRAM peak).  Here a CTA owns a column strip
// [TW pixels x 32 channels] and streams its input rows ONCE through a shared-memory ring of 4-row groups filled by a
// producer warp with cp.async.bulk.tensor (the pad halo and the image border are TMA out-of-bounds zero fill).  A
// consumer thread owns one 16-byte channel vector of one column: per input row it does the horizontal pass from shared
// memory and slides a 4-row register window down for the vertical pass, so no row is filtered twice, the pipeline
// never drains between units of work, and the per-element instruction count drops to ~19.
#include <cuda.h>

#include "common.cuh"

namespace ood {

constexpr int BR_CB = 32, BR_RG = 4;

struct BlurRowsParams {
    void *out_img, *out_y, *out_ys;
    const float *d, *noise, *noise_w, *bias, *s_next;
    int64_t noise_bstride;
    float k[4];
    int batch, oh, ow, C, pad0;
    int tiles_x, cblocks, chunks, chunk_rows, total_units;
    int act;
};

__device__ __forceinline__ uint32_t br_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void br_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nBRW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra BRD;\nbra BRW;\nBRD:\n}\n" ::"r"(br_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 16 bytes of shared memory -> N/2 fp32 pairs
template <typename T>
__device__ __forceinline__ void br_lds(const T *p, float2 *dst);
template <> __device__ __forceinline__ void br_lds<float>(const float *p, float2 *dst) {
    const float4 r = *reinterpret_cast<const float4 *>(p);
    dst[0] = make_float2(r.x, r.y); dst[1] = make_float2(r.z, r.w);
}
template <> __device__ __forceinline__ void br_lds<__nv_bfloat16>(const __nv_bfloat16 *p, float2 *dst) {
    const uint4 r = *reinterpret_cast<const uint4 *>(p);
    dst[0] = unpack_bf16x2(r.x); dst[1] = unpack_bf16x2(r.y); dst[2] = unpack_bf16x2(r.z); dst[3] = unpack_bf16x2(r.w);
}
template <typename T>
__device__ __forceinline__ void br_stg(T *p, const float2 *v);
template <> __device__ __forceinline__ void br_stg<float>(float *p, const float2 *v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
}
template <> __device__ __forceinline__ void br_stg<__nv_bfloat16>(__nv_bfloat16 *p, const float2 *v) {
    uint4 r;
    r.x = pack_bf16x2(v[0].x, v[0].y); r.y = pack_bf16x2(v[1].x, v[1].y);
    r.z = pack_bf16x2(v[2].x, v[2].y); r.w = pack_bf16x2(v[3].x, v[3].y);
    *reinterpret_cast<uint4 *>(p) = r;
}

template <typename T> struct BrCfg {
    static constexpr int N = Vec<T>::N;                 // channels per thread (16 bytes)
    static constexpr int CVECS = BR_CB / N;             // 4 (bf16) or 8 (fp32)
    static constexpr int TW = 256 / CVECS;              // strip width: 64 (bf16) or 32 (fp32)
    static constexpr int IW = TW + 3;
    static constexpr int SLOT_BYTES = BR_RG * IW * BR_CB * (int)sizeof(T);   // 17152 / 17920: multiples of 128
};

// OUT: which outputs exist -- 0 out_img only (no activation), 1 out_y, 2 out_ys, 3 anything (checked at run time)
template <typename T, int NG, int OUT>
__global__ void __maxnreg__(112) blur_rows_kernel(const __grid_constant__ CUtensorMap tm, const BlurRowsParams p) {
    using Cfg = BrCfg<T>;
    constexpr int N = Cfg::N, N2 = N / 2, CVECS = Cfg::CVECS, TW = Cfg::TW, IW = Cfg::IW, SLOT = Cfg::SLOT_BYTES;
    extern __shared__ uint8_t br_raw[];
    // aligned by offset (not by casting the address) so the compiler keeps the shared address space: LDS, not generic LD
    uint8_t *smem = br_raw + ((128u - (br_smem_u32(br_raw) & 127u)) & 127u);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NG * SLOT);
    uint64_t *empty = full + NG;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < NG; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(br_smem_u32(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(br_smem_u32(&empty[i])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    }
    __syncthreads();

    auto decode = [&](int unit, int &b, int &cb, int &tx, int &oy0, int &rows) {
        cb = unit % p.cblocks;
        int r = unit / p.cblocks;
        tx = r % p.tiles_x; r /= p.tiles_x;
        const int ch = r % p.chunks;
        b = r / p.chunks;
        oy0 = ch * p.chunk_rows;
        rows = min(p.chunk_rows, p.oh - oy0);
    };

    if (warp == 8) {
        // ---------------------------------------------------------------- producer
        if (lane != 0) return;
        uint32_t it = 0;
        for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
            int b, cb, tx, oy0, rows;
            decode(unit, b, cb, tx, oy0, rows);
            const int ngroups = (rows + 3 + BR_RG - 1) / BR_RG;
            for (int g = 0; g < ngroups; ++g, ++it) {
                const int slot = it % NG;
                br_wait(&empty[slot], ((it / NG) & 1u) ^ 1u);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(br_smem_u32(&full[slot])), "r"(SLOT) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                    ::"r"(br_smem_u32(smem + slot * SLOT)), "l"(&tm), "r"(br_smem_u32(&full[slot])), "r"(cb * BR_CB),
                    "r"(tx * TW - p.pad0), "r"(oy0 - p.pad0 + g * BR_RG), "r"(b)
                    : "memory");
            }
        }
        return;
    }

    // -------------------------------------------------------------------- consumers
    const int cvec = tid % CVECS, x = tid / CVECS;
    const float nw = (p.noise && p.noise_w) ? *p.noise_w : 0.f;
    constexpr float kS2 = 1.4142135623730951f;
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        int b, cb, tx, oy0, rows;
        decode(unit, b, cb, tx, oy0, rows);
        const int c = cb * BR_CB + cvec * N;
        const int ox = tx * TW + x;
        const bool col_ok = ox < p.ow;
        const int ngroups = (rows + 3 + BR_RG - 1) / BR_RG;

        // lrelu_sqrt2(v*d + nz + bias) = max(t, 0.2 t) with t = v*(d*sqrt2) + (nz*sqrt2 + bias*sqrt2); all arithmetic on
        // channel pairs (FFMA2 / FMUL2 / FADD2)
        float2 d2[N2], b2[N2], sreg[N2];
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            const int cj = c + 2 * j;
            d2[j] = p.d ? make_float2(__ldg(p.d + (int64_t)b * p.C + cj) * kS2, __ldg(p.d + (int64_t)b * p.C + cj + 1) * kS2) : f2(kS2);
            b2[j] = (p.act && p.bias) ? make_float2(__ldg(p.bias + cj) * kS2, __ldg(p.bias + cj + 1) * kS2) : f2(0.f);
            sreg[j] = p.out_ys ? make_float2(__ldg(p.s_next + (int64_t)b * p.C + cj), __ldg(p.s_next + (int64_t)b * p.C + cj + 1)) : f2(1.f);
        }
        const float *nzp = p.noise ? p.noise + b * p.noise_bstride + ox : nullptr;
        const bool w_img = OUT == 0 || (OUT == 3 && p.out_img), w_act = OUT != 0 && (OUT != 3 || p.act);
        const bool w_y = OUT == 1 || (OUT == 3 && p.out_y), w_ys = OUT == 2 || (OUT == 3 && p.out_ys);
        // byte offset of this thread's first output row; advanced by one row per output row
        const int64_t row_bytes = (int64_t)p.ow * p.C * (int)sizeof(T);
        const int64_t off0 = ((((int64_t)b * p.oh + oy0) * p.ow + ox) * p.C + c) * (int)sizeof(T);
        const float2 nwk = f2(nw * kS2);
        const float2 k2[4] = {f2(p.k[0]), f2(p.k[1]), f2(p.k[2]), f2(p.k[3])};

        // noise of the (up to) four output rows of a group, fetched one group ahead: it streams from DRAM, and ncu showed
        // the warps parked on it (long-scoreboard 3.2 stalled warps per issue) when it was loaded in the group that uses it
        auto load_noise = [&](int g, float *nz) {
#pragma unroll
            for (int r = 0; r < BR_RG; ++r) {
                const int oy = oy0 + g * BR_RG + r - 3;
                nz[r] = (nzp && col_ok && oy >= oy0 && oy < oy0 + rows) ? __ldg(nzp + (int64_t)oy * p.ow) : 0.f;
            }
        };
        float nz_next[BR_RG];
        load_noise(0, nz_next);

        float2 hw[4][N2];
        for (int g = 0; g < ngroups; ++g, ++it) {
            const int slot = it % NG;
            float nz[BR_RG];
#pragma unroll
            for (int r = 0; r < BR_RG; ++r) nz[r] = nz_next[r];
            load_noise(g + 1, nz_next);
            br_wait(&full[slot], (it / NG) & 1u);
            const int64_t goff = off0 + (int64_t)(g * BR_RG - 3) * row_bytes;
            const T *grp_in = reinterpret_cast<const T *>(smem + slot * SLOT) + (x * BR_CB + cvec * N);
#pragma unroll
            for (int r = 0; r < BR_RG; ++r) {
                const T *rowp = grp_in + r * IW * BR_CB;
                {
                    float2 v[N2];
                    br_lds<T>(rowp, v);
#pragma unroll
                    for (int j = 0; j < N2; ++j) hw[r][j] = mul2(k2[0], v[j]);
                }
#pragma unroll
                for (int q = 1; q < 4; ++q) {
                    float2 v[N2];
                    br_lds<T>(rowp + q * BR_CB, v);
#pragma unroll
                    for (int j = 0; j < N2; ++j) hw[r][j] = fma2(k2[q], v[j], hw[r][j]);
                }
                const int i = g * BR_RG + r;                // input row of the unit
                if (i >= 3 && i - 3 < rows && col_ok) {
                    float2 v[N2];
#pragma unroll
                    for (int j = 0; j < N2; ++j)
                        v[j] = fma2(k2[3], hw[r][j], fma2(k2[2], hw[(r + 3) & 3][j], fma2(k2[1], hw[(r + 2) & 3][j], mul2(k2[0], hw[(r + 1) & 3][j]))));
                    const int64_t off = goff + r * row_bytes;
                    float2 o[N2];
                    if (w_img) {
#pragma unroll
                        for (int j = 0; j < N2; ++j) o[j] = mul2(v[j], mul2(d2[j], f2(1.f / kS2)));
                        br_stg<T>(reinterpret_cast<T *>((char *)p.out_img + off), o);
                    }
                    if (w_act) {
                        const float2 nz2 = f2(nz[r]);
#pragma unroll
                        for (int j = 0; j < N2; ++j) {
                            const float2 t = fma2(v[j], d2[j], fma2(nz2, nwk, b2[j]));
                            v[j] = max2(t, mul2(t, f2(0.2f)));
                        }
                        if (w_y) br_stg<T>(reinterpret_cast<T *>((char *)p.out_y + off), v);
                        if (w_ys) {
#pragma unroll
                            for (int j = 0; j < N2; ++j) o[j] = mul2(v[j], sreg[j]);
                            br_stg<T>(reinterpret_cast<T *>((char *)p.out_ys + off), o);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(br_smem_u32(&empty[slot])) : "memory");
        }
    }
}

typedef CUresult (*BrEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T, int NG, int OUT>
static int launch_blur_rows(const CUtensorMap &tm, BlurRowsParams &p, cudaStream_t st) {
    using Cfg = BrCfg<T>;
    constexpr int SMEM = NG * Cfg::SLOT_BYTES + 2 * NG * 8 + 128;
    auto kern = blur_rows_kernel<T, NG, OUT>;
    static DeviceOnce attr;
    if (attr.first()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) { set_error("blur_act rows: smem attribute: %s", cudaGetErrorString(e)); return OOD_ERR_CUDA; }
    }
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = std::max(1, std::min(2, (227 * 1024) / (SMEM + 1024)));
    const int grid_max = sms * per_sm;
    // split the strips into row chunks so that the static round-robin over the persistent CTAs ends evenly: pick the
    // chunk count with the smallest makespan  ceil(units / grid) * (chunk_rows + 3)
    p.tiles_x = ceil_div(p.ow, Cfg::TW);
    const int64_t strips = (int64_t)p.tiles_x * p.cblocks * p.batch;
    int best_chunks = 1;
    int64_t best_cost = -1;
    const int max_chunks = std::max(1, p.oh / 29);
    for (int ch = 1; ch <= max_chunks; ++ch) {
        int rows = ceil_div(p.oh, ch);
        rows = (rows + 6) / 4 * 4 - 3;                       // rows + 3 input rows = whole 4-row groups
        const int nch = ceil_div(p.oh, rows);
        const int64_t units = strips * nch;
        const int64_t g = std::min<int64_t>(units, grid_max);
        const int64_t cost = ((units + g - 1) / g) * (rows + 3 + 2);   // +2: per-unit start-up
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_chunks = ch; }
    }
    int rows = ceil_div(p.oh, best_chunks);
    rows = (rows + 6) / 4 * 4 - 3;
    p.chunk_rows = rows;
    p.chunks = ceil_div(p.oh, rows);
    const int64_t total = strips * p.chunks;
    if (total >= (1LL << 31)) return 1;
    p.total_units = (int)total;
    const int grid = (int)std::min<int64_t>(total, grid_max);
    kern<<<grid, 288, SMEM, st>>>(tm, p);
    return check_launch("blur_act rows");
}

// handled = 1 if this path took the configuration (otherwise the caller tries the tile kernel / register-window kernel)
int blur_act_rows(const ood_blur_act_args *a, cudaStream_t st, int *handled) {
    *handled = 0;
    if (a->channels % BR_CB != 0 || !ood_device_is_sm100()) return OOD_OK;
    if (((uintptr_t)a->in % 16) != 0) return OOD_OK;
    if (a->dtype != OOD_F32 && a->in_f32) return OOD_OK;
    const int pad0 = a->pad0 > 0 ? a->pad0 : 1, pad1 = a->pad0 > 0 ? a->pad1 : 1;
    const int oh = a->ih + pad0 + pad1 - 3, ow = a->iw + pad0 + pad1 - 3;
    const int tw = a->dtype == OOD_F32 ? BrCfg<float>::TW : BrCfg<__nv_bfloat16>::TW;
    if (ow < tw || oh < 32) return OOD_OK;                  // small images: the tile kernel wastes fewer lanes
    static BrEncodeFn encode = nullptr;
    if (!encode) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return OOD_OK;
        encode = (BrEncodeFn)ptr;
    }
    const int es = a->dtype == OOD_F32 ? 4 : 2;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)a->channels, (cuuint64_t)a->iw, (cuuint64_t)a->ih, (cuuint64_t)a->batch};
    cuuint64_t strides[3] = {(cuuint64_t)a->channels * es, (cuuint64_t)a->iw * a->channels * es,
                             (cuuint64_t)a->ih * a->iw * a->channels * es};
    cuuint32_t box[4] = {BR_CB, (cuuint32_t)(tw + 3), BR_RG, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, a->dtype == OOD_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                        const_cast<void *>(a->in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OOD_OK;
    BlurRowsParams p{};
    p.out_img = a->out_img; p.out_y = a->out_y; p.out_ys = a->out_ys;
    p.d = a->d; p.noise = a->noise; p.noise_w = a->noise_w; p.bias = a->bias; p.s_next = a->s_next;
    p.noise_bstride = a->noise_bstride;
    for (int i = 0; i < 4; ++i) p.k[i] = a->taps[3 - i];   // correlation with the flipped FIR (upfirdn2d.py:179)
    p.pad0 = pad0;
    p.batch = a->batch; p.oh = oh; p.ow = ow; p.C = a->channels;
    p.cblocks = a->channels / BR_CB;
    p.act = a->act;
    int rc;
    const int out = (!a->act && a->out_img) ? 0
                    : (a->act && !a->out_img && a->out_y && !a->out_ys) ? 1
                    : (a->act && !a->out_img && !a->out_y && a->out_ys) ? 2 : 3;
#define OOD_BR(T) (out == 0 ? launch_blur_rows<T, 6, 0>(tm, p, st) : out == 1 ? launch_blur_rows<T, 6, 1>(tm, p, st) \
                   : out == 2 ? launch_blur_rows<T, 6, 2>(tm, p, st) : launch_blur_rows<T, 6, 3>(tm, p, st))
    if (a->dtype == OOD_F32) rc = OOD_BR(float);
    else rc = OOD_BR(__nv_bfloat16);
#undef OOD_BR
    if (rc == 1) return OOD_OK;                             // too many units: let the tile kernel take it
    *handled = 1;
    return rc;
}

}  // namespace ood
