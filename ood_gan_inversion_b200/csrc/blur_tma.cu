// TMA-staged version of the fused FIR-blur + StyledConv tail (see blur_act.cu for the op and its reference lines).
//
// The register-window kernel in blur_act.cu is latency-bound at 1024 px (ncu: 15 % of DRAM peak, 18 % warps active,
// long-scoreboard stalls): each thread has only a few 16-byte loads in flight.  Here a persistent CTA streams
// [19 x 35 pixels x 32 channels] input tiles through a 2-stage shared-memory ring with cp.async.bulk.tensor (one
// elected thread issues; the image border and the pad-(1,1) halo come from TMA out-of-bounds zero fill), so ~40-85 KB
// per SM are in flight while the previous tile is filtered from shared memory with conflict-free 16-byte LDS.
#include <cuda.h>

#include "common.cuh"

namespace ood {

constexpr int BT_TH = 16, BT_TW = 32, BT_CB = 32;            // output tile, channel block
constexpr int BT_IH = BT_TH + 3, BT_IW = BT_TW + 3;

struct BlurTmaParams {
    void *out_img, *out_y, *out_ys;
    const float *d, *noise, *noise_w, *bias, *s_next;
    int64_t noise_bstride;
    float k[4];
    int batch, oh, ow, C, pad0;
    int tiles_x, tiles_y, cblocks, total_tiles;
    int act;
};

__device__ __forceinline__ uint32_t bt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bt_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nBTW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra BTD;\nbra BTW;\nBTD:\n}\n" ::"r"(bt_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <typename TIN, int N>
__device__ __forceinline__ void lds_n(const TIN *p, float *dst);
template <> __device__ __forceinline__ void lds_n<float, 4>(const float *p, float *dst) {
    const float4 r = *reinterpret_cast<const float4 *>(p);
    dst[0] = r.x; dst[1] = r.y; dst[2] = r.z; dst[3] = r.w;
}
template <> __device__ __forceinline__ void lds_n<float, 8>(const float *p, float *dst) {
    lds_n<float, 4>(p, dst);
    lds_n<float, 4>(p + 4, dst + 4);
}
template <> __device__ __forceinline__ void lds_n<__nv_bfloat16, 8>(const __nv_bfloat16 *p, float *dst) {
    const uint4 r = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        dst[2 * i] = __uint_as_float(w[i] << 16);
        dst[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

template <typename T, typename TIN>
__global__ void __launch_bounds__(256) blur_tma_kernel(const __grid_constant__ CUtensorMap tm, const BlurTmaParams p) {
    constexpr int N = Vec<T>::N;
    constexpr int CVECS = BT_CB / N;                  // 4 (bf16 out) or 8 (fp32 out)
    constexpr int UNITS = BT_TW * CVECS;              // threads per row group
    constexpr int GROUPS = 256 / UNITS;               // 2 or 1
    constexpr int ROWS = BT_TH / GROUPS;              // output rows per thread
    constexpr int STAGE_BYTES = BT_IH * BT_IW * BT_CB * (int)sizeof(TIN);
    constexpr int STAGE_STRIDE = (STAGE_BYTES + 1023) & ~1023;
    extern __shared__ uint8_t bt_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(bt_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * STAGE_STRIDE);

    const int tid = threadIdx.x;
    const int grp = tid / UNITS, u = tid % UNITS;
    const int cvec = u % CVECS, x = u / CVECS;
    const float nw = (p.noise && p.noise_w) ? *p.noise_w : 0.f;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bt_smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bt_smem_u32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    }
    __syncthreads();

    auto issue = [&](int tile, int stage) {
        const int cb = tile % p.cblocks;
        int r = tile / p.cblocks;
        const int tx = r % p.tiles_x; r /= p.tiles_x;
        const int ty = r % p.tiles_y;
        const int b = r / p.tiles_y;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bt_smem_u32(&bars[stage])), "r"(STAGE_BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(bt_smem_u32(smem + stage * STAGE_STRIDE)), "l"(&tm), "r"(bt_smem_u32(&bars[stage])), "r"(cb * BT_CB),
            "r"(tx * BT_TW - p.pad0), "r"(ty * BT_TH - p.pad0), "r"(b)
            : "memory");
    };

    int stage = 0;
    uint32_t parity = 0;            // bit s = phase parity of stage s
    if (tid == 0 && (int)blockIdx.x < p.total_tiles) issue(blockIdx.x, 0);

    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int next = tile + gridDim.x;
        if (tid == 0 && next < p.total_tiles) issue(next, stage ^ 1);     // the other stage was released by the barrier below
        const int cb = tile % p.cblocks;
        int r = tile / p.cblocks;
        const int tx = r % p.tiles_x; r /= p.tiles_x;
        const int ty = r % p.tiles_y;
        const int b = r / p.tiles_y;
        const int c = cb * BT_CB + cvec * N;
        const int ox = tx * BT_TW + x;
        const int oy0 = ty * BT_TH + grp * ROWS;

        float dreg[N], breg[N], sreg[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            dreg[j] = p.d ? __ldg(p.d + (int64_t)b * p.C + c + j) : 1.f;
            breg[j] = (p.act && p.bias) ? __ldg(p.bias + c + j) : 0.f;
            sreg[j] = p.out_ys ? __ldg(p.s_next + (int64_t)b * p.C + c + j) : 1.f;
        }

        bt_wait(&bars[stage], (parity >> stage) & 1u);
        parity ^= 1u << stage;
        const TIN *tile_in = reinterpret_cast<const TIN *>(smem + stage * STAGE_STRIDE);

        float hw[4][N];
#pragma unroll
        for (int i = 0; i < ROWS + 3; ++i) {
            const int ir = grp * ROWS + i;               // input row inside the tile
            const TIN *rowp = tile_in + ((ir * BT_IW + x) * BT_CB + cvec * N);
            float acc[N];
#pragma unroll
            for (int j = 0; j < N; ++j) acc[j] = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[N];
                lds_n<TIN, N>(rowp + q * BT_CB, v);
#pragma unroll
                for (int j = 0; j < N; ++j) acc[j] = fmaf(p.k[q], v[j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < N; ++j) hw[i & 3][j] = acc[j];
            if (i >= 3) {
                const int oy = oy0 + i - 3;
                if (oy < p.oh && ox < p.ow) {
                    float v[N];
#pragma unroll
                    for (int j = 0; j < N; ++j)
                        v[j] = (p.k[0] * hw[(i - 3) & 3][j] + p.k[1] * hw[(i - 2) & 3][j] + p.k[2] * hw[(i - 1) & 3][j] +
                                p.k[3] * hw[i & 3][j]) * dreg[j];
                    const int64_t off = (((int64_t)b * p.oh + oy) * p.ow + ox) * p.C + c;
                    Vec<T> o;
                    if (p.out_img) {
#pragma unroll
                        for (int j = 0; j < N; ++j) o.v[j] = v[j];
                        store_vec<T>((T *)p.out_img + off, o);
                    }
                    if (p.act) {
                        const float nz = p.noise ? nw * __ldg(p.noise + b * p.noise_bstride + (int64_t)oy * p.ow + ox) : 0.f;
#pragma unroll
                        for (int j = 0; j < N; ++j) v[j] = lrelu_sqrt2(v[j] + nz + breg[j]);
                        if (p.out_y) {
#pragma unroll
                            for (int j = 0; j < N; ++j) o.v[j] = v[j];
                            store_vec<T>((T *)p.out_y + off, o);
                        }
                        if (p.out_ys) {
#pragma unroll
                            for (int j = 0; j < N; ++j) o.v[j] = v[j] * sreg[j];
                            store_vec<T>((T *)p.out_ys + off, o);
                        }
                    }
                }
            }
        }
        __syncthreads();            // everyone is done reading this stage: it may be refilled next iteration
        stage ^= 1;
    }
}

typedef CUresult (*BtEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T, typename TIN>
static int launch_blur_tma(const CUtensorMap &tm, const BlurTmaParams &p, cudaStream_t st) {
    constexpr int STAGE_BYTES = BT_IH * BT_IW * BT_CB * (int)sizeof(TIN);
    constexpr int STAGE_STRIDE = (STAGE_BYTES + 1023) & ~1023;
    constexpr int SMEM = 2 * STAGE_STRIDE + 1024 + 64;
    auto kern = blur_tma_kernel<T, TIN>;
    static DeviceOnce attr;
    if (attr.first()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) { set_error("blur_act tma: smem attribute: %s", cudaGetErrorString(e)); return OOD_ERR_CUDA; }
    }
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = SMEM <= 110 * 1024 ? 2 : 1;
    const int grid = std::min(p.total_tiles, sms * per_sm);
    kern<<<grid, 256, SMEM, st>>>(tm, p);
    return check_launch("blur_act tma");
}

// returns 1 if this path cannot take the configuration (caller falls back to the register-window kernel)
int blur_act_tma(const ood_blur_act_args *a, cudaStream_t st, int *handled) {
    *handled = 0;
    if (a->channels % BT_CB != 0 || !ood_device_is_sm100()) return OOD_OK;
    if (((uintptr_t)a->in % 16) != 0) return OOD_OK;
    static BtEncodeFn encode = nullptr;
    if (!encode) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return OOD_OK;
        encode = (BtEncodeFn)ptr;
    }
    const bool in_f32 = a->dtype == OOD_F32 || a->in_f32;
    const int es = in_f32 ? 4 : 2;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)a->channels, (cuuint64_t)a->iw, (cuuint64_t)a->ih, (cuuint64_t)a->batch};
    cuuint64_t strides[3] = {(cuuint64_t)a->channels * es, (cuuint64_t)a->iw * a->channels * es,
                             (cuuint64_t)a->ih * a->iw * a->channels * es};
    cuuint32_t box[4] = {BT_CB, BT_IW, BT_IH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, in_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(a->in),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OOD_OK;       // fall back
    BlurTmaParams p;
    p.out_img = a->out_img; p.out_y = a->out_y; p.out_ys = a->out_ys;
    p.d = a->d; p.noise = a->noise; p.noise_w = a->noise_w; p.bias = a->bias; p.s_next = a->s_next;
    p.noise_bstride = a->noise_bstride;
    for (int i = 0; i < 4; ++i) p.k[i] = a->taps[3 - i];
    p.pad0 = a->pad0 > 0 ? a->pad0 : 1;
    const int pad1 = a->pad0 > 0 ? a->pad1 : 1;
    p.batch = a->batch; p.oh = a->ih + p.pad0 + pad1 - 3; p.ow = a->iw + p.pad0 + pad1 - 3; p.C = a->channels;
    p.tiles_x = ceil_div(p.ow, BT_TW); p.tiles_y = ceil_div(p.oh, BT_TH); p.cblocks = a->channels / BT_CB;
    const int64_t total = (int64_t)p.tiles_x * p.tiles_y * p.cblocks * a->batch;
    if (total >= (1LL << 31)) return OOD_OK;
    p.total_tiles = (int)total;
    p.act = a->act;
    *handled = 1;
    if (a->dtype == OOD_F32) return launch_blur_tma<float, float>(tm, p, st);
    if (a->in_f32) return launch_blur_tma<__nv_bfloat16, float>(tm, p, st);
    return launch_blur_tma<__nv_bfloat16, __nv_bfloat16>(tm, p, st);
}

}  // namespace ood
