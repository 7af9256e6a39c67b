// Weight gradient of the shared-weight 3x3 / 1x1 convolution on the tcgen05 tensor cores (the
// training-side contraction; reference: autograd of the plain Conv2d layers of the AlignNet, src/ops/SAMM/helpers.py:85-109,
// e4e/encoders/helpers.py:426-448, trained by src/models/OOD_faceGAN_model.py:663-789):
//
//     gW[o, i, ky, kx] = sum over (b, y, x) of  g[b, y, x, o] * x[b, y + ky - 1, x + kx - 1, i]          (zero padding)
//
// A GEMM per tap with M = Co, N = Ci and K = ALL PIXELS.  Both operands are NHWC, i.e. [pixel][channel]: the contraction
// dimension is the OUTER one, so the tiles are "MN-major" operands of tcgen05.mma (instruction-descriptor bits 15 / 16) --
// no transposed copy of the activations is made.  A K chunk is a 64-pixel patch [TH x TW] of one image; TMA loads it as
// boxes [64 channels x TW x TH] (128-byte rows, SWIZZLE_128B), the tap's shift is a coordinate offset of the x box and the
// padding is TMA's out-of-bounds zero fill, exactly as in the forward kernel (conv_tc.cu).
//   work item = (M tile of 128 output channels, N tile of up to 256 input channels, tap, K slice); a CTA owns one item:
//   warp 0 TMA producer, warp 1 MMA issuer (one elected lane each), warps 2..5 drain the fp32 accumulator from TMEM to the
//   item's slot of a partial buffer [slice][tap][Co][Ci]; a second kernel sums the slices in a fixed order (deterministic) and
//   writes the PyTorch layout [Co][Ci][kh][kw].
#include <cuda.h>

#include "common.cuh"

namespace ood {
namespace wg {

constexpr int kThreads = 192;
constexpr int BM = 128;
constexpr int BKP = 64;                  // pixels per K chunk

struct WgParams {
    int batch, h, w, cin, cout, taps;    // taps 9 (3x3, pad 1) or 1 (1x1)
    int TW, TH, chunks_x, chunks_y;      // 64-pixel patch and patches per image
    int n_tiles_m, n_tiles_n, BN, slices, chunks_per_slice, total_chunks;
    float *partial;                      // [slices][taps][cout][cin]
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWG_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WG_DONE;\nbra WG_LOOP;\nWG_DONE:\n}\n" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(desc_a),
                 "l"(desc_b), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// MN-major operand, SWIZZLE_128B: 64 channels (128 bytes) contiguous per pixel row, 8-row swizzle atoms of 1024 bytes stacked
// along K (stride byte offset), 64-channel column blocks `lbo` bytes apart along M / N (leading byte offset).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int BN>
struct Cfg {
    static constexpr int kABytes = (BM / 64) * BKP * 128;         // column blocks of [64 px][64 ch]
    static constexpr int kBBytes = (BN / 64) * BKP * 128;
    static constexpr int kStage = kABytes + kBBytes;
    static constexpr int kStages = (192 * 1024 / kStage) > 6 ? 6 : (192 * 1024 / kStage);
    static constexpr int kSmem = kStages * kStage + 1024 + 256;
    static constexpr int kTmemCols = BN < 32 ? 32 : BN;
    // c = f32, a = b = bf16, a_major = b_major = MN (bits 15, 16), N >> 3, M >> 4
    static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgParams p, const uint32_t lbo, const uint32_t sbo) {
    using C = Cfg<BN>;
    constexpr int S = C::kStages;
    extern __shared__ uint8_t wg_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(wg_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + S * C::kABytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S * C::kStage);
    uint64_t *full = bars, *empty = bars + S, *done = bars + 2 * S;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 2 * S + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item of this CTA
    int item = blockIdx.x;
    const int nt = item % p.n_tiles_n; item /= p.n_tiles_n;
    const int mt = item % p.n_tiles_m; item /= p.n_tiles_m;
    const int tap = item % p.taps;     item /= p.taps;
    const int slice = item;
    const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
    const int c_begin = slice * p.chunks_per_slice, c_end = min(c_begin + p.chunks_per_slice, p.total_chunks);
    const int nchunks = max(c_end - c_begin, 0);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const int per_image = p.chunks_x * p.chunks_y;
            for (int c = c_begin; c < c_end; ++c) {
                const int b = c / per_image, r = c - b * per_image;
                const int y0 = (r / p.chunks_x) * p.TH, x0 = (r % p.chunks_x) * p.TW;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], C::kStage);
#pragma unroll
                for (int i = 0; i < BM / 64; ++i)
                    tma_load_4d(sA + stage * C::kABytes + i * (BKP * 128), &tmG, &full[stage], mt * BM + i * 64, x0, y0, b);
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                    tma_load_4d(sB + stage * C::kBBytes + j * (BKP * 128), &tmX, &full[stage], nt * BN + j * 64, x0 + dx, y0 + dy, b);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < nchunks; ++it) {
                mbar_wait(&full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = make_desc_mn(smem_u32(sA + stage * C::kABytes), lbo, sbo);
                const uint64_t db = make_desc_mn(smem_u32(sB + stage * C::kBBytes), lbo, sbo);
#pragma unroll
                for (int k = 0; k < BKP / 16; ++k)          // 16 pixel rows (2048 bytes) per instruction
                    umma_bf16(tmem_base, da + (uint64_t)(k * (2048 >> 4)), db + (uint64_t)(k * (2048 >> 4)), C::kIdesc, (it | k) != 0);
                umma_commit(&empty[stage]);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            umma_commit(done);
        }
    } else {
        // epilogue: TMEM lane quadrant = warp % 4; thread = one output channel (row of gW), 32 input channels per chunk
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int co = mt * BM + row;
        float *dst = p.partial + (((int64_t)slice * p.taps + tap) * p.cout + co) * p.cin + nt * BN;
        if (nchunks > 0) {
            mbar_wait(done, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
            uint32_t r[32];
            if (nchunks > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + ch * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            if (co < p.cout) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(dst + ch * 32 + 4 * j) =
                        make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
    }
}

// out[o][i][tap] = sum over slices of partial[s][tap][o][i]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial, float *__restrict__ out, int slices, int taps, int cout,
                                                            int cin, int cout_real, int cin_real) {
    const int64_t total = (int64_t)cout_real * cin_real * taps;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % taps);
        const int ci = (int)((i / taps) % cin_real);
        const int co = (int)(i / ((int64_t)taps * cin_real));
        float s = 0.f;
        for (int k = 0; k < slices; ++k) s += partial[(((int64_t)k * taps + t) * cout + co) * cin + ci];
        out[i] = s;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int plan(int batch, int h, int w, int cin, int cout, int taps, WgParams &p) {
    p.batch = batch; p.h = h; p.w = w; p.cin = cin; p.cout = cout; p.taps = taps;
    p.TW = std::min(w, BKP);
    if (p.TW <= 0 || BKP % p.TW != 0) return 0;
    p.TH = BKP / p.TW;
    if (w % p.TW != 0 || h % p.TH != 0) return 0;
    p.chunks_x = w / p.TW; p.chunks_y = h / p.TH;
    p.total_chunks = batch * p.chunks_x * p.chunks_y;
    p.BN = cin % 256 == 0 ? 256 : (cin % 128 == 0 ? 128 : 64);
    p.n_tiles_m = cout / BM; p.n_tiles_n = cin / p.BN;
    const int tiles = p.n_tiles_m * p.n_tiles_n * taps;
    // K slices: about two waves of CTAs, at least 8 chunks each
    int slices = std::max(1, (2 * kNumSMs + tiles - 1) / tiles);
    slices = std::min(slices, std::max(1, p.total_chunks / 8));
    p.chunks_per_slice = (p.total_chunks + slices - 1) / slices;
    p.slices = (p.total_chunks + p.chunks_per_slice - 1) / p.chunks_per_slice;
    return 1;
}

}  // namespace wg
}  // namespace ood

extern "C" int64_t ood_conv_wgrad_workspace(int batch, int h, int w, int cin, int cout, int taps) {
    using namespace ood;
    if (batch <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || (taps != 1 && taps != 9) || cout % wg::BM != 0 || cin % 64 != 0) return 0;
    wg::WgParams p{};
    if (!wg::plan(batch, h, w, cin, cout, taps, p)) return 0;
    return (int64_t)p.slices * taps * cout * cin * (int64_t)sizeof(float);
}

extern "C" int ood_conv_wgrad(const void *g, const void *x, float *workspace, float *gw, int batch, int h, int w, int cin, int cout, int taps,
                              int cin_real, int cout_real, void *stream) {
    using namespace ood;
    using namespace ood::wg;
    OOD_REQUIRE(g && x && workspace && gw && batch > 0 && h > 0 && w > 0, "conv_wgrad: bad arguments");
    OOD_REQUIRE(taps == 1 || taps == 9, "conv_wgrad: taps must be 1 (1x1) or 9 (3x3, pad 1)");
    OOD_REQUIRE(cout % BM == 0 && cin % 64 == 0, "conv_wgrad: cout %% 128 and cin %% 64 must be 0 (got %d, %d)", cout, cin);
    OOD_REQUIRE(cin_real > 0 && cin_real <= cin && cout_real > 0 && cout_real <= cout, "conv_wgrad: bad real channel counts");
    OOD_REQUIRE(((uintptr_t)g % 16) == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)workspace % 16) == 0, "conv_wgrad: operands must be 16-byte aligned");
    if (!ood_device_is_sm100()) { set_error("conv_wgrad: needs an sm_100 device"); return OOD_ERR_DEVICE; }
    WgParams p{};
    OOD_REQUIRE(plan(batch, h, w, cin, cout, taps, p), "conv_wgrad: %dx%d is outside the 64-pixel patch plan (w a power-of-two divisor or multiple of 64)", h, w);
    p.partial = workspace;
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) { set_error("conv_wgrad: no cuTensorMapEncodeTiled"); return OOD_ERR_CUDA; }
        encode = (EncodeFn)ptr;
    }
    auto make_map = [&](CUtensorMap &tm, const void *base, int channels) {
        cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
        cuuint64_t strides[3] = {(cuuint64_t)channels * 2, (cuuint64_t)w * channels * 2, (cuuint64_t)h * w * channels * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        return encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap tmG, tmX;
    if (make_map(tmG, g, cout) != CUDA_SUCCESS || make_map(tmX, x, cin) != CUDA_SUCCESS) { set_error("conv_wgrad: tensor map encode failed"); return OOD_ERR_CUDA; }
    // descriptor strides of the MN-major tiles: column blocks of 64 channels are one [64 px][128 B] box apart; 8-row atoms are 1024 B
    uint32_t lbo = BKP * 128, sbo = 1024;
    if (const char *e = getenv("OOD_WGRAD_LBO")) lbo = (uint32_t)atoi(e);
    if (const char *e = getenv("OOD_WGRAD_SBO")) sbo = (uint32_t)atoi(e);
    const int items = p.n_tiles_m * p.n_tiles_n * taps * p.slices;
    cudaStream_t st = (cudaStream_t)stream;
#define OOD_WG(bn)                                                                                                    \
    do {                                                                                                              \
        auto kern = wgrad_tc_kernel<bn>;                                                                              \
        static DeviceOnce attr;                                                                                       \
        if (attr.first()) {                                                                                           \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<bn>::kSmem);  \
            if (e != cudaSuccess) { set_error("conv_wgrad: smem attribute: %s", cudaGetErrorString(e)); return OOD_ERR_CUDA; } \
        }                                                                                                             \
        kern<<<items, kThreads, Cfg<bn>::kSmem, st>>>(tmG, tmX, p, lbo, sbo);                                         \
    } while (0)
    if (p.BN == 256) OOD_WG(256); else if (p.BN == 128) OOD_WG(128); else OOD_WG(64);
#undef OOD_WG
    const int64_t total = (int64_t)cout_real * cin_real * taps;
    wgrad_reduce_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, kNumSMs * 8), 256, 0, st>>>(workspace, gw, p.slices, taps, cout, cin, cout_real, cin_real);
    return check_launch("conv_wgrad", 2);
}
