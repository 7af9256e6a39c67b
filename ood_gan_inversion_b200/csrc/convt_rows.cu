// Row-streaming form of the stride-2 transposed 3x3 convolution for the 64 -> 32 up-sampling layer at 1024 px
// (src/ops/StyleGAN/model.py:246-256: conv_transpose2d(stride 2) of the 512 x 512 x 64 activations; SURVEY appendix A: 96 FLOP/B,
// HBM-bound).  Round 1 ran it as one GEMM over the four output-parity phases with zero-padded shift weights (conv_tc.cu form 5:
// 16/9 of the MACs, every input patch fetched four times from L2, 0.43 of the HBM roofline).  Here, as in conv_rows.cu:
//   * a persistent CTA owns a strip of 32 position rows of one 128-position column block; the nine weight tiles are loaded once;
//   * every INPUT row segment [129 px x 64 ch] is loaded once into a ring; position row oy reads rows oy (dy = 0) and oy - 1 (dy = -1),
//     the horizontal shift dx = -1 is the same buffer with the descriptor start one pixel row earlier;
//   * the four output-parity phases are the N dimension, ordered (0,1),(0,0),(1,0),(1,1) so that every input shift feeds a CONTIGUOUS
//     range of phases: shift (0,0) -> all four (N = 4 Co), (0,-1) -> (0,0),(1,0), (-1,0) -> (0,1),(0,0), (-1,-1) -> (0,0): exactly the
//     nine taps, no multiplications by zero;
//   * a position row's accumulator is two complete output row segments [256 px x Co] (Y = 2 oy and 2 oy + 1): the epilogue writes them
//     through swizzled staging with two bulk tensor stores.
// Only the interior positions (oy < h, ox < w) are computed here; the last output row and column go through the generic tiles
// (conv_common.cuh: make_geom_transposed_part 1 and 2).
#include <cuda.h>

#include "conv_common.cuh"

namespace ood {
namespace trows {

constexpr int kThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two groups of four warps: alternate position rows)
constexpr int kStripRows = 32;
constexpr int CI = 64, CO = 32;
constexpr int ROWB = CI * 2;           // 128-byte pixel rows: SWIZZLE_128B
constexpr int kRowPx = 129;            // 128 positions + the left neighbour (dx = -1)
constexpr int kRowBytes = kRowPx * ROWB;
constexpr int kRowStride = (kRowBytes + 1023) & ~1023;
constexpr int kRing = 6;
constexpr int kWTile = CO * ROWB;      // one tap's [Co x Ci] tile: 4096 bytes
constexpr int kNAcc = 4;               // accumulators of 4*Co = 128 columns
constexpr int kOutRow = 256 * CO * 2;  // one staged output row segment: 256 px x Co bf16
constexpr int kSmem = 9 * kWTile + kRing * kRowStride + 2 * 2 * kOutRow + 1024 + 512;
// block order of the weight tiles in shared memory: [t01 t00 t10 t11 | t02 t12 | t21 t20 | t22]  (tap = ky*3 + kx)
__constant__ int kTapOrder[9] = {1, 0, 3, 4, 2, 5, 7, 6, 8};

struct Params {
    int batch, h, w, tiles_x, strips_y, total_strips;
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nTR_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra TR_DONE;\nbra TR_LOOP;\nTR_DONE:\n}\n" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile("{\n.reg .b64 da, db;\n.reg .pred p;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d),
                 "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {        // K-major, SWIZZLE_128B, 8-row atoms of 1024 bytes
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct Ring {
    int slot;
    uint32_t phase;
    __device__ __forceinline__ void advance() { if (++slot == kRing) { slot = 0; phase ^= 1; } }
    __device__ __forceinline__ Ring next() const { Ring r = *this; r.advance(); return r; }
};

__global__ void __launch_bounds__(kThreads, 1)
convt_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY, const Params p) {
    extern __shared__ uint8_t trows_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(trows_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sW = smem;
    uint8_t *sR = smem + 9 * kWTile;
    uint8_t *sO = sR + kRing * kRowStride;                        // staging: [group][py][256 px][CO] bf16, SWIZZLE_64B
    uint64_t *bars = reinterpret_cast<uint64_t *>(sO + 2 * 2 * kOutRow);
    uint64_t *full = bars, *empty = bars + kRing, *tfull = bars + 2 * kRing, *tempty = tfull + kNAcc, *wbar = tempty + kNAcc;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kRing; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kNAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(kNAcc * 4 * CO) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    auto strip_coords = [&](int strip, int &b, int &ya, int &x0, int &nrows) {
        const int tx = strip % p.tiles_x;
        int r = strip / p.tiles_x;
        const int sy = r % p.strips_y;
        b = r / p.strips_y;
        ya = sy * kStripRows;
        x0 = tx * 128;
        nrows = min(kStripRows, p.h - ya);
    };

    if (warp == 0) {
        // ===================================================== TMA producer: input rows ya - 1 .. ya + nrows - 1 of the strip
        if (elect_one()) {
            mbar_expect_tx(wbar, 9 * kWTile);
            for (int t = 0; t < 9; ++t) tma_load_3d(sW + t * kWTile, &tmB, wbar, 0, 0, kTapOrder[t]);
            Ring ring{0, 0};
            for (int strip = blockIdx.x; strip < p.total_strips; strip += gridDim.x) {
                int b, ya, x0, nrows;
                strip_coords(strip, b, ya, x0, nrows);
                for (int i = 0; i < nrows + 1; ++i) {
                    mbar_wait(&empty[ring.slot], ring.phase ^ 1);
                    mbar_expect_tx(&full[ring.slot], kRowBytes);
                    tma_load_4d(sR + ring.slot * kRowStride, &tmA, &full[ring.slot], 0, x0 - 1, ya - 1 + i, b);      // out-of-range rows / columns: zero fill
                    ring.advance();
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (elect_one()) {
            mbar_wait(wbar, 0);
            tc_fence_after();
            Ring base{0, 0};
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint64_t dproto = make_desc(0);
            const uint32_t desc_hi = (uint32_t)(dproto >> 32);
            const uint32_t ring_lo = (uint32_t)dproto | ((smem_u32(sR) >> 4) & 0x3FFF);
            const uint32_t w_lo = (uint32_t)dproto | ((smem_u32(sW) >> 4) & 0x3FFF);
            constexpr uint32_t kIdescNoN = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
            auto idesc = [](int n) { return kIdescNoN | ((uint32_t)(n >> 3) << 17); };
            for (int strip = blockIdx.x; strip < p.total_strips; strip += gridDim.x) {
                int b, ya, x0, nrows;
                strip_coords(strip, b, ya, x0, nrows);
                for (int j = 0; j < nrows; ++j) {
                    const Ring r0 = base, r1 = r0.next();             // input rows oy - 1 (dy = -1) and oy (dy = 0)
                    if (j == 0) mbar_wait(&full[r0.slot], r0.phase);
                    mbar_wait(&full[r1.slot], r1.phase);
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(acc * 4 * CO);
                    const uint32_t a_up = ring_lo + (uint32_t)r0.slot * (kRowStride >> 4), a_cur = ring_lo + (uint32_t)r1.slot * (kRowStride >> 4);
#pragma unroll
                    for (int k = 0; k < CI / 16; ++k) {
                        const uint32_t ko = (uint32_t)((k * 32) >> 4);
                        // shift (0, 0): every phase; start one pixel row into the buffer (the buffer begins at ox0 - 1)
                        umma_bf16_split(d, a_cur + (ROWB >> 4) + ko, desc_hi, w_lo + ko, desc_hi, idesc(4 * CO), k != 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < CI / 16; ++k) {
                        const uint32_t ko = (uint32_t)((k * 32) >> 4);
                        umma_bf16_split(d + CO, a_cur + ko, desc_hi, w_lo + (uint32_t)((4 * kWTile) >> 4) + ko, desc_hi, idesc(2 * CO), 1u);          // (0,-1): (0,0),(1,0)
                        umma_bf16_split(d, a_up + (ROWB >> 4) + ko, desc_hi, w_lo + (uint32_t)((6 * kWTile) >> 4) + ko, desc_hi, idesc(2 * CO), 1u);  // (-1,0): (0,1),(0,0)
                        umma_bf16_split(d + CO, a_up + ko, desc_hi, w_lo + (uint32_t)((8 * kWTile) >> 4) + ko, desc_hi, idesc(CO), 1u);              // (-1,-1): (0,0)
                    }
                    umma_commit(&empty[r0.slot]);                 // input row oy - 1 has no later consumer
                    if (j == nrows - 1) umma_commit(&empty[r1.slot]);
                    umma_commit(&tfull[acc]);
                    if (++acc == kNAcc) { acc = 0; acc_phase ^= 1; }
                    base.advance();
                }
                base.advance();
            }
        }
    } else {
        // ===================================================== epilogue: two groups of four warps, alternate position rows; a thread owns one
        // position = a 2 x 2 output pixel block; raw accumulators -> bf16 -> swizzled staging -> two bulk stores (rows 2 oy, 2 oy + 1)
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int m = quad * 32 + lane;
        const bool issuer = ((int)threadIdx.x - 64 - grp * 128) == 0;
        const int bar_id = 1 + grp;
        uint8_t *stage = sO + grp * 2 * kOutRow;
        uint32_t nrow0 = 0;
        for (int strip = blockIdx.x; strip < p.total_strips; strip += gridDim.x) {
            int b, ya, x0, nrows;
            strip_coords(strip, b, ya, x0, nrows);
            const int j0 = (int)((nrow0 ^ (uint32_t)grp) & 1u);
            for (int j = j0; j < nrows; j += 2) {
                const uint32_t n = nrow0 + (uint32_t)j;
                const int acc = (int)(n % kNAcc);
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous stores of this group have read the staging
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                mbar_wait(&tfull[acc], (n / kNAcc) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 4 * CO);
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) {                  // column blocks in the order (0,1),(0,0),(1,0),(1,1)
                    uint32_t r[32];
                    tmem_ld32(taddr + ph * CO, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const int py = ph >> 1, px = (ph == 0 || ph == 3) ? 1 : 0;
                    const int xr = 2 * m + px;                        // pixel row of the staged [256 px][CO] tile
                    uint8_t *dst = stage + py * kOutRow + xr * (CO * 2);
                    const uint32_t swz = (uint32_t)((xr >> 1) & 3);
#pragma unroll
                    for (int q = 0; q < CO / 8; ++q)
                        *reinterpret_cast<uint4 *>(dst + (((uint32_t)q ^ swz) * 16)) =
                            make_uint4(pack_bf16x2(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])), pack_bf16x2(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                                       pack_bf16x2(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])), pack_bf16x2(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (issuer) {
                    const int oy = ya + j;
#pragma unroll
                    for (int py = 0; py < 2; ++py)
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"(&tmY), "r"(smem_u32(stage + py * kOutRow)), "r"(0), "r"(2 * x0), "r"(2 * oy + py), "r"(b) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            nrow0 += (uint32_t)nrows;
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kNAcc * 4 * CO) : "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace trows

// Interior (oy < h, ox < w) of the stride-2 transposed convolution; *handled = 0 when the configuration is outside this kernel.
int convt_rows_interior(const ood_conv3x3_args &a, cudaStream_t st, int *handled) {
    using namespace trows;
    *handled = 0;
    static int enabled = -1;
    if (enabled < 0) { const char *e = getenv("OOD_CONVT_ROWS"); enabled = (e && e[0] == '0') ? 0 : 1; }
    if (!enabled || a.transposed != 1 || a.dtype != OOD_BF16 || a.out_dtype == OOD_F16 || a.out_f32 || !a.out_y || a.cin != CI || a.cout != CO || a.groups > 1) return OOD_OK;
    if (a.w % 128 != 0 || a.h < 2 || (int64_t)a.batch * (2 * a.h + 1) * (2 * a.w + 1) >= (1LL << 31)) return OOD_OK;
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return OOD_OK;
        encode = (EncodeFn)ptr;
    }
    CUtensorMap tmA, tmB, tmY;
    {
        cuuint64_t dims[4] = {(cuuint64_t)CI, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.batch};
        cuuint64_t strides[3] = {(cuuint64_t)CI * 2, (cuuint64_t)a.w * CI * 2, (cuuint64_t)a.h * a.w * CI * 2};
        cuuint32_t box[4] = {(cuuint32_t)CI, (cuuint32_t)kRowPx, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(a.in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return OOD_OK;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)CI, (cuuint64_t)CO, 9};
        cuuint64_t strides[2] = {(cuuint64_t)CI * 2, (cuuint64_t)CO * CI * 2};
        cuuint32_t box[3] = {(cuuint32_t)CI, (cuuint32_t)CO, 1};
        cuuint32_t es[3] = {1, 1, 1};
        if (encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(a.weight), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return OOD_OK;
    }
    {
        const int OH = 2 * a.h + 1, OW = 2 * a.w + 1;
        cuuint64_t dims[4] = {(cuuint64_t)CO, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)a.batch};
        cuuint64_t strides[3] = {(cuuint64_t)CO * 2, (cuuint64_t)OW * CO * 2, (cuuint64_t)OH * OW * CO * 2};
        cuuint32_t box[4] = {(cuuint32_t)CO, 256, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        if (encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.out_y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return OOD_OK;
    }
    Params p{};
    p.batch = a.batch; p.h = a.h; p.w = a.w;
    p.tiles_x = a.w / 128;
    p.strips_y = ceil_div(a.h, kStripRows);
    const int64_t total = (int64_t)p.tiles_x * p.strips_y * a.batch;
    {   // small problems: the generic tiles fill the GPU better (OOD_ROWS_MIN_STRIPS overrides, as for conv_rows: the parity tests use it)
        const char *e = getenv("OOD_ROWS_MIN_STRIPS");
        const int64_t min_strips = e ? atoll(e) : kNumSMs;
        if (total >= (1LL << 31) || total < min_strips) return OOD_OK;
    }
    p.total_strips = (int)total;
    static DeviceOnce attr;
    if (attr.first()) {
        cudaError_t e = cudaFuncSetAttribute(convt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) { set_error("conv3x3 transposed rows: smem attribute: %s", cudaGetErrorString(e)); return OOD_ERR_CUDA; }
    }
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *handled = 1;
    convt_rows_kernel<<<std::min(p.total_strips, sms), kThreads, kSmem, st>>>(tmA, tmB, tmY, p);
    return check_launch("conv3x3 transposed rows");
}

}  // namespace ood
