// 3x3 (transposed) convolution as an implicit GEMM on the 5th-generation tensor cores (sm_100a):
//   tcgen05.mma (bf16 x bf16 -> fp32) issued by one thread, accumulators in TMEM (double buffered),
//   operands staged by TMA into 128B/64B-swizzled shared memory through an mbarrier ring,
//   persistent CTAs (one per SM) with warp roles: TMA producer / MMA issuer / 4 epilogue warps.
//
// Reference op: ModulatedConv2d.forward (src/ops/StyleGAN/model.py:233-274) -- there a cuDNN grouped conv over
// B x Co x Ci x 3 x 3 materialised weights.  Here: M = output pixels of a phase (128-pixel rectangular patches so
// that one TMA box [NB,TH,TW,BK] per tap IS the im2col tile; out-of-bounds box elements are zero-filled by TMA,
// which implements the padding), N = Co, K = taps x Ci; weights are shared ([tap][Co][Ci] bf16, K-major), the
// style modulation rides on the activations and demodulation / noise / bias / leaky-ReLU / next-layer style are
// applied in the TMEM->register epilogue.
#include <cuda.h>

#include "conv_common.cuh"

namespace ood {

constexpr int TBM = 128;            // UMMA M
// threads: warp0 TMA, warp1 MMA (+TMEM alloc), then EPW epilogue warps (template parameter of the kernel)

struct TcPhase {
    int oh, ow, oy0, ox0, py, px, ntaps;
    int dy[9], dx[9], wt[9];
    int tiles_x, tiles_y, tiles_b;
    int tile_begin;                 // first global tile id of this phase
};

struct TcParams {
    int batch, h, w, cin, cout, OH, OW, sy, sx;
    int isy, isx;                   // input stride (2 for the strided data-gradient form; the TMA map then carries elementStrides 2)
    int TW, TH, NB;                 // pixel patch of one M tile: NB*TH*TW == 128
    int nphases, n_tiles_n, total_tiles;
    int wtaps;                      // taps per group in the weight pack (9, or 1 for the 1x1 form)
    int fused;                      // transposed == 5: N = 4*cout phase-major columns, outputs scattered to 2x2 pixel blocks
    int groups, gbatch, in_shared;  // grouped form: output image g*gbatch + i uses weights [g*9 + tap] and input image i (shared) or g*gbatch + i
    TcPhase ph[4];
    ConvEpilogue ep;
    int out_bf16;                   // storage type of out_y / out_ys
    int in_f16, out_f16;            // OOD_F16 operands (idesc A/B format) / half-precision outputs (epilogue pack)
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one lane of a converged warp (elect.sync): ptxas knows that code under this predicate runs on a single thread and keeps the
// operands of tcgen05.mma / TMA instructions in uniform registers (with `lane == 0` every MMA cost an ELECT + R2UR.BROADCAST
// sequence of dependent fixed-latency instructions: ~200 cycles per instruction on the row-sliding kernel, ncu round 2)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *holder_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// ---- CTA pairs (cta_group::2): two CTAs of a cluster run ONE 256-row MMA; each holds its own 128 rows of A and HALF of the B tile
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in the CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the bytes land in THIS CTA's shared memory, the transaction count on the LEADER's barrier (bar_cluster)
__device__ __forceinline__ void tma_load_4d_pair(void *dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void *dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *holder_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (when the MMAs issued so far retire) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, rows of SWZ bytes (SWZ = 128 or 64), 8-row swizzle atoms stacked contiguously.
template <int SWZ>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);                 // start address
    d |= (uint64_t)1 << 16;                                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * SWZ) >> 4) << 32;                  // stride byte offset: one 8-row atom
    d |= (uint64_t)1 << 46;                                 // descriptor version (sm_100)
    d |= (uint64_t)(SWZ == 128 ? 2 : 4) << 61;              // SWIZZLE_128B / SWIZZLE_64B
    return d;
}

template <int BN, int BK, bool CTA2 = false, bool TP = false>      // TP: the STATS kernels of the pair form keep a [4 warps][32 px][36] fp32 transpose buffer
struct TcCfg {
    static constexpr int kABytes = TBM * BK * 2;
    static constexpr int kBBytes = (CTA2 ? BN / 2 : BN) * BK * 2;          // CTA pairs: a CTA holds half of the N tile's weight rows (6 stages instead of 4 at BN = 256)
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTpBytes = TP ? 4 * 32 * 36 * 4 : 0;
    static constexpr int kStages = ((200 * 1024 - kTpBytes) / kStageBytes) > 8 ? 8 : ((200 * 1024 - kTpBytes) / kStageBytes);
    static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int kEpiVecs = 7;                       // bias, prelu, d, s_next, rgb_w[3]: BN floats each, double buffered
    static constexpr int kEpiBytes = 2 * kEpiVecs * BN * 4;
    static constexpr int kStatBytes = 4 * BN * 2 * 4 + kTpBytes;        // STATS kernels: [epilogue warp][channel][sum, sum of squares] (+ transpose buffer)
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiBytes;
    static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
};

struct TileCoord {
    int phase, b0, y0, x0, n0;
    int group, bl0;                 // group of this tile and its first image inside the group (b0 = group*gbatch + bl0)
};

__device__ __forceinline__ TileCoord decode_tile(const TcParams &p, int tile, int BN) {
    TileCoord tc;
    int ph = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < p.nphases && tile >= p.ph[i].tile_begin) ph = i;
    const TcPhase &P = p.ph[ph];
    const int local = tile - P.tile_begin;
    const int nt = local % p.n_tiles_n, mt = local / p.n_tiles_n;
    const int tx = mt % P.tiles_x, ty = (mt / P.tiles_x) % P.tiles_y, tb = mt / (P.tiles_x * P.tiles_y);
    const int tiles_bg = P.tiles_b / p.groups;          // M tiles never straddle a group
    tc.group = tb / tiles_bg; tc.bl0 = (tb - tc.group * tiles_bg) * p.NB;
    tc.phase = ph; tc.b0 = tc.group * p.gbatch + tc.bl0; tc.y0 = P.oy0 + ty * p.TH; tc.x0 = P.ox0 + tx * p.TW; tc.n0 = nt * BN;
    return tc;
}

// SEED: fp32 accumulator seed (ConvEpilogue::acc_in).  STATS: per-tile channel moments of the stored output
// (ConvEpilogue::stat_partial) -- the statistics of the InstanceNorm that follows the convolution, without a pass over it.
// EPW: epilogue warps, 4 or 8.  A warp reads the TMEM lane quadrant warp % 4; with 8 warps two warps share a quadrant and
// take alternate 32-column chunks -- for tiles whose K is short (the fused-phase transposed form, K = 4 shifts) the epilogue,
// not the MMA, sets the tile period.
// CTA2: the CTAs 2c, 2c+1 form a cluster and work on the M tiles 2m, 2m+1 of one N tile as ONE tcgen05.mma.cta_group::2 of 256 rows:
// every CTA loads its own A tile and half of the B tile (half the weight bytes per SM from L2 and from shared memory), CTA 0 of
// the pair issues the MMAs for both, each CTA's TMEM receives its own 128 rows and its epilogue warps drain them as before.
template <int BN, int BK, bool SEED = false, bool STATS = false, int EPW = 4, bool CTA2 = false>
__global__ void __launch_bounds__(64 + 32 * EPW, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    using Cfg = TcCfg<BN, BK, CTA2, STATS && CTA2>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + S * Cfg::kABytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S * Cfg::kStageBytes);
    uint64_t *full = bars, *empty = bars + S, *tfull = bars + 2 * S, *tempty = bars + 2 * S + 2;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);
    float *epi_vecs = reinterpret_cast<float *>(smem + S * Cfg::kStageBytes + 256);
    float *stat_red = epi_vecs + 2 * Cfg::kEpiVecs * BN;     // STATS kernels only (the launch adds kStatBytes)
    float *stat_tp = stat_red + 4 * BN * 2;                  // STATS kernels of the pair form: per-warp transpose buffer

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = p.cin / BK;

    const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
    // work items of this CTA: tiles, or (CTA2) pairs of M-adjacent tiles of which this CTA takes the one of its rank
    const int n_items = CTA2 ? p.total_tiles >> 1 : p.total_tiles;
    const int item0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, item_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto item_tile = [&](int q) -> int {
        if (!CTA2) return q;
        int ph = 0;                                     // every phase has an even number of M tiles (host check): pair q of phase ph starts at tile_begin / 2
#pragma unroll
        for (int i = 1; i < 4; ++i)
            if (i < p.nphases && 2 * q >= p.ph[i].tile_begin) ph = i;
        const int lq = q - (p.ph[ph].tile_begin >> 1);
        return p.ph[ph].tile_begin + (2 * (lq / p.n_tiles_n) + (int)cta_rank) * p.n_tiles_n + lq % p.n_tiles_n;
    };

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CTA2 ? 2 * EPW : EPW); }      // pair: both CTAs' epilogue warps arrive on the leader's
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) { if (CTA2) tmem_alloc_pair(tmem_holder, Cfg::kTmemCols); else tmem_alloc(tmem_holder, Cfg::kTmemCols); }
    tc_fence_before();
    if (CTA2) cluster_sync_all(); else __syncthreads();       // pair: the peer's barriers must exist before anything is signalled across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int q = item0; q < n_items; q += item_step) {
                const int tile = item_tile(q);
                const TileCoord tc = decode_tile(p, tile, BN);
                const TcPhase &P = p.ph[tc.phase];
                for (int t = 0; t < P.ntaps; ++t) {
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        if constexpr (CTA2) {
                            // both CTAs' bytes are counted on the leader's barrier (its producer announces the sum)
                            if (cta_rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
                            const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
                            tma_load_4d_pair(sA + stage * Cfg::kABytes, &tmA, lbar, kc * BK, tc.x0 * p.isx + P.dx[t], tc.y0 * p.isy + P.dy[t], tc.b0);
                            tma_load_3d_pair(sB + stage * Cfg::kBBytes, &tmB, lbar, kc * BK, tc.n0 + (int)cta_rank * (BN / 2), P.wt[t]);
                        } else {
                            mbar_expect_tx(&full[stage], Cfg::kStageBytes);
                            tma_load_4d(sA + stage * Cfg::kABytes, &tmA, &full[stage], kc * BK, tc.x0 * p.isx + P.dx[t], tc.y0 * p.isy + P.dy[t],
                                        p.in_shared ? tc.bl0 : tc.b0);
                            tma_load_3d(sB + stage * Cfg::kBBytes, &tmB, &full[stage], kc * BK, tc.n0, P.wt[t] + tc.group * p.wtaps);
                        }
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if ((!CTA2 || cta_rank == 0) && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            // instruction descriptor: A / B format field 1 = bf16, 0 = f16 (bits 7..9 and 10..12); pair: M = 256
            constexpr uint32_t kIdescM = CTA2 ? ((Cfg::kIdesc & ~(0x1Fu << 24)) | ((uint32_t)(2 * TBM >> 4) << 24)) : Cfg::kIdesc;
            const uint32_t idesc = p.in_f16 ? (kIdescM & ~((1u << 7) | (1u << 10))) : kIdescM;
            for (int q = item0; q < n_items; q += item_step) {
                const int tile = item_tile(q);
                const TileCoord tc = decode_tile(p, tile, BN);
                const int kiters = p.ph[tc.phase].ntaps * kchunks;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = make_smem_desc<BK * 2>(smem_u32(sA + stage * Cfg::kABytes));
                    const uint64_t db = make_smem_desc<BK * 2>(smem_u32(sB + stage * Cfg::kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if constexpr (CTA2) umma_bf16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
                        else umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
                    }
                    if constexpr (CTA2) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);          // frees the smem slot (of both CTAs) when these MMAs retire
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                if constexpr (CTA2) umma_commit_pair(&tfull[acc]); else umma_commit(&tfull[acc]);                // accumulator complete -> epilogue (of both CTAs)
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue warps (TMEM lane quadrant = warp % 4)
        const int quad = warp & 3;
        const int chalf = (warp - 2) >> 2;      // EPW == 8: which of the two warps of this quadrant (chunk parity it owns)
        const int row = quad * 32 + lane;
        const int nb = row / (p.TH * p.TW), ty = (row / p.TW) % p.TH, tx = row % p.TW;
        const float nw = (p.ep.noise && p.ep.noise_w) ? *p.ep.noise_w : 0.f;
        int acc = 0;
        uint32_t acc_phase = 0;
        int skn0 = -2, skn1 = -2, skb0 = -2, skb1 = -2;      // what the two staged epilogue-vector buffers hold: (group, N tile), image
        for (int q = item0; q < n_items; q += item_step) {
            const int tile = item_tile(q);
            const TileCoord tc = decode_tile(p, tile, BN);
            const TcPhase &P = p.ph[tc.phase];
            const int b = tc.b0 + nb, oy = tc.y0 + ty, ox = tc.x0 + tx;
            const bool valid = tc.bl0 + nb < p.gbatch && oy < P.oy0 + P.oh && ox < P.ox0 + P.ow;
            const int gofs = tc.group * p.cout;           // per-group bias / PReLU slopes: [groups][Co]
            const int Y = oy * p.sy + P.py, X = ox * p.sx + P.px;
            const int64_t pix = ((int64_t)b * p.OH + Y) * p.OW + X;
            float nz = 0.f;
            if (valid && p.ep.noise) nz = nw * __ldg(p.ep.noise + b * p.ep.noise_bstride + (int64_t)Y * p.OW + X);
            // Per-channel epilogue vectors of this tile staged in shared memory while the MMAs of the tile are still running:
            // read from global memory inside the chunk loop they sat behind the TMEM load (an asm barrier for the compiler)
            // and exposed an L2 round trip per vector per chunk (ncu: the top stall of the epilogue warps).  The per-sample
            // vectors (d, s_next, rgb_w) are tile-uniform only when a tile holds one image (NB == 1); otherwise they stay global.
            float *sv = epi_vecs + acc * (Cfg::kEpiVecs * BN);
            const bool stg = p.NB == 1;
            // ... and only when they differ from what this buffer already holds: short-K launches (the 1x1 convolutions: one k-step per tile) have no
            // MMA time to hide the loads and the barrier behind, and their vectors (bias / PReLU, or one image's d / s_next) rarely change between tiles
            const int key_b = (stg && (p.ep.d || p.ep.out_ys || p.ep.rgb_out)) ? tc.b0 : -1;
            const int key_n = tc.n0 + gofs * 4;              // n0 < cout <= gofs step: distinct per (group, N tile)
            const bool restage = (acc ? skn1 : skn0) != key_n || (acc ? skb1 : skb0) != key_b;
            if (restage) {
                (acc ? skn1 : skn0) = key_n; (acc ? skb1 : skb0) = key_b;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPW) : "memory");      // every epilogue warp is done reading the buffer's previous contents
                const int et = threadIdx.x - 64;
                for (int i = et; i < BN; i += 32 * EPW) {
                    const int n = tc.n0 + i;
                    if (p.ep.bias) sv[i] = __ldg(p.ep.bias + gofs + n);
                    if (p.ep.act == 2) sv[BN + i] = __ldg(p.ep.prelu + gofs + n);
                    if (stg) {
                        if (p.ep.d) sv[2 * BN + i] = __ldg(p.ep.d + (int64_t)tc.b0 * p.cout + n);
                        if (p.ep.out_ys) sv[3 * BN + i] = __ldg(p.ep.s_next + (int64_t)tc.b0 * p.cout + n);
                        if (p.ep.rgb_out) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) sv[(4 + k) * BN + i] = __ldg(p.ep.rgb_w + ((int64_t)tc.b0 * 3 + k) * p.cout + n);
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPW) : "memory");      // the epilogue warps only
            }
            auto vec4 = [&](int kind, const float *gvec, int ch, int j) -> float4 {      // float4 j of chunk ch of a per-sample vector
                return stg ? *reinterpret_cast<const float4 *>(sv + kind * BN + ch * 32 + 4 * j) : __ldg(reinterpret_cast<const float4 *>(gvec) + j);
            };
            float rgb_tail[3] = {0.f, 0.f, 0.f};      // fused ToRGB: bias + upsampled skip, independent of the accumulator
            if (p.ep.rgb_out && valid) {
#pragma unroll
                for (int k = 0; k < 3; ++k) rgb_tail[k] = rgb_finish(p.ep, 0.f, b, k, Y, X, p.OH, p.OW);
            }
            // accumulator seed: this pixel's 32-channel run of the fp32 partial, fetched one chunk ahead of the TMEM read so
            // that its latency hides behind the previous chunk's stores (and, for chunk 0, behind the wait for the MMAs)
            // NHWC: this pixel's channel run, float4 j of chunk c at seed[c*8 + j]; tile order: at seed[(c*8 + j)*128] (lanes contiguous)
            const int sstep = p.ep.tiled ? TBM : 1;
            const float4 *seed = nullptr;
            if (SEED && p.ep.acc_in && valid)
                seed = p.ep.tiled ? reinterpret_cast<const float4 *>(p.ep.acc_in) + (int64_t)tile * (BN / 32) * 8 * TBM + row
                                  : reinterpret_cast<const float4 *>(p.ep.acc_in + pix * p.cout + tc.n0);
            if (SEED && p.ep.acc_in && p.ep.tiled && threadIdx.x == 64) {
                // tile order: the seed of a tile is one contiguous block -- ask L2 for the NEXT tile's block now (a tile period
                // ahead), and for this CTA's first tile at its start; the register prefetch below then only covers an L2 hit
                const int64_t blk = (int64_t)BN * TBM * sizeof(float);
                const char *base = reinterpret_cast<const char *>(p.ep.acc_in);
                if (q == item0)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (int64_t)tile * blk), "r"((uint32_t)blk) : "memory");
                if (q + item_step < n_items)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (int64_t)item_tile(q + item_step) * blk), "r"((uint32_t)blk) : "memory");
            }
            float4 sd[8], sdn[8] = {};
            if (SEED && seed) {
#pragma unroll
                for (int j = 0; j < 8; ++j) sd[j] = __ldg(seed + j * sstep);
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
            float rgbp[3] = {0.f, 0.f, 0.f};
            // one 32-channel chunk; `cur` holds this chunk's seed values, the next chunk's are requested into `nxt` first so that
            // their latency spans the whole chunk (the two arrays swap roles from chunk to chunk: no register copies, which
            // would wait for the loads)
            // (Measured and not kept: requesting the next chunk's TMEM load before processing the current one -- a second 32-register
            // buffer in the plain kernels -- left the half-K convolutions at 384 / 406 us and made the C=128 ones 5-9 % slower.)
            const bool of16 = p.out_f16 != 0;
            auto pk2 = [&](float lo, float hi) -> uint32_t { return of16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); };
            auto do_chunk = [&](const int ch, float4 (&cur)[8], float4 (&nxt)[8]) {
                uint32_t r[32];
                if (SEED && seed && ch + 1 < BN / 32) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) nxt[j] = __ldg(seed + ((ch + 1) * 8 + j) * sstep);
                }
                tmem_ld32(taddr + ch * 32, r);
                tmem_ld_wait();
                float v[32];
                int n = tc.n0 + ch * 32;
                bool cvalid = valid;
                int64_t cpix = pix;
                if (p.fused) {          // column block -> (output parity phase, channel); position (oy, ox) -> pixel (2oy+fy, 2ox+fx)
                    const int phs = n / p.cout, fy = phs >> 1, fx = phs & 1;
                    n -= phs * p.cout;
                    cvalid = valid && oy < p.h + 1 - fy && ox < p.w + 1 - fx;
                    cpix = ((int64_t)b * p.OH + 2 * oy + fy) * p.OW + 2 * ox + fx;
                }
                if (cvalid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (SEED && seed) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[4 * j] += cur[j].x; v[4 * j + 1] += cur[j].y; v[4 * j + 2] += cur[j].z; v[4 * j + 3] += cur[j].w;
                        }
                    }
                    if (p.ep.d) {
                        const float *dp = p.ep.d + (int64_t)b * p.cout + n;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t = vec4(2, dp, ch, j);
                            v[4 * j] *= t.x; v[4 * j + 1] *= t.y; v[4 * j + 2] *= t.z; v[4 * j + 3] *= t.w;
                        }
                    }
                    if (p.ep.bias) {
                        const float4 *bp = reinterpret_cast<const float4 *>(sv + ch * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t = bp[j];
                            v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                        }
                    }
                    if (p.ep.act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = lrelu_sqrt2(v[j] + nz);
                    } else if (p.ep.act == 2) {
                        const float4 *pp = reinterpret_cast<const float4 *>(sv + BN + ch * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t = pp[j];
                            v[4 * j] = apply_act(v[4 * j] + nz, 2, t.x); v[4 * j + 1] = apply_act(v[4 * j + 1] + nz, 2, t.y);
                            v[4 * j + 2] = apply_act(v[4 * j + 2] + nz, 2, t.z); v[4 * j + 3] = apply_act(v[4 * j + 3] + nz, 2, t.w);
                        }
                    } else if (p.ep.noise) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += nz;
                    }
                    if (p.ep.rgb_out) {      // fused ToRGB: 3 dot products over this chunk's 32 channels of the unscaled activation
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const float *wp = p.ep.rgb_w + ((int64_t)b * 3 + k) * p.cout + n;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 t = vec4(4 + k, wp, ch, j);
                                rgbp[k] = fmaf(v[4 * j], t.x, fmaf(v[4 * j + 1], t.y, fmaf(v[4 * j + 2], t.z, fmaf(v[4 * j + 3], t.w, rgbp[k]))));
                            }
                        }
                    }
                    if (p.ep.out_y) {
                        if (p.ep.out_f32) {
                            const int ostep = p.ep.tiled ? TBM : 1;
                            float4 *o = p.ep.tiled ? reinterpret_cast<float4 *>(p.ep.out_y) + ((int64_t)tile * (BN / 32) + ch) * 8 * TBM + row
                                                   : reinterpret_cast<float4 *>((float *)p.ep.out_y + cpix * p.cout + n);
                            if (p.ep.tiled) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) o[j * ostep] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            } else {        // NHWC: a 128-byte run per thread as four 32-byte stores
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    st_global_256(o + 2 * j, __float_as_uint(v[8 * j]), __float_as_uint(v[8 * j + 1]), __float_as_uint(v[8 * j + 2]),
                                                  __float_as_uint(v[8 * j + 3]), __float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]),
                                                  __float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7]));
                            }
                        } else {
                            __nv_bfloat16 *o = (__nv_bfloat16 *)p.ep.out_y + cpix * p.cout + n;      // 64-byte aligned (cout, n % 32 == 0)
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                st_global_256(o + 16 * j, pk2(v[16 * j], v[16 * j + 1]), pk2(v[16 * j + 2], v[16 * j + 3]),
                                              pk2(v[16 * j + 4], v[16 * j + 5]), pk2(v[16 * j + 6], v[16 * j + 7]),
                                              pk2(v[16 * j + 8], v[16 * j + 9]), pk2(v[16 * j + 10], v[16 * j + 11]),
                                              pk2(v[16 * j + 12], v[16 * j + 13]), pk2(v[16 * j + 14], v[16 * j + 15]));
                        }
                    }
                    if (p.ep.out_ys) {
                        const float *sp = p.ep.s_next + (int64_t)b * p.cout + n;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t = vec4(3, sp, ch, j);
                            v[4 * j] *= t.x; v[4 * j + 1] *= t.y; v[4 * j + 2] *= t.z; v[4 * j + 3] *= t.w;
                        }
                        __nv_bfloat16 *o = (__nv_bfloat16 *)p.ep.out_ys + cpix * p.cout + n;
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            st_global_256(o + 16 * j, pk2(v[16 * j], v[16 * j + 1]), pk2(v[16 * j + 2], v[16 * j + 3]),
                                          pk2(v[16 * j + 4], v[16 * j + 5]), pk2(v[16 * j + 6], v[16 * j + 7]),
                                          pk2(v[16 * j + 8], v[16 * j + 9]), pk2(v[16 * j + 10], v[16 * j + 11]),
                                          pk2(v[16 * j + 12], v[16 * j + 13]), pk2(v[16 * j + 14], v[16 * j + 15]));
                    }
                }
                if constexpr (STATS && CTA2) {
                    // moments of this warp's 32 pixels for the chunk's 32 channels, of the values as stored: transposed through shared memory
                    // (lane = pixel writes its 32 channels, pitch 36 floats: conflict-free 16-byte stores and 4-byte column reads; lane =
                    // channel then sums its column in pixel order).  A third of the instructions of the shuffle transpose below, which at
                    // 256 channels / 256 px made the epilogue, not the MMAs, set the tile period (1061 vs 802 us for the plain kernel).
                    float *tp = stat_tp + quad * (32 * 36);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float2 a = !valid ? make_float2(0.f, 0.f) : of16 ? unpack_f16x2(pack_f16x2(v[j], v[j + 1])) : unpack_bf16x2(pack_bf16x2(v[j], v[j + 1]));
                        const float2 c = !valid ? make_float2(0.f, 0.f) : of16 ? unpack_f16x2(pack_f16x2(v[j + 2], v[j + 3])) : unpack_bf16x2(pack_bf16x2(v[j + 2], v[j + 3]));
                        *reinterpret_cast<float4 *>(tp + lane * 36 + j) = make_float4(a.x, a.y, c.x, c.y);
                    }
                    __syncwarp();
                    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
                    for (int pp = 0; pp < 32; pp += 2) {
                        const float t0 = tp[pp * 36 + lane], t1 = tp[(pp + 1) * 36 + lane];
                        s1a += t0; s1b += t1;
                        s2a = fmaf(t0, t0, s2a); s2b = fmaf(t1, t1, s2b);
                    }
                    __syncwarp();
                    *reinterpret_cast<float2 *>(stat_red + ((quad * BN) + ch * 32 + lane) * 2) = make_float2(s1a + s1b, p.ep.stat_sums_only ? 0.f : s2a + s2b);
                } else if constexpr (STATS) {
                    // moments of this warp's 32 pixels for the chunk's 32 channels, of the values as stored (bf16 / f16): a
                    // transpose-reduce over the lanes (31 shuffles per moment) leaves channel `lane` in element 0
                    float q[32];
                    auto lane_transpose_sum = [&]() {
#pragma unroll
                        for (int o = 16; o >= 1; o >>= 1) {
                            const bool up = lane & o;
#pragma unroll
                            for (int i = 0; i < o; ++i) {
                                const float s1 = up ? q[i] : q[i + o], k1 = up ? q[i + o] : q[i];
                                q[i] = k1 + __shfl_xor_sync(0xffffffffu, s1, o);
                            }
                        }
                    };
                    float sum1, sum2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float2 t = !valid ? make_float2(0.f, 0.f) : of16 ? unpack_f16x2(pack_f16x2(v[j], v[j + 1])) : unpack_bf16x2(pack_bf16x2(v[j], v[j + 1]));
                        q[j] = t.x; q[j + 1] = t.y;
                    }
                    if (!p.ep.stat_sums_only) {          // second moment first (q is rebuilt from the same rounded values afterwards)
#pragma unroll
                        for (int j = 0; j < 32; ++j) q[j] *= q[j];
                        lane_transpose_sum();
                        sum2 = q[0];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float2 t = !valid ? make_float2(0.f, 0.f) : of16 ? unpack_f16x2(pack_f16x2(v[j], v[j + 1])) : unpack_bf16x2(pack_bf16x2(v[j], v[j + 1]));
                            q[j] = t.x; q[j + 1] = t.y;
                        }
                    }
                    lane_transpose_sum();
                    sum1 = q[0];
                    *reinterpret_cast<float2 *>(stat_red + ((quad * BN) + ch * 32 + lane) * 2) = make_float2(sum1, sum2);
                }
            };
            if constexpr (SEED) {
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ch += 2) { do_chunk(ch, sd, sdn); do_chunk(ch + 1, sdn, sd); }
            } else if constexpr (EPW == 8) {
#pragma unroll 1
                for (int ch = chalf; ch < BN / 32; ch += 2) do_chunk(ch, sd, sdn);
            } else {
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ++ch) do_chunk(ch, sd, sdn);
            }
            if constexpr (STATS) {
                // combine the four warps in a fixed order: partial[image][m tile of the image][channel] = {sum, sum of squares}
                asm volatile("bar.sync 2, 128;" ::: "memory");
                const int mt = (tile - P.tile_begin) / p.n_tiles_n, mpi = P.tiles_x * P.tiles_y;
                float2 *dst = reinterpret_cast<float2 *>(p.ep.stat_partial) + ((int64_t)tc.b0 * mpi + mt % mpi) * p.cout + tc.n0;
                for (int i = threadIdx.x - 64; i < BN; i += 128) {
                    float2 a = *reinterpret_cast<const float2 *>(stat_red + i * 2);
#pragma unroll
                    for (int w4 = 1; w4 < 4; ++w4) {
                        const float2 t = *reinterpret_cast<const float2 *>(stat_red + (w4 * BN + i) * 2);
                        a.x += t.x; a.y += t.y;
                    }
                    dst[i] = a;
                }
            }
            if (p.ep.rgb_out && valid) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    p.ep.rgb_out[((int64_t)b * 3 + k) * p.OH * p.OW + (int64_t)Y * p.OW + X] = rgbp[k] + rgb_tail[k];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));      // the leader's MMA thread waits for both CTAs' drains
                else mbar_arrive(&tempty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    if (CTA2) cluster_sync_all(); else __syncthreads();       // pair: no CTA may exit while its peer can still signal into it
    if (warp == 1) {
        tc_fence_after();
        if (CTA2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static int pow2_ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }

template <int BN, int BK, bool SEED = false, bool STATS = false, int EPW = 4, bool CTA2 = false>
static int launch_tc(const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &p, cudaStream_t st) {
    using Cfg = TcCfg<BN, BK, CTA2, STATS && CTA2>;
    static_assert(EPW == 4 || (EPW == 8 && !SEED && !STATS), "8 epilogue warps: plain kernels only");
    static_assert(!CTA2 || (EPW == 4 && BK == 64 && BN >= 128), "CTA pairs: wide tiles, four epilogue warps");
    auto kern = conv_tc_kernel<BN, BK, SEED, STATS, EPW, CTA2>;
    constexpr int kSmem = Cfg::kSmemBytes + (STATS ? Cfg::kStatBytes : 0);
    static_assert(kSmem <= 227 * 1024, "shared memory budget");
    static DeviceOnce attr_set;
    if (attr_set.first()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) { set_error("conv3x3 tc: smem attribute: %s", cudaGetErrorString(e)); return OOD_ERR_CUDA; }
    }
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if constexpr (CTA2) {
        const int grid = std::min(p.total_tiles, sms) & ~1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * EPW); cfg.dynamicSmemBytes = kSmem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
        if (e != cudaSuccess) { set_error("conv3x3 tc: CTA-pair launch: %s", cudaGetErrorString(e)); return OOD_ERR_CUDA; }
        return check_launch("conv3x3 tc pair");
    }
    const int grid = std::min(p.total_tiles, sms);
    kern<<<grid, 64 + 32 * EPW, kSmem, st>>>(tmA, tmB, p);
    return check_launch("conv3x3 tc");
}

void in_finalize_launch(const float *partial, float *st2, int64_t P, int C, int nchunks, float eps, int batch, cudaStream_t s);   // alignnet.cu

// Narrower N tiles are tried while a launch has fewer tiles than this.  100, not one per SM: at 100..147 tiles one wave of wide
// tiles on part of the SMs beats two waves of half-width tiles, which also need 1.33x the L2 -> SM operand traffic per FLOP
// (encoder 256 -> 256 at 32 px, 128 tiles: 0.83 -> 0.74 ms per 26 launches; OOD_MIN_TILES overrides).
static int min_tiles() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("OOD_MIN_TILES"); v = e ? atoi(e) : 100; if (v <= 0) v = 100; }
    return v;
}

static bool epw8() {        // experiment switch: OOD_EPW8=0 falls back to four epilogue warps
    static int v = -1;
    if (v < 0) { const char *e = getenv("OOD_EPW8"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// Tile plan of a launch: patch shape, N tile, tile enumeration.  A function of (batch, h, w, cout, form, groups) and of whether
// the call uses seeded / tile-order tensors (those keep the 128- or 256-wide N tiles) -- never of the data or of cin, so two
// launches with the same arguments enumerate the same tiles (what the tile-order fp32 tensors rely on).
static void plan_tiles(const ood_conv3x3_args &a, const ConvGeom &g, TcParams &p, int &BN) {
    const int groups = a.groups > 1 ? a.groups : 1;
    const int ncols = a.transposed == 5 ? 4 * a.cout : a.cout;      // GEMM N
    BN = ncols % 256 == 0 ? 256 : (ncols % 128 == 0 ? 128 : (ncols % 64 == 0 ? 64 : 32));
    p.fused = a.transposed == 5;
    p.batch = g.batch; p.h = g.h; p.w = g.w; p.cin = g.cin; p.cout = g.cout; p.OH = g.OH; p.OW = g.OW; p.sy = g.sy; p.sx = g.sx;
    p.isy = g.isy; p.isx = g.isx;
    int ohm = 0, owm = 0;
    for (int i = 0; i < g.nphases; ++i) { ohm = std::max(ohm, g.ph[i].oh); owm = std::max(owm, g.ph[i].ow); }
    p.TW = std::min(pow2_ceil(owm), TBM);
    p.TH = std::min(pow2_ceil(ohm), TBM / p.TW);
    p.NB = TBM / (p.TW * p.TH);
    p.nphases = g.nphases;
    p.groups = groups; p.gbatch = a.batch / groups; p.in_shared = a.in_shared ? 1 : 0;
    p.wtaps = (a.transposed == 4 || a.transposed == 6) ? 1 : (a.transposed == 5 ? 4 : 9);
    const int bn_min = (a.acc_in || a.tiled || a.stats_out || a.stats_ws) ? 128 : 64;
    int tiles = 0;
    for (;;) {
        p.n_tiles_n = ncols / BN;
        tiles = 0;
        for (int i = 0; i < g.nphases; ++i) {
            TcPhase &P = p.ph[i];
            const ConvPhase &G = g.ph[i];
            P.oh = G.oh; P.ow = G.ow; P.oy0 = G.oy0; P.ox0 = G.ox0; P.py = G.py; P.px = G.px; P.ntaps = G.ntaps;
            for (int t = 0; t < 9; ++t) { P.dy[t] = G.dy[t]; P.dx[t] = G.dx[t]; P.wt[t] = G.wt[t]; }
            P.tiles_x = ceil_div(G.ow, p.TW); P.tiles_y = ceil_div(G.oh, p.TH); P.tiles_b = groups * ceil_div(p.gbatch, p.NB);
            P.tile_begin = tiles;
            tiles += P.tiles_x * P.tiles_y * P.tiles_b * p.n_tiles_n;
        }
        // few-pixel problems (the tails of the encoder's style heads) are bound by streaming the weights: narrower N
        // tiles put more SMs on that stream
        if (tiles >= min_tiles() || BN <= bn_min || a.rgb_out) break;
        BN >>= 1;
    }
    p.total_tiles = tiles;
}

static int conv3x3_tc_geom(const ood_conv3x3_args &a, const ConvGeom &g, cudaStream_t st);
int convt_rows_interior(const ood_conv3x3_args &a, cudaStream_t st, int *handled);      // convt_rows.cu

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int conv3x3_tc(const ood_conv3x3_args &a, cudaStream_t st) {
    // stride-2 transposed convolution of a power-of-two input: interior + last row + last column, each with exact tiles
    static int split_t = -1;
    if (split_t < 0) { const char *e = getenv("OOD_SPLIT_TRANSPOSED"); split_t = (e && e[0] == '0') ? 0 : 1; }
    const bool rows_shape = a.cin == 64 && a.cout == 32 && a.w % 128 == 0 && a.h >= 2;        // convt_rows.cu takes the interior of this layer
    if (a.transposed == 1 && split_t && ((pow2(a.h) && pow2(a.w) && a.h >= 16 && a.w >= 16) || rows_shape) && !a.acc_in && !a.tiled && !a.stats_out && !a.stats_ws && a.groups <= 1) {
        for (int part = 0; part < 3; ++part) {
            if (part == 0) {        // the 64 -> 32 layer at 1024 px: row-streaming kernel for the interior
                int handled = 0;
                const int rc = convt_rows_interior(a, st, &handled);
                if (rc != OOD_OK) return rc;
                if (handled) { g_conv_route = 2; continue; }
            }
            const int rc = conv3x3_tc_geom(a, make_geom_transposed_part(a.batch, a.h, a.w, a.cin, a.cout, part), st);
            if (rc != OOD_OK) return rc;
        }
        return OOD_OK;
    }
    return conv3x3_tc_geom(a, make_geom(a.batch, a.h, a.w, a.cin, a.cout, a.transposed), st);
}

static int conv3x3_tc_geom(const ood_conv3x3_args &a, const ConvGeom &g, cudaStream_t st) {
    OOD_REQUIRE(a.dtype == OOD_BF16 || a.dtype == OOD_F16, "conv3x3 tc: storage type must be bf16 or f16");
    OOD_REQUIRE(a.out_dtype == 0 || a.out_dtype == OOD_BF16 || a.out_dtype == OOD_F16, "conv3x3 tc: out_dtype must be 0, OOD_BF16 or OOD_F16");
    OOD_REQUIRE(a.cin % 32 == 0 && a.cout % 32 == 0, "conv3x3 tc: cin and cout must be multiples of 32 (got %d, %d)", a.cin, a.cout);
    OOD_REQUIRE(((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.weight % 16 == 0), "conv3x3 tc: operands must be 16-byte aligned");
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { set_error("conv3x3 tc: cuTensorMapEncodeTiled is unavailable"); return OOD_ERR_CUDA; }

    const int groups = a.groups > 1 ? a.groups : 1;
    OOD_REQUIRE(a.batch % groups == 0, "conv3x3 tc: batch (%d) must be a multiple of groups (%d)", a.batch, groups);
    OOD_REQUIRE(groups == 1 || (!a.d && !a.noise && !a.out_ys && !a.rgb_out), "conv3x3 tc: the grouped form supports bias / activation epilogues only");
    OOD_REQUIRE(a.transposed != 5 || (groups == 1 && !a.acc_in && !a.tiled && !a.stats_out && !a.stats_ws && !a.rgb_out),
                "conv3x3 tc: the fused-phase transposed form takes no seed / tile-order / statistics / ToRGB options");
    OOD_REQUIRE(!(a.acc_in || a.tiled) || (a.cout % 128 == 0 && a.transposed != 1),
                "conv3x3 tc: acc_in / tiled need cout %% 128 == 0 (got %d) and a single-phase form", a.cout);
    OOD_REQUIRE(!a.tiled || a.acc_in || (a.out_f32 && a.out_y), "conv3x3 tc: tiled = 1 without a tile-order tensor (acc_in, or out_y with out_f32)");
    const int BK = (a.cin % 64 == 0) ? 64 : 32;
    int BN = 0;

    TcParams p{};
    plan_tiles(a, g, p, BN);
    OOD_REQUIRE(!a.rgb_out || (p.n_tiles_n == 1 && !a.transposed), "conv3x3 tc: the fused ToRGB epilogue needs Co == tile N (Co <= 256) and the stride-1 form");
    p.ep = make_epilogue(a, a.out_f32);
    p.out_bf16 = 1;
    p.in_f16 = a.dtype == OOD_F16;
    p.out_f16 = (a.out_dtype ? a.out_dtype : a.dtype) == OOD_F16;
    OOD_REQUIRE(!a.acc_in || (uintptr_t)a.acc_in % 16 == 0, "conv3x3 tc: acc_in must be 16-byte aligned");
    OOD_REQUIRE((uintptr_t)a.out_y % 32 == 0 && (uintptr_t)a.out_ys % 32 == 0, "conv3x3 tc: outputs must be 32-byte aligned (256-bit stores)");

    // CTA pairs (cta_group::2): ungrouped launches of wide tiles with an even number of M tiles in every phase and at least two tiles per SM
    // OOD_CTA2: 0 off, 1 (default) = the 256-wide tiles, 2 = the 128-wide tiles too (measured slower); OOD_CTA2_MIN_TILES: launches with fewer tiles stay single-CTA
    // (read per call: the parity tests switch them inside one process)
    const char *e2 = getenv("OOD_CTA2"), *e2m = getenv("OOD_CTA2_MIN_TILES");
    const int cta2_mode = e2 ? atoi(e2) : 1, cta2_min = e2m ? atoi(e2m) : 2 * kNumSMs;
    bool m_even = true;                                 // per phase: the two CTAs of a pair take M-adjacent tiles of one phase and one N tile
    for (int i = 0; i < p.nphases; ++i) m_even = m_even && (p.ph[i].tiles_x * p.ph[i].tiles_y * p.ph[i].tiles_b) % 2 == 0;
    const bool pair = cta2_mode > 0 && groups == 1 && !p.in_shared && !p.fused && BK == 64 && (BN == 256 || BN == 128) &&
                      m_even && p.total_tiles >= cta2_min && (cta2_mode > 1 || BN == 256);
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)a.cin, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)(p.in_shared ? p.gbatch : a.batch)};
        cuuint64_t strides[3] = {(cuuint64_t)a.cin * 2, (cuuint64_t)a.w * a.cin * 2, (cuuint64_t)a.h * a.w * a.cin * 2};
        // with element stride s the box spans s*(n-1)+1 tensor elements and delivers n of them to shared memory
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(g.isx * (p.TW - 1) + 1), (cuuint32_t)(g.isy * (p.TH - 1) + 1), (cuuint32_t)p.NB};
        cuuint32_t es[4] = {1, (cuuint32_t)g.isx, (cuuint32_t)g.isy, 1};
        CUresult r = encode(&tmA, a.dtype == OOD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(a.in), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3 tc: activation tensor map encode failed (%d)", (int)r); return OOD_ERR_CUDA; }
    }
    {
        const cuuint64_t ncols = (cuuint64_t)(p.fused ? 4 * a.cout : a.cout);
        cuuint64_t dims[3] = {(cuuint64_t)a.cin, ncols, (cuuint64_t)(p.wtaps * groups)};
        cuuint64_t strides[2] = {(cuuint64_t)a.cin * 2, ncols * a.cin * 2};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(pair ? BN / 2 : BN), 1};          // pair: every CTA loads half of the N tile's rows
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&tmB, a.dtype == OOD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(a.weight), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv3x3 tc: weight tensor map encode failed (%d)", (int)r); return OOD_ERR_CUDA; }
    }
    if (a.stats_out || a.stats_ws) {  // fused output statistics: wide tiles, one image per tile, single phase (stats_out == NULL: the per-tile sums only)
        OOD_REQUIRE(a.stats_ws && !a.acc_in && !a.out_ys && !a.out_f32 && a.out_y && groups == 1 && (a.transposed == 0 || a.transposed == 3 || a.transposed == 4 || a.transposed == 6) &&
                    p.NB == 1 && BK == 64 && (BN == 256 || BN == 128),
                    "conv3x3 tc: stats_out needs the stride-1 / stride-2 pad-1 / 1x1 form, bf16 out_y only, cin %% 64 == 0, cout %% 128 == 0 and >= 128 output pixels");
        p.ep.stat_partial = a.stats_ws;
        p.ep.stat_sums_only = a.stats_out == nullptr;
        const int rc = pair ? (BN == 256 ? launch_tc<256, 64, false, true, 4, true>(tmA, tmB, p, st) : launch_tc<128, 64, false, true, 4, true>(tmA, tmB, p, st))
                            : (BN == 256 ? launch_tc<256, 64, false, true>(tmA, tmB, p, st) : launch_tc<128, 64, false, true>(tmA, tmB, p, st));
        if (rc != OOD_OK) return rc;
        if (!a.stats_out) return OOD_OK;
        in_finalize_launch(a.stats_ws, a.stats_out, (int64_t)g.OH * g.OW, a.cout, p.ph[0].tiles_x * p.ph[0].tiles_y, a.stats_eps, a.batch, st);
        return check_launch("conv3x3 tc stats", 1);
    }
    if (a.acc_in) {     // seeded accumulators: built for the wide tiles only (the AlignNet convolutions)
        if (pair) return BN == 256 ? launch_tc<256, 64, true, false, 4, true>(tmA, tmB, p, st) : launch_tc<128, 64, true, false, 4, true>(tmA, tmB, p, st);
        if (BN == 256 && BK == 64) return launch_tc<256, 64, true>(tmA, tmB, p, st);
        if (BN == 128 && BK == 64) return launch_tc<128, 64, true>(tmA, tmB, p, st);
        set_error("conv3x3 tc: acc_in needs cin %% 64 == 0 and a 128- or 256-wide N tile (cout %% 128 == 0), got cin %d cout %d", a.cin, a.cout);
        return OOD_ERR_ARG;
    }
    if (p.fused && epw8()) {
        // eight epilogue warps for the fused-phase transposed form (4 k-iterations per tile: 742 -> 709 us on the 1024 px
        // layer).  Measured and NOT adopted elsewhere: the stride-1 convolutions, including the half-K AlignNet ones, did not
        // move with 8 warps (A/B in profiles/README.md) -- their tile period is not set by epilogue issue rate.
#define OOD_TC_CASE8(bn, bk) if (BN == bn && BK == bk) return launch_tc<bn, bk, false, false, 8>(tmA, tmB, p, st)
        OOD_TC_CASE8(256, 64); OOD_TC_CASE8(128, 64); OOD_TC_CASE8(256, 32); OOD_TC_CASE8(128, 32);
#undef OOD_TC_CASE8
    }
    if (pair) return BN == 256 ? launch_tc<256, 64, false, false, 4, true>(tmA, tmB, p, st) : launch_tc<128, 64, false, false, 4, true>(tmA, tmB, p, st);
#define OOD_TC_CASE(bn, bk) if (BN == bn && BK == bk) return launch_tc<bn, bk>(tmA, tmB, p, st)
    OOD_TC_CASE(256, 64); OOD_TC_CASE(128, 64); OOD_TC_CASE(64, 64); OOD_TC_CASE(32, 64);
    OOD_TC_CASE(256, 32); OOD_TC_CASE(128, 32); OOD_TC_CASE(64, 32); OOD_TC_CASE(32, 32);
#undef OOD_TC_CASE
    set_error("conv3x3 tc: no kernel for BN=%d BK=%d", BN, BK);
    return OOD_ERR_ARG;
}

}  // namespace ood

extern "C" int64_t ood_conv3x3_stats_workspace(int batch, int h, int w, int cin, int cout, int transposed) {
    using namespace ood;
    if (batch <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || cout % 128 != 0 || (transposed != 0 && transposed != 3 && transposed != 4 && transposed != 6)) return 0;
    ood_conv3x3_args a{};
    a.batch = batch; a.h = h; a.w = w; a.cin = cin; a.cout = cout; a.transposed = transposed;
    float dummy = 0.f;
    a.stats_out = &dummy;
    const ConvGeom g = make_geom(batch, h, w, cin, cout, transposed);
    TcParams p{};
    int BN = 0;
    plan_tiles(a, g, p, BN);
    if (p.NB != 1) return 0;
    return (int64_t)batch * p.ph[0].tiles_x * p.ph[0].tiles_y * cout * 2 * (int64_t)sizeof(float);
}

extern "C" int64_t ood_conv3x3_tiled_bytes(int batch, int h, int w, int cin, int cout, int transposed) {
    using namespace ood;
    if (batch <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || cout % 128 != 0 || transposed == 1 || transposed == 5 || transposed < 0 || transposed > 6) return 0;
    ood_conv3x3_args a{};
    a.batch = batch; a.h = h; a.w = w; a.cin = cin; a.cout = cout; a.transposed = transposed; a.tiled = 1;
    const ConvGeom g = make_geom(batch, h, w, cin, cout, transposed);
    TcParams p{};
    int BN = 0;
    plan_tiles(a, g, p, BN);
    return (int64_t)p.total_tiles * TBM * BN * (int64_t)sizeof(float);
}

namespace ood {
int conv3x3_simt(const ood_conv3x3_args &a, cudaStream_t st);
int conv3x3_rows(const ood_conv3x3_args &a, cudaStream_t st, int *handled);
}

extern "C" int ood_conv3x3(const ood_conv3x3_args *a, void *stream) {
    using namespace ood;
    OOD_REQUIRE(a && a->in && a->weight, "conv3x3: null pointer");
    OOD_REQUIRE(a->batch > 0 && a->h > 0 && a->w > 0 && a->cin > 0 && a->cout > 0, "conv3x3: bad sizes");
    OOD_REQUIRE(a->out_y || a->out_ys || a->rgb_out, "conv3x3: no output requested");
    OOD_REQUIRE(!a->rgb_out || (a->rgb_w && a->rgb_bias && a->act == 1 && a->h % 2 == 0 && a->w % 2 == 0), "conv3x3: fused ToRGB needs rgb_w, rgb_bias, act=1 and even sizes");
    OOD_REQUIRE(!a->out_ys || a->s_next, "conv3x3: out_ys needs s_next");
    OOD_REQUIRE(a->transposed >= 0 && a->transposed <= 6, "conv3x3: transposed must be 0..6");
    OOD_REQUIRE((a->transposed != 1 && a->transposed != 5) || (!a->out_ys && !a->act && !a->noise && !a->bias && !a->d),
                "conv3x3: the transposed form writes raw accumulators (the epilogue follows the blur)");
    OOD_REQUIRE(a->transposed != 2 || (a->h % 2 == 1 && a->w % 2 == 1 && a->h >= 3 && a->w >= 3 && !a->noise),
                "conv3x3: the strided data-gradient form needs an odd (2h+1)x(2w+1) input");
    OOD_REQUIRE(!(a->out_f32 && a->out_ys), "conv3x3: out_f32 applies to out_y only");
    OOD_REQUIRE(a->act >= 0 && a->act <= 2 && (a->act != 2 || a->prelu_slope), "conv3x3: act must be 0, 1 or 2 (PReLU needs prelu_slope)");
    cudaStream_t st = (cudaStream_t)stream;
    g_conv_route = 3;
    if (a->impl == 1) return conv3x3_simt(*a, st);
    OOD_REQUIRE(a->impl == 0, "conv3x3: impl must be 0 (tcgen05) or 1 (simt)");
    g_conv_route = 0;
    if (!ood_device_is_sm100()) { set_error("conv3x3: the tcgen05 path needs an sm_100 device"); return OOD_ERR_DEVICE; }
    {   // high-resolution small-channel layers: row-sliding kernel (conv_rows.cu); anything else: the generic tiles below
        int handled = 0;
        const int rc = conv3x3_rows(*a, st, &handled);
        if (handled) { g_conv_route = 1; return rc; }
    }
    return conv3x3_tc(*a, st);
}
