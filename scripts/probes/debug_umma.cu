// Debug probe (NOT linked into libood_b200.so; built by scripts/probe_umma_shift.py into its own library): does a tcgen05 shared-memory descriptor whose start address is shifted by
// whole rows inside a TMA-written swizzled tile address the shifted rows?  Used to validate the sliding-window conv.
// A: [rows][K=BK] bf16 written by TMA (SW128 for BK=64, SW64 for BK=32); B: identity [N=BK][BK] -> D[m][n] = A[m+shift][n].
#include <cuda.h>

#include "common.cuh"

namespace ood {
namespace dbg {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BK>
__global__ void __launch_bounds__(128) umma_shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                         float *out, int shift, int use_base_offset) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    constexpr int ROWB = BK * 2;
    uint8_t *sA = smem;                       // 256 rows
    uint8_t *sB = smem + 256 * ROWB;          // BK rows (N = BK)
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + ((BK * ROWB + 1023) & ~1023));
    uint32_t *holder = reinterpret_cast<uint32_t *>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *holder;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(256 * ROWB + BK * ROWB) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sA)), "l"(&tmA), "r"(smem_u32(&bar[0])), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sB)), "l"(&tmB), "r"(smem_u32(&bar[0])), "r"(0), "r"(0) : "memory");
        asm volatile("{\n.reg .pred p;\nW0:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D0;\nbra W0;\nD0:\n}\n" ::"r"(smem_u32(&bar[0])) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        auto desc = [&](uint32_t addr) {
            uint64_t d = 0;
            d |= (uint64_t)((addr >> 4) & 0x3FFF);
            d |= (uint64_t)1 << 16;
            d |= (uint64_t)((8 * ROWB) >> 4) << 32;
            d |= (uint64_t)1 << 46;
            if (use_base_offset) d |= (uint64_t)((addr >> 7) & 7) << 49;
            d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;
            return d;
        };
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = desc(smem_u32(sA) + shift * ROWB), db = desc(smem_u32(sB));
        for (int k = 0; k < BK / 16; ++k) {
            uint32_t acc = k != 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(&bar[1])) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c = 0; c < BK; ++c) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[row * BK + c] = __uint_as_float(v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

}  // namespace dbg
}  // namespace ood

// a: bf16 [256][bk], b: bf16 [bk][bk], out: fp32 [128][bk]
extern "C" int ood_debug_umma_shift(const void *a, const void *b, float *out, int bk, int shift, int use_base_offset, void *stream) {
    using namespace ood;
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) {
        set_error("debug: no cuTensorMapEncodeTiled");
        return OOD_ERR_CUDA;
    }
    Fn encode = (Fn)ptr;
    OOD_REQUIRE(bk == 64 || bk == 32, "debug: bk must be 32 or 64");
    CUtensorMap tmA, tmB;
    const CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t dA[2] = {(cuuint64_t)bk, 256}, sA[1] = {(cuuint64_t)bk * 2};
    cuuint32_t bA[2] = {(cuuint32_t)bk, 256}, es[2] = {1, 1};
    if (encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)a, dA, sA, bA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { set_error("debug: encode A"); return OOD_ERR_CUDA; }
    cuuint64_t dB[2] = {(cuuint64_t)bk, (cuuint64_t)bk};
    cuuint32_t bB[2] = {(cuuint32_t)bk, (cuuint32_t)bk};
    if (encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)b, dB, sA, bB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { set_error("debug: encode B"); return OOD_ERR_CUDA; }
    if (bk == 64) dbg::umma_shift_kernel<64><<<1, 128, 46 * 1024, (cudaStream_t)stream>>>(tmA, tmB, out, shift, use_base_offset);
    else dbg::umma_shift_kernel<32><<<1, 128, 46 * 1024, (cudaStream_t)stream>>>(tmA, tmB, out, shift, use_base_offset);
    return check_launch("debug_umma_shift");
}
