// Backward of ood_warp_mix and ood_mask_blend (SURVEY.md section 8 row a14 "grid_sample grads", 8b warp_alpha_bwd /
// mask_blend_bwd): thin __global__ wrappers around the portable per-item bodies of samm_bwd.cuh.  First correct path
// (atomics, no staging); the same bodies are checked against torch.autograd on the CPU by tests/test_samm_bwd_cpu.py.
#include "common.cuh"
#include "samm_bwd.cuh"

namespace ood_bwd {
template <> __host__ __device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
}  // namespace ood_bwd

namespace ood {

// threads of a pixel are adjacent (channel groups g = 0..G-1 read adjacent channels), pixels follow
template <typename T>
__global__ void __launch_bounds__(256) warp_mix_bwd_kernel(const T *__restrict__ gen, const float *__restrict__ field,
                                                            const T *__restrict__ gout, float *__restrict__ ggen,
                                                            float *__restrict__ gfield, int H, int W, int C, int G) {
    const int b = blockIdx.y;
    const int64_t items = (int64_t)H * W * G;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
        const int pix = (int)(i / G), g = (int)(i - (int64_t)pix * G);
        ood_bwd::warp_mix_bwd_item<T>(gen, field, gout, ggen, gfield, b, pix, g, G, H, W, C, ood_bwd::DeviceAdd());
    }
}

__global__ void __launch_bounds__(256) mask_blend_bwd_kernel(const ood_bwd::MaskBwdParams mp, const float *__restrict__ xin,
                                                              const float *__restrict__ gen, const float *__restrict__ gout,
                                                              float *__restrict__ gx, float *__restrict__ ggen, int S) {
    const int b = blockIdx.z;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= S || y >= S) return;
    ood_bwd::mask_blend_bwd_item(mp, xin, gen, gout, gx, ggen, b, y, x, S, ood_bwd::DeviceAdd());
}

__global__ void __launch_bounds__(256) field_step_bwd_pass1_kernel(const ood_bwd::FieldBwdArgs a, const float *__restrict__ gacc,
                                                                    float *__restrict__ gf, float *__restrict__ gprev,
                                                                    float *__restrict__ gcoarse) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= a.R || y >= a.R) return;
    ood_bwd::field_step_bwd_pass1_item(a, gacc, gf, gprev, gcoarse, blockIdx.z, y, x, ood_bwd::DeviceAdd());
}

__global__ void __launch_bounds__(256) field_step_bwd_pass2_kernel(const ood_bwd::FieldBwdArgs a, const float *__restrict__ gf,
                                                                    float *__restrict__ gz) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= a.R || y >= a.R) return;
    ood_bwd::field_step_bwd_pass2_item(a, gf, gz, blockIdx.z, y, x);
}

}  // namespace ood

extern "C" int ood_warp_mix_bwd(const void *gen, const float *field, const void *gout, float *ggen, float *gfield, int batch,
                                int h, int w, int channels, int dtype, void *stream) {
    using namespace ood;
    OOD_REQUIRE(gen && field && gout && ggen && gfield && batch > 0 && batch <= 65535 && h > 0 && w > 0 && channels > 0 &&
                    (int64_t)h * w < (1ll << 30),
                "warp_mix_bwd: bad arguments");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16, "warp_mix_bwd: bad dtype");
    int G = 1;
    while (G < 32 && G * 2 <= channels) G *= 2;                 // channel groups per pixel: adjacent threads, adjacent channels
    const int64_t items = (int64_t)h * w * G;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(items, 256), kNumSMs * 32), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32)
        warp_mix_bwd_kernel<float><<<grid, 256, 0, st>>>((const float *)gen, field, (const float *)gout, ggen, gfield, h, w, channels, G);
    else
        warp_mix_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)gen, field, (const __nv_bfloat16 *)gout, ggen,
                                                                 gfield, h, w, channels, G);
    return check_launch("warp_mix_bwd");
}

extern "C" int ood_mask_blend_bwd(const float *const *fields_host, float *const *gfields_host, const int *field_sizes_host,
                                  int n_fields, const float *x, const float *gen, const float *gout, float *gx, float *ggen,
                                  int batch, int size, void *stream) {
    using namespace ood;
    OOD_REQUIRE(fields_host && gfields_host && field_sizes_host && n_fields >= 1 && n_fields <= 4, "mask_blend_bwd: 1..4 fields supported");
    OOD_REQUIRE(x && gen && gout && batch > 0 && batch <= 65535 && size > 0 && size <= 16384, "mask_blend_bwd: bad arguments");
    ood_bwd::MaskBwdParams mp{};
    mp.n = n_fields;
    for (int i = 0; i < n_fields; ++i) {
        OOD_REQUIRE(fields_host[i] && gfields_host[i] && field_sizes_host[i] > 0, "mask_blend_bwd: bad field %d", i);
        mp.f[i] = fields_host[i];
        mp.gf[i] = gfields_host[i];
        mp.r[i] = field_sizes_host[i];
        mp.scale[i] = (float)field_sizes_host[i] / (float)size;
    }
    const dim3 block(32, 8);
    const dim3 grid(ceil_div(size, 32), ceil_div(size, 8), batch);
    mask_blend_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(mp, x, gen, gout, gx, ggen, size);
    return check_launch("mask_blend_bwd");
}

extern "C" int ood_field_step_bwd(const float *z, const float *prev, const float *coarse, const float *gacc, const float *taps_host,
                                  float scale, int batch, int r, int rc, float *gf_workspace, float *gz, float *gprev, float *gcoarse,
                                  void *stream) {
    using namespace ood;
    OOD_REQUIRE(z && gacc && taps_host && gf_workspace && gz && batch > 0 && batch <= 65535 && r > 0 && r <= 32768,
                "field_step_bwd: bad arguments");
    OOD_REQUIRE(!coarse || rc > 0, "field_step_bwd: coarse needs its size");
    OOD_REQUIRE(!gprev || prev, "field_step_bwd: gprev without prev");
    OOD_REQUIRE(!gcoarse || coarse, "field_step_bwd: gcoarse without coarse");
    // the taps as the forward applies them: correlation with the flipped kernel (upfirdn2d.py:179)
    const ood_bwd::FieldBwdArgs a{z, prev, coarse, {taps_host[3], taps_host[2], taps_host[1], taps_host[0]}, scale, r, rc};
    const dim3 block(32, 8);
    const dim3 grid(ceil_div(r, 32), ceil_div(r, 8), batch);
    cudaStream_t st = (cudaStream_t)stream;
    field_step_bwd_pass1_kernel<<<grid, block, 0, st>>>(a, gacc, gf_workspace, gprev, gcoarse);
    field_step_bwd_pass2_kernel<<<grid, block, 0, st>>>(a, gf_workspace, gz);
    return check_launch("field_step_bwd", 2);
}
