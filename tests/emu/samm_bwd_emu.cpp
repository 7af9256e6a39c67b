// Host emulation of the SAMM backward kernels: the per-item bodies of ood_gan_inversion_b200/csrc/samm_bwd.cuh compiled by
// g++ and driven over every (image, pixel, channel group) / (image, pixel) coordinate the __global__ wrappers of samm_bwd.cu
// would visit.  Test infrastructure only (tests/test_samm_bwd_cpu.py builds it into a temporary .so).
#include "../../ood_gan_inversion_b200/csrc/samm_bwd.cuh"

extern "C" void emu_warp_mix_bwd(const float *gen, const float *field, const float *gout, float *ggen, float *gfield, int batch,
                                 int h, int w, int channels, int G) {
    for (int b = 0; b < batch; ++b)
        for (int pix = 0; pix < h * w; ++pix)
            for (int g = 0; g < G; ++g)
                ood_bwd::warp_mix_bwd_item<float>(gen, field, gout, ggen, gfield, b, pix, g, G, h, w, channels, ood_bwd::HostAdd());
}

extern "C" void emu_mask_blend_bwd(const float *const *fields, float *const *gfields, const int *sizes, int n, const float *x,
                                   const float *gen, const float *gout, float *gx, float *ggen, int batch, int size) {
    ood_bwd::MaskBwdParams mp{};
    mp.n = n;
    for (int i = 0; i < n; ++i) {
        mp.f[i] = fields[i];
        mp.gf[i] = gfields[i];
        mp.r[i] = sizes[i];
        mp.scale[i] = (float)sizes[i] / (float)size;
    }
    for (int b = 0; b < batch; ++b)
        for (int y = 0; y < size; ++y)
            for (int x_ = 0; x_ < size; ++x_) ood_bwd::mask_blend_bwd_item(mp, x, gen, gout, gx, ggen, b, y, x_, size, ood_bwd::HostAdd());
}

extern "C" void emu_field_step_bwd(const float *z, const float *prev, const float *coarse, const float *gacc, const float *taps,
                                   float scale, int batch, int r, int rc, float *gf, float *gz, float *gprev, float *gcoarse) {
    ood_bwd::FieldBwdArgs a{z, prev, coarse, {taps[3], taps[2], taps[1], taps[0]}, scale, r, rc};
    for (int b = 0; b < batch; ++b)
        for (int y = 0; y < r; ++y)
            for (int x = 0; x < r; ++x) ood_bwd::field_step_bwd_pass1_item(a, gacc, gf, gprev, gcoarse, b, y, x, ood_bwd::HostAdd());
    for (int b = 0; b < batch; ++b)
        for (int y = 0; y < r; ++y)
            for (int x = 0; x < r; ++x) ood_bwd::field_step_bwd_pass2_item(a, gf, gz, b, y, x);
}
