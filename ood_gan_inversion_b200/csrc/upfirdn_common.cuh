// Parameters shared by the NCHW-plane upfirdn2d kernels (upfirdn2d.cu, plane_fir.cu).
#pragma once
#include "common.cuh"

namespace ood {

struct UpfirdnParams {
    const void *in;
    void *out;
    const float *kernel;
    int64_t planes;
    int in_h, in_w, out_h, out_w;
    int kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0;
    int tiles_x, tiles_y, sih, siw;
};

}  // namespace ood
