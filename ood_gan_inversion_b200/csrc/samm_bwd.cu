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

// Vector form of the warp backward: a thread owns FOUR consecutive channels of one pixel, so every scatter into ggen is one 16-byte
// reduction (red.global.add.v4.f32, sm_90+) instead of four scalar atomics -- the scalar form ran at 0.07-0.13 of the HBM roofline,
// bound by the L2's atomic rate (five atomics per element) -- and the three field-gradient sums are reduced over the threads of a pixel
// with shuffles before one atomic per warp segment.  Same arithmetic per element as warp_mix_bwd_item (samm_bwd.cuh).
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <typename T> __device__ __forceinline__ void ld4(const T *p, float *v);
template <> __device__ __forceinline__ void ld4<float>(const float *p, float *v) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16 *p, float *v) {
    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
    const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename T>
__global__ void __launch_bounds__(256) warp_mix_bwd_vec_kernel(const T *__restrict__ gen, const float *__restrict__ field,
                                                                const T *__restrict__ gout, float *__restrict__ ggen,
                                                                float *__restrict__ gfield, int H, int W, int C, int V) {
    const int b = blockIdx.y;
    const int64_t P = (int64_t)H * W;
    const int64_t items = P * V;
    const int lane = threadIdx.x & 31;
    const int seg = V < 32 ? V : 32;                       // threads of one pixel inside a warp (V is a power of two)
    const float *fb = field + (int64_t)b * 3 * P;
    const T *gb = gen + (int64_t)b * P * C;
    float *ggb = ggen + (int64_t)b * P * C;
    float *gf = gfield + (int64_t)b * 3 * P;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i - lane < items; i += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = i < items;
        float s_ix = 0.f, s_iy = 0.f, s_al = 0.f;
        int pix = 0;
        if (valid) {
            pix = (int)(i / V);
            const int c = (int)(i - (int64_t)pix * V) * 4;
            const int y = pix / W, x = pix - y * W;
            const float f0 = __ldg(fb + pix), f1 = __ldg(fb + P + pix), alpha = __ldg(fb + 2 * P + pix);
            const float gx = ood_bwd::linspace_m1_1(x, W) + f0, gy = ood_bwd::linspace_m1_1(y, H) + f1;
            const float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const int x0 = (int)fx0, y0 = (int)fy0;
            const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
            float go[4], ctr[4], t[4][4];
            ld4<T>(gout + ((int64_t)b * P + pix) * C + c, go);
            ld4<T>(gb + (int64_t)pix * C + c, ctr);
            bool ok[4];
            int64_t off[4];
            float w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int xx = x0 + (q & 1), yy = y0 + (q >> 1);
                ok[q] = xx >= 0 && xx < W && yy >= 0 && yy < H;
                off[q] = ok[q] ? ((int64_t)yy * W + xx) * C : 0;
                w[q] = ((q & 1) ? wx1 : wx0) * ((q >> 1) ? wy1 : wy0);
                if (ok[q]) ld4<T>(gb + off[q] + c, t[q]);
                else t[q][0] = t[q][1] = t[q][2] = t[q][3] = 0.f;
            }
            const float a1 = 1.f - alpha;
            red_add_v4(ggb + (int64_t)pix * C + c, a1 * go[0], a1 * go[1], a1 * go[2], a1 * go[3]);
            float ga[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ga[j] = go[j] * alpha;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (ok[q]) red_add_v4(ggb + off[q] + c, ga[0] * w[q], ga[1] * w[q], ga[2] * w[q], ga[3] * w[q]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float smp = ((w[0] * t[0][j] + w[1] * t[1][j]) + w[2] * t[2][j]) + w[3] * t[3][j];
                s_ix += ga[j] * ((t[1][j] - t[0][j]) * wy0 + (t[3][j] - t[2][j]) * wy1);
                s_iy += ga[j] * ((t[2][j] - t[0][j]) * wx0 + (t[3][j] - t[1][j]) * wx1);
                s_al += go[j] * (smp - ctr[j]);
            }
        }
        for (int o = seg >> 1; o > 0; o >>= 1) {           // the threads of a pixel are `seg` adjacent lanes
            s_ix += __shfl_xor_sync(0xffffffffu, s_ix, o);
            s_iy += __shfl_xor_sync(0xffffffffu, s_iy, o);
            s_al += __shfl_xor_sync(0xffffffffu, s_al, o);
        }
        if (valid && (lane & (seg - 1)) == 0) {
            atomicAdd(gf + pix, s_ix * (0.5f * (float)W));
            atomicAdd(gf + P + pix, s_iy * (0.5f * (float)H));
            atomicAdd(gf + 2 * P + pix, s_al);
        }
    }
}

// The level gradients: S^2 / r^2 full-resolution pixels share one mask cell (1024 at the 32 px level), so plain global atomics
// serialise on a few thousand addresses (round 2 measured 11.4 ms for one call at 1024 px, batch 16).  A block owns a 32 x 32 pixel tile,
// whose bilinear taps fall into a window of at most 10 x 10 cells per level: contributions are accumulated in shared memory
// (shared-memory atomics) and flushed once per cell per block.
constexpr int kMbTile = 32, kMbWin = 12;
struct WindowAdd {
    float *win;                       // [4][kMbWin][kMbWin]
    const float *base[4];             // alpha-gradient plane of this image per level
    int r[4], cy0[4], cx0[4], n;
    __device__ __forceinline__ void operator()(float *p, float v) const {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= n) break;
            const int64_t off = p - base[k];
            if (off >= 0 && off < (int64_t)r[k] * r[k]) {
                const int cy = (int)(off / r[k]), cx = (int)(off - (int64_t)cy * r[k]);
                const int wy = cy - cy0[k], wx = cx - cx0[k];
                if (wy >= 0 && wy < kMbWin && wx >= 0 && wx < kMbWin) atomicAdd(win + (k * kMbWin + wy) * kMbWin + wx, v);
                else atomicAdd(p, v);
                return;
            }
        }
        atomicAdd(p, v);
    }
};

__global__ void __launch_bounds__(256) mask_blend_bwd_kernel(const ood_bwd::MaskBwdParams mp, const float *__restrict__ xin,
                                                              const float *__restrict__ gen, const float *__restrict__ gout,
                                                              float *__restrict__ gx, float *__restrict__ ggen, int S) {
    __shared__ float win[4 * kMbWin * kMbWin];
    const int b = blockIdx.z;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = tid; i < 4 * kMbWin * kMbWin; i += 256) win[i] = 0.f;
    WindowAdd add;
    add.win = win;
    add.n = mp.n;
    const int ty0 = blockIdx.y * kMbTile, tx0 = blockIdx.x * kMbTile;
    for (int k = 0; k < 4; ++k) {
        if (k >= mp.n) { add.base[k] = nullptr; add.r[k] = 0; add.cy0[k] = add.cx0[k] = 0; continue; }
        add.r[k] = mp.r[k];
        add.base[k] = mp.gf[k] + ((int64_t)b * 3 + 2) * mp.r[k] * mp.r[k];
        const ood_bwd::BilinearTap t = ood_bwd::bilinear_tap(mp.r[k], mp.scale[k], ty0, tx0);      // first cell the tile's top-left pixel touches
        add.cy0[k] = t.y0;
        add.cx0[k] = t.x0;
    }
    __syncthreads();
    const int x = tx0 + threadIdx.x;
    for (int yy = threadIdx.y; yy < kMbTile; yy += blockDim.y) {
        const int y = ty0 + yy;
        if (x < S && y < S) ood_bwd::mask_blend_bwd_item(mp, xin, gen, gout, gx, ggen, b, y, x, S, add);
    }
    __syncthreads();
    for (int i = tid; i < 4 * kMbWin * kMbWin; i += 256) {
        const int k = i / (kMbWin * kMbWin), wy = (i / kMbWin) % kMbWin, wx = i % kMbWin;
        if (k >= mp.n) continue;
        const float v = win[i];
        const int cy = add.cy0[k] + wy, cx = add.cx0[k] + wx;
        if (v != 0.f && cy < add.r[k] && cx < add.r[k]) atomicAdd(const_cast<float *>(add.base[k]) + (int64_t)cy * add.r[k] + cx, v);
    }
}

// ---- deterministic two-stage form of the level gradients (no atomics).
// stage 1 (item = full-resolution pixel): gx / ggen as above, and gu_k = dL/d(up(alpha_k))[y,x] written to a workspace plane per level;
// stage 2 (one warp per mask cell): the adjoint of the bilinear up-sampling as a GATHER over the cell's footprint,
//          g_alpha_k[cy,cx] = sum_y wy_k(y -> cy) sum_x wx_k(x -> cx) gu_k[y,x],   fixed summation order.
__global__ void __launch_bounds__(256) mask_blend_bwd_gu_kernel(const ood_bwd::MaskBwdParams mp, const float *__restrict__ xin,
                                                                 const float *__restrict__ gen, const float *__restrict__ gout,
                                                                 float *__restrict__ gx, float *__restrict__ ggen, float *__restrict__ gu, int S, int batch) {
    const int b = blockIdx.z;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= S || y >= S) return;
    const int64_t P = (int64_t)S * S, pix = (int64_t)y * S + x;
    float u[4], A[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= mp.n) break;
        const int r = mp.r[k];
        const float *a = mp.f[k] + ((int64_t)b * 3 + 2) * r * r;
        const ood_bwd::BilinearTap t = ood_bwd::bilinear_tap(r, mp.scale[k], y, x);
        u[k] = t.ly0 * (t.lx0 * a[(int64_t)t.y0 * r + t.x0] + t.lx1 * a[(int64_t)t.y0 * r + t.x1]) +
               t.ly1 * (t.lx0 * a[(int64_t)t.y1 * r + t.x0] + t.lx1 * a[(int64_t)t.y1 * r + t.x1]);
        A[k] = (k == 0) ? u[k] : (u[k] * A[k - 1] + A[k - 1] * (1.f - A[k - 1]));
    }
    float An = A[0];
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (k < mp.n) An = A[k];
    const float Ac = fminf(fmaxf(An, 0.f), 1.f);
    float gA = 0.f;
    for (int c = 0; c < 3; ++c) {
        const int64_t o = ((int64_t)b * 3 + c) * P + pix;
        const float go = gout[o];
        gA += go * (xin[o] - gen[o]);
        if (gx) gx[o] = Ac * go;
        if (ggen) ggen[o] = (1.f - Ac) * go;
    }
    if (!(An >= 0.f && An <= 1.f)) gA = 0.f;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        if (k >= mp.n) continue;
        gu[((int64_t)k * batch + b) * P + pix] = (k == 0) ? gA : gA * A[k - 1];
        if (k > 0) gA = gA * (u[k] + 1.f - 2.f * A[k - 1]);
    }
}

template <int LANES>     // threads per cell: a warp for the coarse levels (footprints of up to 66 x 66 pixels), one thread for the fine ones (11 x 11)
__global__ void __launch_bounds__(256) bilinear_up_adjoint_kernel(const float *__restrict__ gu, float *__restrict__ galpha, int r, float scale, int S) {
    // cell (cy, cx) of image blockIdx.y; gu: this level's plane [B][S][S]; galpha: the alpha channel of [B,3,r,r]
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / LANES, lane = threadIdx.x % LANES;
    if (warp >= r * r) return;
    const int b = blockIdx.y;
    const int cy = warp / r, cx = warp - cy * r;
    const float inv = 1.f / scale;
    // pixels whose taps can touch cell c: source coordinate s(p) = (p + 0.5) * scale - 0.5 in (c - 1, c + 1)  (plus the clamped borders)
    const int ylo = max(0, (int)floorf((cy - 1 + 0.5f) * inv - 0.5f) - 1), yhi = min(S - 1, (int)ceilf((cy + 1 + 0.5f) * inv - 0.5f) + 1);
    const int xlo = max(0, (int)floorf((cx - 1 + 0.5f) * inv - 0.5f) - 1), xhi = min(S - 1, (int)ceilf((cx + 1 + 0.5f) * inv - 0.5f) + 1);
    const float *gp = gu + (int64_t)b * S * S;
    float acc = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
        const ood_bwd::BilinearTap ty = ood_bwd::bilinear_tap(r, scale, y, 0);
        const float wy = (ty.y0 == cy ? ty.ly0 : 0.f) + (ty.y1 == cy ? ty.ly1 : 0.f);
        if (wy == 0.f) continue;
        float row = 0.f;
        for (int x = xlo + lane; x <= xhi; x += LANES) {
            const ood_bwd::BilinearTap tx = ood_bwd::bilinear_tap(r, scale, 0, x);
            const float wx = (tx.x0 == cx ? tx.lx0 : 0.f) + (tx.x1 == cx ? tx.lx1 : 0.f);
            row = fmaf(wx, __ldg(gp + (int64_t)y * S + x), row);
        }
        acc = fmaf(wy, row, acc);
    }
    if (LANES == 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (lane == 0) galpha[((int64_t)b * 3 + 2) * r * r + (int64_t)cy * r + cx] = acc;
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
    cudaStream_t st = (cudaStream_t)stream;
    const int V = channels / 4;
    if (channels % 4 == 0 && V > 0 && (V & (V - 1)) == 0 && ((uintptr_t)gen % 16) == 0 && ((uintptr_t)gout % 16) == 0 && ((uintptr_t)ggen % 16) == 0) {
        // vector form: four channels per thread, 16-byte reductions
        const int64_t vitems = (int64_t)h * w * V;
        dim3 vgrid((unsigned)std::min<int64_t>(ceil_div(vitems, 256), kNumSMs * 32), batch);
        if (dtype == OOD_F32)
            warp_mix_bwd_vec_kernel<float><<<vgrid, 256, 0, st>>>((const float *)gen, field, (const float *)gout, ggen, gfield, h, w, channels, V);
        else
            warp_mix_bwd_vec_kernel<__nv_bfloat16><<<vgrid, 256, 0, st>>>((const __nv_bfloat16 *)gen, field, (const __nv_bfloat16 *)gout, ggen, gfield, h, w,
                                                                          channels, V);
        return check_launch("warp_mix_bwd");
    }
    int G = 1;
    while (G < 32 && G * 2 <= channels) G *= 2;                 // channel groups per pixel: adjacent threads, adjacent channels
    const int64_t items = (int64_t)h * w * G;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(items, 256), kNumSMs * 32), batch);
    if (dtype == OOD_F32)
        warp_mix_bwd_kernel<float><<<grid, 256, 0, st>>>((const float *)gen, field, (const float *)gout, ggen, gfield, h, w, channels, G);
    else
        warp_mix_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)gen, field, (const __nv_bfloat16 *)gout, ggen,
                                                                 gfield, h, w, channels, G);
    return check_launch("warp_mix_bwd");
}

extern "C" int ood_mask_blend_bwd(const float *const *fields_host, float *const *gfields_host, const int *field_sizes_host,
                                  int n_fields, const float *x, const float *gen, const float *gout, float *gx, float *ggen,
                                  float *workspace, int batch, int size, void *stream) {
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
    if (workspace) {        // deterministic two-stage form: n_fields * batch * size^2 floats of workspace; gfields are WRITTEN (alpha channel)
        const dim3 g1(ceil_div(size, 32), ceil_div(size, 8), batch);
        mask_blend_bwd_gu_kernel<<<g1, block, 0, (cudaStream_t)stream>>>(mp, x, gen, gout, gx, ggen, workspace, size, batch);
        for (int i = 0; i < n_fields; ++i) {
            const int r = mp.r[i];
            const float *plane = workspace + (int64_t)i * batch * size * size;
            if ((int64_t)r * 8 >= size)         // footprint <= ~19 pixels wide: one thread per cell
                bilinear_up_adjoint_kernel<1><<<dim3(ceil_div((int64_t)r * r, 256), batch), 256, 0, (cudaStream_t)stream>>>(plane, mp.gf[i], r, mp.scale[i], size);
            else
                bilinear_up_adjoint_kernel<32><<<dim3(ceil_div((int64_t)r * r * 32, 256), batch), 256, 0, (cudaStream_t)stream>>>(plane, mp.gf[i], r, mp.scale[i], size);
        }
        return check_launch("mask_blend_bwd", 1 + n_fields);
    }
    const dim3 grid(ceil_div(size, kMbTile), ceil_div(size, kMbTile), batch);
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
