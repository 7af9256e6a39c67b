// fp32 SIMT implicit-GEMM 3x3 convolution on NHWC activations (parity mode of ModulatedConv2d).
// Reference arithmetic: src/ops/StyleGAN/model.py:233-274, executed there as cuDNN grouped convolutions
// over materialised per-sample weights; here shared weights [tap][Ci][Co], pre-modulated input, fused
// demodulation / noise / bias / leaky-ReLU epilogue.  FFMA only: this is the bit-faithful fp32 path that the
// tcgen05 path (conv_tc.cu) is checked against on the device; it is not the throughput path.
#include "conv_common.cuh"

namespace ood {

constexpr int SBM = 64, SBN = 64, SBK = 16;

struct SimtParams {
    const float *in;
    const float *w;
    ConvGeom g;
    ConvEpilogue ep;
};

__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtParams p) {
    __shared__ __align__(16) float As[SBK][SBM + 4];
    __shared__ __align__(16) float Bs[SBK][SBN];
    const ConvPhase &ph = p.g.ph[blockIdx.z];
    const int m0 = blockIdx.x * SBM;
    if (m0 >= ph.m_total) return;
    const int n0 = blockIdx.y * SBN;
    const int tid = threadIdx.x;
    const int cin = p.g.cin, cout = p.g.cout;

    // A-load role: pixel a_m, k-quad a_kq (4 consecutive input channels)
    const int a_m = tid >> 2, a_kq = tid & 3;
    int ab = 0, aoy = 0, aox = 0;
    const bool a_valid = (m0 + a_m) < ph.m_total;
    if (a_valid) {
        const int m = m0 + a_m;
        ab = m / (ph.oh * ph.ow);
        const int r = m - ab * ph.oh * ph.ow;
        aoy = r / ph.ow;
        aox = r - aoy * ph.ow;
    }
    // B-load role: k row b_k, 4 consecutive output channels at b_n
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < ph.ntaps; ++t) {
        const int iy = aoy * p.g.isy + ph.dy[t], ix = aox * p.g.isx + ph.dx[t];
        const bool in_ok = a_valid && iy >= 0 && iy < p.g.h && ix >= 0 && ix < p.g.w;
        const float *arow = p.in + (((int64_t)ab * p.g.h + iy) * p.g.w + ix) * cin;
        const float *wtap = p.w + (int64_t)ph.wt[t] * cin * cout;
        for (int c0 = 0; c0 < cin; c0 += SBK) {
            float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in_ok) av = __ldg(reinterpret_cast<const float4 *>(arow + c0 + a_kq * 4));
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + b_n < cout) bv = __ldg(reinterpret_cast<const float4 *>(wtap + (int64_t)(c0 + b_k) * cout + n0 + b_n));
            __syncthreads();
            As[a_kq * 4 + 0][a_m] = av.x;
            As[a_kq * 4 + 1][a_m] = av.y;
            As[a_kq * 4 + 2][a_m] = av.z;
            As[a_kq * 4 + 3][a_m] = av.w;
            *reinterpret_cast<float4 *>(&Bs[b_k][b_n]) = bv;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SBK; ++k) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
                const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
            }
        }
    }

    // epilogue
    const int n = n0 + tx * 4;
    if (n >= cout) return;
    const float nw = (p.ep.noise && p.ep.noise_w) ? *p.ep.noise_w : 0.f;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.ep.bias) {
        const float4 bb = *reinterpret_cast<const float4 *>(p.ep.bias + n);
        bias[0] = bb.x; bias[1] = bb.y; bias[2] = bb.z; bias[3] = bb.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= ph.m_total) continue;
        const int b = m / (ph.oh * ph.ow);
        const int r = m - b * ph.oh * ph.ow;
        const int oy = r / ph.ow, ox = r - oy * ph.ow;
        const int Y = oy * p.g.sy + ph.py, X = ox * p.g.sx + ph.px;
        const int64_t pix = ((int64_t)b * p.g.OH + Y) * p.g.OW + X;
        float dd[4] = {1.f, 1.f, 1.f, 1.f};
        if (p.ep.d) {
            const float4 t4 = *reinterpret_cast<const float4 *>(p.ep.d + (int64_t)b * cout + n);
            dd[0] = t4.x; dd[1] = t4.y; dd[2] = t4.z; dd[3] = t4.w;
        }
        const float nz = p.ep.noise ? nw * p.ep.noise[b * p.ep.noise_bstride + (int64_t)Y * p.g.OW + X] : 0.f;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = fmaf(acc[i][j], dd[j], nz) + bias[j];
            v[j] = apply_act(v[j], p.ep.act, p.ep.act == 2 ? p.ep.prelu[n + j] : 0.f);
        }
        if (p.ep.out_y) *reinterpret_cast<float4 *>((float *)p.ep.out_y + pix * cout + n) = make_float4(v[0], v[1], v[2], v[3]);
        if (p.ep.out_ys) {
            const float4 s4 = *reinterpret_cast<const float4 *>(p.ep.s_next + (int64_t)b * cout + n);
            *reinterpret_cast<float4 *>((float *)p.ep.out_ys + pix * cout + n) =
                make_float4(v[0] * s4.x, v[1] * s4.y, v[2] * s4.z, v[3] * s4.w);
        }
    }
}

int conv3x3_simt(const ood_conv3x3_args &a, cudaStream_t st) {
    OOD_REQUIRE(a.dtype == OOD_F32, "conv3x3 simt: storage type must be fp32");
    OOD_REQUIRE(a.cin % SBK == 0 && a.cout % 4 == 0, "conv3x3 simt: cin %% 16 and cout %% 4 must be 0 (got %d, %d)", a.cin, a.cout);
    SimtParams p;
    p.in = (const float *)a.in;
    p.w = (const float *)a.weight;
    p.g = make_geom(a.batch, a.h, a.w, a.cin, a.cout, a.transposed);
    OOD_REQUIRE(!a.rgb_out, "conv3x3 simt: the fused ToRGB epilogue exists on the tcgen05 path only");
    OOD_REQUIRE(a.groups <= 1, "conv3x3 simt: the grouped form exists on the tcgen05 path only");
    OOD_REQUIRE(a.transposed != 5, "conv3x3 simt: the fused-phase transposed form (5) exists on the tcgen05 path only; use form 1");
    OOD_REQUIRE(!a.acc_in && !a.tiled && !a.stats_out && !a.stats_ws, "conv3x3 simt: the accumulator seed (acc_in), the tile-order tensors and the fused statistics exist on the tcgen05 path only");
    p.ep = make_epilogue(a, 1);
    int mmax = 0;
    for (int i = 0; i < p.g.nphases; ++i) mmax = std::max(mmax, p.g.ph[i].m_total);
    dim3 grid(ceil_div(mmax, SBM), ceil_div(a.cout, SBN), p.g.nphases);
    OOD_REQUIRE(grid.y <= 65535, "conv3x3 simt: cout too large");
    conv_simt_kernel<<<grid, 256, 0, st>>>(p);
    return check_launch("conv3x3 simt");
}

}  // namespace ood
