// Glue kernels of the IR-SE bottleneck of the E4E encoder on NHWC activations.
// Reference: src/ops/e4e/encoders/helpers.py:59-76 (SEModule), :476-501 (bottleneck_IR_SE):
//     out = SE(BN2(conv2(PReLU(conv1(BN1(x)))))) + shortcut(x),   SE(v) = v * sigmoid(fc2(relu(fc1(mean_hw(v)))))
// The convolutions run on the tcgen05 kernel (conv_tc.cu) with PReLU / folded-BN2 bias epilogues; what is left is
//   se_gate:     per-image channel means (from ood_in_stats) -> the two tiny fully-connected layers -> gate[b,c]
//   se_residual: out = v * gate + shortcut(x) in one pass, which also emits the NEXT block's BN1(out) so that the
//                normalised copy never costs a pass of its own.
// PyTorch runs this tail as ~8 elementwise / reduction kernels per block (24 blocks).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ood {

__global__ void __launch_bounds__(256) se_gate_kernel(const float *__restrict__ stats, const float *__restrict__ w1,
                                                       const float *__restrict__ w2, float *__restrict__ gate, int C, int Cr) {
    extern __shared__ float sg[];            // mean[C], hid[Cr]
    float *mean = sg, *hid = sg + C;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < C; c += blockDim.x) mean[c] = stats[((int64_t)b * C + c) * 2];
    __syncthreads();
    for (int j = warp; j < Cr; j += blockDim.x / 32) {
        const float *wr = w1 + (int64_t)j * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(wr[c], mean[c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) hid[j] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        const float *wr = w2 + (int64_t)c * Cr;
        float s = 0.f;
        for (int j = 0; j < Cr; ++j) s = fmaf(wr[j], hid[j], s);
        gate[(int64_t)b * C + c] = 1.f / (1.f + __expf(-s));
    }
}

template <typename T> __device__ __forceinline__ void enc_load(const T *p, float2 *dst);
template <> __device__ __forceinline__ void enc_load<float>(const float *p, float2 *dst) {
    const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
    dst[0] = make_float2(r.x, r.y); dst[1] = make_float2(r.z, r.w);
}
template <> __device__ __forceinline__ void enc_load<__nv_bfloat16>(const __nv_bfloat16 *p, float2 *dst) {
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    dst[0] = unpack_bf16x2(r.x); dst[1] = unpack_bf16x2(r.y); dst[2] = unpack_bf16x2(r.z); dst[3] = unpack_bf16x2(r.w);
}
template <> __device__ __forceinline__ void enc_load<__half>(const __half *p, float2 *dst) {
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    dst[0] = unpack_f16x2(r.x); dst[1] = unpack_f16x2(r.y); dst[2] = unpack_f16x2(r.z); dst[3] = unpack_f16x2(r.w);
}
template <typename T> __device__ __forceinline__ void unpack8(const uint4 &r, float2 *dst);
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4 &r, float2 *dst) {
    dst[0] = unpack_bf16x2(r.x); dst[1] = unpack_bf16x2(r.y); dst[2] = unpack_bf16x2(r.z); dst[3] = unpack_bf16x2(r.w);
}
template <> __device__ __forceinline__ void unpack8<__half>(const uint4 &r, float2 *dst) {
    dst[0] = unpack_f16x2(r.x); dst[1] = unpack_f16x2(r.y); dst[2] = unpack_f16x2(r.z); dst[3] = unpack_f16x2(r.w);
}
template <typename T> __device__ __forceinline__ void enc_store(T *p, const float2 *v);
template <> __device__ __forceinline__ void enc_store<__half>(__half *p, const float2 *v) {
    uint4 r;
    r.x = pack_f16x2(v[0].x, v[0].y); r.y = pack_f16x2(v[1].x, v[1].y);
    r.z = pack_f16x2(v[2].x, v[2].y); r.w = pack_f16x2(v[3].x, v[3].y);
    *reinterpret_cast<uint4 *>(p) = r;
}
template <> __device__ __forceinline__ void enc_store<float>(float *p, const float2 *v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
}
template <> __device__ __forceinline__ void enc_store<__nv_bfloat16>(__nv_bfloat16 *p, const float2 *v) {
    uint4 r;
    r.x = pack_bf16x2(v[0].x, v[0].y); r.y = pack_bf16x2(v[1].x, v[1].y);
    r.z = pack_bf16x2(v[2].x, v[2].y); r.w = pack_bf16x2(v[3].x, v[3].y);
    *reinterpret_cast<uint4 *>(p) = r;
}

// A thread owns one 16-byte channel vector and walks down a chunk of pixels (coefficients in registers).
// SF / OF: the shortcut / `out` are fp32 although the activations are T -- the bf16 path keeps the RESIDUAL STREAM in
// fp32 (24 blocks of "round the sum to bf16" cost the pipeline 6 dB of PSNR and half of its max-abs budget), while the
// convolution inputs (t_next) and the residual branch (v) stay bf16.
template <typename T, bool SF, bool OF>
__global__ void __launch_bounds__(256) se_residual_kernel(const T *__restrict__ v, const float *__restrict__ gate,
                                                           const void *__restrict__ sc_, int ss, const float *__restrict__ bn_g,
                                                           const float *__restrict__ bn_h, void *__restrict__ out_,
                                                           T *__restrict__ tn, T *__restrict__ out_lp, int H, int W, int C, int64_t chunk) {
    constexpr int N = Vec<T>::N, N2 = N / 2;
    const int b = blockIdx.y;
    const int cv = C / N;
    const int lanes = blockDim.x / cv;
    const int vec = threadIdx.x % cv, lane = threadIdx.x / cv;
    if (lane >= lanes) return;
    const int c = vec * N;
    const int64_t P = (int64_t)H * W;
    float2 g[N2], bg[N2], bh[N2];
#pragma unroll
    for (int j = 0; j < N2; ++j) {
        g[j] = gate ? make_float2(gate[(int64_t)b * C + c + 2 * j], gate[(int64_t)b * C + c + 2 * j + 1]) : f2(1.f);
        bg[j] = tn ? make_float2(bn_g[c + 2 * j], bn_g[c + 2 * j + 1]) : f2(1.f);
        bh[j] = tn ? make_float2(bn_h[c + 2 * j], bn_h[c + 2 * j + 1]) : f2(0.f);
    }
    const int64_t p0 = (int64_t)blockIdx.x * chunk, p1 = min(p0 + chunk, P);
    const int Ws = W * ss;
#pragma unroll 2
    for (int64_t p = p0 + lane; p < p1; p += lanes) {
        const int64_t off = ((int64_t)b * P + p) * C + c;
        float2 x[N2];
        enc_load<T>(v + off, x);
        if (sc_) {
            int64_t soff = off;
            if (ss != 1) {
                const int y = (int)(p / W), xx = (int)(p - (int64_t)y * W);
                soff = (((int64_t)b * H * ss + (int64_t)y * ss) * Ws + (int64_t)xx * ss) * C + c;
            }
            float2 s[N2];
            if (SF) {
#pragma unroll
                for (int q = 0; q < N / 4; ++q) enc_load<float>((const float *)sc_ + soff + 4 * q, s + 2 * q);
            } else {
                enc_load<T>((const T *)sc_ + soff, s);
            }
#pragma unroll
            for (int j = 0; j < N2; ++j) x[j] = fma2(x[j], g[j], s[j]);
        } else {
#pragma unroll
            for (int j = 0; j < N2; ++j) x[j] = mul2(x[j], g[j]);
        }
        if (out_) {
            if (OF) {
#pragma unroll
                for (int q = 0; q < N / 4; ++q) enc_store<float>((float *)out_ + off + 4 * q, x + 2 * q);
            } else {
                enc_store<T>((T *)out_ + off, x);
            }
        }
        if (out_lp) enc_store<T>(out_lp + off, x);       // the residual stream in the storage type: the tapped feature maps
        if (tn) {
#pragma unroll
            for (int j = 0; j < N2; ++j) x[j] = fma2(x[j], bg[j], bh[j]);
            enc_store<T>(tn + off, x);
        }
    }
}

// out[b,k,y,x] = sum over the nine taps of the per-pixel projections (see ood_b200.h: ood_tap_sum); sc (optional) = channels
// 27..29 of the pixel itself: a 1x1 shortcut convolution that rode along in the projection's spare output channels
__global__ void __launch_bounds__(256) tap_sum_kernel(const float *__restrict__ proj, float *__restrict__ out, float *__restrict__ sc,
                                                       int H, int W, int Cp) {
    const int b = blockIdx.y;
    const int64_t P = (int64_t)H * W;
    const float *pb = proj + (int64_t)b * P * Cp;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < (int)P; pix += gridDim.x * blockDim.x) {     // P < 2^31 (host check)
        const int y = pix / W, x = pix - y * W;
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float *q = pb + (int64_t)(yy * W + xx) * Cp + 3 * t;
            acc[0] += __ldg(q); acc[1] += __ldg(q + 1); acc[2] += __ldg(q + 2);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) out[((int64_t)b * 3 + k) * P + pix] = acc[k];
        if (sc) {
#pragma unroll
            for (int k = 0; k < 3; ++k) sc[((int64_t)b * 3 + k) * P + pix] = __ldg(pb + (int64_t)pix * Cp + 27 + k);
        }
    }
}

// Tiled form (Cp % 4 == 0): a block stages the (8+2) x (32+2) pixel neighbourhood of its 8 x 32 outputs in shared memory
// with coalesced 16-byte loads -- every projection line is fetched once per block instead of by nine different threads, three
// scalars at a time (that version was bound by L1 sector requests at 1.0 TB/s) -- and each thread then sums its 27 values from
// shared memory (row pitch 33 words: conflict-free).
constexpr int TSX = 32, TSY = 8;
template <int TSV>                                  // float4 per pixel: 7 (27 channels used) or 8 (+ the shortcut's 27..29)
__global__ void __launch_bounds__(TSX * TSY) tap_sum_tile_kernel(const float *__restrict__ proj, float *__restrict__ out,
                                                                  float *__restrict__ sc, int H, int W, int Cp) {
    __shared__ float sp[(TSY + 2) * (TSX + 2)][33];
    const int b = blockIdx.z, x0 = blockIdx.x * TSX, y0 = blockIdx.y * TSY;
    const int64_t P = (int64_t)H * W;
    const float *pb = proj + (int64_t)b * P * Cp;
    for (int i = threadIdx.x; i < (TSY + 2) * (TSX + 2) * TSV; i += TSX * TSY) {
        const int pixel = i / TSV, v = i - pixel * TSV;
        const int ty = pixel / (TSX + 2), tx = pixel - ty * (TSX + 2);
        const int y = y0 + ty - 1, x = x0 + tx - 1;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) val = __ldg(reinterpret_cast<const float4 *>(pb + ((int64_t)y * W + x) * Cp) + v);
        float *d = &sp[pixel][4 * v];
        d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
    }
    __syncthreads();
    const int ty = threadIdx.x / TSX, tx = threadIdx.x - ty * TSX;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) return;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t) {                  // same summation order as the scalar kernel (out-of-image taps add zero)
        const float *q = sp[(ty + t / 3) * (TSX + 2) + tx + t % 3] + 3 * t;
        acc[0] += q[0]; acc[1] += q[1]; acc[2] += q[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[((int64_t)b * 3 + k) * P + (int64_t)y * W + x] = acc[k];
    if (TSV == 8 && sc) {
        const float *q = sp[(ty + 1) * (TSX + 2) + tx + 1] + 27;
#pragma unroll
        for (int k = 0; k < 3; ++k) sc[((int64_t)b * 3 + k) * P + (int64_t)y * W + x] = q[k];
    }
}

// The whole squeeze-excite tail of a bottleneck in ONE launch (helpers.py:59-76 + :494-501): channel means of v over the pixels, the
// two-layer gate, out = v * gate + shortcut, the next block's BatchNorm and the tapped copy.  Round 1 ran it as four launches
// (statistics partial + finalise, gate, residual) on tensors of 8-33 MB: 96 launches per step whose ramp-up and tail, not their
// bytes, cost ~1.4 ms.  Here a thread-block CLUSTER of 8 CTAs owns one image: every CTA sums its pixel chunk per channel, the
// partial sums are exchanged through distributed shared memory (cluster barrier, fixed rank order: every CTA computes the same
// gate, bit for bit), and the second pass re-reads v (it was just written by the convolution: an L2 hit).
constexpr int kSeCluster = 8;
constexpr int kSeThreads = 1024;      // many pixel lanes per CTA: the passes are chains of dependent L2 round trips, a thread should own few pixels
template <typename T, bool SF>
__global__ void __cluster_dims__(kSeCluster, 1, 1) __launch_bounds__(kSeThreads)
se_tail_kernel(const T *__restrict__ v, const float *__restrict__ w1, const float *__restrict__ w2, const void *__restrict__ sc_, int ss,
               const float *__restrict__ bn_g, const float *__restrict__ bn_h, float *__restrict__ out, T *__restrict__ tn, T *__restrict__ out_lp,
               int H, int W, int C, int Cr) {
    constexpr int N = Vec<T>::N, N2 = N / 2;
    extern __shared__ float se_s[];                   // part[C] | mean[C] | hid[Cr] | gate[C] | red[lanes][C]
    float *part = se_s, *mean = se_s + C, *hid = se_s + 2 * C, *gate_s = hid + Cr, *red = gate_s + C;
    cg::cluster_group cluster = cg::this_cluster();
    const int b = blockIdx.y, rank = (int)cluster.block_rank();
    const int cv = C / N, lanes = blockDim.x / cv;
    const int vec = threadIdx.x % cv, lane = threadIdx.x / cv;
    const int c = vec * N;
    const int64_t P = (int64_t)H * W;
    const int64_t chunk = (P + kSeCluster - 1) / kSeCluster;
    const int64_t p0 = (int64_t)rank * chunk, p1 = min(p0 + chunk, P);
    // ---- pass 1: per-channel sums of this CTA's pixels
    float2 acc[N2];
#pragma unroll
    for (int j = 0; j < N2; ++j) acc[j] = f2(0.f);
    if (lane < lanes) {
#pragma unroll 4
        for (int64_t p = p0 + lane; p < p1; p += lanes) {
            float2 x[N2];
            enc_load<T>(v + ((int64_t)b * P + p) * C + c, x);
#pragma unroll
            for (int j = 0; j < N2; ++j) acc[j] = add2(acc[j], x[j]);
        }
#pragma unroll
        for (int j = 0; j < N2; ++j) { red[lane * C + c + 2 * j] = acc[j].x; red[lane * C + c + 2 * j + 1] = acc[j].y; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[l * C + i];
        part[i] = s;
    }
    cluster.sync();
    // ---- the gate, computed identically by every CTA of the cluster from the 8 partial sums (fixed order)
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < kSeCluster; ++r) s += cluster.map_shared_rank(part, r)[i];
        mean[i] = s / (float)P;
    }
    cluster.sync();                                   // every remote read of `part` is done (no CTA may run ahead and exit)
    {
        const int wlane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int j = warp; j < Cr; j += blockDim.x / 32) {
            const float *wr = w1 + (int64_t)j * C;
            float s = 0.f;
            for (int i = wlane; i < C; i += 32) s = fmaf(wr[i], mean[i], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (wlane == 0) hid[j] = fmaxf(s, 0.f);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            const float *wr = w2 + (int64_t)i * Cr;
            float s = 0.f;
            for (int j = 0; j < Cr; ++j) s = fmaf(wr[j], hid[j], s);
            gate_s[i] = 1.f / (1.f + __expf(-s));
        }
        __syncthreads();
    }
    // ---- pass 2: out = v * gate + shortcut (fp32 stream), next block's BatchNorm, tapped copy
    if (lane >= lanes) return;
    float2 g[N2], bg[N2], bh[N2];
#pragma unroll
    for (int j = 0; j < N2; ++j) {
        g[j] = make_float2(gate_s[c + 2 * j], gate_s[c + 2 * j + 1]);
        bg[j] = tn ? make_float2(bn_g[c + 2 * j], bn_g[c + 2 * j + 1]) : f2(1.f);
        bh[j] = tn ? make_float2(bn_h[c + 2 * j], bn_h[c + 2 * j + 1]) : f2(0.f);
    }
    const int Ws = W * ss;
#pragma unroll 2
    for (int64_t p = p0 + lane; p < p1; p += lanes) {
        const int64_t off = ((int64_t)b * P + p) * C + c;
        float2 x[N2];
        enc_load<T>(v + off, x);
        int64_t soff = off;
        if (ss != 1) {
            const int y = (int)(p / W), xx = (int)(p - (int64_t)y * W);
            soff = (((int64_t)b * H * ss + (int64_t)y * ss) * Ws + (int64_t)xx * ss) * C + c;
        }
        float2 s[N2];
        if (SF) {
#pragma unroll
            for (int q = 0; q < N / 4; ++q) enc_load<float>((const float *)sc_ + soff + 4 * q, s + 2 * q);
        } else {
            enc_load<T>((const T *)sc_ + soff, s);
        }
#pragma unroll
        for (int j = 0; j < N2; ++j) x[j] = fma2(x[j], g[j], s[j]);
#pragma unroll
        for (int q = 0; q < N / 4; ++q) enc_store<float>(out + off + 4 * q, x + 2 * q);
        if (out_lp) enc_store<T>(out_lp + off, x);
        if (tn) {
#pragma unroll
            for (int j = 0; j < N2; ++j) x[j] = fma2(x[j], bg[j], bh[j]);
            enc_store<T>(tn + off, x);
        }
    }
}

// The squeeze-excite tail when the channel SUMS already exist: the convolution that wrote v left per-tile channel sums in its
// statistics workspace (conv3x3 stats_ws, [B][tiles][C][2], conv_tc.cu STATS), so the pooling pass, the cluster and its barriers
// go away and what is left is one streaming pass on a full-width grid.  A thread owns PER (pixel, 8-channel vector) cells: it
// REQUESTS their v / shortcut values first, then every block derives the gate of its image (tile sums -> mean -> two-layer MLP,
// fixed order: all blocks of an image compute the same bits) while those loads are in flight, then applies it.
constexpr int kSaThreads = 512;
template <typename T, bool SF, int PER>
__global__ void __launch_bounds__(kSaThreads, 2)
se_apply_kernel(const T *__restrict__ v, const float2 *__restrict__ partial, int ntiles, const float *__restrict__ w1, const float *__restrict__ w2,
                const void *__restrict__ sc_, int ss, const float *__restrict__ bn_g, const float *__restrict__ bn_h, float *__restrict__ out,
                T *__restrict__ tn, T *__restrict__ out_lp, int H, int W, int C, int Cr) {
    constexpr int N = Vec<T>::N, N2 = N / 2;
    static_assert(N == 8, "16-bit storage");
    extern __shared__ float sa_s[];                   // mean[C] | hid[Cr] | gate[C] | red[groups][C]
    float *mean = sa_s, *hid = sa_s + C, *gate_s = hid + Cr, *red = gate_s + C;
    const int b = blockIdx.y;
    const int cv = C / N, lanes = kSaThreads / cv;
    const int vec = threadIdx.x % cv, lane = threadIdx.x / cv;
    const int c = vec * N;
    const int P = H * W, Ws = W * ss;
    // ---- 1. request this thread's cells (kept packed until the gate exists: 4 + 8 registers per cell)
    uint4 xr[PER], sr[PER][SF ? 2 : 1];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int p = (blockIdx.x * PER + k) * lanes + lane;
        if (p < P) {
            const int64_t off = ((int64_t)b * P + p) * C + c;
            xr[k] = __ldg(reinterpret_cast<const uint4 *>(v + off));
            int64_t soff = off;
            if (ss != 1) {
                const int y = p / W, xx = p - y * W;
                soff = (((int64_t)b * H * ss + (int64_t)y * ss) * Ws + (int64_t)xx * ss) * C + c;
            }
            if (SF) {
                sr[k][0] = __ldg(reinterpret_cast<const uint4 *>((const float *)sc_ + soff));
                sr[k][SF ? 1 : 0] = __ldg(reinterpret_cast<const uint4 *>((const float *)sc_ + soff + 4));
            } else {
                sr[k][0] = __ldg(reinterpret_cast<const uint4 *>((const T *)sc_ + soff));
            }
        }
    }
    // the gate's weights are cold (every bottleneck has its own): ask L2 for them now, they are needed two barriers from here
    for (int i = threadIdx.x * 32; i < C * Cr; i += kSaThreads * 32) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(w1 + i));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(w2 + i));
    }
    // ---- 2. channel means from the convolution's tile sums
    {
        const int groups = kSaThreads / C, i = threadIdx.x % C, tg = threadIdx.x / C;          // C <= 512 divides 512 (host check)
        float a = 0.f;
#pragma unroll 4
        for (int t = tg; t < ntiles; t += groups) a += __ldg(&partial[((int64_t)b * ntiles + t) * C + i].x);
        red[tg * C + i] = a;
        __syncthreads();
        if (threadIdx.x < C) {
            float m = 0.f;
            for (int g = 0; g < groups; ++g) m += red[g * C + threadIdx.x];
            mean[threadIdx.x] = m / (float)P;
        }
        __syncthreads();
    }
    // ---- 3. the gate (SEModule, helpers.py:59-76)
    {
        const int wlane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int j = warp; j < Cr; j += kSaThreads / 32) {
            const float *wr = w1 + (int64_t)j * C;
            float a = 0.f;
            for (int i = wlane; i < C; i += 32) a = fmaf(wr[i], mean[i], a);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (wlane == 0) hid[j] = fmaxf(a, 0.f);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += kSaThreads) {
            const float *wr = w2 + (int64_t)i * Cr;
            float a = 0.f;
            for (int j = 0; j < Cr; ++j) a = fmaf(wr[j], hid[j], a);
            gate_s[i] = 1.f / (1.f + __expf(-a));
        }
        __syncthreads();
    }
    // ---- 4. out = v * gate + shortcut (fp32 stream), next block's BatchNorm, tapped copy
    float2 g[N2], bg[N2], bh[N2];
#pragma unroll
    for (int j = 0; j < N2; ++j) {
        g[j] = make_float2(gate_s[c + 2 * j], gate_s[c + 2 * j + 1]);
        bg[j] = tn ? make_float2(bn_g[c + 2 * j], bn_g[c + 2 * j + 1]) : f2(1.f);
        bh[j] = tn ? make_float2(bn_h[c + 2 * j], bn_h[c + 2 * j + 1]) : f2(0.f);
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int p = (blockIdx.x * PER + k) * lanes + lane;
        if (p >= P) continue;
        const int64_t off = ((int64_t)b * P + p) * C + c;
        float2 x[N2], s[N2];
        unpack8<T>(xr[k], x);
        if (SF) {
            s[0] = make_float2(__uint_as_float(sr[k][0].x), __uint_as_float(sr[k][0].y));
            s[1] = make_float2(__uint_as_float(sr[k][0].z), __uint_as_float(sr[k][0].w));
            s[2] = make_float2(__uint_as_float(sr[k][SF ? 1 : 0].x), __uint_as_float(sr[k][SF ? 1 : 0].y));
            s[3] = make_float2(__uint_as_float(sr[k][SF ? 1 : 0].z), __uint_as_float(sr[k][SF ? 1 : 0].w));
        } else {
            unpack8<T>(sr[k][0], s);
        }
#pragma unroll
        for (int j = 0; j < N2; ++j) x[j] = fma2(x[j], g[j], s[j]);
#pragma unroll
        for (int q = 0; q < N / 4; ++q) enc_store<float>(out + off + 4 * q, x + 2 * q);
        if (out_lp) enc_store<T>(out_lp + off, x);
        if (tn) {
#pragma unroll
            for (int j = 0; j < N2; ++j) x[j] = fma2(x[j], bg[j], bh[j]);
            enc_store<T>(tn + off, x);
        }
    }
}

// W+ assembly of Encoder4Editing.forward (psp_encoders.py:199-214) and of the arch (OOD_faceGAN_e4e_arch.py:261):
//   w[b,0] = head_0;  w[b,i] = head_0 + head_i for 1 <= i <= stage, head_0 beyond;  out = w + avg[d] + delta[i,d]
// heads: [n_styles][B][D] fp32 (the grouped EqualLinear outputs, one row block per head; heads beyond `stage` are not read).
__global__ void __launch_bounds__(256) latent_assemble_kernel(const float *__restrict__ heads, const float *__restrict__ avg,
                                                               const float *__restrict__ delta, float *__restrict__ out, int B, int n, int D,
                                                               int stage) {
    const int64_t total = (int64_t)B * n * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D), st = (int)((i / D) % n), b = (int)(i / ((int64_t)D * n));
        float v = heads[(int64_t)b * D + d];
        if (st >= 1 && st <= stage) v += heads[((int64_t)st * B + b) * D + d];
        if (avg) v += avg[d];
        if (delta) v += delta[(int64_t)st * D + d];
        out[i] = v;
    }
}

static int launch_tap_sum(const float *proj, float *out, float *sc, int batch, int h, int w, int cp, cudaStream_t st) {
    OOD_REQUIRE(proj && out && batch > 0 && batch <= 65535 && h > 0 && w > 0 && cp >= (sc ? 30 : 27) && (int64_t)h * w < (1LL << 30),
                "tap_sum: bad arguments");
    const int64_t P = (int64_t)h * w;
    if (cp % 4 == 0 && ((uintptr_t)proj % 16) == 0 && ceil_div(h, TSY) <= 65535) {
        dim3 tg(ceil_div(w, TSX), ceil_div(h, TSY), batch);
        if (sc) tap_sum_tile_kernel<8><<<tg, TSX * TSY, 0, st>>>(proj, out, sc, h, w, cp);
        else tap_sum_tile_kernel<7><<<tg, TSX * TSY, 0, st>>>(proj, out, nullptr, h, w, cp);
        return check_launch("tap_sum");
    }
    dim3 grid((unsigned)std::min<int64_t>((P + 255) / 256, kNumSMs * 16), batch);
    tap_sum_kernel<<<grid, 256, 0, st>>>(proj, out, sc, h, w, cp);
    return check_launch("tap_sum");
}

}  // namespace ood

extern "C" int ood_tap_sum(const float *proj, float *out, int batch, int h, int w, int cp, void *stream) {
    return ood::launch_tap_sum(proj, out, nullptr, batch, h, w, cp, (cudaStream_t)stream);
}

extern "C" int ood_tap_sum_shortcut(const float *proj, float *out, float *shortcut, int batch, int h, int w, int cp, void *stream) {
    using namespace ood;
    OOD_REQUIRE(shortcut, "tap_sum_shortcut: null pointer");
    return launch_tap_sum(proj, out, shortcut, batch, h, w, cp, (cudaStream_t)stream);
}

extern "C" int ood_se_gate(const float *stats, const float *w1, const float *w2, float *gate, int batch, int channels,
                           int reduced, void *stream) {
    using namespace ood;
    OOD_REQUIRE(stats && w1 && w2 && gate && batch > 0 && channels > 0 && reduced > 0, "se_gate: bad arguments");
    const size_t smem = (size_t)(channels + reduced) * sizeof(float);
    OOD_REQUIRE(smem <= 48 * 1024, "se_gate: too many channels (%d)", channels);
    se_gate_kernel<<<batch, 256, smem, (cudaStream_t)stream>>>(stats, w1, w2, gate, channels, reduced);
    return check_launch("se_gate");
}

extern "C" int ood_se_residual(const void *v, const float *gate, const void *shortcut, int sc_stride, const float *bn_g,
                               const float *bn_h, void *out, void *t_next, void *out_lp, int batch, int h, int w, int channels, int dtype,
                               int shortcut_f32, int out_f32, void *stream) {
    using namespace ood;
    OOD_REQUIRE(v && (out || t_next || out_lp) && batch > 0 && batch <= 65535 && h > 0 && w > 0, "se_residual: bad arguments");
    OOD_REQUIRE(dtype == OOD_F32 || dtype == OOD_BF16 || dtype == OOD_F16, "se_residual: bad dtype");
    OOD_REQUIRE(sc_stride == 1 || sc_stride == 2, "se_residual: shortcut stride must be 1 or 2");
    OOD_REQUIRE(!t_next || (bn_g && bn_h), "se_residual: t_next needs the affine coefficients");
    const int N = dtype == OOD_F32 ? 4 : 8;
    OOD_REQUIRE(channels % N == 0 && channels / N <= 256, "se_residual: channels (%d) must be a multiple of %d and at most %d", channels, N, 256 * N);
    const int64_t P = (int64_t)h * w;
    const int lanes = std::max(1, 256 / (channels / N));
    const int64_t want_blocks = std::max<int64_t>(1, (int64_t)kNumSMs * pixwalk_blocks_per_sm(8) / batch);
    int64_t chunk = std::max<int64_t>((P + want_blocks - 1) / want_blocks, (int64_t)lanes * 4);
    chunk = (chunk + lanes - 1) / lanes * lanes;
    dim3 grid((unsigned)((P + chunk - 1) / chunk), batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OOD_F32) {
        se_residual_kernel<float, false, false><<<grid, 256, 0, st>>>((const float *)v, gate, shortcut, sc_stride, bn_g, bn_h, out, (float *)t_next, (float *)out_lp, h, w, channels, chunk);
    } else if (dtype == OOD_F16) {
        const __half *vv = (const __half *)v;
        __half *tn = (__half *)t_next;
#define OOD_SE(SF, OF) se_residual_kernel<__half, SF, OF><<<grid, 256, 0, st>>>(vv, gate, shortcut, sc_stride, bn_g, bn_h, out, tn, (__half *)out_lp, h, w, channels, chunk)
        if (shortcut_f32 && out_f32) OOD_SE(true, true);
        else if (shortcut_f32) OOD_SE(true, false);
        else if (out_f32) OOD_SE(false, true);
        else OOD_SE(false, false);
#undef OOD_SE
    } else {
        const __nv_bfloat16 *vv = (const __nv_bfloat16 *)v;
        __nv_bfloat16 *tn = (__nv_bfloat16 *)t_next;
#define OOD_SE(SF, OF) se_residual_kernel<__nv_bfloat16, SF, OF><<<grid, 256, 0, st>>>(vv, gate, shortcut, sc_stride, bn_g, bn_h, out, tn, (__nv_bfloat16 *)out_lp, h, w, channels, chunk)
        if (shortcut_f32 && out_f32) OOD_SE(true, true);
        else if (shortcut_f32) OOD_SE(true, false);
        else if (out_f32) OOD_SE(false, true);
        else OOD_SE(false, false);
#undef OOD_SE
    }
    return check_launch("se_residual");
}

extern "C" int ood_latent_assemble(const float *heads, const float *avg, const float *delta, float *out, int batch, int n_styles, int dim,
                                   int stage, void *stream) {
    using namespace ood;
    OOD_REQUIRE(heads && out && batch > 0 && n_styles > 0 && dim > 0 && stage >= 0, "latent_assemble: bad arguments");
    const int64_t total = (int64_t)batch * n_styles * dim;
    latent_assemble_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(heads, avg, delta, out, batch,
                                                                                                                   n_styles, dim, stage);
    return check_launch("latent_assemble");
}

extern "C" int ood_se_tail(const void *v, const float *w1, const float *w2, const void *shortcut, int sc_stride, const float *bn_g, const float *bn_h,
                           float *out, void *t_next, void *out_lp, int batch, int h, int w, int channels, int reduced, int dtype, int shortcut_f32,
                           void *stream) {
    using namespace ood;
    OOD_REQUIRE(v && w1 && w2 && shortcut && out && batch > 0 && batch <= 65535 && h > 0 && w > 0 && reduced > 0, "se_tail: bad arguments");
    OOD_REQUIRE(dtype == OOD_BF16 || dtype == OOD_F16, "se_tail: storage type must be bf16 or f16 (fp32 residual stream out)");
    OOD_REQUIRE(sc_stride == 1 || sc_stride == 2, "se_tail: shortcut stride must be 1 or 2");
    OOD_REQUIRE(!t_next || (bn_g && bn_h), "se_tail: t_next needs the affine coefficients");
    OOD_REQUIRE(channels % 8 == 0 && channels / 8 <= 256 && 256 % (channels / 8) == 0, "se_tail: channels (%d) must be 8 * a divisor of 256", channels);
    const int lanes = kSeThreads / (channels / 8);
    const size_t smem = (size_t)(3 * channels + reduced + lanes * channels) * sizeof(float);
    OOD_REQUIRE(smem <= 200 * 1024, "se_tail: too many channels (%d)", channels);
    dim3 grid(kSeCluster, batch);
    cudaStream_t st = (cudaStream_t)stream;
#define OOD_SET(T, SF)                                                                                                              \
    do {                                                                                                                            \
        auto kern = se_tail_kernel<T, SF>;                                                                                          \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                   \
        kern<<<grid, kSeThreads, smem, st>>>((const T *)v, w1, w2, shortcut, sc_stride, bn_g, bn_h, out, (T *)t_next, (T *)out_lp, h, w, channels, reduced); \
    } while (0)
    if (dtype == OOD_F16) { if (shortcut_f32) OOD_SET(__half, true); else OOD_SET(__half, false); }
    else { if (shortcut_f32) OOD_SET(__nv_bfloat16, true); else OOD_SET(__nv_bfloat16, false); }
#undef OOD_SET
    return check_launch("se_tail");
}

extern "C" int ood_se_apply(const void *v, const float *tile_sums, int tiles, const float *w1, const float *w2, const void *shortcut, int sc_stride,
                            const float *bn_g, const float *bn_h, float *out, void *t_next, void *out_lp, int batch, int h, int w, int channels,
                            int reduced, int dtype, int shortcut_f32, void *stream) {
    using namespace ood;
    OOD_REQUIRE(v && tile_sums && tiles > 0 && w1 && w2 && shortcut && out && batch > 0 && batch <= 65535 && h > 0 && w > 0 && reduced > 0,
                "se_apply: bad arguments");
    OOD_REQUIRE(dtype == OOD_BF16 || dtype == OOD_F16, "se_apply: storage type must be bf16 or f16 (fp32 residual stream out)");
    OOD_REQUIRE(sc_stride == 1 || sc_stride == 2, "se_apply: shortcut stride must be 1 or 2");
    OOD_REQUIRE(!t_next || (bn_g && bn_h), "se_apply: t_next needs the affine coefficients");
    OOD_REQUIRE(channels >= 64 && channels <= kSaThreads && kSaThreads % channels == 0, "se_apply: channels (%d) must be 64, 128, 256 or 512", channels);
    OOD_REQUIRE((int64_t)h * w < (1LL << 30), "se_apply: image too large");
    const int lanes = kSaThreads / (channels / 8);
    const int P = h * w;
    const size_t smem = (size_t)(2 * channels + reduced + (kSaThreads / channels) * channels) * sizeof(float);
    // two cells per thread while that still is one resident wave (2 blocks of 512 threads per SM), four otherwise
    int per = (int64_t)ceil_div(P, lanes * 2) * batch <= 2 * kNumSMs ? 2 : 4;
    if (const char *e = getenv("OOD_SE_APPLY_PER")) per = atoi(e) == 2 ? 2 : 4;          // measurement switch
    dim3 grid(ceil_div(P, lanes * per), batch);
    cudaStream_t st = (cudaStream_t)stream;
#define OOD_SA(T, SF, PER)                                                                                                               \
    se_apply_kernel<T, SF, PER><<<grid, kSaThreads, smem, st>>>((const T *)v, (const float2 *)tile_sums, tiles, w1, w2, shortcut, sc_stride, bn_g, \
                                                                bn_h, out, (T *)t_next, (T *)out_lp, h, w, channels, reduced)
#define OOD_SA2(T, SF) do { if (per == 2) OOD_SA(T, SF, 2); else OOD_SA(T, SF, 4); } while (0)
    if (dtype == OOD_F16) { if (shortcut_f32) OOD_SA2(__half, true); else OOD_SA2(__half, false); }
    else { if (shortcut_f32) OOD_SA2(__nv_bfloat16, true); else OOD_SA2(__nv_bfloat16, false); }
#undef OOD_SA2
#undef OOD_SA
    return check_launch("se_apply");
}
