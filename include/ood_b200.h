/*
 * ood_b200.h -- C ABI of libood_b200.so: the B200 (sm_100a) hot path of OOD-GAN-inversion.
 *
 * Drop-in boundary.  The reference reaches its native code through two pybind/torch extensions
 * ("boundary #2"):
 *     upfirdn2d_op.upfirdn2d(input[N,H,W,minor], kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
 *                                                  -- /root/reference/src/ops/op/upfirdn2d.cpp:12-23
 *     fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *                                                  -- /root/reference/src/ops/op/fused_bias_act.cpp:11-21
 * and everything else on the path (modulated conv, ToRGB, warp, mask, blend) through ATen library calls
 * made by src/ops/StyleGAN/model.py, src/ops/SAMM/helpers.py and src/archs/OOD_faceGAN_e4e_arch.py.
 * This header replaces both: plain pointers and sizes, no torch types, no exceptions across the boundary.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns and pre-allocates every buffer (no hidden allocation, no hidden synchronisation);
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - return value 0 = launched; negative = error, text in ood_last_error() (thread local);
 *   - dtype: OOD_F32 or OOD_BF16 is the STORAGE type of activations; arithmetic is always fp32
 *     (tensor-core paths: bf16 x bf16 -> fp32 accumulate);
 *   - "NHWC" activations are [B][H][W][C] with C innermost (== torch channels_last of a [B,C,H,W] tensor);
 *     "NCHW" planes are [B*C][H][W].
 */
#ifndef OOD_B200_H_
#define OOD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OOD_F32 0
#define OOD_BF16 1
#define OOD_F16 2   /* IEEE half storage: the encoder path (normalised activations); conv3x3 impl 0, in_stats, se_residual, bicubic_up_add, thumbnail_nhwc, pack_conv_weight */

#define OOD_OK 0
#define OOD_ERR_ARG (-1)      /* bad argument / unsupported configuration (never a silent no-op) */
#define OOD_ERR_CUDA (-2)     /* CUDA runtime / driver error at launch                          */
#define OOD_ERR_DEVICE (-3)   /* device is not sm_100                                            */

int ood_version(void);
const char *ood_last_error(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMEM/TMA paths usable). */
int ood_device_is_sm100(void);
/* number of kernels this library has launched in this process (monotonic; bench.py reports the difference). */
unsigned long long ood_launch_count(void);
/* which kernel family the calling thread's last ood_conv3x3 ran on: 0 generic tcgen05 tiles (conv_tc.cu), 1 row-sliding kernel (conv_rows.cu),
 * 2 row-streaming transposed kernel for the interior (convt_rows.cu), 3 SIMT.  For per-kernel accounting (bench.py), not for control flow. */
int ood_last_conv_route(void);

/* ---- a1. upfirdn2d on NCHW planes: replaces upfirdn2d_op.upfirdn2d (upfirdn2d.cpp:12-23) with minor==1,
 *      which is the only form the Python wrapper issues (upfirdn2d.py:103).  Generic in every parameter
 *      (the reference kernel silently returns garbage for unsupported modes, SURVEY finding 11).
 *      out_h = (in_h*up_y + pad_y0 + pad_y1 - kh)/down_y + 1, likewise out_w.  kernel: fp32 [kh][kw]. */
int ood_upfirdn2d(const void *in, void *out, const float *kernel, int64_t planes, int in_h, int in_w,
                  int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                  int pad_y0, int pad_y1, int dtype, void *stream);

/* ---- a2. fused bias + leaky-ReLU: replaces fused.fused_bias_act (fused_bias_act.cpp:11-21).
 *      act must be 3 (leaky relu).  grad 0: out = lrelu(in + bias[(i/inner)%channels], alpha)*scale
 *      grad 1: out = in * (refer>0 ? 1 : alpha) * scale        (bias ignored; refer = saved forward output)
 *      grad 2: out = 0.                                       bias / refer may be NULL where unused. */
int ood_fused_bias_act(const void *in, const void *bias, const void *refer, void *out, int64_t numel,
                       int channels, int64_t inner, int act, int grad, float alpha, float scale, int dtype,
                       void *stream);
/* grad_bias[c] = sum over batch and inner of g[b][c][inner]; fp32 output.  (fused_act.py:38-43) */
int ood_bias_grad(const void *g, float *grad_bias, int64_t batch, int channels, int64_t inner, int dtype,
                  void *stream);

/* ---- a13: the encoder's input thumbnail.  OOD_faceGAN_e4e_arch.py:256 `F.interpolate(x, (256, 256), mode='bilinear')`
 *      (align_corners=False) fused with the layout change, cast and channel padding of the encoder's first convolution operand:
 *      in fp32 NCHW [B,3,H,W] -> out `dtype` NHWC [B,oh,ow,cp] with channels 3..cp-1 zero (cp = 32).  ATen's arithmetic
 *      (upsample_bilinear2d): src = scale*(dst+0.5)-0.5 clamped at 0; at 1024 -> 256 the mean of the 2x2 centre pixels of a 4x4 block. */
int ood_thumbnail_nhwc(const float *in, void *out, int batch, int channels, int h, int w, int oh, int ow, int cp, int dtype,
                       void *stream);

/* ---- layout: NCHW fp32 <-> NHWC storage type, optional per-(b,c) scale (the style modulation of the NEXT conv).
 *      in_batch_stride 0 broadcasts one sample (ConstantInput, model.py:295-305). */
int ood_nchw_to_nhwc(const float *in, int64_t in_batch_stride, const float *scale_bc, void *out, int batch,
                     int channels, int h, int w, int dtype, void *stream);
int ood_nhwc_to_nchw(const void *in, float *out, int batch, int channels, int h, int w, int dtype, void *stream);
/* out[b,p,c] = in[b,p,c] * scale_bc[b,c]   (NHWC, either storage type) */
int ood_nhwc_scale(const void *in, const float *scale_bc, void *out, int batch, int channels, int64_t pixels,
                   int dtype, void *stream);

/* ---- a4/a5. style modulation and demodulation coefficients (model.py:129-163, 236-241).
 *      s[b,i]  = sum_j latent[b,j]*mod_w[i,j]/sqrt(style_dim) + mod_b[i]
 *      d[b,o]  = conv_scale * rsqrt(conv_scale^2 * sum_i s[b,i]^2 * wsq[o,i] + 1e-8)   (demodulate)
 *              = conv_scale                                                             (wsq == NULL)
 *      wsq[o,i] = sum_k W[o,i,k]^2 (ood_weight_sumsq).  latent rows are latent + b*latent_stride. */
int ood_modulation(const float *latent, int64_t latent_stride, const float *mod_w, const float *mod_b,
                   const float *wsq, float conv_scale, float *s_out, float *d_out, int batch, int style_dim,
                   int cin, int cout, void *stream);
int ood_weight_sumsq(const float *w /*[Co][Ci][k*k]*/, float *wsq /*[Co][Ci]*/, int cout, int cin, int taps,
                     void *stream);
/* weight repack W[Co][Ci][ky][kx] fp32 -> [tap][Co][Ci] in `dtype` (tensor-core B operand, K-major)
 *                                     or -> [tap][Ci][Co] fp32 when ci_major_out != 0 (SIMT path). */
int ood_pack_conv_weight(const float *w, void *out, int cout, int cin, int taps, int ci_major_out, int dtype,
                         void *stream);

/* ---- a5. 3x3 convolution of pre-modulated NHWC activations with SHARED weights.
 *      Uses (W*s*d) (*) x == d . (W (*) (s . x))  (SURVEY.md section 7 step 4): `in` already carries s.
 *      transposed == 0: stride-1, pad-1 conv, out [B,H,W,Co]           (model.py:268-272)
 *      transposed == 1: stride-2 transposed conv, out [B,2H+1,2W+1,Co]  (model.py:246-256), raw accumulators.
 *      transposed == 2: stride-2 valid conv, in [B,2h+1,2w+1,Ci] -> out [B,h,w,Co]: the data gradient of form 1 (autograd of model.py:255).
 *      transposed == 3: stride-2, pad-1 conv, out [B,(H-1)/2+1,(W-1)/2+1,Co]: the encoder's down-sampling convolutions
 *                       (src/ops/e4e/encoders/psp_encoders.py:41-48, helpers.py:488-491).
 *      transposed == 4: 1x1 convolution, weight pack [1][Co][Ci] (bf16, K-major): lateral / feature convolutions
 *                       (psp_encoders.py:153-154, e4e_arch.py feats_conv) and the per-tap projection consumed by ood_tap_sum.
 *      transposed == 6: 1x1 stride-2 convolution, out [B,(H-1)/2+1,(W-1)/2+1,Co], weight pack [1][Co][Ci]: the shortcut
 *                       convolutions of the encoder's down-sampling bottlenecks (e4e helpers.py:483-486).
 *      transposed == 5: form 1 with the four output-parity phases fused into one GEMM (tcgen05 path; raw accumulators like
 *                       form 1).  Weight pack [4 shifts][4*Co][Ci] bf16, shift t = (dy, dx) = (-(t>>1), -(t&1)), row
 *                       (2*py + px)*Co + o holds W[o][:][ky][kx] of the tap of output parity (py, px) that reads input
 *                       (oy+dy, ox+dx) -- ky = py if dy == 0, ky = 2 if dy == -1 and py == 0, else the row is zero; kx alike.
 *                       For the 512 / 1024 px layers (Co <= 64): one read of the input patch per shift instead of per tap,
 *                       2x2 output pixel blocks per position, a quarter of the tiles.
 *      Epilogue (stride-1 only; any pointer may be NULL to skip that term):
 *          v  = acc * d[b,o] + noise_w * noise[b,y,x] + bias[o];  y = act ? lrelu(v,0.2)*sqrt2 : v
 *          out_y  = y            out_ys = y * s_next[b,o]
 *      impl 0: tcgen05/TMEM/TMA implicit GEMM (dtype OOD_BF16 or OOD_F16, weights from ood_pack_conv_weight in the same type)
 *      impl 1: fp32 SIMT implicit GEMM (dtype OOD_F32, weights packed ci_major_out=1).
 *      noise: fp32 [B or 1][H][W] with batch stride noise_bstride (0 = shared).  noise_w: device scalar. */
typedef struct {
    const void *in;        /* NHWC [B,H,W,Ci], pre-modulated */
    const void *weight;    /* packed */
    void *out_y;           /* NHWC [B,OH,OW,Co] or NULL */
    void *out_ys;          /* NHWC or NULL */
    const float *d;        /* [B,Co] or NULL (=1) */
    const float *noise;    /* or NULL */
    int64_t noise_bstride;
    const float *noise_w;  /* device scalar or NULL */
    const float *bias;     /* [Co] or NULL */
    const float *s_next;   /* [B,Co]; required iff out_ys */
    int batch, h, w, cin, cout;
    int transposed;        /* 0 stride-1 conv | 1 stride-2 transposed conv | 2 stride-2 valid conv (data gradient of 1) | 3 stride-2 pad-1 conv | 4 1x1 conv | 5 = 1 with fused phases | 6 1x1 stride-2 conv */
    int act;               /* 0 none | 1 leaky-ReLU(0.2)*sqrt2 | 2 PReLU(prelu_slope[o]) (AlignNet, shared-weight mode) */
    int impl;              /* 0 tcgen05 | 1 simt */
    int dtype;             /* storage type of in / out */
    int out_f32;           /* 1: out_y is fp32 regardless of dtype (raw accumulators of the transposed conv) */
    const float *prelu_slope; /* [Co], act == 2 */
    /* fused ToRGB (tcgen05 path, stride-1, Co <= 256): rgb_out[B,3,H,W] fp32 = sum_o y[o]*rgb_w[b,k,o] + rgb_bias[k] + up2fir(rgb_skip);
     * rgb_w from ood_torgb_weight, rgb_skip [B,3,H/2,W/2] fp32 or NULL, rgb_taps = 1-D up-FIR.  out_y / out_ys may then be NULL. */
    const float *rgb_w, *rgb_bias, *rgb_skip;
    float *rgb_out;
    float rgb_taps[4];
    /* grouped form (tcgen05 path; the 18 style heads of the E4E encoder, psp_encoders.py:34-56,207-214): `batch` = groups * images;
     * output image g*images + i is the convolution of input image g*images + i (or of image i of a [images,...] input shared by
     * all groups when in_shared = 1) with weights [g] of a [groups][9][Co][Ci] pack; bias / prelu_slope are [groups][Co].
     * Epilogue: bias + activation only.  groups <= 1: the plain form. */
    int groups;
    int in_shared;
    /* accumulator seed (tcgen05 path, every form except transposed == 1): fp32 NHWC [B,OH,OW,Co] added to the accumulator
     * before d / bias / noise / activation.  A convolution whose input channels split into a part that changes between calls
     * and a part that does not (AlignNet: cat[IN(cur)-IN(enc), IN(enc)], SAMM/helpers.py:96-101, over the alignment cycles
     * :154-166) runs the constant part once (out_f32 = 1, no epilogue terms) and seeds every later call with it. */
    const float *acc_in;
    /* tiled = 1: the fp32 tensors of this call -- acc_in, and out_y when out_f32 = 1 -- are not NHWC but in the kernel's own
     * tile order (ood_conv3x3_tiled_bytes() bytes; opaque, valid between calls with the same batch, h, w, cout and form).  An
     * epilogue thread owns one pixel and 32 consecutive channels of it, so NHWC fp32 costs 32 separate 16-byte transactions
     * per warp instruction; the tile order makes the same access one contiguous 512-byte run (measured on the 256 px
     * AlignNet level: seeded half convolution 898 us with an NHWC seed against 505 us without any seed). */
    int tiled;
    /* fused output statistics (tcgen05 path; stride-1 / stride-2 pad-1 / 1x1 forms, out_y only (bf16 or f16), cin % 64 == 0, cout % 128 == 0, h*w >= 128):
     * stats_out [B,Co,2] = {mean, rsqrt(biased variance + stats_eps)} over the pixels of every output channel of out_y AS
     * STORED -- the statistics of the InstanceNorm2d that follows the convolution (bottleneck_IR res_layer[4],
     * e4e/encoders/helpers.py:436-441) from the epilogue instead of a pass over out_y.  Deterministic (fixed-order sums).
     * stats_ws: ood_conv3x3_stats_workspace() bytes.  stats_ws without stats_out: only the per-tile sums [B][tiles][Co][2] = {sum, 0}
     * are left in the workspace (no second moment, no finalise launch) -- the pooled sums of the encoder's squeeze-excite tail (ood_se_apply). */
    float *stats_out, *stats_ws;
    float stats_eps;
    /* storage type of out_y / out_ys when it differs from `dtype` (tcgen05 path): 0 = same as dtype, OOD_BF16 or OOD_F16.  The
     * encoder's features are half precision, the pipeline that consumes feats_conv's output is bf16 (e4e_arch.py:109-113). */
    int out_dtype;
} ood_conv3x3_args;
int ood_conv3x3(const ood_conv3x3_args *args_host, void *stream);
/* bytes of a tile-order fp32 tensor for the tcgen05 path (0 when the arguments are outside that path) */
int64_t ood_conv3x3_tiled_bytes(int batch, int h, int w, int cin, int cout, int transposed);
/* bytes of stats_ws (0 when the arguments are outside the fused-statistics envelope) */
int64_t ood_conv3x3_stats_workspace(int batch, int h, int w, int cin, int cout, int transposed);

/* ---- a1+a2+a6 fused: FIR blur (4x4 separable taps, pad (1,1)) of the transposed-conv output
 *      [B,2h+1,2w+1,C] -> [B,2h,2w,C], then the StyledConv epilogue (model.py:257, 283-292, fused_act.py:96):
 *          img = blur(t) * d[b,c]
 *          out_img (optional) = img                      (what the alignment callback sees, model.py:288-290)
 *          v = img + noise_w*noise + bias ; y = lrelu(v)*sqrt2 ; out_y = y ; out_ys = y*s_next[b,c]
 *      taps: 4 fp32 HOST values of the 1-D FIR (already including the gain, e.g. [1,3,3,1]/8*2). */
typedef struct {
    const void *in;       /* NHWC [B,IH,IW,C] */
    int in_f32;           /* 1: `in` is fp32 regardless of dtype */
    void *out_img;        /* NHWC [B,IH-1,IW-1,C] storage dtype, or NULL */
    void *out_y, *out_ys; /* NHWC or NULL */
    const float *d, *noise, *noise_w, *bias, *s_next;
    int64_t noise_bstride;
    float taps[4];
    int batch, ih, iw, channels;
    int act;              /* 0: stop after img (out_y/out_ys must be NULL) */
    int dtype;
    int pad0, pad1;       /* FIR pads (same on both axes); 0,0 means (1,1).  (2,2) is the adjoint of the (1,1) blur: out = in + 1 */
} ood_blur_act_args;
int ood_blur_act(const ood_blur_act_args *args_host, void *stream);

/* ---- a6+a2 on NHWC: y = lrelu(img + noise_w*noise + bias)*sqrt2 ; out_y, out_ys as above.
 *      Used after the alignment callback replaced `img` (e4e_arch.py:224-242: image := aligned + w*noise). */
int ood_noise_act(const void *img, void *out_y, void *out_ys, const float *noise, int64_t noise_bstride,
                  const float *noise_w, const float *bias, const float *s_next, int batch, int64_t pixels,
                  int channels, int dtype, void *stream);

/* ---- a8. ToRGB: 1x1 modulated conv without demodulation + bias + FIR-upsampled skip (model.py:363-372).
 *      wrgb[b,k,c] = W[k,c]*s[b,c]*scale (fp32, from ood_torgb_weight; scale = 1/sqrt(C)); y is the UNscaled NHWC activation.
 *      out[b,k,Y,X] = sum_c y[b,Y,X,c]*wrgb[b,k,c] + bias[k] + up2fir(skip)[b,k,Y,X]     (NCHW fp32)
 *      skip: [B,3,H/2,W/2] fp32 or NULL.  taps_up: 4 host values of the 1-D up-FIR ([1,3,3,1]/8*2). */
int ood_torgb_weight(const float *w /*[3][C]*/, const float *s /*[B][C]*/, float *wrgb, float scale /* 1/sqrt(fan_in) */,
                     int batch, int channels, void *stream);
int ood_torgb(const void *y, const float *wrgb, const float *bias, const float *skip, float *out,
              const float *taps_up_host, int batch, int h, int w, int channels, int dtype, void *stream);

/* ---- a11. alignment-field step (SAMM/helpers.py:104-107, 129-147, 154-166), all fp32 NCHW [B,3,R,R]:
 *      f = blur_{pad(2,1)}([tanh(z0)*scale, tanh(z1)*scale, sigmoid(z2)])   (taps: 4 host values, [1,3,3,1]/8)
 *      prev == NULL: acc = f
 *      else acc = [clip(prev0+f0,+-scale), clip(prev1+f1,+-scale), clip(PRM(prev2, f2),0,1)]
 *      coarse != NULL (last cycle of a finer level): acc2 = clip(PRM(bicubic_up(coarse2), acc2),0,1)
 *      PRM(x,y) = y*x + x*(1-x).  coarse is [B,3,Rc,Rc]; bicubic align_corners=True (helpers.py:69-70). */
int ood_field_step(const float *z, const float *prev, const float *coarse, float *acc, const float *taps_host,
                   float scale, int batch, int r, int rc, const float *z2, const float *coef, void *stream);
/*      z2 / coef [B,3,3] (both or neither): the pre-activation field is z*coef[b,ch,0] + z2*coef[b,ch,1] + coef[b,ch,2].
 *
 *      ood_alignnet_tail (bottleneck_IR(2C -> 3) tail, e4e/encoders/helpers.py:426-448 inside SAMM/helpers.py:97-101):
 *      r2 = conv3x3(PReLU(res; slope[3]); conv_w [3,3,3,3], pad 1), and coef such that
 *          InstanceNorm(r2; in_res_w, in_res_b) + InstanceNorm(shortcut; in_sc_w, in_sc_b) = r2*coef[..0] + shortcut*coef[..1] + coef[..2]
 *      (biased variance, eps) -- feed (z = r2, z2 = shortcut, coef) to ood_field_step.  res / shortcut / r2 fp32 [B,3,R,R]. */
int64_t ood_alignnet_tail_workspace(int batch, int r);
int ood_alignnet_tail(const float *res, const float *shortcut, const float *prelu_slope, const float *conv_w,
                      const float *in_res_w, const float *in_res_b, const float *in_sc_w, const float *in_sc_b, float eps,
                      float *r2, float *workspace, float *coef, int batch, int r, void *stream);

/* ---- a13 (encoder FPN merge, e4e/encoders/helpers.py:504-521): out = bicubic_up(x, align_corners=True) + y on NHWC.
 *      x [B,h,w,C], y / out [B,H,W,C] (y may be NULL).  ATen's channels-last bicubic costs 40 ms per call at B=16. */
int ood_bicubic_up_add(const void *x, const void *y, void *out, int batch, int h, int w, int H, int W, int channels,
                       int dtype, void *stream);

/* ---- a14 ("grid_sample grads"; section 8b warp_alpha_bwd / mask_blend_bwd): backward of ood_warp_mix and ood_mask_blend, i.e. of
 *      F.grid_sample(bilinear, zeros, align_corners=False) + alpha mix (SAMM/helpers.py:168-177) and of the mask pyramid
 *      composition + clip + blend (OOD_faceGAN_e4e_arch.py:315-347), as torch.autograd computes them for the reference.
 *      First correct path (atomics).  ZERO-INITIALISE ggen / gfield / gfields before the call: they are accumulated into.
 *      ood_warp_mix_bwd: gen, gout NHWC [B,H,W,C] (dtype), field fp32 [B,3,H,W] -> ggen fp32 NHWC [B,H,W,C], gfield fp32 [B,3,H,W].
 *      ood_mask_blend_bwd: fields / gfields: n host arrays of device pointers to fp32 [B,3,r_i,r_i] (only channel 2 receives a
 *      gradient); x, gen, gout fp32 [B,3,S,S] -> gx, ggen fp32 [B,3,S,S] (written; either may be NULL). */
int ood_warp_mix_bwd(const void *gen, const float *field, const void *gout, float *ggen, float *gfield, int batch, int h, int w,
                     int channels, int dtype, void *stream);
/*      workspace (fp32, n_fields * B * S * S elements) selects the DETERMINISTIC two-stage form: per-pixel gradients of the up-sampled
 *      masks, then the adjoint of the bilinear up-sampling as a gather per mask cell (one warp per cell, fixed order); NULL = the
 *      atomic form (shared-memory windows per block + global atomics; gfields ZERO-INITIALISED by the caller). */
int ood_mask_blend_bwd(const float *const *fields_host, float *const *gfields_host, const int *field_sizes_host, int n_fields,
                       const float *x, const float *gen, const float *gout, float *gx, float *ggen, float *workspace, int batch, int size,
                       void *stream);

/*      ood_field_step_bwd (section 8b prm_bwd): backward of ood_field_step in its plain form (z only, no folded norms): heads
 *      (tanh * scale, tanh * scale, sigmoid) + 4x4 FIR + accumulate / clip / PRM + bicubic coarse PRM (SAMM/helpers.py:62-77,
 *      149-166).  z, gacc, gz, gf_workspace fp32 [B,3,R,R]; prev / gprev [B,3,R,R] or NULL; coarse / gcoarse [B,3,Rc,Rc] or NULL
 *      (gcoarse ZERO-INITIALISED by the caller: the bicubic gradient is accumulated into its alpha channel). */
int ood_field_step_bwd(const float *z, const float *prev, const float *coarse, const float *gacc, const float *taps_host, float scale,
                       int batch, int r, int rc, float *gf_workspace, float *gz, float *gprev, float *gcoarse, void *stream);

/* ---- section 8f rank 4 (host I/O either side of the path): the byte formats of the reference's inference script, on the
 *      device, so that a batch crosses PCIe as 3 bytes per pixel.  Bit-exact against the reference's arithmetic.
 *      ood_img2tensor_u8: replaces `cv2.imread(f) / 255.0 -> img2tensor(bgr2rgb) -> (t - 0.5) * 2`
 *          (run_ood_faceGAN_inversion.py:158-159, BasicSR/basicsr/utils/img_util.py:10-36):
 *          in uint8 [B][H][W][3] (interleaved, cv2 layout) -> out fp32 planes [B][3][H][W],
 *          out = (float32(double(v) / 255.0) - sub) * mul, channel order reversed if swap_rb.
 *      ood_tensor2img_u8: replaces `tensor2img(t, rgb2bgr, np.uint8, min_max)` (img_util.py:38-94,
 *          run_ood_faceGAN_inversion.py:64-72): in fp32 planes [B][3][H][W] -> out uint8 [B][H][W][3],
 *          out = uint8(rint(((clamp(x, lo, hi) - lo) / (hi - lo)) * 255.0f)) (round half to even), reversed if swap_rb. */
int ood_img2tensor_u8(const uint8_t *in, float *out, int batch, int h, int w, int swap_rb, float sub, float mul, void *stream);
int ood_tensor2img_u8(const float *in, uint8_t *out, int batch, int h, int w, int swap_rb, float lo, float hi, void *stream);

/* ---- a11 (AlignNet head, SAMM/helpers.py:85-109 via bottleneck_IR e4e/encoders/helpers.py:426-448): a 3x3 convolution
 *      2C -> 3 reads its 2C-channel input nine times for three outputs.  Instead ood_conv3x3(transposed = 4) projects every
 *      pixel once onto the 27 (tap, colour) weights, proj [B,H,W,Cp] fp32 with channel 3*t + k (Cp >= 27), and
 *          out[b,k,y,x] = sum_t proj[b, y + t/3 - 1, x + t%3 - 1, 3*t + k]        (zero outside the image)
 *      gathers the nine shifted partial sums.  out fp32 NCHW [B,3,H,W]. */
int ood_tap_sum(const float *proj, float *out, int batch, int h, int w, int cp, void *stream);
/*      ood_tap_sum_shortcut: the same, plus shortcut[b,k,y,x] = proj[b,y,x,27+k] (Cp >= 30): the bottleneck's 1x1 shortcut
 *      convolution 2C -> 3 (e4e/encoders/helpers.py:430-433) computed by the projection in three of its spare output channels,
 *      so its 2C-channel input is not read a second time. */
int ood_tap_sum_shortcut(const float *proj, float *out, float *shortcut, int batch, int h, int w, int cp, void *stream);

/* ---- a13 (encoder trunk, e4e/encoders/helpers.py:59-76 SEModule, :476-501 bottleneck_IR_SE) on NHWC activations.
 *      ood_se_gate:     stats [B,C,2] from ood_in_stats (channel means) -> gate[b,c] = sigmoid(w2 . relu(w1 . mean[b,:]));
 *                       w1 [C/r, C], w2 [C, C/r] fp32 (the two bias-free 1x1 convolutions).
 *      ood_se_residual: out = v * gate[b,c] + shortcut   (shortcut [B, h*s, w*s, C] read at stride s = 1 | 2 -- MaxPool2d(1, s) --
 *                       or NULL; gate NULL = 1), and, when t_next is given, t_next = out * bn_g[c] + bn_h[c]: the next
 *                       block's eval-mode BatchNorm.  Either output may be NULL.  With dtype OOD_BF16, shortcut_f32 / out_f32 = 1
 *                       read the shortcut / write `out` as fp32: the bf16 path keeps the residual stream itself in fp32. */
int ood_se_gate(const float *stats, const float *w1, const float *w2, float *gate, int batch, int channels, int reduced,
                void *stream);
int ood_se_residual(const void *v, const float *gate, const void *shortcut, int sc_stride, const float *bn_g,
                    const float *bn_h, void *out, void *t_next, void *out_lp, int batch, int h, int w, int channels, int dtype,
                    int shortcut_f32, int out_f32, void *stream);
/*      out_lp (or NULL): `out` once more in the storage type `dtype` -- the tapped feature maps (psp_encoders.py:185-192) when the
 *                       residual stream itself is kept in fp32.
 *      ood_latent_assemble: the W+ assembly of Encoder4Editing.forward (psp_encoders.py:199-214) and of the arch (e4e_arch.py:261):
 *                       heads [n_styles][B][D] fp32 = the style heads' outputs; out[b,i,:] = heads[0][b] + (1 <= i <= stage ? heads[i][b] : 0)
 *                       + avg[:] + delta[i,:]  (avg [D], delta [n_styles,D]; either may be NULL). */
/*      ood_se_tail:     ood_in_stats + ood_se_gate + ood_se_residual of one bottleneck in ONE launch (a cluster of 8 thread blocks per image
 *                       exchanges the channel sums through distributed shared memory): out (fp32) = v * sigmoid(w2 . relu(w1 . mean_hw(v)))
 *                       + shortcut, t_next / out_lp as in ood_se_residual.  dtype OOD_BF16 | OOD_F16. */
int ood_se_tail(const void *v, const float *w1, const float *w2, const void *shortcut, int sc_stride, const float *bn_g, const float *bn_h,
                float *out, void *t_next, void *out_lp, int batch, int h, int w, int channels, int reduced, int dtype, int shortcut_f32,
                void *stream);
/*      ood_se_apply:    ood_se_tail when the channel sums of v already exist: tile_sums = the stats_ws of the ood_conv3x3 call that wrote v
 *                       (stats_out = NULL: per-tile sums only), [B][tiles][C][2] with tiles = workspace bytes / (B*C*8).  One streaming pass on a
 *                       full grid (no pooling pass, no cluster); every block derives its image's gate from the tile sums in a fixed order.
 *                       channels in {64,128,256,512}. */
int ood_se_apply(const void *v, const float *tile_sums, int tiles, const float *w1, const float *w2, const void *shortcut, int sc_stride,
                 const float *bn_g, const float *bn_h, float *out, void *t_next, void *out_lp, int batch, int h, int w, int channels,
                 int reduced, int dtype, int shortcut_f32, void *stream);
int ood_latent_assemble(const float *heads, const float *avg, const float *delta, float *out, int batch, int n_styles, int dim,
                        int stage, void *stream);
/*      ood_alignnet_head_weights (a11): per-sample 1x1 projection weights of the AlignNet's 2C -> 3 head with the affine
 *                       InstanceNorm in front of it folded in (SAMM/helpers.py:85-109; e4e/encoders/helpers.py:436-441):
 *                       wps[b,r,c] = w27[r,c] * rstd[b,c]*in_w[c]  (rows 27..29 = w1[r-27,c] when w1 is given), bias[b,r] = sum_c
 *                       w27[r,c] * (in_b[c] - mean[b,c]*rstd[b,c]*in_w[c]);  stats [B,C,2] = {mean, rstd}; w27 [32,C], w1 [3,C] fp32;
 *                       wps [B,32,C] in `dtype` (OOD_BF16 | OOD_F32), bias [B,32] fp32. */
int ood_alignnet_head_weights(const float *stats, const float *in_w, const float *in_b, const float *w27, const float *w1,
                              void *wps, float *bias, int batch, int channels, int dtype, void *stream);

/* ---- a14. backward of the synthesis path for optimisation-based inversion (autograd through model.py:233-372; weights frozen).
 *      ood_act_bwd : gv = gy*sqrt2*(y>0 ? 1 : 0.2) (gate on the saved OUTPUT, fused_bias_act_kernel.cu:36-47);  g = gv*d[b,c];
 *                    gd[b,c] = sum_pix gv*acc with acc = (y/(sqrt2*gate) - noise_w*noise - bias)/d re-derived from y.
 *      ood_dot_reduce: out[b,c] = sum_pix a*b  (style gradient sum_pix gxs*x).
 *      ood_torgb_bwd : g_out = (g_in ? g_in : 0) + sum_k g_rgb[b,k,p]*wrgb[b,k,c]  (NHWC);  g_wrgb[b,c,k] = sum_pix g_rgb[b,k,p]*y[b,p,c].
 *      workspace: ood_bwd_workspace(batch, pixels, channels, k) bytes, k = 1 (act_bwd, dot_reduce) or 3 (torgb_bwd). */
int64_t ood_bwd_workspace(int batch, int64_t pixels, int channels, int k);
int ood_act_bwd(const void *gy, const void *y, const float *d, const float *bias, const float *noise, int64_t noise_bstride,
                const float *noise_w, void *g, float *workspace, float *gd, int batch, int64_t pixels, int channels, int dtype,
                void *stream);
/*      ood_act_bwd_fused: one pass per layer of the latent-gradient chain instead of torgb_bwd + act_bwd + dot_reduce:
 *          gy = g_in * g_scale[b,c] + sum_k g_rgb[b,k,p] * wrgb[b,k,c];   g = gy * lrelu'(y) * sqrt2 * d
 *          sums [B,C,K]: K = 2 {gd (as ood_act_bwd), sum_pix g_in*y} or, with g_rgb, K = 5 {.., sum_pix g_rgb_k * y, k = 0..2}
 *      g_in: the UNSCALED data gradient written by the convolution of the layer above (out_y without s_next; NULL for the top layer),
 *      g_scale [B,C]: that layer's input modulation (NULL = 1); g_rgb [B,3,P] fp32 / wrgb [B,3,C]: this level's ToRGB gradient (or NULL).
 *      workspace: ood_bwd_workspace(batch, pixels, channels, K) bytes.  Deterministic two-stage reductions. */
int ood_act_bwd_fused(const void *g_in, const float *g_scale, const float *g_rgb, const float *wrgb, const void *y, const float *d,
                      const float *bias, const float *noise, int64_t noise_bstride, const float *noise_w, void *g, float *workspace,
                      float *sums, int batch, int64_t pixels, int channels, int dtype, void *stream);
int ood_dot_reduce(const void *a, const void *b, float *workspace, float *out, int batch, int64_t pixels, int channels,
                   int dtype, void *stream);
int ood_torgb_bwd(const float *g_rgb, const float *wrgb, const void *y, const void *g_in, void *g_out, float *workspace,
                  float *g_wrgb, int batch, int64_t pixels, int channels, int dtype, void *stream);

/* ---- a11. AlignNet instance norms on NHWC (SAMM/helpers.py:96-101, e4e/encoders/helpers.py:93-99,426-448).
 *      ood_in_stats: per-(b,c) moments over the pixels.  y == NULL: stats[b][c] = {mean, rstd}.
 *      y != NULL (pair): stats[b][c] = {mean_x, rstd_x, mean_y, rstd_y, rstd_d, rstd_e2} where, with a = IN(x), e = IN(y),
 *      rstd_d = rsqrt(var(a-e)+eps) and rstd_e2 = rsqrt(var(e)+eps): the moments of z0 = cat[a-e, e] without a pass over z0.
 *      workspace: ood_in_stats_workspace() bytes (two-stage deterministic reduction, no float atomics).
 *      ood_alignnet_front: out[.,0:C] = (a-e)*rstd_d*w+bias, out[.,C:2C] = e*rstd_e2*w+bias   (= INaff(z0), block-0 conv input)
 *      ood_alignnet_res0 : out = INaff(t; st2, w, bias) + z0 (z0 re-derived from cur/enc/st6)   (block-0 output, 2C channels)
 *      ood_in_apply      : out = (x-mean)*rstd*w + bias. */
int64_t ood_in_stats_workspace(int batch, int64_t pixels, int channels, int pair);
int ood_in_stats(const void *x, const void *y, float *workspace, float *stats, int batch, int64_t pixels, int channels,
                 float eps, int dtype, void *stream);
int ood_alignnet_front(const void *cur, const void *enc, const float *st6, const float *w, const float *bias, void *out,
                       int batch, int64_t pixels, int channels, int dtype, void *stream);
int ood_alignnet_res0(const void *t, const float *st2, const float *w, const float *bias, const void *cur, const void *enc,
                      const float *st6, void *out, int batch, int64_t pixels, int channels, int dtype, void *stream);
int ood_in_apply(const void *x, const float *st2, const float *w, const float *bias, void *out, int batch, int64_t pixels,
                 int channels, int dtype, void *stream);
/*      ood_alignnet_front_split: the two halves of ood_alignnet_front as separate [B,P,C] tensors; out_hi == NULL skips the
 *      IN(enc) half, which does not depend on `cur` (the second alignment cycle, SAMM/helpers.py:154-166, reuses it -- and the
 *      part of the first convolution computed from it, see ood_conv3x3_args.acc_in).
 *      ood_alignnet_res0_stats: ood_alignnet_res0 that also returns stats_out [B,2C,2] = {mean, rstd} of its OUTPUT as stored
 *      (the statistics of the next InstanceNorm, e4e/encoders/helpers.py:426-448 res_layer[0] of the second block) without a
 *      pass over it; channels / (16 / sizeof(T)) must divide 256.  workspace: ood_alignnet_res0_workspace() bytes. */
int ood_alignnet_front_split(const void *cur, const void *enc, const float *st6, const float *w, const float *bias,
                             void *out_lo, void *out_hi, int batch, int64_t pixels, int channels, int dtype, void *stream);
int64_t ood_alignnet_res0_workspace(int batch, int64_t pixels, int channels, int dtype);
int ood_alignnet_res0_stats(const void *t, const float *st2, const float *w, const float *bias, const void *cur,
                            const void *enc, const float *st6, void *out, float *workspace, float *stats_out, float eps,
                            int batch, int64_t pixels, int channels, int dtype, void *stream);

/* ---- a11. warp + alpha mix (helpers.py:168-177) on NHWC features:
 *      out[b,y,x,:] = bilinear(gen[b], lin_x[x]+dx, lin_y[y]+dy)*alpha + gen[b,y,x,:]*(1-alpha)
 *      lin = linspace(-1,1,R); sampling with zeros padding, align_corners=False.  field fp32 [B,3,R,R]. */
int ood_warp_mix(const void *gen, const float *field, void *out, int batch, int h, int w, int channels,
                 int dtype, void *stream);

/* ---- a12. mask compose + blend (e4e_arch.py:315-347): A = up(a_1); A = up(a_k)*A + A*(1-A); clip;
 *      out = A*x + gen*(1-A).  fields[k]: fp32 [B,3,r_k,r_k] (alpha = channel 2), ascending size;
 *      bilinear align_corners=False to size x size.  x, gen, out: NCHW fp32 [B,3,size,size];
 *      alpha_out: [B,1,size,size] or NULL. */
int ood_mask_blend(const float *const *fields_host, const int *field_sizes_host, int n_fields, const float *x,
                   const float *gen, float *out, float *alpha_out, int batch, int size, void *stream);

/* ---- a14 / f3: elementwise building blocks of the differentiable alignment path (autograd through SAMM/helpers.py:85-109,149-179
 *      and e4e/encoders/helpers.py:426-448) on NHWC activations, dtype OOD_F32 | OOD_BF16.
 *      ood_nhwc_affine2: out[b,p,off_out+c] = a[b,c]*x1[b,p,off1+c] + b[b,c]*x2[b,p,off2+c] + c[b,c] for c < channels, every operand a
 *                        channel slice (offset) of a tensor with `pitch` channels per pixel; x2 / a / b / c may be NULL (no second
 *                        operand, 1, 1, 0).  The combine step of the InstanceNorm backward, residual sums / differences, channel
 *                        concatenation and its adjoint.
 *      ood_prelu:        g == NULL: out = x > 0 ? x : slope[c]*x;  else the backward out = g * (x > 0 ? 1 : slope[c]).
 *      ood_tap_gather:   adjoint of ood_tap_sum: out[b,y,x,3t+k] = g[b,k,y-dy_t,x-dx_t] (zero outside; channels 27..cp-1 zero). */
int ood_nhwc_affine2(const void *x1, int pitch1, int off1, const void *x2, int pitch2, int off2, const float *a, const float *b,
                     const float *c, void *out, int pitch_out, int off_out, int batch, int64_t pixels, int channels, int dtype,
                     void *stream);
int ood_prelu(const void *x, const void *g, const float *slope, void *out, int64_t pixels_total, int channels, int dtype, void *stream);
int ood_tap_gather(const float *g, void *out, int batch, int h, int w, int cp, int dtype, void *stream);

/* ---- f3: weight gradient of the shared-weight convolution (the AlignNet's plain Conv2d layers, SAMM/helpers.py:85-109, trained by
 *      src/models/OOD_faceGAN_model.py:663-789) on tcgen05: gw[o,i,ky,kx] = sum_{b,y,x} g[b,y,x,o] * x[b,y+ky-1,x+kx-1,i] (zero padding),
 *      g NHWC bf16 [B,H,W,cout], x NHWC bf16 [B,H,W,cin], gw fp32 [cout_real][cin_real][taps] (taps = 9: 3x3 pad 1; 1: 1x1), channel
 *      counts beyond *_real are padding.  cout % 128 == 0, cin % 64 == 0, w a divisor or a multiple of 64 (and h of 64 / w).  Deterministic
 *      (K slices summed in a fixed order).  workspace: ood_conv_wgrad_workspace() bytes (0 = outside the envelope). */
int64_t ood_conv_wgrad_workspace(int batch, int h, int w, int cin, int cout, int taps);
int ood_conv_wgrad(const void *g, const void *x, float *workspace, float *gw, int batch, int h, int w, int cin, int cout, int taps,
                   int cin_real, int cout_real, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OOD_B200_H_ */
