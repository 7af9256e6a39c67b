"""ctypes binding of libood_b200.so (the C ABI declared in include/ood_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libood_b200.so')

F32, BF16, F16 = 0, 1, 2

c_void_p, c_int, c_i64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float


class ConvArgs(C.Structure):
    _fields_ = [('inp', c_void_p), ('weight', c_void_p), ('out_y', c_void_p), ('out_ys', c_void_p), ('d', c_void_p),
                ('noise', c_void_p), ('noise_bstride', c_i64), ('noise_w', c_void_p), ('bias', c_void_p),
                ('s_next', c_void_p), ('batch', c_int), ('h', c_int), ('w', c_int), ('cin', c_int), ('cout', c_int),
                ('transposed', c_int), ('act', c_int), ('impl', c_int), ('dtype', c_int), ('out_f32', c_int), ('prelu_slope', c_void_p),
                ('rgb_w', c_void_p), ('rgb_bias', c_void_p), ('rgb_skip', c_void_p), ('rgb_out', c_void_p), ('rgb_taps', c_float * 4),
                ('groups', c_int), ('in_shared', c_int), ('acc_in', c_void_p), ('tiled', c_int), ('stats_out', c_void_p), ('stats_ws', c_void_p), ('stats_eps', c_float), ('out_dtype', c_int)]


class BlurActArgs(C.Structure):
    _fields_ = [('inp', c_void_p), ('in_f32', c_int), ('out_img', c_void_p), ('out_y', c_void_p), ('out_ys', c_void_p),
                ('d', c_void_p), ('noise', c_void_p), ('noise_w', c_void_p), ('bias', c_void_p), ('s_next', c_void_p),
                ('noise_bstride', c_i64), ('taps', c_float * 4), ('batch', c_int), ('ih', c_int), ('iw', c_int),
                ('channels', c_int), ('act', c_int), ('dtype', c_int), ('pad0', c_int), ('pad1', c_int)]


_SIGS = {
    'ood_version': ([], c_int),
    'ood_last_error': ([], C.c_char_p),
    'ood_device_is_sm100': ([], c_int),
    'ood_launch_count': ([], C.c_ulonglong),
    'ood_last_conv_route': ([], c_int),
    'ood_upfirdn2d': ([c_void_p, c_void_p, c_void_p, c_i64] + [c_int] * 12 + [c_int, c_void_p], c_int),
    'ood_fused_bias_act': ([c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64, c_int, c_int, c_float, c_float,
                            c_int, c_void_p], c_int),
    'ood_bias_grad': ([c_void_p, c_void_p, c_i64, c_int, c_i64, c_int, c_void_p], c_int),
    'ood_nchw_to_nhwc': ([c_void_p, c_i64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_thumbnail_nhwc': ([c_void_p, c_void_p] + [c_int] * 8 + [c_void_p], c_int),
    'ood_nhwc_to_nchw': ([c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_nhwc_scale': ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_i64, c_int, c_void_p], c_int),
    'ood_modulation': ([c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                        c_int, c_void_p], c_int),
    'ood_weight_sumsq': ([c_void_p, c_void_p, c_int, c_int, c_int, c_void_p], c_int),
    'ood_pack_conv_weight': ([c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_conv3x3': ([C.POINTER(ConvArgs), c_void_p], c_int),
    'ood_conv3x3_tiled_bytes': ([c_int] * 6, c_i64),
    'ood_conv3x3_stats_workspace': ([c_int] * 6, c_i64),
    'ood_blur_act': ([C.POINTER(BlurActArgs), c_void_p], c_int),
    'ood_noise_act': ([c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int,
                       c_int, c_void_p], c_int),
    'ood_torgb_weight': ([c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p], c_int),
    'ood_torgb': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(c_float), c_int, c_int, c_int, c_int, c_int,
                   c_void_p], c_int),
    'ood_field_step': ([c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(c_float), c_float, c_int, c_int, c_int, c_void_p, c_void_p,
                       c_void_p], c_int),
    'ood_alignnet_tail_workspace': ([c_int, c_int], c_i64),
    'ood_alignnet_tail': ([c_void_p] * 8 + [c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p], c_int),
    'ood_tap_sum': ([c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_tap_sum_shortcut': ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_se_gate': ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p], c_int),
    'ood_se_residual': ([c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                        c_int, c_int, c_int, c_void_p], c_int),
    'ood_se_tail': ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p], c_int),
    'ood_se_apply': ([c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p], c_int),
    'ood_act_bwd_fused': ([c_void_p] * 8 + [c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_void_p], c_int),
    'ood_latent_assemble': ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_alignnet_head_weights': ([c_void_p] * 7 + [c_int, c_int, c_int, c_void_p], c_int),
    'ood_bicubic_up_add': ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_warp_mix_bwd': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_mask_blend_bwd': ([c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p], c_int),
    'ood_field_step_bwd': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p], c_int),
    'ood_img2tensor_u8': ([c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p], c_int),
    'ood_tensor2img_u8': ([c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p], c_int),
    'ood_bwd_workspace': ([c_int, c_i64, c_int, c_int], c_i64),
    'ood_act_bwd': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64,
                     c_int, c_int, c_void_p], c_int),
    'ood_dot_reduce': ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_void_p], c_int),
    'ood_torgb_bwd': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_int,
                       c_void_p], c_int),
    'ood_in_stats_workspace': ([c_int, c_i64, c_int, c_int], c_i64),
    'ood_in_stats': ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_float, c_int, c_void_p], c_int),
    'ood_alignnet_front': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_void_p], c_int),
    'ood_alignnet_res0': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int,
                           c_int, c_void_p], c_int),
    'ood_alignnet_front_split': ([c_void_p] * 7 + [c_int, c_i64, c_int, c_int, c_void_p], c_int),
    'ood_alignnet_res0_workspace': ([c_int, c_i64, c_int, c_int], c_i64),
    'ood_alignnet_res0_stats': ([c_void_p] * 10 + [c_float, c_int, c_i64, c_int, c_int, c_void_p], c_int),
    'ood_in_apply': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_void_p], c_int),
    'ood_warp_mix': ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_nhwc_affine2': ([c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_i64, c_int, c_int,
                          c_void_p], c_int),
    'ood_prelu': ([c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p], c_int),
    'ood_tap_gather': ([c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    'ood_conv_wgrad_workspace': ([c_int] * 6, c_i64),
    'ood_conv_wgrad': ([c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_void_p], c_int),
    'ood_mask_blend': ([C.POINTER(c_void_p), C.POINTER(c_int), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                        c_void_p], c_int),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(ood_gan_inversion_b200 has no CPU or PyTorch fallback)')
        handle = C.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is not exported
            fn.argtypes = args
            fn.restype = res
        _lib = handle
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().ood_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'libood_b200 {what} failed (code {rc}): {msg}')
