"""ctypes binding of libb200robust.so (the C-ABI declared in include/b200r.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200R_LIB") or os.path.join(_HERE, "lib", "libb200robust.so")

_lib = None

c_u8p = C.c_void_p
c_f32p = C.c_void_p
c_i64p = C.c_void_p
c_stream = C.c_void_p
c_host_f3 = C.POINTER(C.c_float)

# name -> (restype, argtypes); mirrors include/b200r.h one to one
SIGNATURES = {
    "b200r_last_error": (C.c_char_p, []),
    "b200r_version": (C.c_int, []),
    "b200r_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b200r_corrupt_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.POINTER(C.c_size_t)]),
    "b200r_corrupt_ext_noise_count": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.POINTER(C.c_size_t)]),
    "b200r_corrupt_u8": (C.c_int, [C.c_int, C.c_int, c_u8p, c_u8p, C.c_int, C.c_int, C.c_int,
                                   C.c_uint64, C.c_uint64, c_f32p, C.c_void_p, C.c_size_t, c_stream]),
    "b200r_set_frost_texture": (C.c_int, [C.c_int, c_u8p, C.c_int, C.c_int]),
    "b200r_normal_strata_table": (C.c_int, [C.c_void_p]),
    "b200r_split_f32_scaled": (C.c_int, [c_f32p, C.c_void_p, C.c_size_t, C.c_float, c_stream]),
    "b200r_model_create": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200r_model_destroy": (C.c_int, [C.c_void_p]),
    "b200r_model_num_classes": (C.c_int, [C.c_void_p]),
    "b200r_model_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200r_model_forward_u8": (C.c_int, [C.c_void_p, c_u8p, c_f32p, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_model_forward_f32": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_model_input_grad": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_stream]),
    "b200r_allreduce_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_stream]),
    "b200r_fab_projection_linf": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_void_p, c_stream]),
    "b200r_fab_combine_linf": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_float, C.c_float, c_stream]),
    "b200r_l1_projection": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_float, c_stream]),
    "b200r_pgd_step_l1": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_float, C.c_float, c_stream]),
    "b200r_u8nhwc_to_f32nchw": (C.c_int, [c_u8p, c_f32p, C.c_int, C.c_int, C.c_int, c_host_f3,
                                          c_host_f3, c_stream]),
    "b200r_normalize_f32nchw": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, c_host_f3,
                                          c_host_f3, C.c_int, c_stream]),
    "b200r_random_start_linf": (C.c_int, [c_f32p, c_f32p, C.c_size_t, C.c_size_t, C.c_float,
                                          C.c_uint64, C.c_uint64, c_f32p, C.c_int, c_stream]),
    "b200r_pgd_step_linf": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_size_t, C.c_size_t, C.c_float,
                                      C.c_float, c_stream]),
    "b200r_random_start_l2": (C.c_int, [c_f32p, c_f32p, C.c_size_t, C.c_size_t, C.c_float, C.c_uint64, C.c_uint64, c_stream]),
    "b200r_pgd_step_l2": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_size_t, C.c_size_t, C.c_float,
                                    C.c_float, c_f32p, c_stream]),
    "b200r_mim_step_linf": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_size_t, C.c_size_t,
                                      C.c_float, C.c_float, C.c_float, c_f32p, c_stream]),
    "b200r_ce_loss_grad": (C.c_int, [c_f32p, c_i64p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_float,
                                     c_stream]),
    "b200r_softmax": (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, c_stream]),
    "b200r_topk_count": (C.c_int, [c_f32p, c_i64p, C.c_int, C.c_int, c_i64p, c_i64p, c_stream]),
    "b200r_split_f32": (C.c_int, [c_f32p, C.c_void_p, C.c_size_t, c_stream]),
    "b200r_merge_f32": (C.c_int, [C.c_void_p, c_f32p, C.c_size_t, c_stream]),
    "b200r_conv2d_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_void_p,
                                    c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_linear": (C.c_int, [C.c_void_p, C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_void_p, c_f32p,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_linear_keep_pre": (C.c_int, [C.c_void_p, C.c_void_p, c_f32p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_stem_im2col_u8": (C.c_int, [c_u8p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_host_f3,
                                       c_host_f3, c_stream]),
    "b200r_stem_im2col_f32": (C.c_int, [c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_host_f3,
                                        c_host_f3, c_stream]),
    "b200r_stem_conv7x7_u8": (C.c_int, [c_u8p, C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        c_host_f3, c_host_f3, C.c_int, C.c_int, c_stream]),
    "b200r_stem_pool_u8_f16": (C.c_int, [c_u8p, C.c_void_p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         c_host_f3, c_host_f3, c_stream]),
    "b200r_stem_pool_split_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_host_f3, c_host_f3, C.c_int, C.c_void_p, C.POINTER(C.c_float)]),
    "b200r_stem_pool_f32_split": (C.c_int, [c_f32p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_stem_pool_u8_split": (C.c_int, [c_u8p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_stem_conv7x7_f32": (C.c_int, [c_f32p, C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         c_host_f3, c_host_f3, C.c_int, C.c_int, c_stream]),
    "b200r_resize_workspace_bytes": (C.c_int, [C.c_int] * 10 + [C.POINTER(C.c_size_t)]),
    "b200r_resize_u8": (C.c_int, [c_u8p, c_u8p] + [C.c_int] * 10 + [C.c_void_p, C.c_size_t, c_stream]),
    "b200r_image_stem3x3s2_u8": (C.c_int, [c_u8p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           c_host_f3, c_host_f3, c_stream]),
    "b200r_image_stem3x3s2_f32": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            c_host_f3, c_host_f3, c_stream]),
    "b200r_maxpool3x3s2_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          c_stream]),
    "b200r_global_avgpool_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            c_stream]),
    "b200r_f32_to_f16": (C.c_int, [c_f32p, C.c_void_p, C.c_size_t, C.c_float, c_stream]),
    "b200r_f16_to_f32": (C.c_int, [C.c_void_p, c_f32p, C.c_size_t, C.c_float, c_stream]),
    "b200r_maxpool3x3s2_nhwc_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_global_avgpool_nhwc_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_conv2d_dgrad3x3s2_nhwc": (C.c_int, [C.c_void_p] * 8 + [C.c_int] * 6 + [c_stream]),
    "b200r_conv2d_dgrad1x1s2_acc_nhwc": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 6 + [c_stream]),
    "b200r_conv2d_dgrad_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_relu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    "b200r_dilate2_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_maxpool3x3s2_nhwc_codes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_maxpool3x3s2_bwd_codes_hi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_maxpool3x3s2_relu_bwd_hi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_maxpool3x3s2_bwd_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                              C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_global_avgpool_bwd_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_stem_col2im_f32": (C.c_int, [C.c_void_p, c_f32p, C.c_int, C.c_int, C.c_int, c_host_f3, c_stream]),
    "b200r_relu_bwd_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    "b200r_dilate2_nhwc_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_maxpool3x3s2_bwd_nhwc_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                                  C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_global_avgpool_bwd_nhwc_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_stem_col2im_f32_f16": (C.c_int, [C.c_void_p, c_f32p, C.c_int, C.c_int, C.c_int, c_host_f3, C.c_float, c_stream]),
}


class B200RError(RuntimeError):
    pass


def load():
    """dlopen the library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200RError(
            "libb200robust.so is not built (%s). Run `python -m robustart_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().b200r_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("b200r: " + msg)
        if rc == -4:
            raise NotImplementedError("b200r: " + msg)
        raise B200RError("b200r error %d: %s" % (rc, msg))


def f3(vals):
    return (C.c_float * 3)(*[float(v) for v in vals])

SIGNATURES.update({
    "b200r_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_float, c_stream]),
    "b200r_patch_gather_u8": (C.c_int, [c_u8p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_host_f3, c_host_f3, c_stream]),
    "b200r_patch_gather_f32": (C.c_int, [c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_host_f3, c_host_f3, c_stream]),
    "b200r_assemble_tokens": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_stream]),
    "b200r_tokens_to_channels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_channels_to_tokens_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_resize_cv_u8": (C.c_int, [c_u8p, c_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, c_f32p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, c_stream]),
    "b200r_act_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, c_stream]),
    "b200r_act_bwd_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, c_stream]),
    "b200r_patch_scatter_f32": (C.c_int, [C.c_void_p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, c_host_f3, c_stream]),
    "b200r_attention_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_stream]),
    "b200r_channel_dot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_planes_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    "b200r_image_stem3x3s2_bwd": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, c_host_f3, C.c_float, c_stream]),
    "b200r_attention_bwd_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "b200r_attention_bwd_ws": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                         c_stream]),
})

SIGNATURES.update({
    "b200r_dwconv_nhwc": (C.c_int, [C.c_void_p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_pointwise_smallk_nhwc": (C.c_int, [C.c_void_p, c_f32p, c_f32p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_channel_scale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_stream]),
    "b200r_image_im2col_u8": (C.c_int, [c_u8p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_host_f3,
                                        c_host_f3, c_stream]),
    "b200r_image_im2col_f32": (C.c_int, [c_f32p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_host_f3,
                                         c_host_f3, c_stream]),
})

SIGNATURES.update({
    "b200r_apgd_step_linf": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_size_t, C.c_size_t, C.c_float, C.c_float, c_stream]),
    "b200r_dlr_loss_grad": (C.c_int, [c_f32p, c_i64p, c_i64p, c_f32p, c_f32p, C.c_int, C.c_int, c_stream]),
    "b200r_square_propose_linf": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            c_host_f3, C.c_float, c_stream]),
    "b200r_masked_rows_copy": (C.c_int, [c_f32p, c_f32p, c_u8p, C.c_size_t, C.c_size_t, c_stream]),
})
