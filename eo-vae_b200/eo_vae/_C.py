"""ctypes binding of libeovae_sm100.so (C ABI declared in include/eovae.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` from the repository root.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EOVAE_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libeovae_sm100.so"))

BF16, F16, F32 = 0, 1, 2
CONV_3X3, CONV_1X1, CONV_3X3_S2 = 0, 1, 2

_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/eovae.h declares (checked by tests/test_abi.py)
SIGNATURES = {
    "eovae_version": (_i, []),
    "eovae_last_error": (C.c_char_p, []),
    "eovae_num_sms": (_i, []),
    "eovae_launch_count": (C.c_ulonglong, []),
    "eovae_set_debug_mode": (None, [_i]),
    "eovae_set_tuning": (None, [_i, _i]),
    "eovae_conv_chunk_bytes": (_i, [_i]),
    "eovae_conv_k_per_tap": (_i, [_i]),
    "eovae_pack_conv_weight": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "eovae_conv2d_gn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "eovae_conv2d_gn_prologue_ok": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "eovae_conv2d": (_i, [_vp, _i, _i, _i, _i, _ll, _i, _vp, _i, _vp, _vp, _i, _ll, _vp, _i, _ll, _i, _f, _vp, _i, _f, _vp,
                          _sz, _vp, _i, _ll, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "eovae_gemm_tn_batched": (_i, [_vp, _ll, _ll, _vp, _ll, _ll, _vp, _i, _ll, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "eovae_conv2d_up2x_ok": (_i, [_i, _i, _i, _i, _i]),
    "eovae_conv2d_up2x_gn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "eovae_pack_conv_weight_up2x": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "eovae_conv2d_up2x": (_i, [_vp, _i, _i, _i, _i, _ll, _vp, _i, _vp, _vp, _i, _ll, _i, _vp, _i, _f, _vp, _sz, _vp]),
    "eovae_conv2d_up2x_dgrad": (_i, [_vp, _i, _i, _i, _i, _ll, _vp, _i, _vp, _i, _ll, _i, _vp]),
    "eovae_conv2d_s2_dgrad": (_i, [_vp, _i, _i, _i, _i, _ll, _vp, _i, _vp, _i, _ll, _i, _vp]),
    "eovae_conv2d_s2_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "eovae_conv2d_s2_wgrad": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "eovae_conv2d_up2x_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "eovae_conv2d_up2x_wgrad": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "eovae_gemm_strided_f32": (_i, [_vp, _ll, _ll, _vp, _ll, _ll, _ll, _vp, _ll, _i, _i, _i, _i, _f, _vp]),
    "eovae_gn_stats_workspace_bytes": (_sz, [_i, _ll, _i, _i]),
    "eovae_gn_stats": (_i, [_vp, _i, _i, _ll, _i, _ll, _i, _f, _vp, _vp, _sz, _vp]),
    "eovae_gn_apply": (_i, [_vp, _i, _ll, _vp, _vp, _vp, _vp, _i, _ll, _i, _ll, _i, _i, _i, _vp]),
    "eovae_nchw_to_nhwc16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "eovae_nhwc_to_nchw_f32": (_i, [_vp, _i, _ll, _vp, _i, _i, _i, _i, _vp]),
    "eovae_upsample2x": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "eovae_softmax_rows": (_i, [_vp, _i, _ll, _vp, _i, _ll, _ll, _i, _vp]),
    "eovae_transpose16": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "eovae_hypernet_backward_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "eovae_hypernet_backward": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _f, _vp, _f, _vp, _i, _vp, _sz, _vp]),
    "eovae_hypernet_forward_taped": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "eovae_hypernet_factorized_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "eovae_hypernet_factorized_forward": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "eovae_hypernet_factorized_backward": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _f, _vp, _f, _vp, _vp,
                                                _sz, _vp]),
    "eovae_wavelength_style_workspace_bytes": (_sz, [_i]),
    "eovae_wavelength_style_forward": (_i, [_vp, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "eovae_wavelength_style_backward": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "eovae_adain_affine_forward": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "eovae_adain_affine_backward": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eovae_sam_loss": (_i, [_vp, _vp, _i, _i, _ll, _f, _vp, _vp, _sz, _vp]),
    "eovae_sam_loss_backward": (_i, [_vp, _vp, _i, _i, _ll, _f, _vp, _vp, _vp]),
    "eovae_grad_diff_loss": (_i, [_vp, _vp, _ll, _i, _i, _vp, _vp, _sz, _vp]),
    "eovae_grad_diff_loss_backward": (_i, [_vp, _vp, _ll, _i, _i, _vp, _vp, _vp]),
    "eovae_focal_freq_loss_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "eovae_focal_freq_loss": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _sz, _vp]),
    "eovae_focal_freq_loss_backward": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "eovae_latent_resize_rot": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp]),
    "eovae_latent_resize_rot_backward": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp]),
    "eovae_area_resize_rot": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp]),
    "eovae_grad_norm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "eovae_adam_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f, _f, _i, _vp, _f, _vp]),
    "eovae_msssim_backward_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "eovae_msssim_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "eovae_conv2d_wgrad_nhwc_ok": (_i, [_i, _i]),
    "eovae_conv2d_wgrad_nhwc_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "eovae_conv2d_wgrad_nhwc": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "eovae_running_stats_update": (_i, [_vp, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eovae_latent_bn_train_forward": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _f, _f, _f, _vp, _i, _ll, _vp, _vp]),
    "eovae_latent_bn_train_backward": (_i, [_vp, _i, _ll, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "eovae_preprocess": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _f, _i, _f, _f, _i, _i, _i, _vp, _vp]),
    "eovae_attention_fused_ok": (_i, [_i, _i]),
    "eovae_attention_fused": (_i, [_vp, _ll, _i, _i, _i, _vp, _ll, _i, _vp]),
    "eovae_softmax_backward": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _ll, _i, _i, _f, _vp]),
    "eovae_reparam_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "eovae_pixel_loss_backward": (_i, [_vp, _vp, _ll, _f, _i, _vp, _vp, _vp]),
    "eovae_transpose16_xshift": (_i, [_vp, _ll, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "eovae_latent_norm": (_i, [_vp, C.POINTER(_ll), _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp]),
    "eovae_latent_denorm": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _vp]),
    "eovae_kl_reparam": (_i, [_vp, C.POINTER(_ll), _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "eovae_hypernet_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "eovae_hypernet_forward": (_i, [_vp, _i, C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "eovae_pack_dyn_weight": (_i, [_vp, _i, _i, _i, _f, _vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _vp]),
    "eovae_l1_charbonnier": (_i, [_vp, _vp, _ll, _f, _vp, _vp, _sz, _vp]),
    "eovae_pack_conv_weight_dgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "eovae_gn_backward_workspace_bytes": (_sz, [_i, _ll, _i, _i]),
    "eovae_gn_backward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "eovae_scatter_stride2": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "eovae_pool2x2_sum": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "eovae_conv2d_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "eovae_conv2d_wgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "eovae_bias_grad_workspace_bytes": (_sz, [_ll, _i]),
    "eovae_bias_grad": (_i, [_vp, _i, _ll, _i, _vp, _i, _vp, _sz, _vp]),
    "eovae_msssim_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "eovae_msssim": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the CDLL; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"eo_vae: CUDA extension {LIB_PATH} not found - run __graft_entry__.build(); "
                "there is no CPU / PyTorch fallback for this path")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError => ABI mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
        if handle.eovae_version() != 2:
            raise RuntimeError("eo_vae: libeovae_sm100.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().eovae_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
