"""ctypes binding of libammc_b200.so (C ABI declared in include/ammc_b200.h).

The library is the only compute backend: if it is missing, cannot be loaded, or the device is not sm_100,
every entry point raises -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMMC_B200_LIB") or os.path.join(_PKG, "libammc_b200.so")     # override: A/B builds only

P, I, L, F, Z = c_void_p, c_int, c_int64, c_float, c_size_t

class ConvLayer(ctypes.Structure):
    """`ammc_conv_layer` of include/ammc_b200.h (field order and types must match)."""
    _fields_ = [("in_planes", c_void_p), ("in_cs", c_int), ("in_c_off", c_int),
                ("wp", c_void_p), ("taps", c_int),
                ("scale", c_void_p), ("shift", c_void_p), ("act", c_int),
                ("out_planes", c_void_p), ("out_cs", c_int), ("out_c_off", c_int),
                ("out_nchw", c_void_p), ("res_nchw", c_void_p), ("cout_valid", c_int),
                ("b", c_int), ("h", c_int), ("w", c_int), ("Cin", c_int), ("Cout", c_int),
                ("up2x", c_int), ("precision", c_int), ("in_fmt", c_int), ("out_fmt", c_int), ("io_bf16", c_int)]


# name -> (restype, argtypes); must list every symbol of include/ammc_b200.h (tests/test_capi.py checks it)
SIGNATURES = {
    "ammc_version": (I, []),
    "ammc_last_error": (c_char_p, []),
    "ammc_device_supported": (I, []),
    "ammc_pipeline_check": (I, [P]),
    "ammc_mem_workspace_bytes": (Z, [I] * 7),
    "ammc_set_addressing_mode": (I, [I]),
    "ammc_mem_fwd": (I, [P] * 6 + [P] * 6 + [P, P, P, I, P] + [P, Z] + [I] * 8 + [P]),
    "ammc_mem_io16_supported": (I, [I] * 7),
    "ammc_mem_fwd_io16": (I, [P] * 6 + [P] * 6 + [P, I, P] + [P, Z] + [I] * 8 + [P]),
    "ammc_cast_f32_bf16": (I, [P, P, L, P]),
    "ammc_mem_prep_bytes": (Z, [I] * 4),
    "ammc_mem_prepare": (I, [P] * 5 + [Z] + [I] * 7 + [P]),
    "ammc_set_front_mode": (I, [I]),
    "ammc_mem_dec_uses_tensor": (I, [I] * 7),
    "ammc_set_dec_mode": (I, [I]),
    "ammc_set_enc_mode": (I, [I]),
    "ammc_addr_padded_items": (I, [I]),
    "ammc_addr_pack_queries": (I, [P, P, P, L, I, P]),
    "ammc_addr_pack_bank": (I, [P] * 6 + [I, I, P]),
    "ammc_addr_filter": (I, [P] * 7 + [L, I, I, I, P]),
    "ammc_quantize_workspace_bytes": (Z, [L, I, I, I]),
    "ammc_quantize_fwd": (I, [P, P] + [P] * 5 + [P, P] + [P, Z] + [L, L, I, I, I] + [P]),
    "ammc_quantize_bwd_workspace_bytes": (Z, [L, I, I]),
    "ammc_quantize_bwd": (I, [P] * 6 + [P, Z] + [L, I, I, I] + [P]),
    "ammc_embed_code": (I, [P, P, P, L, I, I, P]),
    "ammc_ema_update": (I, [P] * 5 + [I, I, F, F, P]),
    "ammc_mem_bwd_workspace_bytes": (Z, [I] * 7),
    "ammc_mem_bwd": (I, [P] * 8 + [P] * 5 + [P, Z] + [I] * 8 + [P]),
    "ammc_pack_conv_weights": (I, [P, P, I, I, P]),
    "ammc_pack_nhwc": (I, [P, P, I, I, I, I, P]),
    "ammc_q_act_bytes": (Z, [L]),
    "ammc_q_weight_bytes": (Z, [I, I]),
    "ammc_pack_nhwc_q": (I, [P, P, I, I, I, I, P]),
    "ammc_pack_nhwc_q_planes": (I, [P, P, P, I, I, I, I, P]),
    "ammc_pack_conv_weights_q": (I, [P, P, I, I, I, P]),
    "ammc_pack_conv_weights_q_pair": (I, [P, P, P, I, I, I, P]),
    "ammc_conv3x3_bn_relu": (I, [P] * 7 + [I] * 7 + [P]),
    "ammc_conv_layer_run": (I, [ctypes.POINTER(ConvLayer), P]),
    "ammc_pack_conv_weights_padded": (I, [P, P, I, I, I, I, I, P]),
    "ammc_pack_convt_weights": (I, [P, P, I, I, P]),
    "ammc_pack_nhwc_padded": (I, [P, P, I, I, I, I, I, P]),
    "ammc_maxpool2_planes": (I, [P, I, I, P, I, I, I, I, P]),
    "ammc_unpack_nhwc": (I, [P, I, I, P, I, I, I, I, P]),
    "ammc_set_conv_pair_mode": (I, [I]),
    "ammc_set_conv_halo_mode": (I, [I]),
    "ammc_pack_conv_weights_1x1": (I, [P, P, I, I, P]),
    "ammc_conv1x1_bn_relu": (I, [P] * 7 + [I] * 7 + [P]),
    "ammc_bn_batch_stats": (I, [P] * 9 + [P, Z] + [I, I, I, I, F, F, I, P]),
    "ammc_bn_apply": (I, [P, P, P, I, P, P, P, P, I, I, I, I, P]),
    "ammc_bn_q_workspace_bytes": (Z, [I]),
    "ammc_bn_batch_stats_q": (I, [P] * 9 + [P, Z] + [I] * 4 + [F, F, P]),
    "ammc_bn_apply_q": (I, [P, P, P, I, P, P, P] + [I] * 4 + [P]),
    "ammc_bn_backward_q": (I, [P] * 6 + [I, I, P, P, P, P, P, Z] + [I] * 4 + [P]),
    "ammc_bn_backward": (I, [P] * 6 + [I, I] + [P] * 4 + [P, Z] + [I, I, I, I, P]),
    "ammc_bn_batch_stats_staged": (I, [P] * 9 + [P, Z] + [I, I, I, I, F, F, I, I, c_double, P]),
    "ammc_bn_backward_staged": (I, [P] * 6 + [I, I] + [P] * 4 + [P, Z] + [I, I, I, I, I, c_double, P]),
    "ammc_pack_planes": (I, [P, P, L, P]),
    "ammc_pack_conv_weights_dgrad": (I, [P, P, I, I, P]),
    "ammc_conv3x3_wgrad": (I, [P, P, P, I, I, I, I, I, I, P]),
    "ammc_conv1x1_wgrad": (I, [P, P, P, I, I, I, I, I, I, P]),
    "ammc_bn_fold": (I, [P] * 4 + [F] + [P, P, I, P]),
    "ammc_preprocess_frames_u8": (I, [P, P, I, I, I, I, I, P]),
    "ammc_preprocess_flow": (I, [P, P, I, I, I, I, I, P]),
    "ammc_cast_bf16_f32": (I, [P, P, L, P]),
    "ammc_frame_losses_workspace_bytes": (Z, [I, I, I]),
    "ammc_frame_losses_fwd": (I, [P, P, P, P, Z, I, I, I, I, P]),
    "ammc_frame_losses_bwd": (I, [P, P, P, P, P, I, I, I, I, P]),
    "ammc_elem_loss_workspace_bytes": (Z, [L]),
    "ammc_elem_loss_fwd": (I, [P, P, P, I, L, P, Z, P]),
    "ammc_elem_loss_bwd": (I, [P, P, P, P, P, I, L, P]),
    "ammc_gen_objective_workspace_bytes": (Z, [I] * 6 + [L, L]),
    "ammc_gen_objective_fwd": (I, [P] * 8 + [I] * 8 + [L, L, I] + [F] * 6 + [P, P, Z, P]),
    "ammc_gen_objective_bwd": (I, [P] * 8 + [I] * 8 + [L, L] + [F] * 6 + [P] * 5 + [P]),
    "ammc_psnr_workspace_bytes": (Z, [I, L]),
    "ammc_psnr_batch": (I, [P, P, P, P, Z, I, L, P]),
    "ammc_score_workspace_bytes": (Z, [L, I]),
    "ammc_auc_workspace_bytes": (Z, [L]),
    "ammc_roc_auc": (I, [P, P, I, P, P, Z, L, P]),
    "ammc_score_reduce": (I, [P, P, P, I, F, F, F, F, P, P, Z, L, P]),
}

_lib = None


def load():
    """Load (once) and return the ctypes handle; raises RuntimeError when the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "ammcnet_aaai2021_b200: %s is missing. Build it with `python -m ammcnet_aaai2021_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise RuntimeError("ammcnet_aaai2021_b200: cannot load %s: %s" % (LIB_PATH, e))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ammc_last_error()
        raise RuntimeError("ammc_b200 %s failed (code %d): %s" % (what, rc, (msg or b"").decode(errors="replace")))


def call(name: str, *args):
    """Call an int-returning entry point and raise RuntimeError on a non-zero code."""
    check(getattr(load(), name)(*args), name)
