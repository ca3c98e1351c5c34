"""ctypes loader for libmonodetr_b200.so (the C-ABI product library).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmonodetr_b200.so")
_lib = None

c_int, c_void_p, c_float = ctypes.c_int, ctypes.c_void_p, ctypes.c_float

# name -> argtypes (restype is always int unless listed in _RESTYPES)
_PTR = c_void_p
SIGNATURES = {
    "mdb_abi_version": [],
    "mdb_error_string": [c_int],
    "mdb_msda_forward_f32": [_PTR] * 5 + [c_int] * 7 + [_PTR, _PTR],
    "mdb_msda_forward_f64": [_PTR] * 5 + [c_int] * 7 + [_PTR, _PTR],
    "mdb_msda_backward_f32": [_PTR] * 6 + [c_int] * 7 + [_PTR] * 4,
    "mdb_msda_backward_f64": [_PTR] * 6 + [c_int] * 7 + [_PTR] * 4,
    "mdb_msda_fused_forward_f32": [_PTR] * 6 + [c_int] * 8 + [_PTR, _PTR],
    "mdb_msda_fused_backward_f32": [_PTR] * 7 + [c_int] * 8 + [_PTR] * 4,
    "mdb_msda_prep_forward_f32": [_PTR] * 4 + [c_int] * 6 + [_PTR] * 3,
    "mdb_msda_prep_backward_f32": [_PTR] * 5 + [c_int] * 6 + [_PTR] * 3,
    "mdb_set_deterministic": [c_int],
    "mdb_get_deterministic": [],
    "mdb_set_precision": [c_int],
    "mdb_get_precision": [],
    "mdb_conv2d_forward_f32": [_PTR] * 5 + [c_int] * 10 + [_PTR],
    "mdb_conv2d_forward_bf16x3": [_PTR] * 5 + [c_int] * 10 + [_PTR],
    "mdb_conv2d_dgrad_bf16x3": [_PTR] * 5 + [c_int] * 10 + [_PTR],
    "mdb_conv2d_forward_workspace_bytes": [c_int] * 12,
    "mdb_set_workspace": [_PTR, ctypes.c_ulonglong],
    "mdb_pack_gemm_weights_bf16x3": [c_int] + [_PTR] * 7 + [c_int, _PTR],
    "mdb_conv2d_dgrad_f32": [_PTR] * 5 + [c_int] * 10 + [_PTR],
    "mdb_conv2d_wgrad_f32": [_PTR] * 4 + [c_int] * 10 + [_PTR],
    "mdb_conv2d_wgrad_bias_f32": [_PTR] * 5 + [c_int] * 10 + [_PTR],
    "mdb_pack_conv_weight_f32": [_PTR] * 3 + [c_int] * 3 + [_PTR],
    "mdb_unpack_conv_wgrad_f32": [_PTR] * 2 + [c_int] * 4 + [_PTR],
    "mdb_pack_conv_weights_multi_f32": [c_int] + [_PTR] * 6 + [_PTR],
    "mdb_unpack_conv_wgrads_multi_f32": [c_int] + [_PTR] * 5 + [_PTR],
    "mdb_colsum_f32": [_PTR] * 2 + [ctypes.c_longlong, c_int, c_int, _PTR],
    "mdb_attention_forward_f32": [_PTR] * 6 + [c_int] * 9 + [c_float, _PTR, ctypes.c_ulonglong, _PTR],
    "mdb_attention_backward_f32": [_PTR] * 11 + [c_int] * 12 + [c_float, _PTR, ctypes.c_ulonglong, _PTR],
    "mdb_add_layernorm_forward_f32": [_PTR] * 7 + [ctypes.c_longlong, c_int, c_float, c_float, _PTR, ctypes.c_ulonglong, _PTR],
    "mdb_add_layernorm_backward_f32": [_PTR] * 10 + [ctypes.c_longlong, c_int, c_float, _PTR, ctypes.c_ulonglong, c_int, _PTR],
    "mdb_groupnorm_forward_f32": [_PTR] * 7 + [c_int] * 4 + [c_float, c_int, _PTR],
    "mdb_groupnorm_backward_f32": [_PTR] * 10 + [c_int] * 5 + [_PTR],
    "mdb_relu_backward_f32": [_PTR] * 3 + [ctypes.c_longlong, c_float, _PTR],
    "mdb_dropout_f32": [_PTR] * 2 + [ctypes.c_longlong, c_float, _PTR, ctypes.c_ulonglong, _PTR],
    "mdb_round_tf32_f32": [_PTR] * 2 + [ctypes.c_longlong, _PTR],
    "mdb_stem_conv7x7_bn_relu_f32": [_PTR] * 5 + [c_int] * 3 + [_PTR],
    "mdb_maxpool3x3s2_nhwc_f32": [_PTR] * 2 + [c_int] * 4 + [_PTR],
    "mdb_depth_sample_forward_f32": [_PTR] * 3 + [c_int] * 4 + [_PTR],
    "mdb_depth_sample_backward_f32": [_PTR] * 3 + [c_int] * 4 + [_PTR],
    "mdb_box_refine_forward_f32": [_PTR] * 3 + [ctypes.c_longlong, c_int, _PTR],
    "mdb_box_refine_backward_f32": [_PTR] * 5 + [ctypes.c_longlong, c_int, _PTR],
    "mdb_head_depth_forward_f32": [_PTR] * 7 + [c_int] * 4 + [_PTR],
    "mdb_head_depth_backward_f32": [_PTR] * 10 + [c_int] * 4 + [_PTR],
    "mdb_depth_tail_forward_f32": [_PTR] * 5 + [ctypes.c_longlong, c_int, c_int, c_int, c_float, _PTR],
    "mdb_depth_tail_backward_f32": [_PTR] * 7 + [ctypes.c_longlong, c_int, c_int, c_int, c_float, _PTR],
    "mdb_mean3_f32": [_PTR] * 4 + [ctypes.c_longlong, _PTR],
    "mdb_scale_f32": [_PTR] * 2 + [ctypes.c_longlong, c_float, _PTR],
    "mdb_sum_mean_squares_forward_f32": [c_int, _PTR, _PTR, _PTR, _PTR],
    "mdb_sum_mean_squares_backward_f32": [c_int, _PTR, _PTR, _PTR, _PTR, _PTR],
    "mdb_adamw_step_f32": [_PTR] * 4 + [ctypes.c_longlong] * 2 + [c_float] * 7 + [_PTR, _PTR],
    "mdb_criterion_prepare": [_PTR, c_int, c_int, _PTR, _PTR, _PTR, _PTR],
    "mdb_criterion_match_f32": [c_int] + [_PTR] * 6 + [c_int] * 5 + [c_float] * 4 + [_PTR] * 3,
    "mdb_criterion_depth_map_f32": [_PTR] + [ctypes.c_longlong] * 3 + [_PTR] * 4 + [c_int] * 5 + [c_float] * 7 + [_PTR] * 4,
    "mdb_criterion_losses_f32": [c_int] + [_PTR] * 17 + [c_int] * 6 + [c_float] * 2 + [_PTR, _PTR, _PTR],
    "mdb_criterion_losses_backward_f32": [c_int] + [_PTR] * 16 + [c_int] * 5 + [c_float] * 2 + [_PTR] * 8,
    "mdb_warp_affine_normalize_u8": [_PTR] * 5 + [c_int] * 3 + [_PTR] * 4,
    "mdb_extract_dets_f32": [_PTR] * 5 + [c_int] * 4 + [_PTR, _PTR],
    "mdb_decode_dets_f32": [_PTR] * 4 + [c_int] * 3 + [c_float, _PTR, _PTR, _PTR],
}
_RESTYPES = {"mdb_error_string": ctypes.c_char_p, "mdb_conv2d_forward_workspace_bytes": ctypes.c_longlong}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m monodetr_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mdb_error_string(rc)
        raise RuntimeError(f"monodetr_b200 {what} failed (code {rc}): {msg.decode() if msg else '?'}")


# number of kernels of THIS library launched so far in this process (bench.py reports the per-step delta)
_launches = 0


def count(n: int = 1):
    global _launches
    _launches += n


def launch_count() -> int:
    return _launches
