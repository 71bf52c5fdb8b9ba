"""ctypes binding of libtcn_b200.so (include/tcn_b200.h).

There is no fallback: if the shared library is missing, or a tensor is not a CUDA tensor, the
call raises.  Build with ``python -m computervision_codes_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TCN_LIB_PATH") or os.path.join(_HERE, "csrc", "libtcn_b200.so")  # override: A/B builds (tools/exp)

_lib = None


class TcnError(RuntimeError):
    pass


class TapGemmArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int), ("x_unpadded", C.c_int),
        ("colscale", C.c_void_p), ("colscale_ld", C.c_int),
        ("wf", C.c_void_p), ("bias", C.c_void_p),
        ("y", C.c_void_p), ("ldy", C.c_int),
        ("residual", C.c_void_p), ("ldr", C.c_int),
        ("relu_mask", C.c_void_p), ("ldm", C.c_int),
        ("meta", C.c_void_p), ("nblk", C.c_int),
        ("c_in", C.c_int), ("n_out", C.c_int), ("ntaps", C.c_int), ("shift", C.c_int * 3),
        ("relu", C.c_int),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint), ("drop_stream", C.c_uint),
        ("in_drop_p", C.c_float),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("g", C.c_void_p), ("ldg", C.c_int), ("g_cols", C.c_int),
        ("x", C.c_void_p), ("ldx", C.c_int), ("x_unpadded", C.c_int),
        ("colscale", C.c_void_p), ("colscale_ld", C.c_int),
        ("meta", C.c_void_p), ("nblk", C.c_int),
        ("n_out", C.c_int), ("c_in", C.c_int), ("ntaps", C.c_int), ("shift", C.c_int * 3),
        ("dw", C.c_void_p), ("db", C.c_void_p),
        ("g_drop_p", C.c_float), ("drop_seed", C.c_uint), ("drop_stream", C.c_uint),
    ]


class LayerFwdArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("h", C.c_void_p),
        ("w1f", C.c_void_p), ("w2f", C.c_void_p), ("b1", C.c_void_p), ("b2", C.c_void_p),
        ("meta", C.c_void_p), ("nblk", C.c_int), ("channels", C.c_int),
        ("shift", C.c_int * 3),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint), ("drop_stream", C.c_uint),
    ]


class LayerFwdTcArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_rows", C.c_longlong), ("y", C.c_void_p), ("h", C.c_void_p),
        ("w1_hi", C.c_void_p), ("w1_lo", C.c_void_p), ("w2_hi", C.c_void_p), ("w2_lo", C.c_void_p),
        ("b1", C.c_void_p), ("b2", C.c_void_p),
        ("meta", C.c_void_p), ("nblk", C.c_int), ("channels", C.c_int),
        ("shift", C.c_int * 3),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint), ("drop_stream", C.c_uint),
        ("masks", C.c_void_p),
    ]


class LayerBwdTcArgs(C.Structure):
    _fields_ = [
        ("gy", C.c_void_p), ("g_rows", C.c_longlong), ("gu", C.c_void_p), ("gx", C.c_void_p),
        ("masks", C.c_void_p),
        ("w2t_hi", C.c_void_p), ("w2t_lo", C.c_void_p), ("w1t_hi", C.c_void_p), ("w1t_lo", C.c_void_p),
        ("meta", C.c_void_p), ("nblk", C.c_int), ("channels", C.c_int),
        ("shift", C.c_int * 3),
        ("drop_p", C.c_float),
    ]


class GemmTcArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int), ("x_rows", C.c_longlong), ("x_unpadded", C.c_int),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("bias", C.c_void_p),
        ("y", C.c_void_p), ("ldy", C.c_int),
        ("residual", C.c_void_p), ("ldr", C.c_int),
        ("relu_mask", C.c_void_p), ("ldm", C.c_int),
        ("meta", C.c_void_p), ("nblk", C.c_int),
        ("c_in", C.c_int), ("n_out", C.c_int), ("ntaps", C.c_int), ("shift", C.c_int * 3),
        ("relu", C.c_int),
        ("colscale", C.c_void_p), ("colscale_ld", C.c_int),
        ("in_drop_p", C.c_float), ("in_drop_rescale", C.c_int),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint), ("drop_stream", C.c_uint),
    ]


class WgradTcArgs(C.Structure):
    _fields_ = [
        ("g", C.c_void_p), ("ldg", C.c_int), ("g_cols", C.c_int), ("g_rows", C.c_longlong),
        ("x", C.c_void_p), ("ldx", C.c_int), ("x_rows", C.c_longlong), ("x_unpadded", C.c_int),
        ("colscale", C.c_void_p), ("colscale_ld", C.c_int),
        ("meta", C.c_void_p), ("nblk", C.c_int),
        ("n_out", C.c_int), ("c_in", C.c_int), ("ntaps", C.c_int), ("shift", C.c_int * 3),
        ("dw", C.c_void_p), ("db", C.c_void_p),
        ("g_drop_p", C.c_float), ("drop_seed", C.c_uint), ("drop_stream", C.c_uint),
    ]


class WgradLayerArgs(C.Structure):
    _fields_ = [
        ("gu", C.c_void_p), ("x", C.c_void_p), ("gy", C.c_void_p), ("h", C.c_void_p), ("rows", C.c_longlong),
        ("masks", C.c_void_p),
        ("meta", C.c_void_p), ("nblk", C.c_int), ("channels", C.c_int),
        ("shift", C.c_int * 3),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint), ("drop_stream", C.c_uint),
        ("dw1", C.c_void_p), ("db1", C.c_void_p), ("dw2", C.c_void_p), ("db2", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong),
        ("flags", C.c_int),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int), ("k", C.c_void_p), ("ldk", C.c_int), ("v", C.c_void_p), ("ldv", C.c_int),
        ("o", C.c_void_p), ("ldo", C.c_int), ("lse", C.c_void_p),
        ("dout", C.c_void_p), ("lddo", C.c_int), ("dq", C.c_void_p), ("lddq", C.c_int),
        ("dk", C.c_void_p), ("lddk", C.c_int), ("dv", C.c_void_p), ("lddv", C.c_int),
        ("seq_lo", C.c_void_p), ("seq_len", C.c_void_p), ("nseq", C.c_int), ("max_len", C.c_int),
        ("heads", C.c_int), ("head_dim", C.c_int), ("scale", C.c_float),
        ("delta", C.c_void_p),
    ]


class AttnTcArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int), ("k", C.c_void_p), ("ldk", C.c_int), ("v", C.c_void_p), ("ldv", C.c_int),
        ("o", C.c_void_p), ("ldo", C.c_int), ("p", C.c_void_p),
        ("dout", C.c_void_p), ("lddo", C.c_int), ("dq", C.c_void_p), ("lddq", C.c_int), ("dk", C.c_void_p),
        ("lddk", C.c_int), ("dv", C.c_void_p), ("lddv", C.c_int), ("dp", C.c_void_p),
        ("seq_lo", C.c_void_p), ("seq_len", C.c_void_p), ("nseq", C.c_int), ("rows", C.c_longlong),
        ("heads", C.c_int), ("head_dim", C.c_int), ("tmax", C.c_int), ("scale", C.c_float),
    ]


class KdAttnArgs(C.Structure):
    _fields_ = [
        ("s", C.c_void_p), ("lds", C.c_int),
        ("tea", C.c_void_p * 3), ("ldt", C.c_int),
        ("z", C.c_void_p * 3), ("ldz", C.c_int),
        ("tsum", C.c_void_p),
        ("gz", C.c_void_p * 3),
        ("gs", C.c_void_p), ("ldgs", C.c_int),
        ("gtea", C.c_void_p * 3),
        ("n_rows", C.c_int), ("feat_dim", C.c_int),
    ]


class ModelConfig(C.Structure):
    _fields_ = [
        ("layers_pg", C.c_int), ("layers_r", C.c_int), ("num_r", C.c_int), ("channels", C.c_int),
        ("in_dim", C.c_int), ("head_sizes", C.c_int * 4), ("causal", C.c_int),
        ("max_rows", C.c_int), ("max_seqs", C.c_int),
    ]


class BceArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("ldl", C.c_int),
        ("labels", C.c_void_p), ("ldlab", C.c_int), ("lab_unpadded", C.c_int),
        ("meta", C.c_void_p), ("nrows", C.c_int), ("ncols", C.c_int), ("zero_cols", C.c_int),
        ("pos_w", C.c_void_p), ("col_scale", C.c_void_p), ("col_unit", C.c_void_p), ("col_head", C.c_void_p),
        ("row_scale", C.c_float),
        ("loss", C.c_void_p),
        ("dl", C.c_void_p), ("lddl", C.c_int), ("grad_scale", C.c_float),
    ]


# name -> (restype, argtypes); every symbol include/tcn_b200.h declares
SIGNATURES = {
    "tcn_version": (C.c_int, []),
    "tcn_last_error": (C.c_char_p, []),
    "tcn_launch_count": (C.c_longlong, []),
    "tcn_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "tcn_prep_weight_floats": (C.c_longlong, [C.c_int] * 4),
    "tcn_prep_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tcn_tapgemm": (C.c_int, [C.POINTER(TapGemmArgs), C.c_void_p]),
    "tcn_wgrad": (C.c_int, [C.POINTER(WgradArgs), C.c_void_p]),
    "tcn_wgrad_tc": (C.c_int, [C.POINTER(WgradTcArgs), C.c_void_p]),
    "tcn_wgrad_tc_pair": (C.c_int, [C.POINTER(WgradTcArgs), C.POINTER(WgradTcArgs), C.c_void_p]),
    "tcn_wgrad_layer_workspace_bytes": (C.c_longlong, [C.c_int]),
    "tcn_wgrad_layer": (C.c_int, [C.POINTER(WgradLayerArgs), C.c_void_p]),
    "tcn_layer_fwd": (C.c_int, [C.POINTER(LayerFwdArgs), C.c_void_p]),
    "tcn_layer_fwd_tc": (C.c_int, [C.POINTER(LayerFwdTcArgs), C.c_void_p]),
    "tcn_layer_bwd_tc": (C.c_int, [C.POINTER(LayerBwdTcArgs), C.c_void_p]),
    "tcn_gemm_tc_supported": (C.c_int, [C.c_int, C.c_int]),
    "tcn_gemm_tc": (C.c_int, [C.POINTER(GemmTcArgs), C.c_void_p]),
    "tcn_split_weight_floats": (C.c_longlong, [C.c_int] * 4),
    "tcn_split_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "tcn_model_create": (C.c_int, [C.POINTER(ModelConfig), C.POINTER(C.c_void_p)]),
    "tcn_model_destroy": (None, [C.c_void_p]),
    "tcn_model_num_params": (C.c_longlong, [C.c_void_p]),
    "tcn_model_num_tensors": (C.c_int, [C.c_void_p]),
    "tcn_model_debug_ptr": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "tcn_model_param_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]),
    "tcn_model_bind": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "tcn_model_set_loss": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tcn_model_set_loss_norm": (C.c_int, [C.c_void_p, C.c_int]),
    "tcn_model_set_dropout": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "tcn_model_set_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint,
                                      C.c_void_p]),
    "tcn_model_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p]),
    "tcn_model_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_void_p]),
    "tcn_model_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.c_void_p]),
    "tcn_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "tcn_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p]),
    "tcn_attn_fwd": (C.c_int, [C.POINTER(AttnArgs), C.c_void_p]),
    "tcn_attn_bwd": (C.c_int, [C.POINTER(AttnArgs), C.c_void_p]),
    "tcn_attn_tc_supported": (C.c_int, [C.c_int] * 5),
    "tcn_attn_fwd_tc": (C.c_int, [C.POINTER(AttnTcArgs), C.c_void_p]),
    "tcn_attn_bwd_tc": (C.c_int, [C.POINTER(AttnTcArgs), C.c_void_p]),
    "tcn_dwconv_gelu_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p]),
    "tcn_dwconv_gelu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "tcn_axpby": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_longlong, C.c_void_p]),
    "tcn_bce_rows": (C.c_int, [C.POINTER(BceArgs), C.c_void_p]),
    "tcn_kd_attn_fwd": (C.c_int, [C.POINTER(KdAttnArgs), C.c_void_p]),
    "tcn_kd_attn_bwd": (C.c_int, [C.POINTER(KdAttnArgs), C.c_void_p]),
    "tcn_kd_kl_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                 C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "tcn_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_float, C.c_void_p, C.c_float,
                          C.c_void_p]),
    "tcn_ce_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                              C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "tcn_ap_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                              C.c_void_p]),
    "tcn_dropout_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                    C.c_uint, C.c_uint, C.c_void_p]),
    "tcn_dropout_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_uint, C.c_uint, C.c_void_p]),
    "tcn_sgd_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "tcn_sgd_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
}


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TcnError(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(run `python -m computervision_codes_b200.build`); there is no CPU or PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().tcn_last_error().decode("utf-8", "replace")
        raise TcnError(f"{what or 'tcn_b200'} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise TcnError("tcn_b200 kernels need CUDA tensors; there is no CPU fallback")
    return t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
