"""ctypes binding of libdpot_b200.so (include/dpot_b200.h).

This is the whole reference-side binding: plain ``ctypes`` over a C ABI, no torch types in any
signature.  The library is loaded lazily (after DataLoader workers fork) and the product
fails loudly when it is missing -- there is no CPU / torch fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdpot_b200.so")

ACT_IDS = {"gelu": 0, "tanh": 1, "sigmoid": 2, "relu": 3, "leaky_relu": 4, "softplus": 5, "ELU": 6, "silu": 7}
ACT_NONE = -1
GEMM_AUTO, GEMM_SIMT, GEMM_TC, GEMM_TC16 = 0, 1, 2, 3
FMT_F32, FMT_HL16, FMT_HL16G32 = 0, 1, 2
A_PLAIN, A_PATCH = 0, 1

c_f32p = C.c_void_p  # device pointers travel as integers


ABI_VERSION = 2          # DPOT_ABI_VERSION of include/dpot_b200.h this binding was written against


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", c_f32p), ("lda", C.c_int64), ("W", c_f32p), ("ldw", C.c_int64), ("C", c_f32p), ("ldc", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("bias", c_f32p),
        ("rowbias", c_f32p), ("rowbias_period", C.c_int32), ("ldrb", C.c_int64),
        ("residual", c_f32p), ("ldr", C.c_int64),
        ("act", C.c_int32),
        ("a_scale", c_f32p), ("a_shift", c_f32p), ("a_rows_per_sample", C.c_int32),
        ("c_scale", c_f32p), ("c_shift", c_f32p), ("c_rows_per_sample", C.c_int32),
        ("c_group", C.c_int32), ("c_group_stride", C.c_int64),
        ("batch", C.c_int32), ("strideA", C.c_int64), ("strideW", C.c_int64), ("strideC", C.c_int64),
        ("strideBias", C.c_int64),
        ("a_mode", C.c_int32), ("pX", C.c_int32), ("pY", C.c_int32), ("pT", C.c_int32), ("pC", C.c_int32),
        ("pP", C.c_int32),
        ("engine", C.c_int32),
        ("out_stats", c_f32p), ("stats_groups", C.c_int32), ("stats_rows_per_sample", C.c_int32),
        ("C_pre", c_f32p), ("dact_src", c_f32p), ("dact", C.c_int32), ("c_mode", C.c_int32),
        ("a_fmt", C.c_int32), ("w_fmt", C.c_int32), ("c_fmt", C.c_int32),
        ("a_lo_off", C.c_int64), ("w_lo_off", C.c_int64), ("c_lo_off", C.c_int64),
        ("a_trans", C.c_int32), ("w_trans", C.c_int32),
        ("k_split", C.c_int32), ("k_chunk", C.c_int64), ("strideC_split", C.c_int64),
        ("ld_pre", C.c_int64), ("stride_pre", C.c_int64), ("ld_dact", C.c_int64), ("stride_dact", C.c_int64),
        ("out_colsum", C.c_void_p),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("X", c_f32p), ("ldx", C.c_int64), ("Y", c_f32p), ("ldy", C.c_int64), ("dW", c_f32p), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("y_scale", c_f32p), ("y_shift", c_f32p), ("y_rows_per_sample", C.c_int32),
        ("batch", C.c_int32), ("strideX", C.c_int64), ("strideY", C.c_int64), ("strideW", C.c_int64),
        ("y_mode", C.c_int32), ("pX", C.c_int32), ("pY", C.c_int32), ("pT", C.c_int32), ("pC", C.c_int32),
        ("pP", C.c_int32),
        ("accumulate", C.c_int32),
    ]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "img_size", "patch_size", "in_channels", "out_channels", "in_timesteps", "out_timesteps",
        "n_blocks", "embed_dim", "out_layer_dim", "depth", "modes", "hidden_dim", "n_cls",
        "normalize", "act", "time_agg")]


BLOCK_FIELDS = ("norm1_w", "norm1_b", "w1", "b1", "w2", "b2", "norm2_w", "norm2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")


class BlockParams(C.Structure):
    _fields_ = [(n, c_f32p) for n in BLOCK_FIELDS]


PARAM_FIELDS = ("pos_embed", "pe0_w", "pe0_b", "pe2_w", "pe2_b", "tagg_w", "tagg_gamma",
                "cls0_w", "cls0_b", "cls2_w", "cls2_b", "cls4_w", "cls4_b",
                "out0_w", "out0_b", "out2_w", "out2_b", "out4_w", "out4_b",
                "mu_w", "mu_b", "sigma_w", "sigma_b")


class Params(C.Structure):
    _fields_ = ([(n, c_f32p) for n in PARAM_FIELDS] + [("blocks", C.POINTER(BlockParams))] +
                [(n, c_f32p) for n in ("grid_x", "grid_y", "grid_t", "temb")])


# name -> (restype, argtypes); every symbol include/dpot_b200.h declares
_i32, _i64, _f, _d, _p = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p
SIGNATURES = {
    "dpot_abi_version": (C.c_int, []),
    "dpot_last_error_string": (C.c_char_p, []),
    "dpot_device_supported": (C.c_int, []),
    "dpot_launch_count": (C.c_longlong, []),
    "dpot_tc_available": (C.c_int, []),
    "dpot_tc_set_flush": (C.c_int, [C.c_int]),
    "dpot_tc_set_trunc": (C.c_int, [C.c_int]),
    "dpot_tc_set_trace": (None, [C.c_void_p]),
    "dpot_gemm": (C.c_int, [C.POINTER(GemmArgs), _p]),
    "dpot_split_f16": (C.c_int, [_p, _i64, _i64, _i32, _p, _p, _i32, _p, _i64, _i64, _p]),
    "dpot_tc16_available": (C.c_int, []),
    "dpot_tc16_set_pair": (None, [_i32]),
    "dpot_out_tail_set_engine": (None, [_i32]),
    "dpot_patch_embed_set_engine": (None, [_i32]),
    "dpot_tc16_set_debug": (None, [_i32]),
    "dpot_tc16_set_precision": (C.c_int, [_i32]),
    "dpot_tc16_set_chain": (C.c_int, [_i32]),
    "dpot_gemm_chained": (C.c_int, [C.POINTER(GemmArgs), _i32, _p, _i64, _p]),
    "dpot_set_pdl": (None, [_i32]),
    "dpot_set_cls_overlap": (None, [_i32]),
    "dpot_set_cls_engine": (None, [_i32]),
    "dpot_set_sm_budget": (C.c_int, [_i32]),
    "dpot_gn_stats": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "dpot_gn_finalize": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _f, _p, _p, _p]),
    "dpot_afno_fft_fwd": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _f, _p]),
    "dpot_afno_fft_fwd16": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p]),
    "dpot_afno_fft_fwd16_gn": (C.c_int, [_p, _p, _p, _p, _i32, _f, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p]),
    "dpot_afno_fft_inv_gn": (C.c_int, [_p, _p, _p, _p, _p, _i32, _f, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "dpot_afno_fused_supported": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "dpot_afno_fused_packed_floats": (C.c_int64, [_i32]),
    "dpot_afno_fused_pack": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _p, _p]),
    "dpot_afno_fused": (C.c_int, [_p, _p, _p, _p, _i32, _f, _i32, _i32, _i32, _i32, _p, _i32, _p, _p, _p, _p]),
    "dpot_afno_fused_gn2": (C.c_int, [_p, _p, _p, _p, _i32, _f, _i32, _i32, _i32, _i32, _p, _i32, _p, _p, _p, _p, _p, _p, _f, _p]),
    "dpot_afno_set_fused": (None, [_i32]),
    "dpot_afno_set_fused_gn2": (C.c_int, [_i32]),
    "dpot_afno_fused_set_trace": (None, [C.c_void_p]),
    "dpot_split_f16_gn": (C.c_int, [_p, _i64, _i64, _i32, _p, _p, _p, _i32, _f, _i32, _p, _i64, _i64, _p]),
    "dpot_afno_fft_inv": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _i32, _f, _p]),
    "dpot_wgrad": (C.c_int, [C.POINTER(WgradArgs), _p]),
    "dpot_colsum": (C.c_int, [_p, _i64, _i32, _i32, _p, _i32, _p]),
    "dpot_transpose": (C.c_int, [_p, _i64, _p, _i64, _i32, _i32, _i32, _i64, _i64, _p]),
    "dpot_gn_apply": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p]),
    "dpot_gn_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _f, _p, _p, _p, _p, _p]),
    "dpot_act_bwd": (C.c_int, [_p, _p, _i32, _i64, _p, _p]),
    "dpot_pixel_shuffle": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "dpot_unpack_afno_grad": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p]),
    "dpot_pack_afno": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p]),
    "dpot_pack_patch": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "dpot_fold_timeagg": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "dpot_pack_out": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "dpot_out_tail": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _i32, _p, _p]),
    "dpot_spatial_mean": (C.c_int, [_p, _i32, _i32, _i32, _p, _p]),
    "dpot_spatial_mean16": (C.c_int, [_p, _i32, _i32, _i32, _p, _p]),
    "dpot_spatial_mean16s": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p]),
    "dpot_input_stats": (C.c_int, [_p, _i32, _i64, _i32, _i32, _p, _p, _p, _p]),
    "dpot_window_advance": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p]),
    "dpot_ring_insert": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _p]),
    "dpot_patch_embed": (C.c_int, [_p, _i32, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p]),
    "dpot_adam_step": (C.c_int, [_p, _p, _p, _p, _p, _i64, _d, _d, _d, _d, _d, _i32, _i32, _d, _p]),
    "dpot_adam_step_multi": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _d, _d, _d, _d, _d, _p, _i32, _d, _p]),
    "dpot_lamb_step_multi": (C.c_int, [_p, _p, _p, _p, _p, _i32, _d, _d, _d, _d, _d, _d, _p, _i32, _i32, _p, _p, _p]),
    "dpot_adam_step_multi_clip": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _d, _d, _d, _d, _d, _p, _i32, _d, _p, _d, _p]),
    "dpot_grad_sqnorm": (C.c_int, [_p, _p, _i32, _p, _p]),
    "dpot_chan_sumsq": (C.c_int, [_p, _i32, _i64, _i32, _p, _p]),
    "dpot_noise_inject": (C.c_int, [_p, _i32, _i64, _i32, _f, C.c_uint64, C.c_uint64, _p, _p, _p]),
    "dpot_noise_inject_bwd": (C.c_int, [_p, _p, _i32, _i64, _i32, _f, C.c_uint64, C.c_uint64, _p, _p, _p, _p]),
    "dpot_lp_loss": (C.c_int, [_p, _p, _p, _i32, _i64, _i32, _i32, _p, _p, _p, _i32, _p]),
    "dpot_lp_loss_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i64, _i32, _i32, _p, _p]),
    "dpot_packed_floats": (C.c_int64, [C.POINTER(Config)]),
    "dpot_workspace_floats": (C.c_int64, [C.POINTER(Config), _i32]),
    "dpot_pack_weights": (C.c_int, [C.POINTER(Config), C.POINTER(Params), _p, _p]),
    "dpot_forward": (C.c_int, [C.POINTER(Config), C.POINTER(Params), _p, _p, _i32, _p, _p, _p, _i32, _p]),
    "dpot_forward_ring": (C.c_int, [C.POINTER(Config), C.POINTER(Params), _p, _p, _i32, _i32, _p, _p, _p, _i32, _p]),
    "dpot_rollout_step": (C.c_int, [C.POINTER(Config), C.POINTER(Params), _p, _p, _i32, _i32, _p, _p, _p, _i32, _p, _i32, _i32, _p]),
    "dpot_out_tail_tc": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _i32, _p, _p, _p,
                                   _i32, _i32, _i32, _i32, _p]),
    "dpot_out_tail_tc_supported": (C.c_int, [_i32, _i32, _i32]),
    "dpot_afno_fft_fwd16w": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _f, _p, _p]),
    "dpot_assemble_batch": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "dpot_train_supported": (C.c_int, [C.POINTER(Config)]),
    "dpot_train_tape_floats": (C.c_int64, [C.POINTER(Config), _i32]),
    "dpot_train_scratch_floats": (C.c_int64, [C.POINTER(Config), _i32]),
    "dpot_train_wprep_floats": (C.c_int64, [C.POINTER(Config)]),
    "dpot_train_prepare": (C.c_int, [C.POINTER(Config), C.POINTER(Params), _p, _p, _p, _p]),
    "dpot_train_forward": (C.c_int, [C.POINTER(Config), C.POINTER(Params), _p, _p, _p, _i32, _p, _p, _p, _p, _p]),
    "dpot_train_backward": (C.c_int, [C.POINTER(Config), C.POINTER(Params), _p, _p, _p, _i32, _p, _p, _p, _p,
                                      C.POINTER(Params), _p, _p, _p]),
    "dpot_out_tail_ring": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _i32, _p, _p, _p,
                                     _i32, _i32, _i32, _i32, _p]),
}

_lib = None
_lock = threading.Lock()


class DpotLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load (once) and type the shared library.  Raises DpotLibraryError when it is absent:
    the hot path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DpotLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m dpot_b200.build` (nvcc, sm_100a). "
                "dpot_b200 has no CPU or eager-PyTorch fallback for the hot path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.dpot_abi_version() != ABI_VERSION:
            raise DpotLibraryError("libdpot_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().dpot_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libdpot_b200 {what} failed (rc={rc}): {msg}")


def ptr(t) -> int:
    """data pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
