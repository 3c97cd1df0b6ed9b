"""ctypes binding of libsta_b200.so (include/sta_b200.h).

There is no fallback: if the shared object is missing or a call returns non-zero, a RuntimeError is raised.
The structs below mirror include/sta_b200.h field for field.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsta_b200.so"

EXPORTED_SYMBOLS = (
    "sta_version",
    "sta_last_error",
    "sta_device_error",
    "sta_sattn_fwd",
    "sta_sattn_bwd",
    "sta_xattn_fwd",
    "sta_xattn_bwd",
    "sta_groupnorm_fwd",
    "sta_groupnorm_bwd",
    "sta_add_layernorm_fwd",
    "sta_add_layernorm_bwd",
    "sta_geglu_fwd",
    "sta_geglu_bwd",
    "sta_upsample2x_fwd",
    "sta_upsample2x_bwd",
    "sta_plms_step_fwd",
    "sta_plms_step_bwd",
    "sta_probe_gemm",
    "sta_probe_tmem_bw",
    "sta_debug_read",
)


class SattnFwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("lse", C.c_void_p),
        ("batch", C.c_int32), ("n", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("q_token_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("k_token_stride", C.c_int64), ("k_batch_stride", C.c_int64),
        ("v_token_stride", C.c_int64), ("v_batch_stride", C.c_int64),
        ("o_token_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("scale", C.c_float),
    ]


class SattnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("d_out", C.c_void_p),
        ("lse", C.c_void_p), ("d_q", C.c_void_p), ("d_k", C.c_void_p), ("d_v", C.c_void_p),
        ("dq_accum", C.c_void_p), ("delta", C.c_void_p),
        ("batch", C.c_int32), ("n", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("q_token_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("k_token_stride", C.c_int64), ("k_batch_stride", C.c_int64),
        ("v_token_stride", C.c_int64), ("v_batch_stride", C.c_int64),
        ("o_token_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("do_token_stride", C.c_int64), ("do_batch_stride", C.c_int64),
        ("scale", C.c_float), ("dqkv_token_stride", C.c_int64),
    ]


class XattnFwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k_ctx", C.c_void_p), ("v_ctx", C.c_void_p), ("mask", C.c_void_p), ("coef", C.c_void_p),
        ("out", C.c_void_p), ("lse", C.c_void_p),
        ("prompts", C.c_int32), ("n", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("n_obj", C.c_int32), ("ctx_len", C.c_int32),
        ("q_token_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("o_token_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("scale", C.c_float),
    ]


class XattnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k_ctx", C.c_void_p), ("v_ctx", C.c_void_p), ("mask", C.c_void_p), ("coef", C.c_void_p),
        ("lse", C.c_void_p), ("d_out", C.c_void_p), ("d_q", C.c_void_p), ("d_coef", C.c_void_p),
        ("prompts", C.c_int32), ("n", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("n_obj", C.c_int32), ("ctx_len", C.c_int32),
        ("q_token_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("do_token_stride", C.c_int64), ("do_batch_stride", C.c_int64),
        ("scale", C.c_float),
        ("out", C.c_void_p), ("o_token_stride", C.c_int64), ("o_batch_stride", C.c_int64),
    ]


class GroupNormArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("d_out", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("out", C.c_void_p),
        ("stats", C.c_void_p), ("bwd_stats", C.c_void_p),
        ("batch", C.c_int32), ("hw", C.c_int32), ("channels", C.c_int32), ("silu", C.c_int32), ("eps", C.c_float),
        ("x_bias", C.c_void_p), ("x_bias_stride", C.c_int64), ("d_res", C.c_void_p), ("d_res_stride", C.c_int64),
        ("x1", C.c_void_p), ("x_cat", C.c_void_p), ("out1", C.c_void_p), ("c_split", C.c_int32),
    ]


class AddLayerNormArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("sum_out", C.c_void_p), ("y", C.c_void_p), ("stats", C.c_void_p),
        ("rows", C.c_int32), ("channels", C.c_int32), ("eps", C.c_float),
    ]


class AddLayerNormBwdArgs(C.Structure):
    _fields_ = [
        ("d_y", C.c_void_p), ("d_sum", C.c_void_p), ("xs", C.c_void_p), ("stats", C.c_void_p), ("gamma", C.c_void_p),
        ("d_x", C.c_void_p), ("rows", C.c_int32), ("channels", C.c_int32),
    ]


class GegluArgs(C.Structure):
    _fields_ = [("proj", C.c_void_p), ("d_out", C.c_void_p), ("out", C.c_void_p), ("rows", C.c_int32), ("inner", C.c_int32)]


class Upsample2xArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("channels", C.c_int32)]


class PlmsStepArgs(C.Structure):
    _fields_ = [
        ("eps", C.c_void_p), ("x", C.c_void_p), ("old", C.c_void_p * 3), ("e_t", C.c_void_p), ("x_prev", C.c_void_p),
        ("pred_x0", C.c_void_p), ("prompts", C.c_int32), ("elems", C.c_int64), ("guidance", C.c_float), ("w_e", C.c_float),
        ("w_old", C.c_float * 3), ("a_x", C.c_float), ("a_e", C.c_float), ("p_x", C.c_float), ("p_e", C.c_float),
    ]


class PlmsStepBwdArgs(C.Structure):
    _fields_ = [
        ("g_x_prev", C.c_void_p), ("g_e_t", C.c_void_p), ("g_eps", C.c_void_p), ("g_x", C.c_void_p), ("g_old", C.c_void_p * 3),
        ("prompts", C.c_int32), ("elems", C.c_int64), ("guidance", C.c_float), ("w_e", C.c_float), ("w_old", C.c_float * 3),
        ("a_x", C.c_float), ("a_e", C.c_float),
    ]


class ProbeArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_int32), ("a_tensor_rows", C.c_int32), ("a_cols", C.c_int32),
        ("a_in_tmem", C.c_int32),
        ("b", C.c_void_p), ("b_rows", C.c_int32), ("b_tensor_rows", C.c_int32), ("b_cols", C.c_int32),
        ("a_desc_hi", C.c_uint64), ("b_desc_hi", C.c_uint64),
        ("nk", C.c_int32), ("a_off", C.c_uint32 * 16), ("b_off", C.c_uint32 * 16),
        ("idesc", C.c_uint32), ("n", C.c_int32), ("out", C.c_void_p), ("smem_dump", C.c_void_p),
        ("dump_bytes", C.c_int32), ("reps", C.c_int32), ("cycles", C.c_void_p),
    ]


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; raise loudly if it has not been built (`python -m ...build` / build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("STA_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the attention kernels."
        )
    lib = C.CDLL(str(path))
    lib.sta_version.restype = C.c_int
    lib.sta_last_error.restype = C.c_char_p
    lib.sta_device_error.argtypes = [C.POINTER(C.c_uint), C.c_int]
    for name, argt in (
        ("sta_sattn_fwd", SattnFwdArgs), ("sta_sattn_bwd", SattnBwdArgs),
        ("sta_xattn_fwd", XattnFwdArgs), ("sta_xattn_bwd", XattnBwdArgs),
        ("sta_probe_gemm", ProbeArgs), ("sta_groupnorm_fwd", GroupNormArgs), ("sta_groupnorm_bwd", GroupNormArgs),
        ("sta_add_layernorm_fwd", AddLayerNormArgs), ("sta_add_layernorm_bwd", AddLayerNormBwdArgs),
        ("sta_geglu_fwd", GegluArgs), ("sta_geglu_bwd", GegluArgs),
        ("sta_upsample2x_fwd", Upsample2xArgs), ("sta_upsample2x_bwd", Upsample2xArgs),
        ("sta_plms_step_fwd", PlmsStepArgs), ("sta_plms_step_bwd", PlmsStepBwdArgs),
    ):
        if not hasattr(lib, name):  # reported by tests/test_cabi.py; calling it raises AttributeError
            continue
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(argt), C.c_void_p]
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sta_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def device_error(clear: bool = True) -> int:
    """Return the device error word (0 = none).  Synchronises."""
    code = C.c_uint(0)
    load().sta_device_error(C.byref(code), 1 if clear else 0)
    return int(code.value)
