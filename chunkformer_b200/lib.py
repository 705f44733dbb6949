"""ctypes binding of the C-ABI shared library (include/chunkformer_b200.h).

The library is built in-tree by chunkformer_b200.build (nvcc, sm_100a).  There is no fallback: if the shared object is
missing or a compute entry point reports an error, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

from . import build as _build

CF_F32, CF_BF16 = 0, 1
CF_OK, CF_ERR_INVALID, CF_ERR_CUDA, CF_ERR_STATE, CF_ERR_WORKSPACE = 0, -1, -2, -3, -4
EPI_BF16, EPI_GLU, EPI_F32, EPI_ARGMAX = 0, 1, 2, 4
ACT_NONE, ACT_RELU, ACT_SILU = 0, 1, 2
FAMILY_OTHER, FAMILY_FFN_W1, FAMILY_FFN_W2, FAMILY_FFN_FUSED = 0, 1, 2, 3


class CfConfig(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("d_model", "heads", "ffn", "layers", "kernel", "vocab", "feat_dim", "has_cmvn", "conv_norm")]


# every symbol include/chunkformer_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "cf_create": (c_int, [POINTER(CfConfig), c_int, POINTER(c_void_p)]),
    "cf_destroy": (None, [c_void_p]),
    "cf_last_error": (c_char_p, [c_void_p]),
    "cf_version": (c_char_p, []),
    "cf_launch_count": (ctypes.c_longlong, []),
    "cf_encode_feature_events": (c_int, [c_void_p, c_int, POINTER(c_int64), POINTER(c_void_p)]),
    "cf_fbank_num_frames": (c_int64, [c_int64, c_int, c_int, c_int]),
    "cf_fbank": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cf_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "cf_kernel_timing_begin": (c_int, [c_void_p, ctypes.c_uint]),
    "cf_kernel_timing_end": (c_int, [c_void_p, c_int, POINTER(ctypes.c_double), POINTER(c_int)]),
    "cf_load_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int, POINTER(c_int64)]),
    "cf_finalize_weights": (c_int, [c_void_p]),
    "cf_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int64),
                               POINTER(c_void_p)]),
    "cf_plan_create_padded": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_void_p)]),
    "cf_plan_destroy": (None, [c_void_p]),
    "cf_plan_num_chunks": (c_int, [c_void_p]),
    "cf_plan_rows": (c_int, [c_void_p]),
    "cf_plan_tables": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32)]),
    "cf_plan_masks": (c_int, [c_void_p, POINTER(c_uint8), POINTER(c_uint8)]),
    "cf_plan_chunk_table": (c_int, [c_void_p, POINTER(c_int32)]),
    "cf_workspace_bytes": (c_size_t, [c_void_p, c_void_p]),
    "cf_plan_pin": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cf_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                          c_size_t, c_void_p]),
    "cf_encode_streams": (c_int, [c_void_p, c_int, c_int, c_int]),
    "cf_encode_output_rows": (c_int64, [c_void_p]),
    "cf_ctc_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int]),
    "cf_ctc_greedy": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cf_ctc_compact_workspace_bytes": (c_size_t, [c_int64]),
    "cf_ctc_compact": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "cf_rnnt_create": (c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    "cf_rnnt_destroy": (None, [c_void_p]),
    "cf_rnnt_last_error": (c_char_p, [c_void_p]),
    "cf_rnnt_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "cf_rnnt_load_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int, POINTER(c_int64)]),
    "cf_rnnt_finalize_weights": (c_int, [c_void_p]),
    "cf_rnnt_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int]),
    "cf_rnnt_greedy": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_int64), POINTER(c_int32), c_int, c_int, c_int, c_void_p,
                               c_void_p, c_void_p, POINTER(c_int64), c_void_p, c_size_t, c_void_p]),
    "cf_op_gemm": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                           c_int64, c_float, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "cf_op_gemm_ln": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_float,
                              c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                              c_void_p, c_int, c_void_p]),
    "cf_op_ffn": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_float,
                          c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "cf_op_layernorm": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int64, c_void_p]),
    "cf_op_dwconv": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                             c_int, c_void_p]),
    "cf_op_frontend_conv": (c_int, [c_int, c_int, c_void_p, POINTER(c_int64), POINTER(c_int32), c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "cf_op_attention": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_void_p]),
}

_LIB = None


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building first if the .so is absent or stale and nvcc is available) and type every entry point."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("CHUNKFORMER_B200_LIB") or _build.LIB      # A/B builds of the same tree (tools only)
    if build_if_missing and path == _build.LIB and (not os.path.exists(path) or _build.needs_build()):
        try:
            _build.build()
        except Exception as e:  # no nvcc on this machine: use the shipped .so if there is one
            if not os.path.exists(path):
                raise RuntimeError(f"chunkformer_b200: native library missing and cannot be built: {e}") from e
    if not os.path.exists(path):
        raise RuntimeError(f"chunkformer_b200: native library not found at {path} (no CPU fallback exists)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, handle=None, what: str = ""):
    if rc != 0:
        msg = load().cf_last_error(handle).decode("utf-8", "replace")
        raise RuntimeError(f"chunkformer_b200 {what} failed ({rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())
