"""ctypes binding of libaga_b200.so (the C ABI declared in include/aga_b200.h).

There is no CPU or eager fallback: every op of this package goes through this library, and loading
fails loudly when the shared object has not been built (run ``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C <pkg>/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libaga_b200.so")

AGA_OK = 0
AGA_F32, AGA_BF16 = 0, 1
EXPORT_NONE, EXPORT_LOGITS, EXPORT_PROBS = 0, 1, 2
ATTN_AUTO, ATTN_SIMT, ATTN_TCGEN05 = 0, 1, 2


class AgaError(RuntimeError):
    pass


class AttnParams(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("impl", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("Tq", C.c_int32), ("Tk", C.c_int32),
        ("causal", C.c_int32), ("export_kind", C.c_int32), ("export_lo", C.c_int32), ("export_hi", C.c_int32),
        ("q_stride_b", C.c_int64), ("q_stride_t", C.c_int64),
        ("k_stride_b", C.c_int64), ("k_stride_t", C.c_int64),
        ("v_stride_b", C.c_int64), ("v_stride_t", C.c_int64),
        ("o_stride_b", C.c_int64), ("o_stride_t", C.c_int64),
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("lse", C.c_void_p), ("head_sel", C.c_void_p), ("export_buf", C.c_void_p), ("kv_len", C.c_void_p),
        ("guided_pattern", C.c_void_p), ("guided_part", C.c_void_p), ("guided_early", C.c_int32),
    ]


class AttnBwdParams(C.Structure):
    _fields_ = [
        ("fwd", AttnParams),
        ("dout", C.c_void_p), ("d_export", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("d_guided_part", C.c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/aga_b200.h declares (tests check this)
_SIGNATURES = {
    "aga_version": (C.c_int, []),
    "aga_status_str": (C.c_char_p, [C.c_int]),
    "aga_last_cuda_error": (C.c_int, []),
    "aga_launch_count": (C.c_uint64, []),
    "aga_device_is_sm100": (C.c_int, [C.c_int]),
    "aga_logmel_packed_filter_bytes": (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    "aga_logmel_pack_filters": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_logmel_workspace_bytes": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_size_t)]),
    "aga_logmel_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_logmel_filters_banded": (C.c_int, [C.c_void_p, C.c_int]),
    "aga_logmel_tc_packed_bytes": (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    "aga_logmel_tc_build_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "aga_logmel_tc_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_logmel_tc_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_attn_fwd_workspace_bytes": (C.c_int, [C.POINTER(AttnParams), C.POINTER(C.c_size_t)]),
    "aga_attn_fwd": (C.c_int, [C.POINTER(AttnParams), C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_attn_bwd_workspace_bytes": (C.c_int, [C.POINTER(AttnBwdParams), C.POINTER(C.c_size_t)]),
    "aga_attn_bwd": (C.c_int, [C.POINTER(AttnBwdParams), C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_guided_loss_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "aga_guided_loss_fwd_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_attention_pattern": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                        C.c_void_p]),
    "aga_head_vote": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aga_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aga_layernorm_pair_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "aga_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aga_layernorm_bwd_acc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aga_ls_ce_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_float,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aga_ls_ce_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_float,
                                C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "aga_linear_residual_workspace_bytes": (C.c_int, [C.POINTER(C.c_size_t)]),
    "aga_linear_residual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                      C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_gemm_gelu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int,
                                C.c_void_p]),
    "aga_gelu_bwd_colsum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aga_gelu_bwd_colsum_acc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aga_wgrad_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "aga_flat_grad_norm_workspace_bytes": (C.c_int, [C.POINTER(C.c_size_t)]),
    "aga_flat_grad_norm": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "aga_flat_adamw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises AgaError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AgaError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. There is no fallback path; "
                "build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


TORCH_EXT_PATH = os.path.join(_PKG_DIR, "aga_torch.so")
_torch_ops = None


def torch_ops():
    """``torch.ops.aga``: the TORCH_LIBRARY operators of csrc/torch_ext/aga_torch.cpp (a thin C++ layer over the C ABI
    above: tensors in, raw pointers + the current CUDA stream down, status codes -> exceptions).  Every per-step call of
    ``ops`` goes through it; this ctypes binding keeps the host-pointer setup calls and the symbol tests."""
    global _torch_ops
    if _torch_ops is None:
        import torch
        if not os.path.exists(TORCH_EXT_PATH):
            raise AgaError(f"{TORCH_EXT_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        lib()  # libaga_b200.so first (the extension links against it)
        torch.ops.load_library(TORCH_EXT_PATH)
        _torch_ops = _Ops(torch.ops.aga)
    return _torch_ops


class _Ops:
    """``torch.ops.aga`` with the extension's TORCH_CHECK failures (status codes of the C ABI, wrong devices / dtypes)
    re-raised as :class:`AgaError`, the package's one error type."""

    def __init__(self, ns):
        self._ns = ns

    def __getattr__(self, name):
        fn = getattr(self._ns, name)

        def call(*args):
            try:
                return fn(*args)
            except AgaError:
                raise
            except (RuntimeError, NotImplementedError) as e:
                raise AgaError(str(e).split("\n")[0]) from None

        call.__name__ = name
        self.__dict__[name] = call
        return call


def check(status: int, what: str) -> None:
    if status != AGA_OK:
        l = lib()
        msg = l.aga_status_str(status).decode()
        extra = f" (cudaError {l.aga_last_cuda_error()})" if status == -3 else ""
        raise AgaError(f"{what} failed: {msg}{extra}")


def launch_count() -> int:
    return int(lib().aga_launch_count())
