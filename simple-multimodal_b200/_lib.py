"""ctypes binding of libb200fusion.so (include/b200_fusion.h).

The library is the only compute backend: if it is missing or a call fails, this module raises --
there is no PyTorch/CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200fusion.so")

F32, BF16 = 0, 1
EPI_RELU, EPI_OUT_F32, EPI_ACCUM = 1, 2, 4


class B200FusionError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("a_layout", C.c_int32), ("b_layout", C.c_int32),
                ("A", C.c_void_p), ("lda", C.c_int64),
                ("B", C.c_void_p), ("ldb", C.c_int64),
                ("C", C.c_void_p), ("ldc", C.c_int64),
                ("bias", C.c_void_p),
                ("residual", C.c_void_p), ("ldr", C.c_int64),
                ("relu_mask", C.c_void_p), ("ldm", C.c_int64),
                ("alpha", C.c_float), ("flags", C.c_int32), ("dtype", C.c_int32), ("split_k", C.c_int32),
                ("dropout_p", C.c_float), ("drop_seed_lo", C.c_uint32), ("drop_seed_hi", C.c_uint32),
                ("colsum", C.c_void_p),
                ("sign_bits_out", C.c_void_p), ("sign_bits", C.c_void_p), ("ldsb", C.c_int64)]


class AttnArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32), ("D", C.c_int32),
                ("Q", C.c_void_p), ("ldq", C.c_int64),
                ("K", C.c_void_p), ("ldk", C.c_int64),
                ("V", C.c_void_p), ("ldv", C.c_int64),
                ("O", C.c_void_p), ("ldo", C.c_int64),
                ("LSE", C.c_void_p), ("scale", C.c_float), ("dtype", C.c_int32),
                ("dO", C.c_void_p), ("lddo", C.c_int64),
                ("dQ", C.c_void_p), ("lddq", C.c_int64),
                ("dK", C.c_void_p), ("lddk", C.c_int64),
                ("dV", C.c_void_p), ("lddv", C.c_int64),
                ("delta", C.c_void_p),
                ("dbq", C.c_void_p), ("dbk", C.c_void_p), ("dbv", C.c_void_p),
                ("dropout_p", C.c_float), ("drop_seed_lo", C.c_uint32), ("drop_seed_hi", C.c_uint32),
                ("pool_sum", C.c_void_p),
                ("bwd_ws", C.c_void_p), ("bwd_ws_bytes", C.c_int64)]


_lib = None

# Optional live per-entry-point profile (tools/step_profile.py): when CALL_PROFILE is a list, every compute entry point is
# bracketed by CUDA events on the launching stream and (name, start, stop) is appended.  Off (None) in normal use.
CALL_PROFILE = None
_NOT_KERNELS = ("b200f_last_error", "b200f_launch_count", "b200f_infonce_workspace_bytes", "b200f_version", "b200f_debug_set", "b200f_attn_pool_parts")


class _ProfiledLib:
    def __getattr__(self, name):
        fn = getattr(_lib, name)
        if name in _NOT_KERNELS:
            return fn

        def call(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            CALL_PROFILE.append((name, e0, e1))
            return rc
        return call


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib if CALL_PROFILE is None else _ProfiledLib()
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200FusionError(
                f"{LIB_PATH} not found: build it with `python simple-multimodal_b200/build.py` "
                "(there is no fallback path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.b200f_last_error.restype = C.c_char_p
        _lib.b200f_launch_count.restype = C.c_ulonglong
        _lib.b200f_infonce_workspace_bytes.restype = C.c_size_t
        _lib.b200f_attn_pool_parts.restype = C.c_int32
        _lib.b200f_attn_bwd_ws_bytes.restype = C.c_int64
        for kv in filter(None, os.environ.get("B200F_DEBUG_SET", "").split(",")):      # A/B experiments only: "key=value,key=value"
            key, val = kv.split("=")                                                      # (b200f_debug_set, csrc/api.cu)
            _lib.b200f_debug_set(int(key), int(val))
    return _lib if CALL_PROFILE is None else _ProfiledLib()


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b200f_last_error().decode(errors="replace")
        raise B200FusionError(f"{what} failed with status {rc}: {msg}")


def stream_ptr() -> C.c_void_p:
    """the CURRENT device's current stream: every entry point launches in the current CUDA context, so callers working on another
    device than the current one must switch first (`require_cuda` checks that operands live on the current device)"""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise B200FusionError(f"unsupported dtype {t}: the fusion kernels take float32 or bfloat16")


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def require_cuda(*tensors) -> None:
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise B200FusionError("b200 fusion kernels need CUDA tensors (no CPU fallback exists)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise B200FusionError(f"operand on cuda:{t.device.index} but the current device is cuda:{cur}: the kernels launch on the current "
                                  "device's stream -- wrap the call in torch.cuda.device(tensor.device)")


def launch_count() -> int:
    return int(lib().b200f_launch_count())
