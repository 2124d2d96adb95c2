"""ctypes binding of libpg_b200.so (the C ABI declared in include/pg_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PG_LIB_PATH") or os.path.join(_HERE, "libpg_b200.so")

PG_F32, PG_BF16, PG_F16 = 0, 1, 2
EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_RES, EPI_GEGLU = 0, 1, 2, 3, 4, 5
MAX_DECODE_BATCH = 8            # rows one launch of the GEMV kernels can take
# from this batch size on the decode step runs the tensor-core (swap-AB skinny GEMM) path: the GEMV kernels turn
# ALU-bound as rows are added (batch 8: 4.34 ms/step vs ~2 ms), tools/kernel_sweep.py SWEEP_B=...
BATCHED_DECODE_MIN = int(os.environ.get("PG_BATCHED_MIN", "4"))

DTYPE_CODE = {torch.float32: PG_F32, torch.bfloat16: PG_BF16, torch.float16: PG_F16}
# attention-mask element kinds understood by pg_decode_inputs
MASK_KIND = {torch.int64: 0, torch.float32: 1, torch.int32: 2, torch.bfloat16: 3, torch.float16: 4, torch.uint8: 5,
             torch.bool: 5, torch.float64: 6}

_vp, _i, _ll, _f, _ull = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong

# name -> argtypes (restype is int unless listed in _RESTYPE)
SIGNATURES = {
    "pg_last_error": [],
    "pg_abi_version": [],
    "pg_launch_count": [],
    "pg_embed_merge": [_vp, _vp, _vp, _vp, _i, _i, _ll, _ll, _ll, _i, _f, _f, _vp, _i, _vp],
    "pg_rmsnorm": [_vp, _vp, _vp, _i, _i, _f, _i, _vp],
    "pg_layernorm": [_vp, _vp, _vp, _vp, _i, _i, _f, _i, _vp],
    "pg_im2col": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "pg_resample_u8": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "pg_u8_to_chw": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "pg_gemm": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "pg_set_workspace": [_vp, _ll],
    "pg_rope_append": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "pg_attention": [_vp, _i, _vp, _i, _vp, _vp, _i, _ll, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i,
                     _f, _i, _i, _vp],
    "pg_attention_tc": [_vp, _i, _vp, _ll, _i, _i, _vp, _vp, _ll, _i, _i, _i, _i, _ll, _vp, _i, _i, _vp, _i, _i, _i,
                        _i, _i, _i, _i, _f, _i, _i, _vp],
    "pg_set_next_prefetch": [_vp, _ll],
    "pg_decode_qkv": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _f,
                      _i, _vp, _vp, _i, _vp],
    "pg_decode_attention_ws_floats": [_i, _i, _i, _i],
    "pg_decode_attention": [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _i,
                            _i, _vp],
    "pg_gemv_res": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    "pg_decode_gateup": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _i, _vp],
    "pg_decode_lmhead": [_vp, _vp, _vp, _vp, _i, _i, _ll, _f, _vp, _vp, _vp, _i, _vp],
    "pg_step_advance": [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp],
    "pg_decode_inputs": [_vp, _vp, _vp, _i, _vp, _i, _ll, _vp, _i, _vp],
    "pg_argmax": [_vp, _vp, _vp, _i, _ll, _vp],
    "pg_top_p_sample": [_vp, _vp, _vp, _i, _ll, _f, _f, _ull, _vp, _vp, _vp],
    "pg_tp_begin_step": [_vp, _vp],
    "pg_tp_push": [_vp, _ll, _vp, _vp],
    "pg_rmsnorm_reduce": [_vp, _vp, _vp, _vp, _i, _i, _f, _vp, _i, _vp],
    "pg_tp_keys_push": [_vp, _i, _ll, _vp, _vp],
    "pg_kv_gather": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
}
_RESTYPE = {"pg_last_error": C.c_char_p, "pg_launch_count": _ull, "pg_decode_attention_ws_floats": _ll}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C .../csrc`).  pg_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = handle
    return _lib


def last_error() -> str:
    return (lib().pg_last_error() or b"").decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"pg_b200 {what} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def exref(ex):
    """`const pg_tp_exchange*` argument: address of a pg_b200.dist.Exchange (None -> NULL = not tensor parallel)."""
    return None if ex is None else C.addressof(ex)


def stream() -> int:
    """Raw handle of torch's current stream on the current device (the private accessor is ~20x cheaper than building
    a torch.cuda.Stream object per kernel launch; fall back to the public API if it ever goes away)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().pg_launch_count())
