"""ctypes binding of libvqb200.so (the C ABI declared in include/vqb200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, a RuntimeError
is raised.  PyTorch only provides device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VQB200_LIB") or os.path.join(_HERE, "libvqb200.so")   # VQB200_LIB: an experiment build of the same ABI

ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_TCGEN05_BF16, ENGINE_TCGEN05_TF32 = 0, 1, 2, 3, 4
ENGINES = {"auto": ENGINE_AUTO, "simt": ENGINE_SIMT, "tcgen05": ENGINE_TCGEN05, "tcgen05_bf16": ENGINE_TCGEN05_BF16,
           "tcgen05_tf32": ENGINE_TCGEN05_TF32}

_p = C.c_void_p
_i32, _i64, _f32, _sz = C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/vqb200.h declares
SIGNATURES = {
    "vqb200_abi_version": (C.c_int, []),
    "vqb200_error_string": (C.c_char_p, [C.c_int]),
    "vqb200_last_cuda_error": (C.c_int, []),
    "vqb200_launch_count": (C.c_uint64, []),
    "vqb200_codebook_bytes": (_sz, [_i32, _i32]),
    "vqb200_forward_scratch_bytes": (_sz, [_i64, _i32, _i32]),
    "vqb200_stats_bytes": (_sz, [_i32, _i32]),
    "vqb200_codebook_prepare": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "vqb200_quantize_forward": (C.c_int, [_p, _i64, _i32, _i32, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _i32, _p]),
    "vqb200_ema_update": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _f32, _f32, _f32, _p, _p]),
    "vqb200_quantize_backward": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i64, _i64, _i64, _i64, _p]),
    "vqb200_quantize_step": (C.c_int, [_p, _i64, _i32, _i32, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                       _i32, _i32, _f32, _f32, _f32, _p]),
    "vqb200_ema_update_p2p": (C.c_int, [_p, _p, _i32, _i32, C.c_uint32, _p, _p, _p, _i32, _i32, _f32, _f32, _f32, _p, _p]),
    "vqb200_quantize_step_peers": (C.c_int, [_p, _i64, _i32, _i32, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i32,
                                             _f32, _f32, _f32, _p, _p, _p, _p, _i32, _i32, _p]),
    "vqb200_repack_rows": (C.c_int, [_p, _p, _i64, _i32, _i64, _i64, _i64, _i64, _i32, _p]),
    "vqb200_embed_code": (C.c_int, [_p, _i64, _p, _i32, _i32, _p, _p, _p]),
    "vqb200_pack_indices": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p]),
    "vqb200_unpack_indices": (C.c_int, [_p, _i64, _i32, _p, _p]),
    "vqb200_debug_tc_scores": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "vqb200_debug_tc_scores_ex": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _i32, _p]),
    "vqb200_tc_split": (C.c_int, []),
    "vqb200_tc_supported": (C.c_int, [_p, _i64, _i32, _i32, _i64, _i64, _i64, _i64]),
    "vqb200_debug_tc_profile": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _i32, _p]),
    "vqb200_tc_profile_slots": (C.c_int, []),
    "vqb200_debug_pingpong": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "vqb200_debug_tc_kernel": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _p, _i32, _p]),
    "vqb200_host_ctx_create": (C.c_int, [_i64, _i32, _i32, C.POINTER(_p)]),
    "vqb200_host_ctx_set_stream": (C.c_int, [_p, _p]),
    "vqb200_host_ctx_destroy": (None, [_p]),
    "vqb200_host_quantize": (C.c_int, [_p, _p, _i64, _p, _p, _p, _f32, _f32, _f32, _i32, _p, _p, _p, _i32]),
    "vqb200_stats_exchange_peers": (C.c_int, [_p, _i64, _p, _p, _p, _p, _i32, _i32, _p]),
    "vqb200_host_quantize_stats": (C.c_int, [_p, _p, _i64, _p, _p, _p, _p, _p, _i32]),
}

_lib = None


def load():
    """Load libvqb200.so (building it first if nvcc is available and it is missing/stale)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build_native
            build_native.build()
        except Exception as exc:  # pragma: no cover - build box always has nvcc
            raise RuntimeError(
                f"libvqb200.so is missing ({LIB_PATH}) and could not be built: {exc}. "
                "The B200 quantizer has no CPU / PyTorch fallback.") from exc
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:
        raise RuntimeError(f"cannot load {LIB_PATH}: {exc}. The B200 quantizer has no fallback path.") from exc
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.vqb200_abi_version() != 1:
        raise RuntimeError("libvqb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        lib = load()
        msg = lib.vqb200_error_string(rc).decode()
        if rc == -3:
            msg += f" [cudaError {lib.vqb200_last_cuda_error()}]"
        raise RuntimeError(f"{what} failed: {msg}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())
