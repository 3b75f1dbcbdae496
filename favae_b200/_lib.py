"""ctypes binding of libfavae_b200.so (include/favae_b200.h).

There is deliberately no fallback: if the library is missing it is built in-tree with nvcc,
and if that is impossible, or a call fails, a RuntimeError is raised.  Nothing here (or
anywhere in favae_b200/) imports ``oracle/`` or computes on the CPU.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

from . import _build

_c = ctypes
_lock = threading.Lock()
_lib = None

_vp, _i64, _i32, _f32, _f64, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float, _c.c_double, _c.c_size_t

SIGNATURES = {
    'favae_abi_version': (_i32, []),
    'favae_last_error': (_c.c_char_p, []),
    'favae_launch_count': (_c.c_longlong, []),
    'favae_vq_prepare_rows': (_i32, [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _vp]),
    'favae_vq_search_exact': (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp]),
    'favae_vq_search_tc_workspace_bytes': (_sz, [_i64, _i64, _i32]),
    'favae_vq_search_tc': (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _sz, _vp, _vp, _vp]),
    'favae_vq_search_tc_overflow_rows': (_i32, [_vp, _i64, _i64, _i32, _vp]),
    'favae_vq_gather_st': (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _vp]),
    'favae_vq_code_stats': (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    'favae_vq_ema_update_cosine': (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    'favae_vq_ema_update_euclid': (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _vp, _vp]),
    'favae_vq_backward': (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _vp, _vp]),
    'favae_vq_gather_rows': (_i32, [_vp, _vp, _i64, _i64, _i32, _i64, _vp, _vp]),
    'favae_ffl_supported': (_i32, [_i32, _i32]),
    'favae_ffl_forward': (_i32, [_vp, _vp, _i64, _i32, _i32, _f32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'favae_sum_scaled': (_i32, [_vp, _i64, _f64, _vp, _vp]),
    'favae_scale_inplace': (_i32, [_vp, _vp, _i64, _vp, _vp]),
    'favae_blur_forward': (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    'favae_blur_partials': (_i64, [_i64, _i32, _i32]),
    'favae_blur_backward': (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
}


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) the C-ABI library.  Raises on any failure."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _build.needs_build():
            _build.build(verbose=False)
        lib = ctypes.CDLL(_build.LIB)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        if lib.favae_abi_version() != 1:
            raise RuntimeError('favae_b200: ABI version mismatch, rebuild the library')
        _lib = lib
        return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {lib.favae_last_error().decode()}')


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('favae_b200 runs on CUDA tensors only (there is no CPU path)')
        if t.dtype != torch.float32 and t.dtype != torch.int64:
            raise RuntimeError(f'favae_b200 expects float32 tensors, got {t.dtype}')


def launch_count() -> int:
    return int(load().favae_launch_count())
