"""ctypes binding of libfavae_b200.so (include/favae_b200.h).

There is deliberately no fallback: if the library is missing it is built in-tree with nvcc,
and if that is impossible, or a call fails, a RuntimeError is raised.  Nothing here (or
anywhere in favae_b200/) imports ``oracle/`` or computes on the CPU.
"""
from __future__ import annotations

import contextlib
import ctypes
import fcntl
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
    'favae_vq_code_stats': (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    'favae_vq_ema_update_cosine': (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _vp]),
    'favae_vq_ema_update_euclid': (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _vp, _vp]),
    'favae_vq_backward': (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _vp, _vp]),
    'favae_vq_gather_rows': (_i32, [_vp, _vp, _i64, _i64, _i32, _i64, _vp, _vp]),
    'favae_ffl_supported': (_i32, [_i32, _i32]),
    'favae_ffl_forward': (_i32, [_vp, _vp, _i64, _i32, _i32, _f32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'favae_ffl_generic_workspace_bytes': (_sz, [_i64, _i32, _i32]),
    'favae_ffl_forward_generic': (_i32, [_vp, _vp, _i64, _i32, _i32, _f32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'favae_sum_scaled': (_i32, [_vp, _i64, _f64, _vp, _vp]),
    'favae_scale_inplace': (_i32, [_vp, _vp, _i64, _vp, _vp]),
    'favae_blur_forward': (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    'favae_blur_partials': (_i64, [_i64, _i32, _i32]),
    'favae_blur_backward': (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _f32, _vp, _vp, _vp, _vp, _vp]),
    'favae_blur_backward_pair': (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'favae_blur_fast_supported': (_i32, [_i32, _i32, _i32]),
    'favae_blur_diff_forward': (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
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
            # every rank of a DDP / accelerate job gets here at once: serialise the build across
            # PROCESSES with a file lock (the winner builds to a temporary name and renames it,
            # _build.build), and re-check once the lock is held
            with open(os.path.join(_build.HERE, '.build.lock'), 'w') as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                try:
                    if _build.needs_build():
                        _build.build(verbose=False)
                finally:
                    fcntl.flock(lock, fcntl.LOCK_UN)
        lib = ctypes.CDLL(_build.LIB)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        if lib.favae_abi_version() != 2:
            raise RuntimeError('favae_b200: ABI version mismatch, rebuild the library')
        _lib = lib
        return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream(device=None):
    """Raw handle of torch's current stream on ``device`` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


_NULL = contextlib.nullcontext()


def on_device_of(*tensors):
    """Context that makes the tensors' device the current one for the launches inside it, after
    checking that they all live on ONE CUDA device.  (The C ABI launches on the current device and
    ``stream()`` returns that device's current stream; a tensor on cuda:1 while cuda:0 is current
    would otherwise be handed to a kernel on the wrong device.)"""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('favae_b200 runs on CUDA tensors only (there is no CPU path)')
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f'favae_b200: tensors on different devices ({dev} and {t.device})')
    if dev is None or dev.index == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(dev)


_trace = None


def start_trace(names):
    """Measurement hook (bench.py): bracket every call of the named entry points with CUDA events on
    the launching stream, so a kernel family can be timed inside a real step without the benchmark
    re-implementing the call sequence."""
    global _trace
    _trace = {n: [] for n in names}


def stop_trace():
    """-> {name: [(milliseconds, args), ...]} (synchronises)."""
    global _trace
    tr, _trace = _trace, None
    torch.cuda.synchronize()
    return {n: [(a.elapsed_time(b), args) for a, b, args in v] for n, v in (tr or {}).items()}


def call(name: str, *args):
    lib = load()
    tr = _trace
    if tr is not None and name in tr:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(lib, name)(*args)
        b.record()
        tr[name].append((a, b, args))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {lib.favae_last_error().decode()}')


def require_cuda(*tensors):
    """Only the device is checked here: half / bfloat16 inputs (autocast) are cast with ``.float()``
    by the callers, as the reference does (l2_quantize.py:393,266 ``x.float()`` under
    ``autocast(enabled=False)``), so autograd casts the gradient back."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('favae_b200 runs on CUDA tensors only (there is no CPU path)')


def launch_count() -> int:
    return int(load().favae_launch_count())
