"""Drop-in replacement for ``focal_frequency_loss.FocalFrequencyLoss`` (pip
focal-frequency-loss==0.3.0; imported at /root/reference/favae_scripts/train_favae.py:27 and
instantiated at :313,318,326) running as ONE fused sm_100a kernel per call.

The reference transforms pred and target separately (2 complex FFTs), stacks real/imag,
and runs ~15 elementwise / reduction kernels plus two host syncs (the ``.item()`` asserts).
Here ``favae_ffl_forward`` reads pred and target once, transforms the *difference* with a
half-spectrum real FFT held in shared memory, reduces the dynamic-weight statistics,
applies the weight and the inverse transform in the same kernel and writes the gradients,
so the backward pass is a no-op scale (include/favae_b200.h).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib

__all__ = ['FocalFrequencyLoss', 'expected_upstream_scale']


import contextlib
import threading

_tls = threading.local()


@contextlib.contextmanager
def expected_upstream_scale(scale):
    """Tell the loss that its result will be multiplied by ``scale`` before ``backward`` (the
    feature-loss wrappers divide by the number of levels, losses/vqgan_losses.py:28,48).  The
    forward kernel then writes gradients already multiplied by ``scale`` and the backward pass
    stays a no-op instead of re-scaling both gradient maps in HBM.  Purely an optimisation: any
    other upstream gradient is still handled exactly."""
    prev = getattr(_tls, 'scale', 1.0)
    _tls.scale = float(scale)
    try:
        yield
    finally:
        _tls.scale = prev


def _run_ffl(pred, target, loss_weight, alpha, log_matrix, batch_matrix, mean_count, gscale, need_p, need_t):
    """One favae_ffl_forward call (two for batch_matrix): returns (loss (1,), grad_pred, grad_target).
    ``target`` None: ``pred`` is already the difference map and grad_pred is dL/d(difference)."""
    maps = pred.numel() // (pred.shape[-1] * pred.shape[-2])
    h, w = pred.shape[-2], pred.shape[-1]
    gp = torch.empty_like(pred) if need_p else None
    gt = torch.empty_like(target) if need_t else None
    dev = pred.device
    map_loss = torch.empty((max(maps, 1),), device=dev, dtype=torch.float32)
    st = _lib.stream()
    lib = _lib.load()
    if lib.favae_ffl_supported(h, w):
        def run(gs, ml, a, b, mmax, fmax):
            _lib.call('favae_ffl_forward', _lib.ptr(pred), _lib.ptr(target), maps, h, w, alpha, int(log_matrix),
                      gs, _lib.ptr(ml), _lib.ptr(a), _lib.ptr(b), _lib.ptr(mmax), _lib.ptr(fmax), st)
    else:
        # any other H x W (the reference's fft2 takes any size): the direct-DFT kernels of ffl_generic.cu
        ws_bytes = lib.favae_ffl_generic_workspace_bytes(maps, h, w)
        if ws_bytes == 0 and maps > 0:
            raise NotImplementedError(f'favae_b200: spectrum-loss maps larger than 2048 per side are not supported, '
                                      f'got {h}x{w}')
        ws = torch.empty((max(ws_bytes, 16),), device=dev, dtype=torch.uint8)

        def run(gs, ml, a, b, mmax, fmax):
            _lib.call('favae_ffl_forward_generic', _lib.ptr(pred), _lib.ptr(target), maps, h, w, alpha,
                      int(log_matrix), gs, _lib.ptr(ml), _lib.ptr(a), _lib.ptr(b), _lib.ptr(mmax), _lib.ptr(fmax),
                      _lib.ptr(ws), st)
    if batch_matrix:
        map_max = torch.empty((max(maps, 1),), device=dev, dtype=torch.float32)
        run(0.0, map_loss, None, None, map_max, None)
        gmax = map_max[:maps].amax().reshape(1) if maps else torch.ones(1, device=dev)
        run(gscale, map_loss, gp, gt, None, gmax)
    else:
        run(gscale, map_loss, gp, gt, None, None)
    loss = torch.empty((1,), device=dev, dtype=torch.float32)
    _lib.call('favae_sum_scaled', _lib.ptr(map_loss), maps, loss_weight / mean_count, _lib.ptr(loss), st)
    return loss, gp, gt


class _FFLFunction(torch.autograd.Function):
    """The forward kernel already writes the gradients (scaled by the announced upstream factor), so
    the first backward is a device-side "multiply by go / announced" that exits without touching
    memory when that ratio is 1.  The buffers are handed to autograd once; a second backward through
    the same graph (retain_graph) recomputes them from the saved inputs instead of un-scaling."""

    @staticmethod
    def forward(ctx, pred, target, loss_weight, alpha, log_matrix, batch_matrix, mean_count, grad_mode):
        need_p = grad_mode and ctx.needs_input_grad[0]
        need_t = grad_mode and ctx.needs_input_grad[1]
        expected = getattr(_tls, 'scale', 1.0)
        ctx.cfg = (loss_weight, alpha, log_matrix, batch_matrix, mean_count,
                   2.0 * loss_weight / mean_count * expected, need_p, need_t)
        with _lib.on_device_of(pred, target):
            loss, gp, gt = _run_ffl(pred, target, *ctx.cfg)
        if need_p or need_t:
            ctx.save_for_backward(pred, target)
        ctx.grads = (gp, gt)
        ctx.expected = expected
        return loss.reshape(())

    @staticmethod
    def backward(ctx, go):
        if ctx.cfg[6] is False and ctx.cfg[7] is False:
            return (None,) * 8
        pred, target = ctx.saved_tensors
        with _lib.on_device_of(pred, target):
            if ctx.grads is None:                # re-entry: the first backward consumed the buffers
                _, gp, gt = _run_ffl(pred, target, *ctx.cfg)
            else:
                gp, gt = ctx.grads
            ctx.grads = None
            s = go.detach().to(torch.float32).reshape(1)
            if ctx.expected != 1.0:
                s = s / ctx.expected             # exactly 1.0 when the announced scale was applied
            s = s.contiguous()
            a, b = (gp, gt) if gp is not None else (gt, None)
            _lib.call('favae_scale_inplace', _lib.ptr(a), _lib.ptr(b), a.numel(), _lib.ptr(s), _lib.stream())
        return gp, gt, None, None, None, None, None, None


class FocalFrequencyLoss(nn.Module):
    """Same constructor and ``forward(pred, target, matrix=None)`` as the pip package.

    The reference only ever uses ``(loss_weight=w, alpha=1.0)``; ``patch_factor``,
    ``ave_spectrum``, ``log_matrix`` and ``batch_matrix`` are supported as well.  A predefined
    ``matrix`` is not (FA-VAE never passes one) and raises NotImplementedError.
    """

    def __init__(self, loss_weight=1.0, alpha=1.0, patch_factor=1, ave_spectrum=False,
                 log_matrix=False, batch_matrix=False):
        super().__init__()
        self.loss_weight = loss_weight
        self.alpha = alpha
        self.patch_factor = patch_factor
        self.ave_spectrum = ave_spectrum
        self.log_matrix = log_matrix
        self.batch_matrix = batch_matrix

    def _patches(self, x):
        # tensor2freq's crop-and-stack: (N,C,H,W) -> (N, P, C, H/p, W/p)
        pf = self.patch_factor
        n, c, h, w = x.shape
        assert h % pf == 0 and w % pf == 0, 'Patch factor should be divisible by image height and width'
        if pf == 1:
            return x.unsqueeze(1)
        ph, pw = h // pf, w // pf
        return x.view(n, c, pf, ph, pf, pw).permute(0, 2, 4, 1, 3, 5).reshape(n, pf * pf, c, ph, pw)

    def forward(self, pred, target, matrix=None, **kwargs):
        if matrix is not None:
            raise NotImplementedError('favae_b200: a predefined spectrum weight matrix is not supported')
        from .gaussian_blur import LazyBlur
        if isinstance(pred, LazyBlur):
            pred = pred.materialize()
        if isinstance(target, LazyBlur):
            target = target.materialize()
        _lib.require_cuda(pred, target)
        if pred.shape != target.shape or pred.dim() != 4:
            raise RuntimeError(f'expected two (N,C,H,W) tensors of equal shape, got {tuple(pred.shape)} '
                               f'and {tuple(target.shape)}')
        p, t = self._patches(pred.float()), self._patches(target.float())
        if self.ave_spectrum:
            # the FFT is linear: the mean spectrum is the spectrum of the batch mean
            p, t = p.mean(0, keepdim=True), t.mean(0, keepdim=True)
        p, t = p.contiguous(), t.contiguous()
        if p.data_ptr() % 16:            # float4 access in the kernel
            p = p.clone()
        if t.data_ptr() % 16:
            t = t.clone()
        return _FFLFunction.apply(p, t, float(self.loss_weight), float(self.alpha), bool(self.log_matrix),
                                  bool(self.batch_matrix), float(p.numel()), torch.is_grad_enabled())
