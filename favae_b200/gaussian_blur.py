"""Reflect-padded Gaussian blur with a learnable sigma -- replaces the five identical
``_gaussian_blur`` bodies of the reference (``models/vqgan_fcm.py:20-41``,
``models/codec.py:255-277, 625-646, 947-968, 1076-1097``) and ``T.GaussianBlur`` at
``losses/vqgan_losses.py:35``.  One separable shared-memory kernel forward, one adjoint kernel
and one sigma-gradient kernel backward; the Gaussian taps are built on the device from the
sigma scalar, so there is no CPU ``linspace`` + H2D copy per call as in the reference."""
from __future__ import annotations

import torch

from . import _lib

__all__ = ['gaussian_blur_reflect', 'install_reference_blur']


_SIGMA_CACHE = {}


class _BlurFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, sigma, kernel_size):
        h, w = x.shape[-2:]
        maps = x.numel() // (h * w)
        y = torch.empty_like(x)
        _lib.call('favae_blur_forward', _lib.ptr(x), maps, h, w, kernel_size, _lib.ptr(sigma), _lib.ptr(y),
                  _lib.stream())
        ctx.save_for_backward(x, sigma)
        ctx.kernel_size = kernel_size
        return y

    @staticmethod
    def backward(ctx, gy):
        x, sigma = ctx.saved_tensors
        need_x, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gy = gy.contiguous()
        h, w = x.shape[-2:]
        maps = x.numel() // (h * w)
        # the fused adjoint + sigma-gradient kernel always produces gx
        gx = torch.empty_like(x) if (need_x or need_s) else None
        gs = partials = None
        if need_s:
            gs = torch.empty((1,), device=x.device, dtype=torch.float32)
            partials = torch.empty((max(int(_lib.load().favae_blur_partials(maps, h, w)), 1),),
                                   device=x.device, dtype=torch.float32)
        if need_x or need_s:
            _lib.call('favae_blur_backward', _lib.ptr(gy), _lib.ptr(x), maps, h, w, ctx.kernel_size,
                      _lib.ptr(sigma), _lib.ptr(gx), _lib.ptr(gs), _lib.ptr(partials), _lib.stream())
        return (gx if need_x else None), (gs.reshape(sigma.shape) if gs is not None else None), None


def gaussian_blur_reflect(x, sigma, kernel_size):
    """``x`` (..., H, W) float32 CUDA; ``sigma`` a float or a one-element CUDA tensor (gradient
    flows to it); ``kernel_size`` odd, ``kernel_size // 2 < min(H, W)``."""
    _lib.require_cuda(x)
    if not torch.is_tensor(sigma):
        key = (x.device, float(sigma))
        if key not in _SIGMA_CACHE:                      # fixed-sigma SL path: one H2D copy ever
            _SIGMA_CACHE[key] = torch.tensor(float(sigma), device=x.device, dtype=torch.float32)
        sigma = _SIGMA_CACHE[key]
    if sigma.numel() != 1:
        raise RuntimeError('sigma must hold one value')
    if not sigma.is_cuda:
        sigma = sigma.to(x.device)
    sig = sigma.float()
    if not sig.is_contiguous():
        sig = sig.contiguous()
    return _BlurFunction.apply(x.float().contiguous(), sig, int(kernel_size))


def _blur_method(self, x, i, device=None):
    """Signature of the reference ``_gaussian_blur(self, x, i[, device])``."""
    return gaussian_blur_reflect(x, self.sigmas[i], self.kernel_size)


def install_reference_blur(*classes):
    """Replace ``_gaussian_blur`` on reference classes (VQGANFCM, EncoderGauss, Decoder*Gauss)."""
    for c in classes:
        c._gaussian_blur = _blur_method
