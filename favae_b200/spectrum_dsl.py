"""Fused Dynamic-Spectrum-Loss level: ``ffl(blur(dec, sigma_dec), blur(enc, sigma_enc))`` as ONE
autograd op that never materialises the blurred feature maps.

Reference data flow (per FCM level): ``_gaussian_blur`` on the encoder feature
(``models/codec.py:284-309``) and on the decoder feature (``:655-686, 978-999, 1105-1123``;
``models/vqgan_fcm.py:131-134`` for the shared-sigma variant), both blurred maps returned from
``VQGANFCM.forward`` and consumed only by ``recon_ffl_features_loss`` -> ``ffl(de_feat[i],
en_feat[i])`` (``losses/vqgan_losses.py:25``).  The spectrum loss is a function of
``d = pred - target`` alone and its gradient with respect to the target is minus that with respect
to pred, so:

  forward   d  = B_dec(dec) - B_enc(enc)              favae_blur_diff_forward   12 B/element
            G  = dL/dd, loss                          favae_ffl_forward(d, NULL) 8 B/element (in place)
  backward  g_dec, g_sigma_dec = +adjoint/sigma(G)    favae_blur_backward_pair  20 B/element
            g_enc, g_sigma_enc = -adjoint/sigma(G)    (both sides in one pass over G)

40 bytes per feature element instead of 56 for blur, blur, loss (16), adjoint+sigma, adjoint+sigma
on materialised maps, one map-sized scratch buffer instead of four, and 3 launches instead of 5.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from .focal_frequency_loss import _run_ffl, _tls
from .gaussian_blur import LazyBlur, _sigma_tensor, blur_backward

__all__ = ['dsl_level_loss', 'fusable']


def fusable(ffl, pred, target):
    """True when ``ffl(pred, target)`` can run as the fused op: both arguments are pending blurs
    (LazyBlur handles) of equal shape and kernel size on shapes the streaming blur and the spectrum
    kernel take, and the loss uses per-map weights on whole maps."""
    from .focal_frequency_loss import FocalFrequencyLoss
    if not (isinstance(pred, LazyBlur) and isinstance(target, LazyBlur)):
        return False
    if pred._lazy_value is not None or target._lazy_value is not None:
        return False                                   # somebody already paid for the blurred map
    if type(ffl) is not FocalFrequencyLoss or ffl.patch_factor != 1 or ffl.ave_spectrum or ffl.batch_matrix:
        return False
    if pred.shape != target.shape or pred.dim() != 4 or pred._lazy_k != target._lazy_k:
        return False
    if pred.device != target.device or not pred.is_cuda:
        return False
    h, w = pred.shape[-2:]
    lib = _lib.load()
    return bool(lib.favae_ffl_supported(h, w)) and bool(lib.favae_blur_fast_supported(h, w, pred._lazy_k))


class _DSLFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, enc, sigma_enc, dec, sigma_dec, ksize, loss_weight, alpha, log_matrix, grad_mode):
        h, w = enc.shape[-2:]
        maps = enc.numel() // (h * w)
        need = [grad_mode and n for n in ctx.needs_input_grad[:4]]
        want_grad = any(need)
        expected = getattr(_tls, 'scale', 1.0)
        ctx.cfg = (ksize, loss_weight, alpha, log_matrix, float(enc.numel()), expected)
        ctx.need = need
        with _lib.on_device_of(enc, dec, sigma_enc, sigma_dec):
            loss, g = _DSLFunction._difference_and_loss(enc, sigma_enc, dec, sigma_dec, ctx.cfg, want_grad)
        if want_grad:
            ctx.save_for_backward(enc, sigma_enc, dec, sigma_dec)
        ctx.g = g
        return loss.reshape(())

    @staticmethod
    def _difference_and_loss(enc, sigma_enc, dec, sigma_dec, cfg, want_grad):
        ksize, loss_weight, alpha, log_matrix, count, expected = cfg
        h, w = enc.shape[-2:]
        maps = enc.numel() // (h * w)
        d = torch.empty_like(enc)
        _lib.call('favae_blur_diff_forward', _lib.ptr(enc), _lib.ptr(dec), maps, h, w, ksize,
                  _lib.ptr(sigma_enc), _lib.ptr(sigma_dec), _lib.ptr(d), _lib.stream())
        # the spectrum kernel holds a whole map in shared memory before it writes that map's gradient,
        # so G overwrites d in place
        maps_loss = torch.empty((max(maps, 1),), device=enc.device, dtype=torch.float32)
        gscale = 2.0 * loss_weight / count * expected
        _lib.call('favae_ffl_forward', _lib.ptr(d), None, maps, h, w, alpha, int(log_matrix),
                  gscale if want_grad else 0.0, _lib.ptr(maps_loss), _lib.ptr(d) if want_grad else None, None,
                  None, None, _lib.stream())
        loss = torch.empty((1,), device=enc.device, dtype=torch.float32)
        _lib.call('favae_sum_scaled', _lib.ptr(maps_loss), maps, loss_weight / count, _lib.ptr(loss), _lib.stream())
        return loss, (d if want_grad else None)

    @staticmethod
    def backward(ctx, go):
        if not any(ctx.need):
            return (None,) * 9
        enc, sigma_enc, dec, sigma_dec = ctx.saved_tensors
        ksize = ctx.cfg[0]
        with _lib.on_device_of(enc, dec, sigma_enc, sigma_dec):
            g = ctx.g
            s = go.detach().to(torch.float32).reshape(1)
            if ctx.cfg[5] != 1.0:
                s = s / ctx.cfg[5]               # exactly 1.0 when the announced scale was applied
            s = s.contiguous()
            # the upstream factor (exactly 1 when the announced scale was applied) rides on the adjoint
            # kernels' output scale: G itself is never re-scaled in HBM.
            # ffl(pred = blur(dec), target = blur(enc)): +G flows to the decoder side, -G to the encoder side
            if any(ctx.need[:2]) and any(ctx.need[2:]) and os.environ.get('FAVAE_DSL_PAIR', '1') not in ('', '0'):
                # both sides in one pass over G (the normal case: features and sigmas of both sides train)
                h, w = enc.shape[-2:]
                maps = enc.numel() // (h * w)
                g_enc, g_dec = torch.empty_like(enc), torch.empty_like(dec)
                gs = torch.empty((2,), device=enc.device, dtype=torch.float32)
                nparts = max(int(_lib.load().favae_blur_partials(maps, h, w)), 1)
                partials = torch.empty((2 * nparts,), device=enc.device, dtype=torch.float32)
                _lib.call('favae_blur_backward_pair', _lib.ptr(g), _lib.ptr(enc), _lib.ptr(dec), maps, h, w, ksize,
                          _lib.ptr(sigma_enc), _lib.ptr(sigma_dec), _lib.ptr(s), _lib.ptr(g_enc), _lib.ptr(g_dec),
                          gs.data_ptr(), gs.data_ptr() + 4, _lib.ptr(partials), _lib.stream())
                gs_enc, gs_dec = gs[0:1], gs[1:2]
                if not ctx.need[0]:
                    g_enc = None
                if not ctx.need[1]:
                    gs_enc = None
                if not ctx.need[2]:
                    g_dec = None
                if not ctx.need[3]:
                    gs_dec = None
            else:
                g_dec, gs_dec = blur_backward(g, dec, sigma_dec, ksize, ctx.need[2], ctx.need[3], 1.0, s)
                g_enc, gs_enc = blur_backward(g, enc, sigma_enc, ksize, ctx.need[0], ctx.need[1], -1.0, s)
            ctx.g = g                            # G is untouched: a second backward can reuse it
        gs_enc = gs_enc.reshape(sigma_enc.shape) if gs_enc is not None else None
        gs_dec = gs_dec.reshape(sigma_dec.shape) if gs_dec is not None else None
        return g_enc, gs_enc, g_dec, gs_dec, None, None, None, None, None


def dsl_level_loss(ffl, pred, target):
    """``ffl(pred, target)`` for two LazyBlur handles (check ``fusable`` first): pred = the decoder
    side, target = the encoder side, as in ``ffl(de_feat[i], en_feat[i])``."""
    dec, sigma_dec, ksize = pred.source()
    enc, sigma_enc, _ = target.source()
    _lib.require_cuda(enc, dec)
    enc, dec = enc.float().contiguous(), dec.float().contiguous()
    if enc.data_ptr() % 16:
        enc = enc.clone()
    if dec.data_ptr() % 16:
        dec = dec.clone()
    return _DSLFunction.apply(enc, _sigma_tensor(sigma_enc, enc.device), dec, _sigma_tensor(sigma_dec, dec.device),
                              int(ksize), float(ffl.loss_weight), float(ffl.alpha), bool(ffl.log_matrix),
                              torch.is_grad_enabled())
