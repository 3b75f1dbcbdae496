"""Reflect-padded Gaussian blur with a learnable sigma -- replaces the five identical
``_gaussian_blur`` bodies of the reference (``models/vqgan_fcm.py:20-41``,
``models/codec.py:255-277, 625-646, 947-968, 1076-1097``) and ``T.GaussianBlur`` at
``losses/vqgan_losses.py:35``.  One separable shared-memory kernel forward, one fused adjoint +
sigma-gradient kernel backward; the Gaussian taps are built on the device from the sigma scalar,
so there is no CPU ``linspace`` + H2D copy per call as in the reference.

``lazy_gaussian_blur`` returns a :class:`LazyBlur` handle instead of the blurred map.  The
reference only ever hands the blurred FCM features to ``recon_ffl_features_loss``
(``favae_scripts/train_favae.py:96``), whose spectrum loss depends on the difference of the two
blurred maps alone; :mod:`favae_b200.vqgan_losses` recognises a pair of handles and runs the fused
``blur -> difference -> spectrum loss`` op (:mod:`favae_b200.spectrum_dsl`) without ever writing
the blurred maps to HBM.  Anything else that touches a handle (any torch function or method)
materialises it with the ordinary differentiable blur, so laziness is never observable."""
from __future__ import annotations

import torch

from . import _lib

__all__ = ['gaussian_blur_reflect', 'lazy_gaussian_blur', 'LazyBlur', 'install_reference_blur', 'sigma_element']


_SIGMA_CACHE = {}


class _BlurFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, sigma, kernel_size):
        h, w = x.shape[-2:]
        maps = x.numel() // (h * w)
        y = torch.empty_like(x)
        with _lib.on_device_of(x, sigma):
            _lib.call('favae_blur_forward', _lib.ptr(x), maps, h, w, kernel_size, _lib.ptr(sigma), _lib.ptr(y),
                      _lib.stream())
        ctx.save_for_backward(x, sigma)
        ctx.kernel_size = kernel_size
        return y

    @staticmethod
    def backward(ctx, gy):
        x, sigma = ctx.saved_tensors
        need_x, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx, gs = blur_backward(gy.contiguous(), x, sigma, ctx.kernel_size, need_x, need_s, 1.0)
        return gx, (gs.reshape(sigma.shape) if gs is not None else None), None


def blur_backward(gy, x, sigma, kernel_size, need_x, need_s, out_scale, scale_dev=None):
    """(s * blur^T(gy), s * d<gy, blur(x)>/dsigma) with s = out_scale * scale_dev[0] through
    favae_blur_backward; the fused adjoint + sigma-gradient kernel always produces gx."""
    if not (need_x or need_s):
        return None, None
    h, w = x.shape[-2:]
    maps = x.numel() // (h * w)
    gx = torch.empty_like(x)
    gs = partials = None
    with _lib.on_device_of(gy, x, sigma):
        if need_s:
            gs = torch.empty((1,), device=x.device, dtype=torch.float32)
            partials = torch.empty((max(int(_lib.load().favae_blur_partials(maps, h, w)), 1),),
                                   device=x.device, dtype=torch.float32)
        _lib.call('favae_blur_backward', _lib.ptr(gy), _lib.ptr(x), maps, h, w, kernel_size, _lib.ptr(sigma),
                  float(out_scale), _lib.ptr(scale_dev), _lib.ptr(gx), _lib.ptr(gs), _lib.ptr(partials),
                  _lib.stream())
    return (gx if need_x else None), gs


def _sigma_tensor(sigma, device):
    if not torch.is_tensor(sigma):
        key = (device, float(sigma))
        if key not in _SIGMA_CACHE:                      # fixed-sigma SL path: one H2D copy ever
            _SIGMA_CACHE[key] = torch.tensor(float(sigma), device=device, dtype=torch.float32)
        sigma = _SIGMA_CACHE[key]
    if sigma.numel() != 1:
        raise RuntimeError('sigma must hold one value')
    if not sigma.is_cuda:
        sigma = sigma.to(device)
    sig = sigma.float()
    if not sig.is_contiguous():
        sig = sig.contiguous()
    return sig


def gaussian_blur_reflect(x, sigma, kernel_size):
    """``x`` (..., H, W) CUDA (cast to float32 like the reference's ``F.conv2d`` under autocast is
    not: see INTEGRATION.md); ``sigma`` a float or a one-element CUDA tensor (gradient flows to it);
    ``kernel_size`` odd, ``kernel_size // 2 < min(H, W)``."""
    if isinstance(x, LazyBlur):
        x = x.materialize()
    _lib.require_cuda(x)
    return _BlurFunction.apply(x.float().contiguous(), _sigma_tensor(sigma, x.device), int(kernel_size))


class LazyBlur(torch.Tensor):
    """Handle for ``gaussian_blur_reflect(x, sigma, kernel_size)`` that has not been computed yet.

    It is a Tensor subclass with the metadata of ``x`` (shape, dtype, device), so it travels through
    lists, module outputs and DDP unharmed.  Metadata queries are answered directly; every other
    torch function or method applied to it first replaces it by the blurred tensor (computed once,
    with autograd to ``x`` and ``sigma``)."""

    _PASS = None

    @staticmethod
    def __new__(cls, x, sigma, kernel_size):
        r = torch.Tensor._make_subclass(cls, x.detach(), False)
        r._lazy_x, r._lazy_sigma, r._lazy_k = x, sigma, int(kernel_size)
        r._lazy_version = x._version
        r._lazy_value = None
        return r

    def source(self):
        """(x, sigma, kernel_size) of the pending blur; raises if x was modified in place since."""
        if self._lazy_x._version != self._lazy_version:
            raise RuntimeError('favae_b200: a feature map was modified in place after its (lazy) Gaussian blur '
                               'was requested; the eager op would have blurred the old values')
        return self._lazy_x, self._lazy_sigma, self._lazy_k

    def materialize(self):
        if self._lazy_value is None:
            x, sigma, k = self.source()
            self._lazy_value = gaussian_blur_reflect(x, sigma, k)
        return self._lazy_value

    @classmethod
    def _passthrough(cls):
        if cls._PASS is None:
            T = torch.Tensor
            cls._PASS = {T.shape.__get__, T.size, T.dim, T.ndim.__get__, T.device.__get__, T.dtype.__get__,
                         T.is_cuda.__get__, T.numel, T.is_floating_point, T.layout.__get__,
                         T.is_contiguous, T.stride, T.element_size, T.is_sparse.__get__, T.is_complex,
                         T.is_quantized.__get__, T.is_meta.__get__, T.nelement, T.__len__}
        return cls._PASS

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in cls._passthrough():
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if func == torch.Tensor.requires_grad.__get__:
            s = args[0]
            return s._lazy_x.requires_grad or (torch.is_tensor(s._lazy_sigma) and s._lazy_sigma.requires_grad)

        def fix(v):
            if isinstance(v, LazyBlur):
                return v.materialize()
            if isinstance(v, (list, tuple)):
                return type(v)(fix(u) for u in v)
            return v
        with torch._C.DisableTorchFunctionSubclass():
            return func(*[fix(a) for a in args], **{k: fix(v) for k, v in kwargs.items()})

    def __repr__(self):
        return (f'LazyBlur(shape={tuple(self.shape)}, kernel_size={self._lazy_k}, '
                f'materialized={self._lazy_value is not None})')


def lazy_gaussian_blur(x, sigma, kernel_size):
    """Deferred ``gaussian_blur_reflect`` (see the module docstring)."""
    if isinstance(x, LazyBlur):
        x = x.materialize()
    _lib.require_cuda(x)
    return LazyBlur(x, sigma, kernel_size)


def sigma_element(sigmas, i):
    """``sigmas[i]`` of a sigma vector (``nn.Parameter`` of one value per FCM level, models/vqgan_fcm.py:67,76,
    models/codec.py:215,575,898,1027) taken through ONE ``unbind`` per (vector, version): autograd then builds
    the vector's gradient with a single ``stack`` instead of a zero-filled vector, a copy and an accumulation
    per level (22 tiny launches per step for the two sigma vectors of the DSL model).  The views are cached
    per vector object and dropped as soon as it is written (optimizer step) or the grad mode changes."""
    if not torch.is_tensor(sigmas) or sigmas.dim() != 1:
        return sigmas[i]
    # (autograd runs a node's backward on the stream of its forward: views made on another stream -- e.g.
    # before a CUDA graph capture -- are not reused)
    stream = torch.cuda.current_stream(sigmas.device).cuda_stream if sigmas.is_cuda else 0
    key = (sigmas._version, torch.is_grad_enabled() and sigmas.requires_grad, sigmas.data_ptr(), stream)
    cached = _UNBOUND.get(id(sigmas))
    if cached is None or cached[0] is not sigmas or cached[1] != key:
        if len(_UNBOUND) > 64:
            _UNBOUND.clear()
        # the stale views go first: they keep the vector's AccumulateGrad node (and its stream) alive, and the
        # new views must not be wired to it
        _UNBOUND.pop(id(sigmas), None)
        cached = None                            # (the local reference to the stale views as well)
        cached = _UNBOUND[id(sigmas)] = (sigmas, key, sigmas.unbind(0))
    return cached[2][i]


_UNBOUND = {}


def _blur_method(self, x, i, device=None):
    """Signature of the reference ``_gaussian_blur(self, x, i[, device])``.  Returns the deferred
    handle: the only consumer in the reference is the DSL loss wrapper.  ``FAVAE_LAZY_BLUR=0`` selects
    the eager op instead (needed under DDP with ``find_unused_parameters=True``, which walks the
    autograd graph of the module outputs and cannot see through a handle)."""
    import os
    if os.environ.get('FAVAE_LAZY_BLUR', '1') in ('', '0'):
        return gaussian_blur_reflect(x, sigma_element(self.sigmas, i), self.kernel_size)
    return lazy_gaussian_blur(x, sigma_element(self.sigmas, i), self.kernel_size)


def install_reference_blur(*classes):
    """Replace ``_gaussian_blur`` on reference classes (VQGANFCM, EncoderGauss, Decoder*Gauss)."""
    for c in classes:
        c._gaussian_blur = _blur_method
