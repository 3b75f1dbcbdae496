"""Drop-in replacement for ``/root/reference/losses/vqgan_losses.py`` (same three functions,
same arguments, same ``(loss (1,), [per-level scalars])`` returns, same in-place reversal of
the caller's ``de_feat`` list).  The reference evaluates every level twice (:25-26, :45-46);
here each level is evaluated once and the same tensor is returned in the list.

When the feature lists hold pending blurs (``LazyBlur`` handles, which is what the patched
``_gaussian_blur`` methods return), ``recon_ffl_features_loss`` runs every level as the fused
blur -> difference -> spectrum-loss op of :mod:`favae_b200.spectrum_dsl`."""
from __future__ import annotations

import contextlib
import os

import torch

from .focal_frequency_loss import expected_upstream_scale
from .gaussian_blur import gaussian_blur_reflect
from .spectrum_dsl import dsl_level_loss, fusable

__all__ = ['recon_ffl_loss', 'recon_ffl_features_loss', 'recon_sl_gaussian_features_loss']


_SMALL_STREAMS = {}


def _small_level_stream(device):
    key = torch.device(device).index
    if key not in _SMALL_STREAMS:
        _SMALL_STREAMS[key] = torch.cuda.Stream(device)
    return _SMALL_STREAMS[key]


def _small_levels(feats):
    """Indices of the levels worth moving off the main stream: CUDA tensors holding less than 1/8 of the
    largest level's elements (FAVAE_LEVEL_STREAMS=0 disables the side stream)."""
    if os.environ.get('FAVAE_LEVEL_STREAMS', '1') in ('', '0') or len(feats) < 2:
        return []
    try:
        sizes = [t.numel() for t in feats]
        if not all(t.is_cuda for t in feats):
            return []
    except AttributeError:
        return []
    big = max(sizes)
    return [i for i, s_ in enumerate(sizes) if 8 * s_ < big]


def _mean_of_levels(losses, n, device):
    """``(zeros(1) + l0 + l1 + ...) / n`` of the reference (:21-28) as stack, sum, divide: three small
    launches instead of n + 2, same (1,) shape."""
    if not losses:
        return torch.zeros(1, device=device)
    return torch.stack(losses).sum().reshape(1) / n


def recon_ffl_loss(ffl, x, x_recon):                                   # :13-14
    return ffl(x_recon, x)


def recon_ffl_features_loss(ffl, en_feat, de_feat, device):            # :18-30
    de_feat.reverse()
    losses = []
    n = len(en_feat)
    # `loss / len` (:28): the float32 factor autograd will hand back, announced to the loss so that its
    # forward kernel writes final gradients (no host-device traffic here: the step stays graph-capturable)
    small = _small_levels(en_feat)
    main = side = None
    if small:
        # The small levels (the three 16 x 16 levels of the f = 16 model hold 4 % of the feature elements)
        # are short single-wave kernels: they run on a side stream next to the big level instead of in
        # front of it, forward and -- because autograd replays an op on the stream of its forward --
        # backward.  Plain stream fork / join: capturable in a CUDA graph as parallel branches.
        dev = en_feat[small[0]].device
        main = torch.cuda.current_stream(dev)
        side = _small_level_stream(dev)
        side.wait_stream(main)
    levels = [None] * n
    with expected_upstream_scale(float(torch.tensor(1.0) / n)):
        for on_side in (True, False):
            ctx = torch.cuda.stream(side) if (on_side and small) else contextlib.nullcontext()
            with ctx:
                for i in range(n):
                    if (i in small) != on_side:
                        continue
                    if fusable(ffl, de_feat[i], en_feat[i]):
                        levels[i] = dsl_level_loss(ffl, de_feat[i], en_feat[i])
                    else:
                        levels[i] = ffl(de_feat[i], en_feat[i])
    if small:
        main.wait_stream(side)
        for i in small:
            levels[i].record_stream(main)
    losses = levels
    return _mean_of_levels(losses, n, device), losses


def recon_sl_gaussian_features_loss(ffl, gaussian_kernel, gaussian_sigma, en_feat, de_feat, device):  # :34-50
    de_feat.reverse()
    losses = []
    n = len(en_feat)
    with expected_upstream_scale(float(torch.tensor(1.0) / n)):
        for i in range(len(en_feat)):
            e = gaussian_blur_reflect(en_feat[i], float(gaussian_sigma), gaussian_kernel)
            d = gaussian_blur_reflect(de_feat[i], float(gaussian_sigma), gaussian_kernel)
            losses.append(ffl(d, e))
    return _mean_of_levels(losses, n, device), losses
