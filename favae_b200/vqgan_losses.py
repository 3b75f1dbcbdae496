"""Drop-in replacement for ``/root/reference/losses/vqgan_losses.py`` (same three functions,
same arguments, same ``(loss (1,), [per-level scalars])`` returns, same in-place reversal of
the caller's ``de_feat`` list).  The reference evaluates every level twice (:25-26, :45-46);
here each level is evaluated once and the same tensor is returned in the list.

When the feature lists hold pending blurs (``LazyBlur`` handles, which is what the patched
``_gaussian_blur`` methods return), ``recon_ffl_features_loss`` runs every level as the fused
blur -> difference -> spectrum-loss op of :mod:`favae_b200.spectrum_dsl`."""
from __future__ import annotations

import torch

from .focal_frequency_loss import expected_upstream_scale
from .gaussian_blur import gaussian_blur_reflect
from .spectrum_dsl import dsl_level_loss, fusable

__all__ = ['recon_ffl_loss', 'recon_ffl_features_loss', 'recon_sl_gaussian_features_loss']


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
    with expected_upstream_scale(float(torch.tensor(1.0) / n)):
        for i in range(len(en_feat)):
            if fusable(ffl, de_feat[i], en_feat[i]):
                level = dsl_level_loss(ffl, de_feat[i], en_feat[i])
            else:
                level = ffl(de_feat[i], en_feat[i])
            losses.append(level)
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
