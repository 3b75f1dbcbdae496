"""CPU restatement of ``/root/reference/losses/vqgan_losses.py`` (TEST INFRASTRUCTURE ONLY).

* ``recon_ffl_loss``                  :13-14
* ``recon_ffl_features_loss``         :18-30   (reverses the caller's ``de_feat`` in place,
                                               evaluates every level twice, divides by len)
* ``recon_sl_gaussian_features_loss`` :34-50   (fixed-sigma blur of all 8 maps first)
"""
from __future__ import annotations

import torch

from .blur_oracle import gaussian_blur_reflect


def recon_ffl_loss(ffl, x, x_recon):
    return ffl(x_recon, x)


def recon_ffl_features_loss(ffl, en_feat, de_feat, device='cpu'):
    de_feat.reverse()
    loss = torch.zeros(1, device=device)
    losses = []
    for e, d in zip(en_feat, de_feat):
        loss = loss + ffl(d, e)
        losses.append(ffl(d, e))
    return loss / len(en_feat), losses


def recon_sl_gaussian_features_loss(ffl, gaussian_kernel, gaussian_sigma, en_feat, de_feat,
                                    device='cpu'):
    de_feat.reverse()
    en_b = [gaussian_blur_reflect(f, float(gaussian_sigma), gaussian_kernel) for f in en_feat]
    de_b = [gaussian_blur_reflect(f, float(gaussian_sigma), gaussian_kernel) for f in de_feat]
    loss = torch.zeros(1, device=device)
    losses = []
    for e, d in zip(en_b, de_b):
        loss = loss + ffl(d, e)
        losses.append(ffl(d, e))
    return loss / len(en_feat), losses
