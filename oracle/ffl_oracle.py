"""CPU restatement of ``focal_frequency_loss.FocalFrequencyLoss`` (TEST INFRASTRUCTURE ONLY).

**PARITY UNPINNED.**  The reference imports this class from the PyPI package
``focal-frequency-loss==0.3.0`` (``/root/reference/environment.yaml:139``;
``favae_scripts/train_favae.py:27``; instantiated ``:313,318,326`` with
``loss_weight=w, alpha=1.0``).  The wheel is not vendored under /root/reference and
cannot be installed offline, so this file restates the package's published
algorithm (upstream ``focal_frequency_loss/focal_frequency_loss.py``):

  tensor2freq : split into ``patch_factor**2`` patches, stack on dim 1,
                ``fft2(norm='ortho')`` -> (N, P, C, H, W, 2) real/imag
  loss        : w = sqrt(dRe^2 + dIm^2) ** alpha            [log(1+.) if log_matrix]
                w = w / max_{H,W} w   (or the global max if batch_matrix)
                NaN -> 0, clamp to [0,1], detached
                loss = mean(w * (dRe^2 + dIm^2)) * loss_weight
  ave_spectrum: both spectra averaged over the batch dim first.

It is anchored on the analytic known-answer tests in ``tests/test_oracle_ffl.py``
and on ``torch.fft`` for the transform itself.
"""
from __future__ import annotations

import torch


def tensor2freq(x: torch.Tensor, patch_factor: int = 1) -> torch.Tensor:
    _, _, h, w = x.shape
    if h % patch_factor or w % patch_factor:
        raise AssertionError('Patch factor should be divisible by image height and width')
    ph, pw = h // patch_factor, w // patch_factor
    patches = [x[:, :, i * ph:(i + 1) * ph, j * pw:(j + 1) * pw]
               for i in range(patch_factor) for j in range(patch_factor)]
    y = torch.stack(patches, 1)
    f = torch.fft.fft2(y, norm='ortho')
    return torch.stack([f.real, f.imag], -1)


def spectrum_weight(diff_freq: torch.Tensor, alpha=1.0, log_matrix=False,
                    batch_matrix=False) -> torch.Tensor:
    sq = diff_freq ** 2
    m = torch.sqrt(sq[..., 0] + sq[..., 1]) ** alpha
    if log_matrix:
        m = torch.log(m + 1.0)
    if batch_matrix:
        m = m / m.max()
    else:
        m = m / m.amax(dim=(-2, -1), keepdim=True)
    m = torch.where(torch.isnan(m), torch.zeros_like(m), m)
    return m.clamp(0.0, 1.0).detach()


def focal_frequency_loss(pred, target, *, loss_weight=1.0, alpha=1.0, patch_factor=1,
                         ave_spectrum=False, log_matrix=False, batch_matrix=False,
                         matrix=None) -> torch.Tensor:
    pf = tensor2freq(pred, patch_factor)
    tf = tensor2freq(target, patch_factor)
    if ave_spectrum:
        pf = pf.mean(0, keepdim=True)
        tf = tf.mean(0, keepdim=True)
    diff = pf - tf
    w = matrix.detach() if matrix is not None else spectrum_weight(
        diff, alpha, log_matrix, batch_matrix)
    if not (w.min().item() >= 0 and w.max().item() <= 1):
        raise AssertionError('The values of spectrum weight matrix should be in the range [0, 1]')
    sq = diff ** 2
    return (w * (sq[..., 0] + sq[..., 1])).mean() * loss_weight


class FocalFrequencyLossOracle(torch.nn.Module):
    """Callable with the package's constructor/forward signature."""

    def __init__(self, loss_weight=1.0, alpha=1.0, patch_factor=1, ave_spectrum=False,
                 log_matrix=False, batch_matrix=False):
        super().__init__()
        self.kw = dict(loss_weight=loss_weight, alpha=alpha, patch_factor=patch_factor,
                       ave_spectrum=ave_spectrum, log_matrix=log_matrix,
                       batch_matrix=batch_matrix)

    def forward(self, pred, target, matrix=None, **kwargs):
        return focal_frequency_loss(pred, target, matrix=matrix, **self.kw)


def closed_form_grad(pred, target, loss_weight=1.0, alpha=1.0):
    """Analytic gradient used by the CUDA path:
    dL/dpred = lw * 2/numel * Re ifft2_ortho(w . D),  dL/dtarget = -dL/dpred."""
    d = (pred - target).double()
    D = torch.fft.fft2(d, norm='ortho')
    A = D.abs()
    w = A ** alpha / A.amax(dim=(-2, -1), keepdim=True)
    w = torch.where(torch.isnan(w), torch.zeros_like(w), w).clamp(0, 1)
    g = torch.fft.ifft2(w * D, norm='ortho').real * (2.0 * loss_weight / d.numel())
    return g
