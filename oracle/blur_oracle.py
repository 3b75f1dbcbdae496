"""CPU restatement of the reference Gaussian blur (TEST INFRASTRUCTURE ONLY).

Follows ``/root/reference/models/vqgan_fcm.py:20-41`` (five identical copies in
``models/codec.py:255-277, 625-646, 947-968, 1076-1097``): a k x k outer-product
Gaussian built from sigma, reflect padding of k//2, depthwise cross-correlation.
``torchvision.transforms.GaussianBlur`` with a fixed sigma
(``losses/vqgan_losses.py:35``) performs the same arithmetic.
Pinned by ``tests/golden/blur_*.npz``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gaussian_kernel1d(kernel_size: int, sigma, dtype=torch.float32) -> torch.Tensor:
    """vqgan_fcm.py:20-26.  The reference builds the taps in float32 (``torch.linspace`` default);
    with ``dtype=torch.float64`` the whole construction runs in double precision, which is what the
    parity tests use as the true value (a float32 ``linspace`` divided by a 0-dim float64 sigma stays
    float32 under torch's type promotion, so the taps and their sigma-derivative would otherwise carry
    float32 rounding into the "fp64" oracle)."""
    half = (kernel_size - 1) * 0.5
    x = torch.linspace(-half, half, steps=kernel_size, dtype=dtype)
    if torch.is_tensor(sigma):
        x = x.to(sigma.device)
        sigma = sigma.to(dtype)
    pdf = torch.exp(-0.5 * (x / sigma) ** 2)
    return pdf / pdf.sum()


def gaussian_blur_reflect(x: torch.Tensor, sigma, kernel_size: int) -> torch.Tensor:
    k1 = gaussian_kernel1d(kernel_size, sigma, x.dtype).to(x.device)
    k2 = k1[:, None] @ k1[None, :]
    c = x.shape[-3]
    p = kernel_size // 2
    xp = F.pad(x, [p, p, p, p], mode='reflect')
    return F.conv2d(xp, k2.repeat(c, 1, 1, 1), groups=c)
