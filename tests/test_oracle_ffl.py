"""Known-answer tests for the FFL restatement (parity unpinned: the focal-frequency-loss
wheel is absent) plus blur and wrapper fixtures."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import blur_oracle as bo
from oracle import ffl_oracle as fo
from oracle import wrappers_oracle as wo

LW = 0.37


def _pair(shape=(2, 3, 16, 16), seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)


def test_kat0_identical_inputs_give_zero():
    p, _ = _pair()
    assert float(fo.focal_frequency_loss(p, p.clone(), loss_weight=LW)) == 0.0


def test_kat1_constant_offset():
    p, _ = _pair()
    c = 0.75
    v = float(fo.focal_frequency_loss(p + c, p, loss_weight=LW))
    assert v == pytest.approx(LW * c * c, rel=1e-5)


def test_kat2_single_cosine():
    H = W = 32
    y, x = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    A, u, v = 1.3, 3, 5
    d = A * torch.cos(2 * math.pi * (u * x / W + v * y / H))
    t = torch.zeros(1, 1, H, W)
    val = float(fo.focal_frequency_loss(t + d, t, loss_weight=LW))
    assert val == pytest.approx(LW * A * A / 2, rel=1e-4)


def test_kat3_parseval_upper_bound():
    p, t = _pair()
    assert float(fo.focal_frequency_loss(p, t, loss_weight=LW)) <= LW * float(((p - t) ** 2).mean()) * (1 + 1e-6)


def test_kat4_single_pixel():
    t = torch.zeros(1, 1, 16, 16); p = t.clone(); p[0, 0, 3, 7] = 2.0
    assert float(fo.focal_frequency_loss(p, t, loss_weight=LW)) == pytest.approx(LW * 4.0 / 256, rel=1e-5)


def test_kat5_quadratic_scaling():
    p, t = _pair()
    a = float(fo.focal_frequency_loss(t + 3.0 * (p - t), t, loss_weight=LW))
    b = float(fo.focal_frequency_loss(p, t, loss_weight=LW))
    assert a == pytest.approx(9.0 * b, rel=1e-5)


def test_kat6_gradient_identity():
    p, t = _pair(seed=3)
    p = p.double().requires_grad_(True); t = t.double().requires_grad_(True)
    fo.focal_frequency_loss(p, t, loss_weight=LW).backward()
    g = fo.closed_form_grad(p.detach(), t.detach(), loss_weight=LW)
    torch.testing.assert_close(p.grad, g, rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(t.grad, -g, rtol=1e-9, atol=1e-12)


def test_kat7_shift_invariance():
    p, t = _pair(); s = torch.randn_like(p)
    a = float(fo.focal_frequency_loss(p + s, t + s, loss_weight=LW))
    b = float(fo.focal_frequency_loss(p, t, loss_weight=LW))
    assert a == pytest.approx(b, rel=1e-4)


def test_option_flags_run():
    p, t = _pair(shape=(2, 2, 8, 8))
    for kw in (dict(patch_factor=2), dict(ave_spectrum=True), dict(log_matrix=True),
               dict(batch_matrix=True), dict(alpha=2.0)):
        v = fo.focal_frequency_loss(p, t, **kw)
        assert torch.isfinite(v) and float(v) > 0


def test_blur_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'blur_cases.npz'))
    for i in range(int(g['n'])):
        x = torch.from_numpy(g[f'x{i}']).requires_grad_(True)
        sig = torch.tensor(float(g[f'sigma{i}']), requires_grad=True)
        y = bo.gaussian_blur_reflect(x, sig, int(g[f'k{i}']))
        torch.testing.assert_close(y, torch.from_numpy(g[f'y{i}']), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(y, torch.from_numpy(g[f'tv{i}']), rtol=1e-5, atol=1e-6)
        (y * torch.from_numpy(g[f'go{i}'])).sum().backward()
        torch.testing.assert_close(x.grad, torch.from_numpy(g[f'gx{i}']), rtol=1e-5, atol=1e-6)
        assert float(sig.grad) == pytest.approx(float(g[f'gsig{i}']), rel=1e-3, abs=1e-5)


def test_wrappers_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'wrappers.npz'))
    ffl = fo.FocalFrequencyLossOracle(loss_weight=0.01, alpha=1.0)
    en = [torch.from_numpy(g[f'en{i}']) for i in range(4)]
    de = [torch.from_numpy(g[f'de{i}']) for i in range(4)]
    d1 = list(de)
    loss, lst = wo.recon_ffl_features_loss(ffl, list(en), d1)
    assert d1[0] is de[-1] and int(g['dsl_reversed_inplace']) == 1
    assert loss.shape == (1,)
    np.testing.assert_allclose(loss.numpy(), g['dsl_loss'], rtol=1e-6)
    np.testing.assert_allclose([float(v) for v in lst], g['dsl_list'], rtol=1e-6)
    loss, lst = wo.recon_sl_gaussian_features_loss(ffl, 5, 3, list(en), list(de))
    np.testing.assert_allclose(loss.numpy(), g['sl_loss'], rtol=1e-5)
    np.testing.assert_allclose([float(v) for v in lst], g['sl_list'], rtol=1e-5)
    v = wo.recon_ffl_loss(ffl, torch.from_numpy(g['img_x']), torch.from_numpy(g['img_xr']))
    np.testing.assert_allclose(float(v), float(g['img_loss']), rtol=1e-6)
