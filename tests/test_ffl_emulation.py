"""Host emulation of the CUDA spectrum-loss phases (favae_b200/csrc/ffl_driver.cuh compiled
by g++, threads replaced by loops) against the oracle.  Covers every kernel configuration,
ragged map counts and the alpha / log_matrix variants without needing a GPU."""
import ctypes
import os
import subprocess

import pytest
import torch

from oracle import ffl_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def emul():
    src = os.path.join(HERE, 'emul', 'ffl_emul.cpp')
    so = os.path.join(HERE, 'emul', 'libffl_emul.so')
    deps = [src] + [os.path.join(HERE, '..', 'favae_b200', 'csrc', f)
                    for f in ('ffl_core.cuh', 'ffl_driver.cuh', 'ffl_configs.cuh')]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', so, src])
    lib = ctypes.CDLL(so)
    lib.ffl_emul.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                             ctypes.c_float, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize('n,maps,alpha,logm', [
    (8, 5, 1.0, 0), (8, 70, 1.0, 0), (16, 3, 1.0, 0), (16, 33, 1.0, 0), (32, 9, 1.0, 0),
    (64, 2, 1.0, 0), (128, 1, 1.0, 0), (256, 1, 1.0, 0), (2564, 1, 1.0, 0), (512, 1, 1.0, 0),
    (16, 4, 2.0, 0), (32, 4, 0.5, 0), (16, 4, 1.0, 1), (64, 1, 1.5, 1)])
def test_emulated_kernel_matches_oracle(emul, n, maps, alpha, logm):
    cfg, n = n, (256 if n == 2564 else n)       # 2564: the 4-CTA-cluster configuration of N = 256
    g = torch.Generator().manual_seed(n + maps)
    p = torch.randn(maps, 1, n, n, generator=g)
    t = torch.randn(maps, 1, n, n, generator=g)
    gp = torch.full_like(p, float('nan')); gt = torch.full_like(p, float('nan'))
    ml = torch.full((maps,), float('nan'))
    lw = 0.5
    gs = 2 * lw / p.numel() / (n * n)
    assert emul.ffl_emul(cfg, p.data_ptr(), t.data_ptr(), maps, alpha, logm, gs,
                         gp.data_ptr(), gt.data_ptr(), ml.data_ptr(), None, None) == 0
    pd = p.double().requires_grad_(True); td = t.double().requires_grad_(True)
    ref = fo.focal_frequency_loss(pd, td, loss_weight=lw, alpha=alpha, log_matrix=bool(logm))
    ref.backward()
    loss = ml.double().sum() * lw / p.numel()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    scale = pd.grad.abs().max()
    assert (gp.double() - pd.grad).abs().max() <= 1e-5 * scale
    assert (gt.double() - td.grad).abs().max() <= 1e-5 * scale


@pytest.mark.parametrize('n,maps', [(16, 33), (64, 3), (256, 3), (512, 1)])
def test_emulated_single_input_form_in_place(emul, n, maps):
    """target == NULL: pred holds the difference map and the gradient G = dL/d(difference) overwrites it
    in place (the fused DSL level); N = 256 also runs the pipelined loads of the next map."""
    g = torch.Generator().manual_seed(n)
    p = torch.randn(maps, 1, n, n, generator=g)
    t = torch.randn(maps, 1, n, n, generator=g)
    d = (p - t).contiguous()
    ml = torch.full((maps,), float('nan'))
    lw = 0.5
    gs = 2 * lw / p.numel() / (n * n)
    assert emul.ffl_emul(n, d.data_ptr(), None, maps, 1.0, 0, gs, d.data_ptr(), None, ml.data_ptr(), None, None) == 0
    pd = p.double().requires_grad_(True)
    ref = fo.focal_frequency_loss(pd, t.double(), loss_weight=lw)
    ref.backward()
    assert abs((ml.double().sum() * lw / p.numel()).item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (d.double() - pd.grad).abs().max() <= 1e-5 * pd.grad.abs().max()


def test_emulated_identical_inputs(emul):
    p = torch.randn(3, 1, 16, 16)
    gp = torch.full_like(p, float('nan')); ml = torch.full((3,), float('nan'))
    assert emul.ffl_emul(16, p.data_ptr(), p.data_ptr(), 3, 1.0, 0, 1.0, gp.data_ptr(), None,
                         ml.data_ptr(), None, None) == 0
    assert torch.all(ml == 0) and torch.all(gp == 0)      # NaN -> 0 rule of the weight matrix


def test_emulated_batch_matrix(emul):
    """batch_matrix=True: one global maximum (two launches: statistics, then weighting)."""
    n, maps, lw = 16, 6, 0.3
    g = torch.Generator().manual_seed(11)
    p = torch.randn(maps, 1, n, n, generator=g) * torch.arange(1, maps + 1).view(-1, 1, 1, 1)
    t = torch.randn(maps, 1, n, n, generator=g)
    ml = torch.empty(maps); mm = torch.empty(maps)
    assert emul.ffl_emul(n, p.data_ptr(), t.data_ptr(), maps, 1.0, 0, 0.0, None, None,
                         ml.data_ptr(), mm.data_ptr(), None) == 0
    gmax = mm.max().reshape(1).contiguous()
    gp = torch.empty_like(p)
    gs = 2 * lw / p.numel() / (n * n)
    assert emul.ffl_emul(n, p.data_ptr(), t.data_ptr(), maps, 1.0, 0, gs, gp.data_ptr(), None,
                         ml.data_ptr(), None, gmax.data_ptr()) == 0
    pd = p.double().requires_grad_(True)
    ref = fo.focal_frequency_loss(pd, t.double(), loss_weight=lw, batch_matrix=True)
    ref.backward()
    loss = ml.double().sum() * lw / p.numel()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (gp.double() - pd.grad).abs().max() <= 1e-5 * pd.grad.abs().max()
