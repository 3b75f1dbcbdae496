"""Measured error of the blur sigma gradient (fp32 kernels) against the fp64 oracle, with the
conditioning of the sum it is made of: gsigma = sum_j gy_j * u_j, u = (V'H + VH') x.  The taps
k' = dk/dsigma sum to zero, so each u_j is itself a cancelling sum; the running-error scale of the whole
double sum is A = sum_j |gy_j| ((|V'|H + V|H'|)|x|)_j (tests/test_gpu_losses.py::sigma_grad_reference):
an fp32 evaluation cannot be closer than ~eps32 * A.  `cond` = A / |gsigma|.
Run on a GPU box:  python profiles/tools/sigma_grad_error.py > profiles/sigma_grad_error_r2.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from favae_b200 import gaussian_blur_reflect          # noqa: E402
from oracle import blur_oracle as bo                  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
from test_gpu_losses import sigma_grad_reference       # noqa: E402

EPS = 2.0 ** -24
print(f'{"shape":>20} {"k":>3} {"sigma":>5} {"input":>7} {"rel err":>10} {"cond":>9} {"err/(eps32*A)":>17}')
for shape, k in [((1, 4, 256, 256), 9), ((2, 8, 16, 16), 9), ((1, 3, 64, 64), 3), ((1, 2, 40, 70), 15),
                 ((1, 2, 33, 31), 5), ((2, 3, 128, 128), 5), ((1, 1, 512, 512), 11), ((3, 2, 64, 64), 9),
                 ((5, 7, 16, 16), 11), ((2, 3, 12, 12), 9), ((8, 32, 256, 256), 9)]:
    for sigma in (3.0, 0.7):
        for kind in ('random', 'smooth'):
            g = torch.Generator().manual_seed(k)
            x = torch.randn(*shape, generator=g)
            go = torch.randn(*shape, generator=g)
            if kind == 'smooth':        # positively correlated terms: the well-conditioned case
                x = x.abs() + 1.0
                go = go.abs() + 0.5
            if x.numel() > 4e6:
                ref, sum_abs = sigma_grad_reference(x.cuda(), go.cuda(), sigma, k)
            else:
                ref, sum_abs = sigma_grad_reference(x, go, sigma, k)
            xg = x.cuda().requires_grad_(True)
            sg = torch.tensor(sigma, device='cuda', requires_grad=True)
            (gaussian_blur_reflect(xg, sg, k) * go.cuda()).sum().backward()
            err = abs(float(sg.grad) - ref)
            print(f'{str(shape):>20} {k:3d} {sigma:5.1f} {kind:>7} {err / abs(ref):10.2e} {sum_abs / abs(ref):9.1f} '
                  f'{err / (EPS * sum_abs):17.2f}')
