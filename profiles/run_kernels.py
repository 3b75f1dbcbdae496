"""Micro-driver for profiling: runs each hot kernel a few times on BASELINE-shaped inputs and prints
CUDA-event timings with the algorithmic GB/s / TFLOP/s.  Used under ncu:

    ncu --set full --import-source on -k regex:<name> -s <skip> -c 1 -o gpurun_out/x python profiles/run_kernels.py --only blur_bwd
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import favae_b200  # noqa: E402
from favae_b200 import _lib  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='all')
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--n-lat', type=int, default=0)
    args = ap.parse_args()
    B, it = args.batch, args.iters
    dev = 'cuda'
    st = _lib.stream

    def want(name):
        return args.only in ('all', name)

    shape = (B, 128, 256, 256)
    E = B * 128 * 256 * 256
    if want('ffl') or want('ffl_diff') or want('blur_fwd') or want('blur_bwd') or want('blur_diff') or want('blur_pair'):
        p = torch.randn(shape, device=dev); t = torch.randn(shape, device=dev)
        gp = torch.empty_like(p); gt = torch.empty_like(p)
        sig = torch.tensor(3.0, device=dev)
    if want('ffl_diff'):
        ml = torch.empty(B * 128, device=dev)
        ms = timed(lambda: _lib.call('favae_ffl_forward', p.data_ptr(), None, B * 128, 256, 256, 1.0, 0,
                                     1e-3, ml.data_ptr(), gp.data_ptr(), None, None, None, st()), it)
        print(f'ffl_256 d -> G          {ms:8.3f} ms  {8 * E / ms / 1e6:8.1f} GB/s algorithmic (8 B/elem)')
    if want('ffl'):
        ml = torch.empty(B * 128, device=dev)
        ms = timed(lambda: _lib.call('favae_ffl_forward', p.data_ptr(), t.data_ptr(), B * 128, 256, 256, 1.0, 0,
                                     1e-3, ml.data_ptr(), gp.data_ptr(), gt.data_ptr(), None, None, st()), it)
        print(f'ffl_256 fwd+grad   {ms:8.3f} ms  {16 * E / ms / 1e6:8.1f} GB/s algorithmic (16 B/elem)')
        ms = timed(lambda: _lib.call('favae_ffl_forward', p.data_ptr(), t.data_ptr(), B * 128, 256, 256, 1.0, 0,
                                     1e-3, ml.data_ptr(), None, None, None, None, st()), it)
        print(f'ffl_256 loss only  {ms:8.3f} ms  {8 * E / ms / 1e6:8.1f} GB/s algorithmic (8 B/elem)')
        # fused DSL level: single-input form (pred holds the difference map), gradient written in place
        d = p.clone()
        ms = timed(lambda: _lib.call('favae_ffl_forward', d.data_ptr(), None, B * 128, 256, 256, 1.0, 0,
                                     1e-3, ml.data_ptr(), d.data_ptr(), None, None, None, st()), it)
        print(f'ffl_256 d -> G in place {ms:8.3f} ms  {8 * E / ms / 1e6:8.1f} GB/s algorithmic (8 B/elem)')
        ms = timed(lambda: _lib.call('favae_ffl_forward', p.data_ptr(), None, B * 128, 256, 256, 1.0, 0,
                                     1e-3, ml.data_ptr(), gp.data_ptr(), None, None, None, st()), it)
        print(f'ffl_256 d -> G          {ms:8.3f} ms  {8 * E / ms / 1e6:8.1f} GB/s algorithmic (8 B/elem)')
        del d
        p16 = torch.randn(B * 512, 1, 16, 16, device=dev); t16 = torch.randn_like(p16)
        g16 = torch.empty_like(p16); h16 = torch.empty_like(p16); ml16 = torch.empty(B * 512, device=dev)
        ms = timed(lambda: _lib.call('favae_ffl_forward', p16.data_ptr(), t16.data_ptr(), B * 512, 16, 16, 1.0, 0,
                                     1e-3, ml16.data_ptr(), g16.data_ptr(), h16.data_ptr(), None, None, st()), it)
        print(f'ffl_16 fwd+grad    {ms:8.3f} ms  {16 * p16.numel() / ms / 1e6:8.1f} GB/s algorithmic')
    if want('blur_fwd'):
        ms = timed(lambda: _lib.call('favae_blur_forward', p.data_ptr(), B * 128, 256, 256, 9, sig.data_ptr(),
                                     gp.data_ptr(), st()), it)
        print(f'blur fwd k9        {ms:8.3f} ms  {8 * E / ms / 1e6:8.1f} GB/s (8 B/elem)')
    if want('blur_diff'):
        sig2 = torch.tensor(2.5, device=dev)
        ms = timed(lambda: _lib.call('favae_blur_diff_forward', p.data_ptr(), t.data_ptr(), B * 128, 256, 256, 9,
                                     sig.data_ptr(), sig2.data_ptr(), gp.data_ptr(), st()), it)
        print(f'blur diff k9       {ms:8.3f} ms  {12 * E / ms / 1e6:8.1f} GB/s (12 B/elem)')
    if want('blur_pair'):
        gs2 = torch.empty(2, device=dev); sig2 = torch.tensor(2.5, device=dev)
        n_p = int(_lib.load().favae_blur_partials(B * 128, 256, 256))
        parts2 = torch.empty(2 * n_p, device=dev)
        g = torch.randn(shape, device=dev)
        ms = timed(lambda: _lib.call('favae_blur_backward_pair', g.data_ptr(), p.data_ptr(), t.data_ptr(), B * 128, 256, 256,
                                     9, sig.data_ptr(), sig2.data_ptr(), None, gp.data_ptr(), gt.data_ptr(),
                                     gs2.data_ptr(), gs2.data_ptr() + 4, parts2.data_ptr(), st()), it)
        print(f'blur bwd pair k9   {ms:8.3f} ms  {20 * E / ms / 1e6:8.1f} GB/s (20 B/elem)')
        del g
    if want('blur_bwd'):
        gs = torch.empty(1, device=dev)
        parts = torch.empty(int(_lib.load().favae_blur_partials(B * 128, 256, 256)), device=dev)
        ms = timed(lambda: _lib.call('favae_blur_backward', t.data_ptr(), p.data_ptr(), B * 128, 256, 256, 9,
                                     sig.data_ptr(), 1.0, None, gp.data_ptr(), gs.data_ptr(), parts.data_ptr(), st()), it)
        print(f'blur bwd+sigma k9  {ms:8.3f} ms  {12 * E / ms / 1e6:8.1f} GB/s (12 B/elem)')
        ms = timed(lambda: _lib.call('favae_blur_backward', t.data_ptr(), p.data_ptr(), B * 128, 256, 256, 9,
                                     sig.data_ptr(), 1.0, None, gp.data_ptr(), None, None, st()), it)
        print(f'blur bwd k9        {ms:8.3f} ms  {8 * E / ms / 1e6:8.1f} GB/s (8 B/elem)')
    if want('small'):
        # the three 16 x 16 feature levels of the f=16 model: B*512, B*512, B*256 maps
        for ch in (512, 256):
            maps = B * ch
            xs = torch.randn(maps, 16, 16, device=dev); gs_ = torch.randn(maps, 16, 16, device=dev)
            ys = torch.empty_like(xs); sg = torch.tensor(3.0, device=dev); g1 = torch.empty(1, device=dev)
            parts = torch.empty(int(_lib.load().favae_blur_partials(maps, 16, 16)), device=dev)
            ml = torch.empty(maps, device=dev); gp2 = torch.empty_like(xs); gt2 = torch.empty_like(xs)
            ms = timed(lambda: _lib.call('favae_blur_forward', xs.data_ptr(), maps, 16, 16, 9, sg.data_ptr(),
                                         ys.data_ptr(), st()), it)
            print(f'blur16 fwd  maps={maps:6d} {ms * 1e3:8.1f} us')
            ms = timed(lambda: _lib.call('favae_blur_backward', gs_.data_ptr(), xs.data_ptr(), maps, 16, 16, 9,
                                         sg.data_ptr(), 1.0, None, ys.data_ptr(), g1.data_ptr(), parts.data_ptr(), st()), it)
            print(f'blur16 bwd+sigma maps={maps:6d} {ms * 1e3:8.1f} us')
            ms = timed(lambda: _lib.call('favae_ffl_forward', xs.data_ptr(), gs_.data_ptr(), maps, 16, 16, 1.0, 0,
                                         1e-3, ml.data_ptr(), gp2.data_ptr(), gt2.data_ptr(), None, None, st()), it)
            print(f'ffl16 fwd+grad maps={maps:6d} {ms * 1e3:8.1f} us')
    if want('f4'):
        # the f = 4 model (BASELINE configs[3]): 64 x 64 feature maps, gaussian_kernel 3; level 0 is
        # (B, 512, 64, 64) per side
        maps = B * 512
        E4 = maps * 64 * 64
        xs = torch.randn(maps, 64, 64, device=dev); gs_ = torch.randn(maps, 64, 64, device=dev)
        ys = torch.empty_like(xs); sg = torch.tensor(3.0, device=dev); g1 = torch.empty(1, device=dev)
        parts = torch.empty(int(_lib.load().favae_blur_partials(maps, 64, 64)), device=dev)
        ml = torch.empty(maps, device=dev); gp2 = torch.empty_like(xs); gt2 = torch.empty_like(xs)
        for ks in (3, 9):
            ms = timed(lambda: _lib.call('favae_blur_forward', xs.data_ptr(), maps, 64, 64, ks, sg.data_ptr(),
                                         ys.data_ptr(), st()), it)
            print(f'blur64 fwd k{ks}        {ms:8.3f} ms  {8 * E4 / ms / 1e6:8.1f} GB/s (8 B/elem)')
            ms = timed(lambda: _lib.call('favae_blur_backward', gs_.data_ptr(), xs.data_ptr(), maps, 64, 64, ks,
                                         sg.data_ptr(), 1.0, None, ys.data_ptr(), g1.data_ptr(), parts.data_ptr(), st()), it)
            print(f'blur64 bwd+sigma k{ks}  {ms:8.3f} ms  {12 * E4 / ms / 1e6:8.1f} GB/s (12 B/elem)')
        ms = timed(lambda: _lib.call('favae_ffl_forward', xs.data_ptr(), gs_.data_ptr(), maps, 64, 64, 1.0, 0,
                                     1e-3, ml.data_ptr(), gp2.data_ptr(), gt2.data_ptr(), None, None, st()), it)
        print(f'ffl_64 fwd+grad       {ms:8.3f} ms  {16 * E4 / ms / 1e6:8.1f} GB/s algorithmic (16 B/elem)')
        for n, side in ((B * 256, 128), (B * 1024, 32), (B * 32, 512)):
            a = torch.randn(n, side, side, device=dev); b_ = torch.randn(n, side, side, device=dev)
            ga = torch.empty_like(a); gb = torch.empty_like(a); mls = torch.empty(n, device=dev)
            ms = timed(lambda: _lib.call('favae_ffl_forward', a.data_ptr(), b_.data_ptr(), n, side, side, 1.0, 0,
                                         1e-3, mls.data_ptr(), ga.data_ptr(), gb.data_ptr(), None, None, st()), it)
            print(f'ffl_{side} fwd+grad      {ms:8.3f} ms  {16 * a.numel() / ms / 1e6:8.1f} GB/s algorithmic (16 B/elem)')
    if want('vq'):
        K, D = 16384, 256
        for n in ([args.n_lat] if args.n_lat else [B * 256, 8192, 65536, 262144]):
            x = torch.randn(n, D, device=dev)
            e = torch.nn.functional.normalize(torch.randn(K, D, device=dev), dim=-1)
            xn = torch.empty(n, D, device=dev); xh = torch.empty(n, D, device=dev, dtype=torch.float16)
            en = torch.empty(K, D, device=dev); eh = torch.empty(K, D, device=dev, dtype=torch.float16)
            _lib.call('favae_vq_prepare_rows', x.data_ptr(), n, D, 1, 1, xn.data_ptr(), xh.data_ptr(), None, st())
            _lib.call('favae_vq_prepare_rows', e.data_ptr(), K, D, 1, 1, en.data_ptr(), eh.data_ptr(), None, st())
            nbytes = _lib.load().favae_vq_search_tc_workspace_bytes(n, K, D)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            idx = torch.empty(n, device=dev, dtype=torch.int64); keys = torch.empty(n, device=dev, dtype=torch.int64)
            ms = timed(lambda: _lib.call('favae_vq_search_tc', xh.data_ptr(), eh.data_ptr(), xn.data_ptr(),
                                         en.data_ptr(), n, K, D, ws.data_ptr(), nbytes, keys.data_ptr(),
                                         idx.data_ptr(), st()), it)
            import ctypes
            cnt = ctypes.c_int(-1)
            _lib.call('favae_vq_search_tc_overflow_rows', ws.data_ptr(), n, K, D, ctypes.addressof(cnt))
            print(f'vq_search_tc n={n:7d}  {ms:8.3f} ms  {2.0 * n * K * D / ms / 1e9:8.1f} TFLOP/s algorithmic '
                  f'(search + rescore + fallback; {cnt.value} latents took the exhaustive fallback)')


if __name__ == '__main__':
    main()
