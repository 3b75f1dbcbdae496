"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference):

    python -m oracle.make_golden

Imports ``models/l2_quantize.py``, ``models/vqgan_fcm.py`` and
``losses/vqgan_losses.py`` from /root/reference and records inputs, outputs, updated
buffers and autograd gradients.  The FFL callable handed to the loss wrappers is the
restatement in ``oracle/ffl_oracle.py`` (the ``focal-frequency-loss`` wheel is absent:
that half stays "parity unpinned").  The fixtures travel with the repo; /root/reference
does not exist on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _import_reference():
    warnings.filterwarnings('ignore')
    sys.path.insert(0, REF)
    from models import l2_quantize            # noqa: E402
    from losses import vqgan_losses            # noqa: E402
    from models.vqgan_fcm import VQGANFCM      # noqa: E402
    sys.path.pop(0)
    return l2_quantize, vqgan_losses, VQGANFCM


def _np(t):
    return t.detach().cpu().numpy().copy()      # copy: buffers are updated in place later


def vq_case(l2q, name, *, K, D, B, h, w, cosine, steps, commit=1.0, dim=None, seed=0, heads=1):
    torch.manual_seed(seed)
    dim = D * heads if dim is None else dim
    vq = l2q.VectorQuantize(dim=dim, codebook_size=K, codebook_dim=D, accept_image_fmap=True,
                            use_cosine_sim=cosine, commitment_weight=commit, heads=heads)
    vq.train()
    rec = {'K': K, 'D': D, 'dim': dim, 'cosine': int(cosine), 'commit': commit, 'steps': steps, 'heads': heads,
           'embed0': _np(vq._codebook.embed[0]), 'cluster0': _np(vq._codebook.cluster_size[0])}
    if dim != D * heads:
        rec.update(pin_w=_np(vq.project_in.weight), pin_b=_np(vq.project_in.bias),
                   pout_w=_np(vq.project_out.weight), pout_b=_np(vq.project_out.bias))
    if not cosine:
        rec['embed_avg0'] = _np(vq._codebook.embed_avg[0])
    g = torch.Generator().manual_seed(1234 + seed)
    for s in range(steps):
        x = torch.randn(B, dim, h, w, generator=g)
        if s == 0 and cosine and dim == D and heads == 1:
            # plant exact ties: two latents sit exactly on duplicated codes
            vq._codebook.embed[0, 5] = vq._codebook.embed[0, 3]
            rec['embed0'] = _np(vq._codebook.embed[0])
            x[0, :, 0, 0] = vq._codebook.embed[0, 3] * 2.5
        x.requires_grad_(True)
        gq = torch.randn(B, dim, h, w, generator=g)
        q, ind, loss = vq(x)
        (q * gq).sum().add(loss.sum() * 0.7).backward()
        rec[f'x{s}'] = _np(x); rec[f'gq{s}'] = _np(gq)
        rec[f'q{s}'] = _np(q); rec[f'ind{s}'] = _np(ind); rec[f'loss{s}'] = _np(loss)
        rec[f'gx{s}'] = _np(x.grad)
        rec[f'embed{s + 1}'] = _np(vq._codebook.embed[0])
        rec[f'cluster{s + 1}'] = _np(vq._codebook.cluster_size[0])
    vq.eval()
    x = torch.randn(B, dim, h, w, generator=g)
    q, ind, loss = vq(x)
    rec.update(x_eval=_np(x), q_eval=_np(q), ind_eval=_np(ind), loss_eval=_np(loss))
    if dim == D and heads == 1:
        ids = torch.randint(0, K, (B, h * w), generator=g)
        rec.update(entry_ids=_np(ids), entry=_np(vq.get_codebook_entry(ids, (B, h, w, D))))
    np.savez_compressed(os.path.join(OUT, f'vq_{name}.npz'), **rec)
    print('wrote', name)


def vq_big_case(l2q, name, *, K, D=256, B=1, h=8, w=8, seed=11, ortho=0.0):
    """Production codebook shapes (K x 256) at small N.  The codebook is not stored (16 MB at
    K = 16384): it is the module's own seeded initialisation, which the test regenerates with the same
    torch CPU generator and checks against the recorded digest; of the updated codebook only the rows
    of the codes that were hit (plus a few that were not) are kept."""
    torch.manual_seed(seed)
    vq = l2q.VectorQuantize(dim=D, codebook_size=K, accept_image_fmap=True, use_cosine_sim=True,
                            commitment_weight=1.0, orthogonal_reg_weight=ortho,
                            orthogonal_reg_max_codes=128 if ortho else None)
    vq.train()
    e0 = vq._codebook.embed.detach()[0].clone()
    rec = {'K': K, 'D': D, 'seed': seed, 'ortho': ortho,
           'embed0_sum': np.float64(e0.double().sum().item()), 'embed0_abs': np.float64(e0.double().abs().sum().item()),
           'embed0_row7': _np(e0[7])}
    g = torch.Generator().manual_seed(4321 + seed)
    x = torch.randn(B, D, h, w, generator=g).requires_grad_(True)
    gq = torch.randn(B, D, h, w, generator=g)
    q, ind, loss = vq(x)
    (q * gq).sum().add(loss.sum() * 0.7).backward()
    e1 = vq._codebook.embed.detach()[0]
    hit = torch.unique(ind)
    rows = torch.cat([hit, torch.tensor([0, 1, K // 2, K - 1])]).unique()
    rec.update(x=_np(x), gq=_np(gq), q=_np(q), ind=_np(ind), loss=_np(loss), gx=_np(x.grad),
               rows=_np(rows), embed1_rows=_np(e1[rows]), cluster1_rows=_np(vq._codebook.cluster_size[0][rows]),
               embed1_sum=np.float64(e1.double().sum().item()),
               cluster1_sum=np.float64(vq._codebook.cluster_size.double().sum().item()))
    if ortho:
        ge = vq._codebook.embed.grad[0]
        rec.update(gembed_rows=_np(ge[rows]), gembed_abs=np.float64(ge.double().abs().sum().item()))
    np.savez_compressed(os.path.join(OUT, f'vq_{name}.npz'), **rec)
    print('wrote', name, 'hit codes', hit.numel())


def _ddp_worker(rank, world, port, K, D, B, h, w, q):
    import torch.distributed as dist
    warnings.filterwarnings('ignore')
    sys.path.insert(0, REF)
    from models import l2_quantize as l2q
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(7)
    vq = l2q.VectorQuantize(dim=D, codebook_size=K, accept_image_fmap=True, use_cosine_sim=True,
                            sync_codebook=True).train()
    embed0 = _np(vq._codebook.embed[0])
    g = torch.Generator().manual_seed(1234 + rank)
    out = {'embed0': embed0}
    for s in range(2):
        x = torch.randn(B, D, h, w, generator=g)
        qz, ind, loss = vq(x)
        out[f'x{s}'] = _np(x); out[f'ind{s}'] = _np(ind); out[f'loss{s}'] = _np(loss)
        out[f'embed{s + 1}'] = _np(vq._codebook.embed[0])
        out[f'cluster{s + 1}'] = _np(vq._codebook.cluster_size[0])
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def vq_ddp_case(name, K=64, D=32, B=2, h=4, w=4, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, 29611, K, D, B, h, w, q))
             for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get() for _ in range(world))
    [p.join() for p in procs]
    rec = {'K': K, 'D': D, 'world': world}
    for r, out in res.items():
        for k, v in out.items():
            rec[f'r{r}_{k}'] = v
    np.savez_compressed(os.path.join(OUT, f'vq_{name}.npz'), **rec)
    print('wrote', name)


def blur_cases(VQGANFCM):
    import torchvision.transforms as T
    rec = {}
    g = torch.Generator().manual_seed(99)
    i = 0
    for (C, H, W) in [(3, 12, 12), (2, 16, 9)]:
        for k in (3, 5, 9):
            for sigma in (3.0, 0.7):
                if k // 2 >= min(H, W):
                    continue
                x = torch.randn(2, C, H, W, generator=g, requires_grad=True)
                go = torch.randn(2, C, H, W, generator=g)
                sig = torch.nn.Parameter(torch.tensor([sigma] * 4))
                shim = types.SimpleNamespace(kernel_size=k, sigmas=sig, padding=[k // 2] * 4)
                shim._get_gaussian_kernel1d = types.MethodType(VQGANFCM._get_gaussian_kernel1d, shim)
                shim._get_gaussian_kernel2d = types.MethodType(VQGANFCM._get_gaussian_kernel2d, shim)
                y = VQGANFCM._gaussian_blur(shim, x, 1, device='cpu')
                (y * go).sum().backward()
                tv = T.GaussianBlur(kernel_size=(k, k), sigma=sigma)(x.detach())
                rec.update({f'x{i}': _np(x), f'go{i}': _np(go), f'y{i}': _np(y),
                            f'gx{i}': _np(x.grad), f'gsig{i}': _np(sig.grad[1]),
                            f'tv{i}': _np(tv), f'k{i}': k, f'sigma{i}': sigma})
                i += 1
    rec['n'] = i
    np.savez_compressed(os.path.join(OUT, 'blur_cases.npz'), **rec)
    print('wrote blur', i)


def wrapper_cases(vl):
    from oracle.ffl_oracle import FocalFrequencyLossOracle
    g = torch.Generator().manual_seed(5)
    shapes = [(2, 4, 16, 16), (2, 6, 8, 8), (2, 6, 8, 8), (2, 3, 8, 8)]
    en = [torch.randn(*s, generator=g) for s in shapes]
    de = [torch.randn(*s, generator=g) for s in reversed(shapes)]
    ffl = FocalFrequencyLossOracle(loss_weight=0.01, alpha=1.0)
    rec = {}
    for i, (e, d) in enumerate(zip(en, de)):
        rec[f'en{i}'] = _np(e); rec[f'de{i}'] = _np(d)
    de1 = list(de)
    loss, lst = vl.recon_ffl_features_loss(ffl, list(en), de1, 'cpu')
    rec['dsl_loss'] = _np(loss); rec['dsl_list'] = np.array([float(v) for v in lst])
    rec['dsl_reversed_inplace'] = int(de1[0] is de[-1])
    de2 = list(de)
    loss, lst = vl.recon_sl_gaussian_features_loss(ffl, 5, 3, list(en), de2, 'cpu')
    rec['sl_loss'] = _np(loss); rec['sl_list'] = np.array([float(v) for v in lst])
    x = torch.randn(2, 3, 16, 16, generator=g); xr = torch.randn(2, 3, 16, 16, generator=g)
    rec['img_x'] = _np(x); rec['img_xr'] = _np(xr)
    rec['img_loss'] = _np(vl.recon_ffl_loss(ffl, x, xr))
    np.savez_compressed(os.path.join(OUT, 'wrappers.npz'), **rec)
    print('wrote wrappers')


def main():
    os.makedirs(OUT, exist_ok=True)
    l2q, vl, VQGANFCM = _import_reference()
    vq_case(l2q, 'cos_small', K=64, D=32, B=2, h=4, w=4, cosine=True, steps=2, commit=0.25)
    vq_case(l2q, 'cos_mid', K=512, D=64, B=2, h=8, w=8, cosine=True, steps=3, commit=1.0, seed=1)
    vq_case(l2q, 'cos_proj', K=128, D=32, B=2, h=8, w=8, cosine=True, steps=2, dim=3, seed=2)
    vq_case(l2q, 'euclid_small', K=64, D=32, B=2, h=4, w=4, cosine=False, steps=2, seed=3)
    vq_case(l2q, 'cos_heads2', K=64, D=32, B=2, h=4, w=4, cosine=True, steps=2, seed=4, heads=2)
    vq_big_case(l2q, 'big_k1024', K=1024)
    vq_big_case(l2q, 'big_k16384', K=16384, B=2, seed=12)
    vq_big_case(l2q, 'big_ortho', K=1024, seed=13, ortho=10.0)
    vq_ddp_case('cos_ddp2')
    blur_cases(VQGANFCM)
    wrapper_cases(vl)


if __name__ == '__main__':
    main()
