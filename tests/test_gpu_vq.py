"""GPU parity of the vector quantizer (through the drop-in module and the C ABI underneath)
against the golden fixtures recorded from the reference and against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import vq_oracle as vo

pytestmark = pytest.mark.gpu


def _t(a, dev='cuda'):
    return torch.from_numpy(np.asarray(a)).to(dev)


def _assert_indices(idx_gpu, idx_ref, flat, embed, cosine=True, tol=1e-6):
    """Bit-exact, except documented near-ties: where the two disagree the reference scores of the
    two codes (fp64) must be within `tol`."""
    idx_gpu = idx_gpu.reshape(-1).cpu(); idx_ref = idx_ref.reshape(-1).cpu()
    bad = (idx_gpu != idx_ref).nonzero().flatten()
    if bad.numel() == 0:
        return 0
    x = flat.double().cpu()[bad]; e = embed.double().cpu()
    if cosine:
        x = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        e = e / e.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        sg = (x * e[idx_gpu[bad]]).sum(-1); sr = (x * e[idx_ref[bad]]).sum(-1)
    else:
        sg = -(x - e[idx_gpu[bad]]).norm(dim=-1); sr = -(x - e[idx_ref[bad]]).norm(dim=-1)
    gap = (sg - sr).abs().max().item()
    assert gap <= tol, f'{bad.numel()} index mismatches, worst score gap {gap:.3e} > {tol}'
    return bad.numel()


def _module(g, sync=False):
    from favae_b200 import VectorQuantize
    vq = VectorQuantize(dim=int(g['dim']), codebook_size=int(g['K']), codebook_dim=int(g['D']),
                        accept_image_fmap=True, use_cosine_sim=bool(g['cosine']),
                        commitment_weight=float(g['commit']), sync_codebook=sync,
                        heads=int(g['heads']) if 'heads' in g else 1).cuda()
    sd = {'_codebook.initted': torch.ones(1), '_codebook.cluster_size': _t(g['cluster0'], 'cpu')[None],
          '_codebook.embed': _t(g['embed0'], 'cpu')[None]}
    if not bool(g['cosine']):
        sd['_codebook.embed_avg'] = _t(g['embed_avg0'], 'cpu')[None]
    if int(g['dim']) != int(g['D']) * (int(g['heads']) if 'heads' in g else 1):
        sd.update({'project_in.weight': _t(g['pin_w'], 'cpu'), 'project_in.bias': _t(g['pin_b'], 'cpu'),
                   'project_out.weight': _t(g['pout_w'], 'cpu'), 'project_out.bias': _t(g['pout_b'], 'cpu')})
    missing = vq.load_state_dict(sd, strict=True)      # same keys as the reference module
    assert not missing.missing_keys and not missing.unexpected_keys
    return vq


@pytest.mark.parametrize('name', ['cos_small', 'cos_mid', 'cos_proj', 'euclid_small', 'cos_heads2'])
@pytest.mark.parametrize('mode', ['exact', 'auto'])
def test_golden_replay(golden_dir, name, mode, monkeypatch):
    monkeypatch.setenv('FAVAE_VQ_SEARCH', mode)
    g = np.load(os.path.join(golden_dir, f'vq_{name}.npz'))
    vq = _module(g).train()
    cosine, D = bool(g['cosine']), int(g['D'])
    heads = int(g['heads']) if 'heads' in g else 1
    proj = int(g['dim']) != D or heads > 1          # no flat (N, D) view of x to re-score near-ties with
    for s in range(int(g['steps'])):
        x = _t(g[f'x{s}']).requires_grad_(True)
        embed_before = vq._codebook.embed[0].clone()
        q, ind, loss = vq(x)
        assert q.shape == x.shape and ind.dtype == torch.int64 and loss.shape == (1,)
        (q * _t(g[f'gq{s}'])).sum().add(loss.sum() * 0.7).backward()
        if not proj:
            flat = x.detach().permute(0, 2, 3, 1).reshape(-1, D)
            n_bad = _assert_indices(ind, _t(g[f'ind{s}']), flat, embed_before, cosine)
        else:
            n_bad = int((ind.cpu() != _t(g[f'ind{s}'], 'cpu')).sum())
            assert n_bad == 0
        if n_bad == 0:
            torch.testing.assert_close(q, _t(g[f'q{s}']), rtol=1e-4, atol=1e-6)
            torch.testing.assert_close(loss, _t(g[f'loss{s}']), rtol=1e-4, atol=1e-8)
            torch.testing.assert_close(x.grad, _t(g[f'gx{s}']), rtol=1e-4, atol=1e-7)
            torch.testing.assert_close(vq._codebook.embed[0], _t(g[f'embed{s + 1}']), rtol=1e-4, atol=1e-6)
            torch.testing.assert_close(vq._codebook.cluster_size[0], _t(g[f'cluster{s + 1}']), rtol=1e-5, atol=1e-7)
        # keep later steps aligned with the fixture even if a near-tie flipped
        vq._codebook.embed.copy_(_t(g[f'embed{s + 1}'])[None])
        vq._codebook.cluster_size.copy_(_t(g[f'cluster{s + 1}'])[None])
    vq.eval()
    q, ind, loss = vq(_t(g['x_eval']))
    assert float(loss) == 0.0 and not loss.requires_grad
    if not proj:
        assert torch.equal(ind.cpu(), _t(g['ind_eval'], 'cpu'))
        torch.testing.assert_close(q, _t(g['q_eval']), rtol=0, atol=0)      # pure gather: exact
        B, _, h, w = g['x_eval'].shape
        e = vq.get_codebook_entry(_t(g['entry_ids']), (B, h, w, D))
        torch.testing.assert_close(e, _t(g['entry']), rtol=0, atol=0)
    else:
        torch.testing.assert_close(q, _t(g['q_eval']), rtol=1e-4, atol=1e-6)


def test_exact_tie_goes_to_lowest_index(golden_dir, monkeypatch):
    g = np.load(os.path.join(golden_dir, 'vq_cos_small.npz'))
    for mode in ('exact', 'auto'):
        monkeypatch.setenv('FAVAE_VQ_SEARCH', mode)
        vq = _module(g).eval()
        _, ind, _ = vq(_t(g['x0']))
        assert int(ind[0, 0, 0]) == 3          # codes 3 and 5 are identical (make_golden.py)


@pytest.mark.parametrize('K,B,hw', [(1024, 2, 16), (16384, 8, 16), (8192, 1, 64)])
@pytest.mark.parametrize('mode', ['exact', 'auto'])
def test_training_step_vs_oracle(K, B, hw, mode, monkeypatch):
    """BASELINE configs: f=16 (16x16 latents) K=1024 / K=16384, f=4 (64x64 latents) K=8192; D=256."""
    monkeypatch.setenv('FAVAE_VQ_SEARCH', mode)
    from favae_b200 import VectorQuantize
    D = 256
    torch.manual_seed(K + B)
    vq = VectorQuantize(dim=D, codebook_size=K, accept_image_fmap=True, use_cosine_sim=True,
                        commitment_weight=0.25).cuda().train()
    embed0 = vq._codebook.embed[0].clone().cpu()
    x = torch.randn(B, D, hw, hw, generator=torch.Generator().manual_seed(1234))
    xg = x.cuda().requires_grad_(True)
    q, ind, loss = vq(xg)
    gq = torch.randn(q.shape, generator=torch.Generator().manual_seed(5))
    (q * gq.cuda()).sum().add(loss.sum()).backward()
    r = vo.vector_quantize_forward(x, embed0, torch.zeros(K), training=True, commitment_weight=0.25)
    flat = x.permute(0, 2, 3, 1).reshape(-1, D)
    n_bad = _assert_indices(ind, r['embed_ind'], flat, embed0)
    assert n_bad <= max(2, ind.numel() // 2000)
    if n_bad == 0:
        torch.testing.assert_close(q.cpu(), r['quantize'], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(loss.cpu(), r['loss'], rtol=1e-4, atol=1e-8)
        gx = vo.vector_quantize_backward(r['flat'], r['q_flat'], gq.permute(0, 2, 3, 1).reshape(-1, D), 1.0, 0.25)
        torch.testing.assert_close(xg.grad.cpu().permute(0, 2, 3, 1).reshape(-1, D), gx, rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(vq._codebook.embed[0].cpu(), r['new_embed'], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(vq._codebook.cluster_size[0].cpu(), r['new_cluster_size'], rtol=1e-5, atol=1e-7)


def test_size_independent_properties():
    """Full-size check without the oracle: idempotence and agreement of both search paths."""
    from favae_b200 import VectorQuantize
    torch.manual_seed(0)
    vq = VectorQuantize(dim=256, codebook_size=16384, accept_image_fmap=True, use_cosine_sim=True).cuda().eval()
    x = torch.randn(32, 256, 16, 16, device='cuda')
    q, ind, _ = vq(x)
    q2, ind2, _ = vq(q)                     # codes are fixed points of the quantizer
    assert torch.equal(ind, ind2) and torch.equal(q, q2)
    assert torch.equal(q, vq.get_codebook_entry(ind.reshape(32, -1), (32, 16, 16, 256)))
    os.environ['FAVAE_VQ_SEARCH'] = 'exact'
    try:
        _, ind3, _ = vq(x)
    finally:
        os.environ.pop('FAVAE_VQ_SEARCH')
    assert (ind3 != ind).sum().item() <= 2


def test_edge_cases():
    from favae_b200 import VectorQuantize
    vq = VectorQuantize(dim=64, codebook_size=128, accept_image_fmap=True, use_cosine_sim=True).cuda().train()
    x = torch.zeros(1, 64, 3, 5, device='cuda')           # all-zero latents, ragged 3x5 grid
    q, ind, loss = vq(x)
    assert torch.all(ind == 0) and torch.isfinite(loss).all()
    with pytest.raises(RuntimeError):
        vq(torch.zeros(1, 64, 2, 2))                       # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        VectorQuantize(dim=64, codebook_size=128, heads=2, separate_codebook_per_head=True)
    with pytest.raises(NotImplementedError):
        VectorQuantize(dim=64, codebook_size=128, sample_codebook_temp=0.5)


def test_kmeans_init_and_dead_code_expiry():
    """Cold options of the reference (kmeans :124-164, expire_codes_ :369-389) built on the search and
    statistics kernels.  Sampling is RNG dependent, so properties are checked instead of fixtures."""
    from favae_b200 import VectorQuantize
    torch.manual_seed(0)
    K, D = 32, 64
    centers = torch.nn.functional.normalize(torch.randn(K, D, device='cuda'), dim=-1)
    x = (centers[torch.randint(0, K, (4 * 16 * 16,), device='cuda')] + 0.01 * torch.randn(1024, D, device='cuda'))
    x = x.view(4, 16, 16, D).permute(0, 3, 1, 2).contiguous()
    vq = VectorQuantize(dim=D, codebook_size=K, accept_image_fmap=True, use_cosine_sim=True, kmeans_init=True,
                        kmeans_iters=10).cuda().train()
    assert float(vq._codebook.initted) == 0.0 and float(vq._codebook.embed.abs().sum()) == 0.0
    q, ind, loss = vq(x)
    assert float(vq._codebook.initted) == 1.0
    torch.testing.assert_close(vq._codebook.embed[0].norm(dim=-1).clamp(max=1.0001),
                               vq._codebook.embed[0].norm(dim=-1))          # EMA of unit vectors: norms <= 1
    # k-means on well separated clusters: most latents end within a small angle of their code
    # (random seeding can merge two clusters, as in the reference)
    qn = torch.nn.functional.normalize(q.permute(0, 2, 3, 1).reshape(-1, D), dim=-1)
    xn = torch.nn.functional.normalize(x.permute(0, 2, 3, 1).reshape(-1, D), dim=-1)
    cos = (qn * xn).sum(-1)
    assert float(cos.median()) > 0.98 and float(cos.mean()) > 0.9
    assert float(loss) < 0.5
    # dead-code expiry: codes that never win are re-seeded with normalised latents of the batch
    vq2 = VectorQuantize(dim=D, codebook_size=128, accept_image_fmap=True, use_cosine_sim=True,
                         threshold_ema_dead_code=2).cuda().train()
    before = vq2._codebook.embed[0].clone()
    vq2(x)
    dead = vq2._codebook.cluster_size[0] < 2
    assert int(dead.sum()) > 0
    after = vq2._codebook.embed[0]
    sims = (after[dead] @ xn.t()).amax(dim=-1)
    torch.testing.assert_close(sims, torch.ones_like(sims), rtol=0, atol=1e-5)   # each is a batch latent
    assert not torch.equal(after[dead], before[dead])


def _tc_search(x, embed):
    """Call the C ABI directly: prepare rows, tensor-core search, exact search."""
    from favae_b200 import _lib
    n, d = x.shape; k = embed.shape[0]
    outs = {}
    bufs = {}
    for name, t, rows in (('x', x, n), ('e', embed, k)):
        f32 = torch.empty(rows, d, device='cuda'); f16 = torch.empty(rows, d, device='cuda', dtype=torch.float16)
        _lib.call('favae_vq_prepare_rows', t.data_ptr(), rows, d, 1, 1, f32.data_ptr(), f16.data_ptr(), None,
                  _lib.stream())
        bufs[name] = (f32, f16)
    keys = torch.empty(n, device='cuda', dtype=torch.int64)
    for mode in ('tc', 'exact'):
        idx = torch.full((n,), -7, device='cuda', dtype=torch.int64)
        if mode == 'tc':
            nbytes = _lib.load().favae_vq_search_tc_workspace_bytes(n, k, d)
            assert nbytes > 0
            ws = torch.empty(nbytes, device='cuda', dtype=torch.uint8)
            _lib.call('favae_vq_search_tc', bufs['x'][1].data_ptr(), bufs['e'][1].data_ptr(),
                      bufs['x'][0].data_ptr(), bufs['e'][0].data_ptr(), n, k, d, ws.data_ptr(), nbytes,
                      keys.data_ptr(), idx.data_ptr(), _lib.stream())
        else:
            _lib.call('favae_vq_search_exact', bufs['x'][0].data_ptr(), bufs['e'][0].data_ptr(), None, n, k, d, 0,
                      keys.data_ptr(), idx.data_ptr(), _lib.stream())
        torch.cuda.synchronize()
        outs[mode] = idx
    return outs


@pytest.mark.parametrize('n,k,d', [(128, 256, 64), (100, 512, 128), (1000, 1024, 256), (4096, 16384, 256),
                                   (37, 256, 192), (20000, 2048, 256)])
def test_tensor_core_search_matches_exact_search(n, k, d):
    torch.manual_seed(n + k)
    x = torch.randn(n, d, device='cuda')
    embed = torch.nn.functional.normalize(torch.randn(k, d, device='cuda'), dim=-1)
    o = _tc_search(x, embed)
    assert int(o['tc'].min()) >= 0 and int(o['tc'].max()) < k
    _assert_indices(o['tc'], o['exact'], x, embed)
    assert (o['tc'] != o['exact']).sum().item() <= max(1, n // 2000)


def test_tensor_core_search_ties_duplicates_and_overflow():
    torch.manual_seed(3)
    k, d = 512, 64
    embed = torch.nn.functional.normalize(torch.randn(k, d, device='cuda'), dim=-1)
    embed[[5, 100, 300]] = embed[3].clone()         # 4 identical codes: inside the candidate band
    embed[200:220] = embed[77].clone()              # 21 identical codes: candidate list overflows
    x = torch.randn(300, d, device='cuda')
    x[0] = embed[3] * 2.5                           # exact 4-way tie -> lowest index 3
    x[1] = embed[77] * 0.3                          # 21-way tie -> exhaustive fallback -> 77
    x[2] = 0.0                                      # all-zero latent: every code ties -> 0
    x[3] = embed[300] + 1e-4 * torch.randn(d, device='cuda')
    o = _tc_search(x, embed)
    assert o['tc'][0].item() == 3 and o['tc'][1].item() == 77 and o['tc'][2].item() == 0
    assert o['tc'][3].item() == 3
    assert torch.equal(o['tc'][:4], o['exact'][:4])
    _assert_indices(o['tc'], o['exact'], x, embed)
