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


def _bad_rows(idx_gpu, idx_ref):
    """(rows whose index differs, codes touched by those rows) -- everything else must still match."""
    a, b = idx_gpu.reshape(-1).cpu(), idx_ref.reshape(-1).cpu()
    bad = (a != b).nonzero().flatten()
    codes = torch.cat([a[bad], b[bad]]).unique()
    return bad, codes


def _compare_step(q, loss, gx, embed, cluster, ref, bad, codes, D):
    """Compare one training step with the reference values in ``ref`` (dict of CPU tensors: q, loss, gx
    as (B,D,h,w); embed (K,D); cluster (K,)).  A near-tie flip (documented in BASELINE.json) changes
    only the flipped latents' rows of q / gx and the two codes involved: those are masked out and every
    other row and code is held to the same tolerance as without a flip."""
    def rows(t):
        return t.detach().cpu().permute(0, 2, 3, 1).reshape(-1, D)
    keep = torch.ones(rows(q).shape[0], dtype=torch.bool)
    keep[bad] = False
    ckeep = torch.ones(ref['embed'].shape[0], dtype=torch.bool)
    ckeep[codes] = False
    torch.testing.assert_close(rows(q)[keep], rows(ref['q'])[keep], rtol=1e-4, atol=1e-6)
    if gx is not None:
        # the commitment term of flipped rows enters only their own gradient rows
        torch.testing.assert_close(rows(gx)[keep], rows(ref['gx'])[keep], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(embed.cpu()[ckeep], ref['embed'][ckeep], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(cluster.cpu()[ckeep], ref['cluster'][ckeep], rtol=1e-5, atol=1e-7)
    if bad.numel() == 0:
        torch.testing.assert_close(loss.detach().cpu(), ref['loss'], rtol=1e-4, atol=1e-8)
    else:
        # a flipped latent sits within 1e-6 (cosine) of both codes, but its squared distance to the
        # un-normalised code may differ by O(1/N) of the mean: bound instead of skipping
        n = keep.numel()
        torch.testing.assert_close(loss.detach().cpu(), ref['loss'], rtol=4.0 * bad.numel() / n + 1e-4, atol=1e-8)


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
        if proj:
            torch.testing.assert_close(q, _t(g[f'q{s}']), rtol=1e-4, atol=1e-6)
            torch.testing.assert_close(loss, _t(g[f'loss{s}']), rtol=1e-4, atol=1e-8)
            torch.testing.assert_close(x.grad, _t(g[f'gx{s}']), rtol=1e-4, atol=1e-7)
            torch.testing.assert_close(vq._codebook.embed[0], _t(g[f'embed{s + 1}']), rtol=1e-4, atol=1e-6)
            torch.testing.assert_close(vq._codebook.cluster_size[0], _t(g[f'cluster{s + 1}']), rtol=1e-5, atol=1e-7)
        else:
            bad, codes = _bad_rows(ind, _t(g[f'ind{s}']))
            ref = dict(q=_t(g[f'q{s}'], 'cpu'), loss=_t(g[f'loss{s}'], 'cpu'), gx=_t(g[f'gx{s}'], 'cpu'),
                       embed=_t(g[f'embed{s + 1}'], 'cpu'), cluster=_t(g[f'cluster{s + 1}'], 'cpu'))
            _compare_step(q, loss, x.grad, vq._codebook.embed[0], vq._codebook.cluster_size[0], ref, bad, codes, D)
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
    bad, codes = _bad_rows(ind, r['embed_ind'])
    gx = vo.vector_quantize_backward(r['flat'], r['q_flat'], gq.permute(0, 2, 3, 1).reshape(-1, D), 1.0, 0.25)
    gx = gx.reshape(B, hw, hw, D).permute(0, 3, 1, 2)
    ref = dict(q=r['quantize'], loss=r['loss'], gx=gx, embed=r['new_embed'], cluster=r['new_cluster_size'])
    _compare_step(q, loss, xg.grad, vq._codebook.embed[0], vq._codebook.cluster_size[0], ref, bad, codes, D)


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


def test_fallback_search_single_and_many_overflow_rows():
    """The exhaustive fallback splits the codebook into slices across blocks (one overflowing latent must
    not serialise on one block) and meets in a 64-bit atomicMax: one, a few and many overflow rows."""
    torch.manual_seed(9)
    k, d = 16384, 256
    embed = torch.nn.functional.normalize(torch.randn(k, d, device='cuda'), dim=-1)
    embed[9000:9200] = embed[4242].clone()           # 201 identical codes: any latent near them overflows
    for n_over in (1, 5, 700):
        x = torch.randn(1024, d, device='cuda')
        rows = torch.randperm(1024)[:n_over]
        x[rows] = embed[4242] * 1.7 + 1e-5 * torch.randn(n_over, d, device='cuda')
        x[rows[0]] = 0.0                             # all-zero latent: every code ties -> index 0
        o = _tc_search(x, embed)
        assert o['tc'][rows[0]].item() == 0
        if n_over > 1:
            assert torch.all(o['tc'][rows[1:]] == 4242)
        _assert_indices(o['tc'], o['exact'], x, embed)


def _big_codebook(g):
    """Seeded initial codebook of a vq_big_* fixture (see tests/test_oracle_vq.py::big_case_codebook)."""
    torch.manual_seed(int(g['seed']))
    e = torch.empty(1, int(g['K']), int(g['D']))
    torch.nn.init.kaiming_uniform_(e)
    e = torch.nn.functional.normalize(e, p=2, dim=-1)
    assert float(e.double().sum()) == pytest.approx(float(g['embed0_sum']), rel=1e-12, abs=1e-9)
    assert torch.equal(e[0, 7], _t(g['embed0_row7'], 'cpu'))
    return e


@pytest.mark.parametrize('name', ['big_k1024', 'big_k16384', 'big_ortho'])
@pytest.mark.parametrize('mode', ['exact', 'auto'])
def test_reference_fixtures_at_production_width(golden_dir, name, mode, monkeypatch):
    """D = 256 with K = 1024 / 16384 -- the shapes the tcgen05 search serves -- against values recorded
    from the reference itself; big_ortho adds orthogonal_reg_weight = 10 (learnable codebook, :563-577)."""
    monkeypatch.setenv('FAVAE_VQ_SEARCH', mode)
    from favae_b200 import VectorQuantize
    g = np.load(os.path.join(golden_dir, f'vq_{name}.npz'))
    K, D, ortho = int(g['K']), int(g['D']), float(g['ortho'])
    vq = VectorQuantize(dim=D, codebook_size=K, accept_image_fmap=True, use_cosine_sim=True,
                        commitment_weight=1.0, orthogonal_reg_weight=ortho,
                        orthogonal_reg_max_codes=128 if ortho else None).cuda().train()
    embed0 = _big_codebook(g)
    with torch.no_grad():
        vq._codebook.embed.copy_(embed0)
    assert isinstance(vq._codebook.embed, torch.nn.Parameter) == (ortho > 0)
    x = _t(g['x']).requires_grad_(True)
    q, ind, loss = vq(x)
    (q * _t(g['gq'])).sum().add(loss.sum() * 0.7).backward()
    flat = x.detach().permute(0, 2, 3, 1).reshape(-1, D)
    _assert_indices(ind, _t(g['ind']), flat, embed0[0])
    bad, codes = _bad_rows(ind, _t(g['ind']))
    assert bad.numel() <= 1
    rows = _t(g['rows'], 'cpu')
    keep = ~torch.isin(rows, codes)
    e1 = vq._codebook.embed.detach()[0].cpu()
    torch.testing.assert_close(e1[rows][keep], _t(g['embed1_rows'], 'cpu')[keep], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(vq._codebook.cluster_size[0].cpu()[rows][keep], _t(g['cluster1_rows'], 'cpu')[keep],
                               rtol=1e-5, atol=1e-7)
    if bad.numel() == 0:
        torch.testing.assert_close(q, _t(g['q']), rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(loss, _t(g['loss']), rtol=1e-4, atol=1e-8)
        torch.testing.assert_close(x.grad, _t(g['gx']), rtol=1e-4, atol=1e-7)
        assert float(e1.double().sum()) == pytest.approx(float(g['embed1_sum']), rel=1e-5, abs=1e-3)
    if ortho:
        ge = vq._codebook.embed.grad[0].cpu()
        ref = _t(g['gembed_rows'], 'cpu')
        assert (ge[rows][keep] - ref[keep]).abs().max() <= 1e-4 * ref.abs().max()
        assert float(ge.double().abs().sum()) == pytest.approx(float(g['gembed_abs']), rel=1e-4)
        # active-codes-only indexes the head axis in the reference: index error for any code id > 0
        vq.orthogonal_reg_active_codes_only = True
        with pytest.raises(IndexError):
            vq(x.detach())


def test_deterministic_statistics_mode(monkeypatch):
    """FAVAE_VQ_DETERMINISTIC=1: bins / embed_sum are summed in latent order by one warp per code, so the
    EMA result is bit-identical from run to run; it agrees with the atomic scatter-add to rounding."""
    from favae_b200 import VectorQuantize
    x = torch.randn(8, 256, 16, 16, device='cuda', generator=torch.Generator('cuda').manual_seed(3))
    x[1] = x[0]                      # many latents per code: long per-code sums
    x[2:6] = x[0] * 0.5

    def run(det):
        monkeypatch.setenv('FAVAE_VQ_DETERMINISTIC', '1' if det else '0')
        torch.manual_seed(0)
        vq = VectorQuantize(dim=256, codebook_size=1024, accept_image_fmap=True, use_cosine_sim=True).cuda().train()
        for _ in range(3):
            vq(x)
        return vq._codebook.embed[0].clone(), vq._codebook.cluster_size[0].clone()
    a, ca = run(True)
    b, cb = run(True)
    assert torch.equal(a, b) and torch.equal(ca, cb)
    c, cc = run(False)
    torch.testing.assert_close(a, c, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ca, cc, rtol=0, atol=0)      # counts are exact either way


def test_cached_codebook_rows_are_bit_identical_and_invalidate(monkeypatch):
    """The EMA kernel emits l2norm(updated codebook) for the next search; those rows must equal a fresh
    preparation bit for bit, and any torch-side write to ``embed`` must invalidate them."""
    from favae_b200 import VectorQuantize
    monkeypatch.setenv('FAVAE_VQ_DETERMINISTIC', '1')     # bitwise comparison of two runs below
    torch.manual_seed(1)
    vq = VectorQuantize(dim=256, codebook_size=1024, accept_image_fmap=True, use_cosine_sim=True).cuda().train()
    cb = vq._codebook
    x = torch.randn(4, 256, 16, 16, device='cuda')
    vq(x); vq(x)
    prep = cb.__dict__['_prep']
    en, eh = cb._prepare(cb.embed.detach()[0], 1024, 1, True, True)
    assert torch.equal(prep['en'], en) and torch.equal(prep['eh'], eh)
    _, ind_cached, _ = vq(x)
    # same call sequence with the cache disabled gives the same indices and codebook
    torch.manual_seed(1)
    vq2 = VectorQuantize(dim=256, codebook_size=1024, accept_image_fmap=True, use_cosine_sim=True).cuda().train()
    os.environ['FAVAE_VQ_CACHE'] = '0'
    try:
        vq2(x); vq2(x)
        _, ind_plain, _ = vq2(x)
    finally:
        os.environ.pop('FAVAE_VQ_CACHE')
    assert torch.equal(ind_cached, ind_plain)
    assert torch.equal(cb.embed, vq2._codebook.embed)
    # a write through torch (load_state_dict, copy_, an optimiser step) bumps the version counter
    with torch.no_grad():
        cb.embed.copy_(torch.nn.functional.normalize(torch.randn_like(cb.embed), dim=-1))
    vq.eval()
    _, ind_new, _ = vq(x)
    vq3 = VectorQuantize(dim=256, codebook_size=1024, accept_image_fmap=True, use_cosine_sim=True).cuda().eval()
    vq3.load_state_dict(vq.state_dict())
    _, ind_ref, _ = vq3(x)
    assert torch.equal(ind_new, ind_ref)


def test_half_precision_inputs_under_autocast():
    """cat_scripts/train_cat.py runs the model under autocast, so the quantizer, the losses and the blur
    receive fp16 / bf16 tensors; the reference casts with .float() (l2_quantize.py:393)."""
    from favae_b200 import FocalFrequencyLoss, VectorQuantize, gaussian_blur_reflect
    torch.manual_seed(2)
    vq = VectorQuantize(dim=64, codebook_size=256, accept_image_fmap=True, use_cosine_sim=True).cuda().train()
    x32 = torch.randn(2, 64, 8, 8, device='cuda')
    for dt in (torch.float16, torch.bfloat16):
        xh = x32.to(dt).requires_grad_(True)
        with torch.autocast('cuda', dtype=dt):
            q, ind, loss = vq(xh)
        assert q.dtype == torch.float32 and loss.dtype == torch.float32
        (q.sum() + loss.sum()).backward()
        assert xh.grad is not None and xh.grad.dtype == dt
        vq.eval()
        _, ind_ref, _ = vq(xh.detach().float())
        vq.train()
        p = torch.randn(1, 2, 32, 32, device='cuda', dtype=dt, requires_grad=True)
        t = torch.randn(1, 2, 32, 32, device='cuda', dtype=dt)
        l = FocalFrequencyLoss()(gaussian_blur_reflect(p, 2.0, 5), t)
        l.backward()
        ref = FocalFrequencyLoss()(gaussian_blur_reflect(p.detach().float(), 2.0, 5), t.float())
        assert float(l) == pytest.approx(float(ref), rel=1e-6)
        assert p.grad.dtype == dt


def test_tensors_on_a_non_current_device(monkeypatch):
    """Calls follow the tensors' device, not the current one (advisor finding, round 1)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    monkeypatch.setenv('FAVAE_VQ_DETERMINISTIC', '1')     # the two devices' codebooks are compared bit for bit
    from favae_b200 import FocalFrequencyLoss, VectorQuantize, gaussian_blur_reflect
    torch.manual_seed(4)
    vq0 = VectorQuantize(dim=256, codebook_size=1024, accept_image_fmap=True, use_cosine_sim=True).train()
    import copy
    vq1 = copy.deepcopy(vq0).to('cuda:1')
    vq0 = vq0.to('cuda:0')
    x = torch.randn(2, 256, 16, 16)
    p = torch.randn(1, 2, 256, 256); t = torch.randn(1, 2, 256, 256)
    torch.cuda.set_device(0)
    q0, i0, l0 = vq0(x.to('cuda:0'))
    q1, i1, l1 = vq1(x.to('cuda:1'))          # tensors on cuda:1 while cuda:0 is current
    assert torch.equal(i0.cpu(), i1.cpu()) and torch.equal(q0.cpu(), q1.cpu())
    assert torch.equal(vq0._codebook.embed.cpu(), vq1._codebook.embed.cpu())
    f0 = FocalFrequencyLoss()(gaussian_blur_reflect(p.to('cuda:0'), 3.0, 9), t.to('cuda:0'))
    f1 = FocalFrequencyLoss()(gaussian_blur_reflect(p.to('cuda:1'), 3.0, 9), t.to('cuda:1'))
    assert float(f0) == float(f1)
    with pytest.raises(RuntimeError):
        FocalFrequencyLoss()(p.to('cuda:0'), t.to('cuda:1'))


@pytest.mark.gpu
def test_ema_on_the_side_stream_matches_the_inline_update(monkeypatch):
    """The training call leaves its tail (statistics -> EMA) on a side stream; FAVAE_EMA_STREAM=0 keeps it on the
    caller's stream.  Same codebook, indices and loss either way, bit for bit, over several steps -- and the
    pending update is applied before anybody can look at the codebook."""
    from favae_b200 import VectorQuantize
    monkeypatch.setenv('FAVAE_VQ_DETERMINISTIC', '1')                 # ordered statistics: bit-comparable runs
    torch.manual_seed(3)
    a = VectorQuantize(dim=256, codebook_size=1024, accept_image_fmap=True, use_cosine_sim=True).cuda().train()
    b = VectorQuantize(dim=256, codebook_size=1024, accept_image_fmap=True, use_cosine_sim=True).cuda().train()
    b.load_state_dict(a.state_dict())
    g = torch.Generator(device='cuda').manual_seed(4)
    for step in range(3):
        x = torch.randn(4, 256, 16, 16, device='cuda', generator=g)
        monkeypatch.setenv('FAVAE_EMA_STREAM', '1')
        qa, ia, la = a(x)
        assert a._codebook.__dict__.get('_pending') is not None        # update in flight on the side stream
        monkeypatch.setenv('FAVAE_EMA_STREAM', '0')
        qb, ib, lb = b(x)
        assert b._codebook.__dict__.get('_pending') is None
        assert torch.equal(ia, ib) and torch.equal(qa, qb) and torch.equal(la, lb)
        assert torch.equal(a._codebook.embed, b._codebook.embed)       # the access waits for the side stream
        assert a._codebook.__dict__.get('_pending') is None
        assert torch.equal(a._codebook.cluster_size, b._codebook.cluster_size)
    with torch.no_grad():
        a.eval(); b.eval()
        x = torch.randn(2, 256, 16, 16, device='cuda', generator=g)
        assert torch.equal(a(x)[1], b(x)[1])
