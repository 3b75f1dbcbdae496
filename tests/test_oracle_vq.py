"""Oracle (oracle/vq_oracle.py) pinned against fixtures recorded from the reference
models/l2_quantize.py (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import vq_oracle as vo


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f'vq_{name}.npz'))


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize('name', ['cos_small', 'cos_mid', 'euclid_small'])
def test_vq_oracle_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    cosine = bool(g['cosine'])
    embed, cluster = _t(g['embed0']), _t(g['cluster0'])
    embed_avg = _t(g['embed_avg0']) if not cosine else None
    for s in range(int(g['steps'])):
        x = _t(g[f'x{s}'])
        r = vo.vector_quantize_forward(x, embed, cluster, training=True,
                                       commitment_weight=float(g['commit']),
                                       use_cosine_sim=cosine, embed_avg=embed_avg)
        assert torch.equal(r['embed_ind'], _t(g[f'ind{s}'])), f'indices differ at step {s}'
        torch.testing.assert_close(r['quantize'], _t(g[f'q{s}']), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(r['loss'], _t(g[f'loss{s}']), rtol=1e-5, atol=1e-8)
        torch.testing.assert_close(r['new_embed'], _t(g[f'embed{s + 1}']), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(r['new_cluster_size'], _t(g[f'cluster{s + 1}']), rtol=1e-6, atol=1e-7)
        gq = _t(g[f'gq{s}']).permute(0, 2, 3, 1).reshape(-1, x.shape[1])
        gx = vo.vector_quantize_backward(r['flat'], r['q_flat'], gq, 0.7, float(g['commit']))
        gx = gx.reshape(x.shape[0], x.shape[2], x.shape[3], -1).permute(0, 3, 1, 2)
        torch.testing.assert_close(gx, _t(g[f'gx{s}']), rtol=1e-5, atol=1e-7)
        embed, cluster = r['new_embed'], r['new_cluster_size']
    embed = _t(g[f"embed{int(g['steps'])}"])     # isolate eval from EMA rounding drift
    r = vo.vector_quantize_forward(_t(g['x_eval']), embed, cluster, training=False,
                                   use_cosine_sim=cosine, embed_avg=embed_avg)
    assert torch.equal(r['embed_ind'], _t(g['ind_eval']))
    torch.testing.assert_close(r['quantize'], _t(g['q_eval']), rtol=0, atol=0)
    assert float(r['loss']) == 0.0
    B, _, h, w = g['x_eval'].shape
    e = vo.codebook_entry(_t(g['entry_ids']), embed, (B, h, w, int(g['D'])))
    torch.testing.assert_close(e, _t(g['entry']), rtol=0, atol=0)


def test_tie_goes_to_first_index(golden_dir):
    g = _load(golden_dir, 'cos_small')
    # codes 3 and 5 were made identical and latent (0,0,0) placed on them (make_golden.py)
    assert np.array_equal(g['embed0'][3], g['embed0'][5])
    assert int(g['ind0'][0, 0, 0]) == 3
    idx, _, _ = vo.cosine_search(_t(g['x0']).permute(0, 2, 3, 1).reshape(-1, 32), _t(g['embed0']))
    assert int(idx[0]) == 3


def test_vq_oracle_ddp_fixture(golden_dir):
    """2-rank gloo run of the reference with sync_codebook=True: the oracle with a summing
    all_reduce over both ranks' stats reproduces each rank's buffers."""
    g = _load(golden_dir, 'cos_ddp2')
    D = int(g['D'])
    embed = [_t(g['r0_embed0']), _t(g['r1_embed0'])]
    assert torch.equal(embed[0], embed[1])
    cluster = [torch.zeros(int(g['K'])) for _ in range(2)]
    for s in range(2):
        flats = [_t(g[f'r{r}_x{s}']).permute(0, 2, 3, 1).reshape(-1, D) for r in range(2)]
        # local stats of both ranks, summed (what all_reduce would produce)
        parts = []
        for r in range(2):
            idx, xn, _ = vo.cosine_search(flats[r], embed[r])
            bins = torch.bincount(idx, minlength=int(g['K'])).float()
            es = torch.zeros_like(embed[r]).index_add_(0, idx, xn)
            parts.append((bins, es))
        tot_bins = parts[0][0] + parts[1][0]
        tot_es = parts[0][1] + parts[1][1]
        for r in range(2):
            calls = iter([tot_bins, tot_es])

            def fake_all_reduce(t, _it=calls):
                t.copy_(next(_it))
            q, idx, ne, nc = vo.cosine_codebook_forward(flats[r], embed[r], cluster[r], training=True,
                                                        all_reduce=fake_all_reduce)
            assert torch.equal(idx.reshape(g[f'r{r}_ind{s}'].shape), _t(g[f'r{r}_ind{s}']))
            torch.testing.assert_close(ne, _t(g[f'r{r}_embed{s + 1}']), rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(nc, _t(g[f'r{r}_cluster{s + 1}']), rtol=1e-6, atol=1e-7)
            embed[r], cluster[r] = ne, nc


def big_case_codebook(g):
    """Regenerate the seeded initial codebook of a vq_big_* fixture (the reference module's own
    initialisation: kaiming_uniform -> l2norm under torch.manual_seed(seed)) and check it against the
    recorded digest, so the 16 MB codebook does not have to be stored."""
    torch.manual_seed(int(g['seed']))
    e = torch.empty(1, int(g['K']), int(g['D']))
    torch.nn.init.kaiming_uniform_(e)
    e = torch.nn.functional.normalize(e, p=2, dim=-1)[0]
    assert float(e.double().sum()) == pytest.approx(float(g['embed0_sum']), rel=1e-12, abs=1e-9)
    assert float(e.double().abs().sum()) == pytest.approx(float(g['embed0_abs']), rel=1e-12)
    assert torch.equal(e[7], _t(g['embed0_row7']))
    return e


@pytest.mark.parametrize('name', ['big_k1024', 'big_k16384', 'big_ortho'])
def test_vq_oracle_matches_reference_at_production_width(golden_dir, name):
    """D = 256, K in {1024, 16384}: the shapes the tensor-core search serves, recorded from the
    reference itself (oracle/make_golden.py: vq_big_case)."""
    g = _load(golden_dir, name)
    K = int(g['K'])
    embed = big_case_codebook(g)
    x = _t(g['x'])
    r = vo.vector_quantize_forward(x, embed, torch.zeros(K), training=True, commitment_weight=1.0)
    assert torch.equal(r['embed_ind'], _t(g['ind']))
    torch.testing.assert_close(r['quantize'], _t(g['q']), rtol=1e-6, atol=1e-7)
    rows = _t(g['rows'])
    torch.testing.assert_close(r['new_embed'][rows], _t(g['embed1_rows']), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(r['new_cluster_size'][rows], _t(g['cluster1_rows']), rtol=1e-6, atol=1e-7)
    assert float(r['new_embed'].double().sum()) == pytest.approx(float(g['embed1_sum']), rel=1e-6, abs=1e-4)
    assert float(r['new_cluster_size'].double().sum()) == pytest.approx(float(g['cluster1_sum']), rel=1e-6)
    loss = r['loss']
    if float(g['ortho']) > 0:
        # the reference measures the (1, K, D) codebook along dim 0: the whole updated codebook enters
        e1 = r['new_embed'].clone().requires_grad_(True)
        lo = vo.orthogonal_loss(e1) * float(g['ortho'])
        lo.backward()
        loss = loss + lo.detach()
        # the fixture's backward was (q * gq).sum() + 0.7 * loss
        torch.testing.assert_close(0.7 * e1.grad[rows], _t(g['gembed_rows']), rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(loss, _t(g['loss']), rtol=1e-5, atol=1e-8)
