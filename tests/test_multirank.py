"""N>1 path.  CPU (gloo, world size 2): the flat [bins | embed_sum] buffer and its single all-reduce
reproduce what the reference's two all-reduces produced in a real 2-rank gloo run of the reference
(tests/golden/vq_cos_ddp2.npz).  GPU (NCCL, needs 2 devices): the drop-in module end to end."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'vq_cos_ddp2.npz')


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _cpu_worker(rank, world, port, q):
    import torch.distributed as dist
    from favae_b200 import _dist
    from oracle import vq_oracle as vo
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = np.load(GOLDEN)
        K, D = int(g['K']), int(g['D'])
        embed = torch.from_numpy(g[f'r{rank}_embed0']); cluster = torch.zeros(K)
        ok = True
        for s in range(2):
            x = torch.from_numpy(g[f'r{rank}_x{s}'])
            flat = x.permute(0, 2, 3, 1).reshape(-1, D)
            idx, xn, en = vo.cosine_search(flat, embed)
            bins = torch.bincount(idx, minlength=K).float()
            esum = torch.zeros(K, D).index_add_(0, idx, xn)
            stats = _dist.pack_stats(bins, esum)
            assert stats.shape == (K * (D + 1),)
            _dist.all_reduce_stats(stats)                      # ONE collective
            rb, re = _dist.unpack_stats(stats, K, D)
            # EMA from the reduced statistics (l2_quantize.py:421-438)
            cluster = cluster * 0.8 + rb * 0.2
            zero = rb == 0
            en_new = torch.nn.functional.normalize(re / rb.masked_fill(zero, 1.0)[:, None], dim=-1)
            embed = embed * 0.8 + torch.where(zero[:, None], en, en_new) * 0.2
            ok &= torch.equal(idx.reshape(g[f'r{rank}_ind{s}'].shape), torch.from_numpy(g[f'r{rank}_ind{s}']))
            ok &= torch.allclose(embed, torch.from_numpy(g[f'r{rank}_embed{s + 1}']), rtol=1e-5, atol=1e-6)
            ok &= torch.allclose(cluster, torch.from_numpy(g[f'r{rank}_cluster{s + 1}']), rtol=1e-6, atol=1e-7)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_flat_stats_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == {0: True, 1: True}


def _gpu_worker(rank, world, port, q):
    import torch.distributed as dist
    from favae_b200 import VectorQuantize
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    import datetime
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank),
                            timeout=datetime.timedelta(seconds=120))
    try:
        g = np.load(GOLDEN)
        K, D = int(g['K']), int(g['D'])
        vq = VectorQuantize(dim=D, codebook_size=K, accept_image_fmap=True, use_cosine_sim=True,
                            sync_codebook=True).cuda().train()
        vq._codebook.embed.copy_(torch.from_numpy(g[f'r{rank}_embed0'])[None])
        ok = True
        for s in range(2):
            _, ind, _ = vq(torch.from_numpy(g[f'r{rank}_x{s}']).cuda())
            # the statistics all-reduce runs on a side stream and the EMA is deferred until somebody
            # looks at the codebook (the attribute accesses below)
            ok &= vq._codebook.__dict__.get('_pending') is not None
            ok &= torch.equal(ind.cpu(), torch.from_numpy(g[f'r{rank}_ind{s}']))
            ok &= torch.allclose(vq._codebook.embed[0].cpu(), torch.from_numpy(g[f'r{rank}_embed{s + 1}']), rtol=1e-4, atol=1e-6)
            ok &= torch.allclose(vq._codebook.cluster_size[0].cpu(), torch.from_numpy(g[f'r{rank}_cluster{s + 1}']), rtol=1e-5, atol=1e-7)
            ok &= vq._codebook.__dict__.get('_pending') is None
        # every rank must hold the SAME codebook, bit for bit (the reference's invariant, :419,427)
        mine = torch.cat([vq._codebook.embed.reshape(-1), vq._codebook.cluster_size.reshape(-1)])
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= torch.equal(lo, hi)
        sd = vq.state_dict()
        ok &= torch.equal(sd['_codebook.embed'], vq._codebook.embed)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sync_codebook_nccl_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run under gpurun --gpus 2)')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=200) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == {0: True, 1: True}
