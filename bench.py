#!/usr/bin/env python
"""Benchmark of the FA-VAE hot path (VQ search + spectrum losses) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload f16|f4|celeba]
                    [--impl reference] [--microbench]

One "step" = the hot path of one FA-VAE 256^2 training step over a per-GPU batch of B synthetic
images, i.e. exactly the calls that favae_scripts/train_favae.py:75-116 makes into
models/l2_quantize.py, losses/vqgan_losses.py and the five _gaussian_blur copies, in the order the
reference model issues them:

  stage 0  encoder features -> deferred learnable-sigma blurs (codec.py:284-309)
           quantizer(z) forward (search, gather, straight-through, commitment loss, code statistics
           [one side-stream all-reduce when N>1], EMA update)
           decoder features -> deferred blurs (codec.py:978-999)
           recon_ffl_loss(x, x_recon)
           recon_ffl_features_loss over the 4 feature levels (fused blur -> difference -> spectrum
           loss per level), backward of everything (to features, sigmas, latents)
  stage 1  quantizer(z) again in train mode under no_grad (the reference re-runs the encoder for the
           discriminator step, vqgan_fcm.py:138-146) -> second EMA update

Workloads (BASELINE.json): f16 = configs[2] (default; the configuration the metric is quoted on),
f4 = configs[3], celeba = configs[1].  The conv backbone, LPIPS and the discriminator are out of
scope: the feature maps they would produce are synthetic tensors.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU modules
(baseline/_ref: models/l2_quantize.py, losses/vqgan_losses.py, VQGANFCM._gaussian_blur; the absent
pip spectrum loss is the oracle port) on the host cores.  `--microbench` prints BASELINE configs[4]
(kernel sweep) as a table instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIGMA0 = 3.0
FFL_W, DSL_W, COMMIT_W = 1.0, 0.01, 1.0
IMG = 256
# levels: (C, H, W) of enc feature level i; the matching dec feature is de_feat[3 - i] before the
# wrapper reverses the list
WORKLOADS = {
    'f16': dict(
        name='BASELINE configs[2]: ImageNet f=16 256^2, codebook 16384x256 cosine-sim quantizer (stage 0 fwd+bwd '
             '+ stage 1), image FFL, non-pair-wise DSL gaussian_kernel 9 over 4 FCM feature levels',
        K=16384, dim=256, cdim=256, lat=16, ksize=9,
        levels=[(128, 256, 256), (512, 16, 16), (512, 16, 16), (256, 16, 16)]),
    'f4': dict(
        name='BASELINE configs[3]: ImageNet f=4 256^2 (train_favae_other_datasets_public.sh:26-30), 64x64 latent '
             'grid, embed_dim 3 -> codebook 8192x256 cosine-sim quantizer, image FFL, DSL gaussian_kernel 3',
        K=8192, dim=3, cdim=256, lat=64, ksize=3,
        levels=[(128, 256, 256), (512, 64, 64), (512, 64, 64), (3, 64, 64)]),
    'celeba': dict(
        name='BASELINE configs[1]: CelebA-HQ f=16 256^2 FCM(Res) + non-pair-wise DSL (train_favae_celeba.sh:55-60), '
             'codebook 1024x256 cosine-sim quantizer, image FFL, gaussian_kernel 9',
        K=1024, dim=256, cdim=256, lat=16, ksize=9,
        levels=[(128, 256, 256), (512, 16, 16), (512, 16, 16), (256, 16, 16)]),
}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p['hbm_gbs']), tf=float(p['bf16_tflops']),
                    tf_sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), src='measured')
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src='fallback')


def make_inputs(wl, batch, seed, device, pin=False):
    g = torch.Generator().manual_seed(seed)

    def r(*shape):
        t = torch.randn(*shape, generator=g)
        if pin:
            return t.pin_memory()
        return t.to(device)
    inp = {'z': r(batch, wl['dim'], wl['lat'], wl['lat']), 'x': r(batch, 3, IMG, IMG), 'x_recon': r(batch, 3, IMG, IMG)}
    inp['enc'] = [r(batch, c, h, w) for (c, h, w) in wl['levels']]
    inp['dec'] = [r(batch, c, h, w) for (c, h, w) in reversed(wl['levels'])]
    return inp


def input_bytes(inp):
    n = sum(t.numel() for t in (inp['z'], inp['x'], inp['x_recon']))
    n += sum(t.numel() for t in inp['enc']) + sum(t.numel() for t in inp['dec'])
    return 4 * n


def feature_elements(wl):
    return sum(c * h * w for (c, h, w) in wl['levels'])


class HotPath:
    """The reference-facing objects one training process holds (B200 implementation)."""

    def __init__(self, wl, device, sync_codebook):
        import favae_b200
        from favae_b200 import vqgan_losses
        torch.manual_seed(0)
        self.wl = wl
        self.fb = favae_b200
        self.vl = vqgan_losses
        self.vq = favae_b200.VectorQuantize(dim=wl['dim'], codebook_size=wl['K'], codebook_dim=wl['cdim'],
                                            accept_image_fmap=True, use_cosine_sim=True,
                                            sync_codebook=sync_codebook, commitment_weight=COMMIT_W).to(device).train()
        self.ffl = favae_b200.FocalFrequencyLoss(loss_weight=FFL_W, alpha=1.0)
        self.dsl = favae_b200.FocalFrequencyLoss(loss_weight=DSL_W, alpha=1.0)
        self.enc_sigmas = torch.nn.Parameter(torch.full((4,), SIGMA0, device=device))
        self.dec_sigmas = torch.nn.Parameter(torch.full((4,), SIGMA0, device=device))
        self.device = device

    def step(self, inp):
        # the patched reference's _gaussian_blur (favae_b200.patch_reference): a deferred blur
        blur = self.fb.lazy_gaussian_blur
        k = self.wl['ksize']
        # fresh leaves every step: in training these are activations, their gradients flow on to
        # the backbone instead of being accumulated into a persistent .grad
        z = inp['z'].detach().requires_grad_(True)
        x_recon = inp['x_recon'].detach().requires_grad_(True)
        enc = [t.detach().requires_grad_(True) for t in inp['enc']]
        dec = [t.detach().requires_grad_(True) for t in inp['dec']]
        # ---- stage 0, in the order of VQGANFCM.forward(stage=0): encoder (features + blurs),
        # quantizer, decoder (features + blurs), losses (train_favae.py:75-99)
        # (sigma_element(sigmas, i) is what the patched _gaussian_blur does for the reference's `self.sigmas[i]`)
        sig = self.fb.gaussian_blur.sigma_element
        enc_b = [blur(enc[i], sig(self.enc_sigmas, i), k) for i in range(4)]
        _, _, loss_q = self.vq(z)
        dec_b = [blur(dec[i], sig(self.dec_sigmas, i), k) for i in range(4)]
        # (1,)-shaped terms added up as train_favae.py:79-101 does
        loss = COMMIT_W * loss_q
        loss = loss + self.vl.recon_ffl_loss(self.ffl, inp['x'], x_recon)
        loss_dsl, _ = self.vl.recon_ffl_features_loss(self.dsl, enc_b, dec_b, self.device)
        loss = loss + loss_dsl
        loss.backward()
        # ---- stage 1
        with torch.no_grad():
            self.vq(z.detach())
        self.enc_sigmas.grad = None
        self.dec_sigmas.grad = None
        return loss.detach()


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own modules on the host cores
# --------------------------------------------------------------------------------------------
class ReferenceHotPath:
    """Same step through the UNMODIFIED reference modules shipped in baseline/_ref (VectorQuantize,
    recon_ffl_loss / recon_ffl_features_loss, VQGANFCM._gaussian_blur) on CPU.  The pip spectrum loss
    (absent offline) is oracle/ffl_oracle.py.  Falls back to the all-oracle port when the tree is
    missing (kind says which)."""

    def __init__(self, wl):
        from oracle import ffl_oracle as fo
        from oracle import reference_tree
        self.wl = wl
        torch.manual_seed(0)
        mods = reference_tree.import_reference()
        self.ffl = fo.FocalFrequencyLossOracle(loss_weight=FFL_W)
        self.dsl = fo.FocalFrequencyLossOracle(loss_weight=DSL_W)
        self.enc_sigmas = torch.full((4,), SIGMA0, requires_grad=True)
        self.dec_sigmas = torch.full((4,), SIGMA0, requires_grad=True)
        if mods is not None:
            l2q, vl, fcm = mods
            self.kind = 'reference+ffl-port'
            self.vq = l2q.VectorQuantize(dim=wl['dim'], codebook_size=wl['K'], codebook_dim=wl['cdim'],
                                         accept_image_fmap=True, use_cosine_sim=True,
                                         commitment_weight=COMMIT_W).train()
            self.vl = vl
            k = wl['ksize']
            V = fcm.VQGANFCM

            def shim(sig):
                s = types.SimpleNamespace(kernel_size=k, sigmas=sig, padding=[k // 2] * 4)
                s._get_gaussian_kernel1d = types.MethodType(V._get_gaussian_kernel1d, s)
                s._get_gaussian_kernel2d = types.MethodType(V._get_gaussian_kernel2d, s)
                return s
            se, sd = shim(self.enc_sigmas), shim(self.dec_sigmas)
            self.blur_e = lambda x, i: V._gaussian_blur(se, x, i, device='cpu')
            self.blur_d = lambda x, i: V._gaussian_blur(sd, x, i, device='cpu')
        else:
            from oracle import blur_oracle as bo
            from oracle import wrappers_oracle as wo
            import torch.nn.functional as F
            self.kind = 'port'
            e = torch.empty(wl['K'], wl['cdim'])
            torch.nn.init.kaiming_uniform_(e)
            self.state = dict(embed=F.normalize(e, dim=-1), cluster=torch.zeros(wl['K']))
            self.vq = None
            self.vl = types.SimpleNamespace(recon_ffl_loss=wo.recon_ffl_loss,
                                            recon_ffl_features_loss=lambda f, e_, d_, dev: wo.recon_ffl_features_loss(f, e_, d_))
            self.blur_e = lambda x, i: bo.gaussian_blur_reflect(x, self.enc_sigmas[i], wl['ksize'])
            self.blur_d = lambda x, i: bo.gaussian_blur_reflect(x, self.dec_sigmas[i], wl['ksize'])

    def _quantize(self, z, train_grad):
        if self.vq is not None:
            return self.vq(z)[2]
        from oracle import vq_oracle as vo
        r = vo.vector_quantize_forward(z.detach(), self.state['embed'], self.state['cluster'], training=True,
                                       commitment_weight=COMMIT_W)
        self.state['embed'], self.state['cluster'] = r['new_embed'], r['new_cluster_size']
        if not train_grad:
            return None
        flat = z.permute(0, 2, 3, 1).reshape(-1, z.shape[1])
        return ((r['q_flat'] - flat) ** 2).mean().reshape(1) * COMMIT_W

    def step(self, inp):
        z = inp['z'].clone().requires_grad_(True)
        x_recon = inp['x_recon'].clone().requires_grad_(True)
        enc = [t.clone().requires_grad_(True) for t in inp['enc']]
        dec = [t.clone().requires_grad_(True) for t in inp['dec']]
        enc_b = [self.blur_e(enc[i], i) for i in range(4)]
        loss_q = self._quantize(z, True)
        dec_b = [self.blur_d(dec[i], i) for i in range(4)]
        loss = COMMIT_W * loss_q.sum() + self.vl.recon_ffl_loss(self.ffl, inp['x'], x_recon)
        loss_dsl, _ = self.vl.recon_ffl_features_loss(self.dsl, enc_b, dec_b, 'cpu')
        (loss + loss_dsl.sum()).backward()
        with torch.no_grad():
            self._quantize(z.detach(), False)
        self.enc_sigmas.grad = None
        self.dec_sigmas.grad = None
        return float(loss)


def time_cpu(wl, batch, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    inp = make_inputs(wl, batch, 1234, 'cpu')
    hp = ReferenceHotPath(wl)
    for _ in range(warmup):
        hp.step(inp)
    t0 = time.perf_counter()
    for _ in range(steps):
        hp.step(inp)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, hp.kind


def bind_near_gpu(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of
    the end-to-end measurement are allocated on the NUMA node behind the GPU's PCIe root (with 8 ranks
    the H2D streams otherwise cross the socket interconnect).  Returns the CPU count, or None."""
    try:
        import pynvml as n
        n.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        ids = [v for v in vis.split(',') if v.strip().isdigit()]
        phys = int(ids[index]) if index < len(ids) else index
        h = n.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = n.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:                                    # noqa: BLE001 - placement is an optimisation only
        pass
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled from a thread
    every 10 ms (no process start-up latency: a 50-step run lasts ~0.2 s); `nvidia-smi -lms` is the
    fallback when the NVML binding is missing."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    BITS = {'sw_power_cap': 0x4, 'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.sm, self.mx, self.reasons, self.stop_flag, self.thread = None, [], None, set(), False, None
        self.power = []

    def _poll(self):
        n, h = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                self.power.append(n.nvmlDeviceGetPowerUsage(h) / 1e3)
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.reasons.update(k for k, bit in self.BITS.items() if r & bit)
            except Exception:                            # noqa: BLE001 - a failed sample is just skipped
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            ids = [v for v in vis.split(',') if v.strip().isdigit()]
            phys = int(ids[self.index]) if self.index < len(ids) else self.index
            h = n.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            self.nvml = (n, h)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:                                # noqa: BLE001 - fall back to nvidia-smi
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.mx,
                    'reasons': sorted(self.reasons), 'samples': len(sm), 'source': 'nvml, 10 ms period',
                    'power_w_max': max(self.power) if self.power else None}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm), 'source': 'nvidia-smi -lms 50'}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    batch = 1
    steps = max(1, min(args.steps, 20))           # bounded sample: well under a second of host work per step
    warm = min(args.warmup, 1)
    value, ms, kind = time_cpu(wl, batch, steps, warm)
    line = base_line(args, wl, world, value, ms)
    what = ('the reference\'s own models/l2_quantize.py, losses/vqgan_losses.py and VQGANFCM._gaussian_blur '
            '(baseline/_ref, unmodified) with the absent pip spectrum loss restated in oracle/ffl_oracle.py'
            if kind != 'port' else 'oracle/ torch-CPU port of the reference (baseline/_ref missing)')
    line.update({'impl': 'reference', 'dtype': 'f32', 'gpu_launches': 0, 'steps': steps, 'warmup': warm,
                 'cpu_baseline': {'value': value, 'unit': 'img/s', 'cores': os.cpu_count(), 'kind': kind,
                                  'sample': f'{steps} steps of the full hot-path step at batch {batch}: {what}, '
                                            f'all host threads'},
                 'e2e': {'value': value, 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})
    print(json.dumps(line), flush=True)


def base_line(args, wl, world, value, ms):
    mb = 8.0 * feature_elements(wl) / 1e6
    fam = 'f=4' if wl['lat'] == 64 else 'f=16'
    return {'metric': f'FA-VAE 256^2 {fam} hot-path (VQ search + spectrum losses) train img/s',
            'value': value, 'unit': 'img/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'per_gpu_batch': args.batch, 'global_batch': args.batch * world,
                       'parallelism': f'dp{world}', 'l2_policy': 'inputs larger than L2 '
                       f'({args.batch * mb:.0f} MB of feature maps per step)'}}


def check_rank_parity(hp, world, rank, device):
    """Every rank must end with the SAME codebook, bit for bit (the reference's invariant: identical
    all-reduced statistics, identical EMA, l2_quantize.py:419-438), and ranks 0/1 must reproduce the
    reference's own 2-rank run (tests/golden/vq_cos_ddp2.npz, recorded from a gloo run of the
    unmodified reference)."""
    import numpy as np
    import torch.distributed as dist
    cb = hp.vq._codebook
    mine = torch.cat([cb.embed.detach().reshape(-1), cb.cluster_size.reshape(-1)])
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        return f'codebooks differ across ranks ({int((lo != hi).sum())} elements)'
    path = os.path.join(ROOT, 'tests', 'golden', 'vq_cos_ddp2.npz')
    if not os.path.exists(path):
        return 'ok (golden replay skipped: fixture missing)'
    import favae_b200
    g = np.load(path)
    K, D = int(g['K']), int(g['D'])
    status = torch.zeros(1, device=device)
    vq = favae_b200.VectorQuantize(dim=D, codebook_size=K, accept_image_fmap=True, use_cosine_sim=True,
                                   sync_codebook=True).to(device).train()
    r = min(rank, 1)
    with torch.no_grad():
        vq._codebook.embed.copy_(torch.from_numpy(g[f'r{r}_embed0'])[None])
    bad = 0
    for s in range(2):
        x = torch.from_numpy(g[f'r{r}_x{s}']).to(device)
        if rank >= 2:
            # the fixture is a 2-rank run: ranks >= 2 join each statistics all-reduce with zeros, which
            # keeps the sums those of the reference's run
            stats = torch.zeros(K * (D + 1), device=device)
            dist.all_reduce(stats)
            continue
        _, ind, _ = vq(x)
        bad += int(not torch.equal(ind.cpu(), torch.from_numpy(g[f'r{r}_ind{s}'])))
        bad += int(not torch.allclose(vq._codebook.embed[0].cpu(), torch.from_numpy(g[f'r{r}_embed{s + 1}']),
                                      rtol=1e-4, atol=1e-6))
        bad += int(not torch.allclose(vq._codebook.cluster_size[0].cpu(),
                                      torch.from_numpy(g[f'r{r}_cluster{s + 1}']), rtol=1e-5, atol=1e-7))
    status += bad
    dist.all_reduce(status)
    return 'ok' if int(status) == 0 else f'golden 2-rank replay failed ({int(status)} checks)'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=32, help='images per GPU')
    ap.add_argument('--workload', default='f16', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='favae_b200', choices=['favae_b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-rooflines', action='store_true', help='skip the stand-alone kernel timings')
    ap.add_argument('--microbench', action='store_true', help='BASELINE configs[4]: kernel sweep table')
    ap.add_argument('--no-graph', dest='graph', action='store_false',
                    help='time eager steps.  Default: the step (forward, backward, both quantizer calls, the '
                         'statistics all-reduces when N > 1) is captured in ONE CUDA graph and replays are timed; '
                         'the per-kernel trace and the launch count are then taken from eager steps')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    wl = WORKLOADS[args.workload]

    if args.impl == 'reference':
        return run_reference(args, wl, rank, world)

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: favae_b200 has no CPU path')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if args.microbench:
        return microbench(device)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=device, timeout=datetime.timedelta(seconds=240))
    from favae_b200 import _lib
    pk = peaks()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, ms, kind = time_cpu(wl, 1, 2, 1)
        cpu = {'value': v, 'unit': 'img/s', 'cores': os.cpu_count(), 'kind': kind,
               'sample': '2 steps (after 1 warm-up) of the full hot-path step at batch 1 on the host cores '
                         '(see --impl reference for what runs), all host threads'}

    hp = HotPath(wl, device, sync_codebook=world > 1)
    inp = make_inputs(wl, args.batch, 1234 + rank, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        hp.step(inp)
    barrier()
    graph, graph_note = None, 'eager steps (--no-graph)'
    if args.graph:
        # the step has no host synchronisation and no data-dependent host control flow, so
        # forward + backward + both quantizer calls (and, for N > 1, the two side-stream all-reduces of the
        # code statistics) replay as one graph launch.  The capture must end with the side stream joined,
        # so the captured step applies the second EMA update at its end instead of deferring it into the
        # next step.
        eager_step = hp.step

        def closed_step(inp_):
            out = eager_step(inp_)
            hp.vq._codebook._flush()
            return out
        try:
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    closed_step(inp)
            torch.cuda.current_stream(device).wait_stream(side)
            barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_loss = closed_step(inp)
            hp.step = lambda _inp: (graph.replay(), graph_loss)[1]
            for _ in range(3):
                hp.step(inp)
            graph_note = 'whole step replayed as ONE CUDA graph (roofline_kernels / gpu_launches from eager steps)'
        except Exception as exc:                          # noqa: BLE001 - fall back to eager, and say so
            graph = None
            hp.step = eager_step
            graph_note = f'eager steps: CUDA graph capture failed ({type(exc).__name__}: {str(exc)[:120]})'
            print(graph_note, file=sys.stderr, flush=True)
            hp.vq._codebook.__dict__['_pending'] = None   # an event recorded inside the aborted capture is void
            torch.cuda.synchronize()
        ok = torch.tensor([1.0 if graph is not None else 0.0], device=device)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)     # all ranks replay, or none does
        if float(ok) == 0.0 and graph is not None:
            graph = None
            hp.step = eager_step
            graph_note = 'eager steps: CUDA graph capture failed on another rank'
        barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    traced = ['favae_ffl_forward', 'favae_blur_diff_forward', 'favae_blur_backward', 'favae_blur_backward_pair',
              'favae_blur_forward', 'favae_vq_search_tc']
    if rank == 0 and graph is None:
        _lib.start_trace(traced)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        hp.step(inp)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    trace = _lib.stop_trace() if rank == 0 and graph is None else {}
    clocks = sampler.stop() if rank == 0 else None
    trace_steps = args.steps
    if graph is not None:
        launches = (launches if launches else 0)
        hp.step = eager_step                       # per-kernel timings and the launch count: eager steps
        l0 = _lib.launch_count()
        _lib.start_trace(traced)
        trace_steps = 10
        for _ in range(trace_steps):
            hp.step(inp)
        trace = _lib.stop_trace()
        launches = (_lib.launch_count() - l0) * args.steps // trace_steps
    t = torch.tensor([ms_total], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = args.batch * world * 1e3 / ms_step

    rank_parity = None
    if world > 1:
        rank_parity = check_rank_parity(hp, world, rank, device)

    # ---- end to end: pinned host inputs -> H2D -> step -> loss.item()
    numa = bind_near_gpu(local)          # before the pinned allocation: first touch decides the NUMA node
    host = make_inputs(wl, args.batch, 4321 + rank, device, pin=True)
    copy_stream = torch.cuda.Stream(device)

    def upload():
        """H2D copy of one step's inputs from pinned host memory, on the copy stream."""
        with torch.cuda.stream(copy_stream):
            dev = {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else
                       [t.to(device, non_blocking=True) for t in v]) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dev, ev

    def e2e_run(n):
        """n steps, each with its own H2D copy and a D2H read of the loss; the copy of step i+1 is
        issued before step i computes (double buffering), as a training input pipeline would."""
        cur = torch.cuda.current_stream(device)
        dev, ev = upload()
        for i in range(n):
            cur.wait_event(ev)
            now = dev
            for v in now.values():
                for t_ in (v if isinstance(v, list) else [v]):
                    t_.record_stream(cur)
            if i + 1 < n:
                dev, ev = upload()
            hp.step(now).item()

    e2e_run(2)
    barrier()
    n_e2e = max(2, min(args.steps, 20))
    e0.record()
    e2e_run(n_e2e)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = args.batch * world * 1e3 / (float(t) / n_e2e)

    if rank == 0:
        line = base_line(args, wl, world, value, ms_step)
        groups = kernel_groups(trace, wl, args, pk, trace_steps)
        hbm_groups = [g_ for g_ in groups if 'bytes_per_launch' in g_]
        dom = max(hbm_groups, key=lambda g_: g_['ms_per_step']) if hbm_groups else None
        step_bytes = 16.0 * feature_elements(wl) * args.batch
        step_gbs = step_bytes / (ms_step * 1e-3) / 1e9
        line.update({
            'e2e': {'value': e2e_value, 'unit': 'img/s', 'h2d_bytes_per_step': input_bytes(host),
                    'd2h_bytes_per_step': 4, 'steps': n_e2e, 'host_cpus_bound': numa},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': None if dom is None else {
                'kernel': dom['kernel'], 'bound': 'hbm', 'achieved': dom['achieved'], 'peak': pk['hbm'],
                'unit': 'GB/s', 'frac': dom['achieved'] / pk['hbm'], 'traffic': dom.get('traffic'),
                'traffic_source': dom.get('traffic_source'), 'peak_source': pk['src'],
                'algorithmic_bytes_per_launch': dom['bytes_per_launch'], 'ms_per_launch': dom['ms_per_launch'],
                'launches_per_step': dom['launches_per_step'], 'ms_per_step': dom['ms_per_step'],
                'timing': 'CUDA events around every launch of this entry point inside the timed steps '
                          '(favae_b200._lib.start_trace)'},
            'roofline_kernels': groups,
            'roofline_step': {'bound': 'hbm', 'algorithmic_bytes_per_step': step_bytes, 'achieved': step_gbs,
                              'peak': pk['hbm'], 'unit': 'GB/s', 'frac': step_gbs / pk['hbm'],
                              'note': 'SURVEY 8(d): 16 B per feature element (read enc/dec, write both gradients), '
                                      'blur fused = 0 extra bytes; whole step incl. quantizer and image FFL'},
            'cpu_baseline': cpu,
        })
        line['config']['execution'] = graph_note
        if rank_parity is not None:
            line['rank_parity'] = rank_parity
        if not args.no_kernel_rooflines:
            n_lat = args.batch * wl['lat'] * wl['lat']
            vq_ms = time_vq_search(n_lat, wl['K'], wl['cdim'], device)
            big_ms = time_vq_search(262144, wl['K'], wl['cdim'], device)

            def vq_line(n, ms):
                flops = 2.0 * n * wl['K'] * wl['cdim']
                hbm_ms = (4.0 * (n * wl['cdim'] + wl['K'] * wl['cdim']) + 8.0 * n) / (pk['hbm'] * 1e9) * 1e3
                tc_ms = flops / (pk['tf'] * 1e12) * 1e3
                return {'n_latents': n, 'ms_per_call': ms, 'achieved': flops / (ms * 1e-3) / 1e12,
                        'frac': flops / (ms * 1e-3) / 1e12 / pk['tf'], 'hbm_bound_ms': hbm_ms,
                        'roofline_ms': max(hbm_ms, tc_ms), 'frac_of_roofline': max(hbm_ms, tc_ms) / ms}
            w_ = vq_line(n_lat, vq_ms)
            line['roofline_vq'] = {
                'kernel': f'favae_vq_search_tc: tcgen05 cta_group::2 search + exact re-score + fallback, '
                          f'{wl["K"]} x {wl["cdim"]} codebook; algorithmic 2*N*K*D flops',
                'bound': 'tensor', 'peak': pk['tf'], 'unit': 'TFLOP/s', 'peak_source': pk['src'],
                'achieved': w_['achieved'], 'frac': w_['frac'],           # quoted at the WORKLOAD's N
                'workload': w_, 'microbench_large_n': vq_line(262144, big_ms)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: the CUDA graph holds captured NCCL kernels; destroying the process group with the
        # graph still alive left the NCCL watchdog of one rank stuck (observed at N = 2).  Drop the graph
        # first, agree that everybody is done, and leave without running destructors.
        failed = rank_parity is not None and not rank_parity.startswith('ok')
        hp.step = None
        if graph is not None:
            graph.reset()
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        if failed and rank == 0:
            print(f'rank parity check failed: {rank_parity}', file=sys.stderr, flush=True)
        os._exit(1 if failed else 0)


# ncu --set full captures of this round (profiles/ncu_r2_summary.md, last section): dram__bytes_read + write per element
NCU_TRAFFIC = {
    'blur_backward': (12.00, 'ncu --set full, blur_adjsig_kernel<9,64>: 561.9 MB read + 242.9 MB written for 1024 maps '
                             'of 256^2 = 12.00 B/element (profiles/ncu_r2_blur_bwd_raw.csv)'),
    'blur_diff': (12.30, 'ncu --set full, blur_diff_kernel<9,64>: 586.5 MB read + 239.2 MB written for 1024 maps of '
                         '256^2 = 12.30 B/element (profiles/ncu_r2b_diff_raw.csv)'),
    'ffl_diff': (7.30, 'ncu --set full, ffl_kernel<256> single-input form: 268.7 MB read + 221.4 MB written for 1024 maps '
                       'of 256^2 = 7.30 B/element, part of the last gradient rows still in L2 when the kernel ends '
                       '(profiles/ncu_r2b_ffldiff_raw.csv)'),
    'blur_pair': (19.64, 'ncu --set full, blur_adjsig_pair_kernel<9,128>: 813.7 MB read + 504.0 MB written for 1024 maps '
                         'of 256^2 = 19.64 B/element (profiles/ncu_r2b_pair_raw.csv)'),
    'ffl2': (15.48, 'ncu --set full, ffl_kernel<256> two-input form: 15.48 B/element (profiles/ncu_r1_summary.md)')}


def kernel_groups(trace, wl, args, pk, steps):
    """Per kernel family of the level-0 maps: launches, mean device time per launch (CUDA events
    inside the timed steps), algorithmic bytes (SURVEY 8d per-element figures), achieved GB/s."""
    steps = max(steps, 1)
    out = []

    def add(name, kernel, rows, bytes_per_launch):
        if not rows:
            return None
        ms = sum(rows) / len(rows)
        g = {'name': name, 'kernel': kernel, 'launches_per_step': len(rows) / steps, 'ms_per_launch': ms,
             'ms_per_step': sum(rows) / steps, 'bytes_per_launch': bytes_per_launch,
             'achieved': bytes_per_launch / (ms * 1e-3) / 1e9, 'frac': bytes_per_launch / (ms * 1e-3) / 1e9 / pk['hbm']}
        out.append(g)
        return g
    c0, h0, w0 = wl['levels'][0]
    e_l0 = args.batch * c0 * h0 * w0
    maps_l0 = args.batch * c0
    # favae_ffl_forward(pred, target, maps, h, w, ...): args[1] is None for the single-input (difference) form
    ffl = trace.get('favae_ffl_forward', [])
    l0_diff = [ms for ms, a in ffl if a[2] == maps_l0 and a[3] == h0 and a[1] is None]
    l0_two = [ms for ms, a in ffl if a[2] == maps_l0 and a[3] == h0 and a[1] is not None]
    g = add('ffl_level0_difference', f'ffl_kernel<{h0}> on the level-0 difference map: read d, write G in place '
            '(8 B/element)', l0_diff, 8.0 * e_l0)
    if g and h0 == 256:
        g['traffic'] = NCU_TRAFFIC['ffl_diff'][0] * e_l0
        g['traffic_source'] = NCU_TRAFFIC['ffl_diff'][1]
    g = add('ffl_level0_two_inputs', f'ffl_kernel<{h0}> (level-0 spectrum loss, read pred/target, write both '
            'gradients: 16 B/element)', l0_two, 16.0 * e_l0)
    if g:
        g['traffic'] = 15.48 * e_l0
        g['traffic_source'] = NCU_TRAFFIC['ffl2'][1]
    bd = [ms for ms, a in trace.get('favae_blur_diff_forward', []) if a[2] == maps_l0 and a[3] == h0]
    g = add('blur_difference_level0', f'blur_diff_kernel<{wl["ksize"]},64>: d = B(dec) - B(enc), read enc/dec, write d '
            '(12 B/element)', bd, 12.0 * e_l0)
    if g and wl['ksize'] == 9:
        g['traffic'] = NCU_TRAFFIC['blur_diff'][0] * e_l0
        g['traffic_source'] = NCU_TRAFFIC['blur_diff'][1]
    # favae_blur_backward(gy, x, maps, h, w, ks, sigma, scale, gx, gsigma, partials, stream)
    bb = [ms for ms, a in trace.get('favae_blur_backward', []) if a[2] == maps_l0 and a[3] == h0]
    g = add('blur_adjoint_sigma_level0', f'blur_adjsig_kernel<{wl["ksize"]},64>: adjoint + sigma gradient, read G and x, '
            'write gx (12 B/element)', bb, 12.0 * e_l0)
    if g and wl['ksize'] == 9:
        g['traffic'] = NCU_TRAFFIC['blur_backward'][0] * e_l0
        g['traffic_source'] = NCU_TRAFFIC['blur_backward'][1]
    # favae_blur_backward_pair(gy, x_enc, x_dec, maps, h, w, ks, ...)
    bp = [ms for ms, a in trace.get('favae_blur_backward_pair', []) if a[3] == maps_l0 and a[4] == h0]
    g = add('blur_adjoint_sigma_pair_level0', f'blur_adjsig_pair_kernel<{wl["ksize"]},{128 if h0 >= 256 else 64}>: both sides '
            'of the level in one pass over G: read G, enc, dec, write both gradients (20 B/element)', bp, 20.0 * e_l0)
    if g and wl['ksize'] == 9 and h0 == 256:
        g['traffic'] = NCU_TRAFFIC['blur_pair'][0] * e_l0
        g['traffic_source'] = NCU_TRAFFIC['blur_pair'][1]
    bf = [ms for ms, a in trace.get('favae_blur_forward', []) if a[1] == maps_l0 and a[2] == h0]
    add('blur_forward_level0', 'blur_fast_kernel forward (8 B/element)', bf, 8.0 * e_l0)
    vq = [ms for ms, a in trace.get('favae_vq_search_tc', [])]
    if vq:
        out.append({'name': 'vq_search_tc', 'kernel': 'favae_vq_search_tc (search + re-score + fallback)',
                    'launches_per_step': len(vq) / steps, 'ms_per_launch': sum(vq) / len(vq),
                    'ms_per_step': sum(vq) / steps, 'bound': 'tensor'})
    return out


def time_vq_search(n_lat, k_codes, dim, device, iters=5):
    """Device time of favae_vq_search_tc (tensor-core search + exact re-score + fallback) on n_lat
    synthetic latents against a k_codes x dim codebook, inputs prepared once."""
    from favae_b200 import _lib
    x = torch.randn(n_lat, dim, device=device)
    e = torch.nn.functional.normalize(torch.randn(k_codes, dim, device=device), dim=-1)
    xn = torch.empty(n_lat, dim, device=device); xh = torch.empty(n_lat, dim, device=device, dtype=torch.float16)
    en = torch.empty(k_codes, dim, device=device); eh = torch.empty(k_codes, dim, device=device, dtype=torch.float16)
    st = _lib.stream()
    _lib.call('favae_vq_prepare_rows', x.data_ptr(), n_lat, dim, 1, 1, xn.data_ptr(), xh.data_ptr(), None, st)
    _lib.call('favae_vq_prepare_rows', e.data_ptr(), k_codes, dim, 1, 1, en.data_ptr(), eh.data_ptr(), None, st)
    nbytes = _lib.load().favae_vq_search_tc_workspace_bytes(n_lat, k_codes, dim)
    ws = torch.empty(nbytes, device=device, dtype=torch.uint8)
    idx = torch.empty(n_lat, device=device, dtype=torch.int64)
    keys = torch.empty(n_lat, device=device, dtype=torch.int64)

    def run():
        _lib.call('favae_vq_search_tc', xh.data_ptr(), eh.data_ptr(), xn.data_ptr(), en.data_ptr(), n_lat, k_codes,
                  dim, ws.data_ptr(), nbytes, keys.data_ptr(), idx.data_ptr(), _lib.stream())
    for _ in range(20):
        run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        run()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def microbench(device):
    """BASELINE configs[4]: VQ search sweep over N latents x K in {1024, 8192, 16384} codes and batched
    2D-FFT spectrum loss over 64^2..512^2 maps, each with its roofline fraction and the reference's
    CPU path timed beside it (bounded samples; the sample is named per row)."""
    import favae_b200
    from oracle import ffl_oracle as fo
    pk = peaks()
    torch.set_num_threads(os.cpu_count() or 1)
    rows = []
    print(f'# microbench (BASELINE configs[4]); peaks: HBM {pk["hbm"]:.0f} GB/s, bf16 {pk["tf"]:.0f} TFLOP/s ({pk["src"]}); '
          f'host cores {os.cpu_count()}')
    print(f'# VQ search: favae_vq_search_tc end to end (tcgen05 search + exact re-score + fallback), D = 256')
    print(f'{"K":>6} {"N":>7} {"ms":>9} {"TFLOP/s":>9} {"of tensor peak":>14} {"of roofline":>11} {"cpu ms":>10} {"gpu/cpu":>9}  cpu sample')
    D = 256
    for K in (1024, 8192, 16384):
        e_cpu = torch.nn.functional.normalize(torch.randn(K, D), dim=-1)
        for N in (256, 512, 2048, 4096, 16384, 65536, 262144):
            ms = time_vq_search(N, K, D, device, iters=10 if N <= 65536 else 3)
            flops = 2.0 * N * K * D
            tf = flops / (ms * 1e-3) / 1e12
            hbm_ms = (4.0 * (N * D + K * D) + 8.0 * N) / (pk['hbm'] * 1e9) * 1e3
            tc_ms = flops / (pk['tf'] * 1e12) * 1e3
            # reference CPU path (l2_quantize.py:403-411): l2norm both, einsum, argmax -- on <= 4096 rows
            ns = min(N, 4096)
            xs = torch.randn(ns, D)
            best = 1e30
            for _ in range(2):
                t0 = time.perf_counter()
                xn = torch.nn.functional.normalize(xs, dim=-1)
                en = torch.nn.functional.normalize(e_cpu, dim=-1)
                torch.einsum('h n d, h c d -> h n c', xn[None], en[None]).argmax(dim=-1)
                best = min(best, time.perf_counter() - t0)
            cpu_ms = best * 1e3 * (N / ns)
            rows.append({'kind': 'vq_search', 'K': K, 'N': N, 'ms': ms, 'tflops': tf, 'frac_tensor': tf / pk['tf'],
                         'frac_roofline': max(hbm_ms, tc_ms) / ms, 'cpu_ms': cpu_ms, 'cpu_rows_timed': ns})
            print(f'{K:6d} {N:7d} {ms:9.4f} {tf:9.1f} {tf / pk["tf"]:14.3f} {max(hbm_ms, tc_ms) / ms:11.3f} {cpu_ms:10.2f} '
                  f'{cpu_ms / ms:9.0f}  {ns} rows' + (' (scaled)' if ns < N else ''))
    print('# spectrum loss: favae_ffl_forward, loss + both gradients (16 B/element), one fused kernel per call (C ABI, CUDA events)')
    print(f'{"B":>3} {"C":>4} {"H":>4} {"ms":>9} {"GB/s":>8} {"of HBM peak":>11} {"cpu ms":>10} {"gpu/cpu":>9}  cpu sample')
    from favae_b200 import _lib
    for (C, H) in ((128, 64), (128, 128), (128, 256), (32, 512), (3, 256), (512, 16), (512, 64)):
        for B in (1, 8, 32):
            if B * C * H * H * 4 > 3 << 30:
                continue
            p = torch.randn(B, C, H, H, device=device)
            t = torch.randn(B, C, H, H, device=device)
            gp, gt = torch.empty_like(p), torch.empty_like(t)
            ml = torch.empty(B * C, device=device)
            gs = 2 * 0.01 / p.numel()

            def run():
                _lib.call('favae_ffl_forward', p.data_ptr(), t.data_ptr(), B * C, H, H, 1.0, 0, gs, ml.data_ptr(),
                          gp.data_ptr(), gt.data_ptr(), None, None, _lib.stream())
            for _ in range(5):
                run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(20):
                run()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            gbs = 16.0 * B * C * H * H / (ms * 1e-3) / 1e9
            cpu_ms = None
            sample = ''
            if B == 1:
                cs = min(C, 16)                       # bounded: <= 16 maps of the CPU restatement, scaled
                pc = torch.randn(1, cs, H, H, requires_grad=True); tcpu = torch.randn(1, cs, H, H, requires_grad=True)
                t0 = time.perf_counter()
                fo.focal_frequency_loss(pc, tcpu, loss_weight=0.01).backward()
                cpu_ms = (time.perf_counter() - t0) * 1e3 * (C / cs)
                sample = f'{cs} maps' + (' (scaled)' if cs < C else '')
            rows.append({'kind': 'spectrum_loss', 'B': B, 'C': C, 'H': H, 'ms': ms, 'gbs': gbs, 'frac_hbm': gbs / pk['hbm'],
                         'cpu_ms': cpu_ms})
            print(f'{B:3d} {C:4d} {H:4d} {ms:9.4f} {gbs:8.0f} {gbs / pk["hbm"]:11.3f} '
                  + (f'{cpu_ms:10.1f} {cpu_ms / ms:9.0f}  {sample}' if cpu_ms is not None else f'{"-":>10} {"-":>9}'))
            del p, t, gp, gt
    print(json.dumps({'microbench': rows}))


if __name__ == '__main__':
    main()
