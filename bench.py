#!/usr/bin/env python
"""Benchmark of the FA-VAE hot path (VQ search + spectrum losses) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = the hot path of one FA-VAE f=16 256^2 training step over a per-GPU batch of B
synthetic images (BASELINE.json configs[2]: codebook 16384 x 256, cosine-sim quantizer,
non-pair-wise DSL with gaussian_kernel 9, image-level FFL), i.e. exactly the calls that
favae_scripts/train_favae.py:75-116 makes into models/l2_quantize.py and losses/vqgan_losses.py:

  stage 0  quantizer(z) forward+backward (search, gather, straight-through, commitment loss,
           code statistics [one all-reduce when N>1], EMA update)
           recon_ffl_loss(x, x_recon)                         fwd+bwd
           8 learnable-sigma 9x9 blurs of the FCM features    fwd+bwd (to features and sigmas)
           recon_ffl_features_loss over the 4 feature levels  fwd+bwd
  stage 1  quantizer(z) again in train mode under no_grad (the reference re-runs the encoder
           for the discriminator step, vqgan_fcm.py:138-146) -> second EMA update

The conv backbone, LPIPS and the discriminator are out of scope and are
not in the step: the feature maps they would produce are synthetic tensors.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of the same step
(the reference is PyTorch code; its hot path restated in oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_CODES, DIM, KSIZE, SIGMA0 = 16384, 256, 9, 3.0
FFL_W, DSL_W, COMMIT_W = 1.0, 0.01, 1.0
IMG = 256
# (C, H, W) of enc feature level i; the matching dec feature is de_feat[3 - i] before the
# wrapper reverses the list (f = 16)
LEVELS = [(128, 256, 256), (512, 16, 16), (512, 16, 16), (256, 16, 16)]


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p['hbm_gbs']), tf=float(p['bf16_tflops']),
                    tf_sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), src='measured')
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src='fallback')


def make_inputs(batch, seed, device, pin=False):
    g = torch.Generator().manual_seed(seed)
    def r(*shape):
        t = torch.randn(*shape, generator=g)
        if pin:
            return t.pin_memory()
        return t.to(device)
    inp = {'z': r(batch, DIM, 16, 16), 'x': r(batch, 3, IMG, IMG), 'x_recon': r(batch, 3, IMG, IMG)}
    inp['enc'] = [r(batch, c, h, w) for (c, h, w) in LEVELS]
    inp['dec'] = [r(batch, c, h, w) for (c, h, w) in reversed(LEVELS)]
    return inp


def input_bytes(inp):
    n = sum(t.numel() for t in (inp['z'], inp['x'], inp['x_recon']))
    n += sum(t.numel() for t in inp['enc']) + sum(t.numel() for t in inp['dec'])
    return 4 * n


class HotPath:
    """The reference-facing objects one training process holds (B200 implementation)."""

    def __init__(self, device, sync_codebook):
        import favae_b200
        from favae_b200 import vqgan_losses
        torch.manual_seed(0)
        self.fb = favae_b200
        self.vl = vqgan_losses
        self.vq = favae_b200.VectorQuantize(dim=DIM, codebook_size=K_CODES, accept_image_fmap=True,
                                            use_cosine_sim=True, sync_codebook=sync_codebook,
                                            commitment_weight=COMMIT_W).to(device).train()
        self.ffl = favae_b200.FocalFrequencyLoss(loss_weight=FFL_W, alpha=1.0)
        self.dsl = favae_b200.FocalFrequencyLoss(loss_weight=DSL_W, alpha=1.0)
        self.enc_sigmas = torch.nn.Parameter(torch.full((4,), SIGMA0, device=device))
        self.dec_sigmas = torch.nn.Parameter(torch.full((4,), SIGMA0, device=device))
        self.device = device
        self.l0_events = None

    def step(self, inp):
        blur = self.fb.gaussian_blur_reflect
        # fresh leaves every step: in training these are activations, their gradients flow on to
        # the backbone instead of being accumulated into a persistent .grad
        z = inp['z'].detach().requires_grad_(True)
        x_recon = inp['x_recon'].detach().requires_grad_(True)
        enc = [t.detach().requires_grad_(True) for t in inp['enc']]
        dec = [t.detach().requires_grad_(True) for t in inp['dec']]
        # ---- stage 0
        _, _, loss_q = self.vq(z)
        loss = COMMIT_W * loss_q.sum()
        loss = loss + self.vl.recon_ffl_loss(self.ffl, inp['x'], x_recon)
        enc_b = [blur(enc[i], self.enc_sigmas[i], KSIZE) for i in range(4)]
        dec_b = [blur(dec[i], self.dec_sigmas[i], KSIZE) for i in range(4)]
        ev = self.l0_events
        if ev is not None:
            # bracket the level-0 spectrum-loss call (the dominant kernel) with CUDA events
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with self.fb.focal_frequency_loss.expected_upstream_scale(0.25):   # as the wrapper does
                a.record()
                l0 = self.dsl(dec_b[3], enc_b[0])
                b.record()
                ev.append((a, b))
                rest = [self.dsl(dec_b[2 - i], enc_b[1 + i]) for i in range(3)]
            loss_dsl = (l0 + rest[0] + rest[1] + rest[2]).reshape(1) * 0.25
        else:
            loss_dsl, _ = self.vl.recon_ffl_features_loss(self.dsl, enc_b, dec_b, self.device)
        loss = loss + loss_dsl.sum()
        loss.backward()
        # ---- stage 1
        with torch.no_grad():
            self.vq(z.detach())
        self.enc_sigmas.grad = None
        self.dec_sigmas.grad = None
        return loss.detach()


def cpu_step(inp, state):
    """The same step on the CPU oracle (oracle/ restates the reference; see its headers)."""
    from oracle import blur_oracle as bo
    from oracle import ffl_oracle as fo
    from oracle import vq_oracle as vo
    from oracle import wrappers_oracle as wo
    z = inp['z'].clone().requires_grad_(True)
    x_recon = inp['x_recon'].clone().requires_grad_(True)
    enc = [t.clone().requires_grad_(True) for t in inp['enc']]
    dec = [t.clone().requires_grad_(True) for t in inp['dec']]
    es = state['enc_sigmas'].clone().requires_grad_(True)
    ds = state['dec_sigmas'].clone().requires_grad_(True)
    r = vo.vector_quantize_forward(z.detach(), state['embed'], state['cluster'], training=True,
                                   commitment_weight=COMMIT_W)
    flat = z.permute(0, 2, 3, 1).reshape(-1, DIM)
    loss_q = ((r['q_flat'] - flat) ** 2).mean() * COMMIT_W
    ffl = fo.FocalFrequencyLossOracle(loss_weight=FFL_W)
    dsl = fo.FocalFrequencyLossOracle(loss_weight=DSL_W)
    loss = COMMIT_W * loss_q + wo.recon_ffl_loss(ffl, inp['x'], x_recon)
    enc_b = [bo.gaussian_blur_reflect(enc[i], es[i], KSIZE) for i in range(4)]
    dec_b = [bo.gaussian_blur_reflect(dec[i], ds[i], KSIZE) for i in range(4)]
    loss_dsl, _ = wo.recon_ffl_features_loss(dsl, enc_b, dec_b)
    (loss + loss_dsl.sum()).backward()
    state['embed'], state['cluster'] = r['new_embed'], r['new_cluster_size']
    r2 = vo.vector_quantize_forward(z.detach(), state['embed'], state['cluster'], training=True,
                                    commitment_weight=COMMIT_W)
    state['embed'], state['cluster'] = r2['new_embed'], r2['new_cluster_size']
    return float(loss)


def cpu_state():
    import torch.nn.functional as F
    torch.manual_seed(0)
    e = torch.empty(K_CODES, DIM)
    torch.nn.init.kaiming_uniform_(e)
    return dict(embed=F.normalize(e, dim=-1), cluster=torch.zeros(K_CODES),
                enc_sigmas=torch.full((4,), SIGMA0), dec_sigmas=torch.full((4,), SIGMA0))


def time_cpu(batch, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    inp = make_inputs(batch, 1234, 'cpu')
    st = cpu_state()
    for _ in range(warmup):
        cpu_step(inp, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(inp, st)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


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

    def _poll(self):
        n, h = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
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
                    'reasons': sorted(self.reasons), 'samples': len(sm), 'source': 'nvml, 10 ms period'}
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    batch = 1
    steps = max(1, min(args.steps, 20))           # bounded sample: ~0.35 s of host work per step
    value, ms = time_cpu(batch, steps, min(args.warmup, 1))
    line = base_line(args, world, value, ms, impl='reference')
    line.update({'impl': 'reference', 'dtype': 'f32', 'gpu_launches': 0, 'steps': steps, 'warmup': min(args.warmup, 1),
                 'cpu_baseline': {'value': value, 'unit': 'img/s', 'cores': os.cpu_count(), 'kind': 'port',
                                  'sample': f'{steps} steps of the full hot-path step at batch {batch} '
                                            f'(oracle/ torch-CPU port of the reference, all host threads)'},
                 'e2e': {'value': value, 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})
    print(json.dumps(line), flush=True)


def base_line(args, world, value, ms, impl='favae_b200'):
    return {'metric': 'FA-VAE 256^2 f=16 hot-path (VQ search + spectrum losses) train img/s',
            'value': value, 'unit': 'img/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'BASELINE configs[2]: ImageNet f=16 256^2, codebook 16384x256 cosine-sim '
                                   'quantizer (stage 0 fwd+bwd + stage 1), image FFL, non-pair-wise DSL '
                                   'gaussian_kernel 9 over 4 FCM feature levels',
                       'per_gpu_batch': args.batch, 'global_batch': args.batch * world,
                       'parallelism': f'dp{world}', 'l2_policy': 'inputs larger than L2 '
                       f'({args.batch * 71.4:.0f} MB of feature maps per step)'}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=32, help='images per GPU')
    ap.add_argument('--impl', default='favae_b200', choices=['favae_b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        return run_reference(args, rank, world)

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: favae_b200 has no CPU path')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    from favae_b200 import _lib
    pk = peaks()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, ms = time_cpu(1, 2, 1)
        cpu = {'value': v, 'unit': 'img/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': '2 steps (after 1 warm-up) of the full hot-path step at batch 1 on the oracle/ '
                         'torch-CPU port of the reference, all host threads'}

    hp = HotPath(device, sync_codebook=world > 1)
    inp = make_inputs(args.batch, 1234 + rank, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        hp.step(inp)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    hp.l0_events = []
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
    l0_ms = sum(a.elapsed_time(b) for a, b in hp.l0_events) / max(len(hp.l0_events), 1)
    hp.l0_events = None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = args.batch * world * 1e3 / ms_step

    # ---- end to end: pinned host inputs -> H2D -> step -> loss.item()
    numa = bind_near_gpu(local)          # before the pinned allocation: first touch decides the NUMA node
    host = make_inputs(args.batch, 4321 + rank, device, pin=True)
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

    # ---- kernel-level rooflines measured alone (burst peaks): VQ search, level-0 blur
    n_lat = args.batch * 256
    vq_ms = time_vq_search(n_lat, device)
    big_n = 262144
    vq_big_ms = time_vq_search(big_n, device)
    blur_ms = time_blur(args.batch, device)

    def vq_line(n, ms):
        flops = 2.0 * n * K_CODES * DIM
        return {'n_latents': n, 'ms_per_call': ms, 'achieved': flops / (ms * 1e-3) / 1e12,
                'frac': flops / (ms * 1e-3) / 1e12 / pk['tf'],
                'hbm_bound_ms': (4.0 * (n * DIM + K_CODES * DIM) + 8.0 * n) / (pk['hbm'] * 1e9) * 1e3}

    if rank == 0:
        e_l0 = args.batch * 128 * 256 * 256
        achieved = 16.0 * e_l0 / (l0_ms * 1e-3) / 1e9
        line = base_line(args, world, value, ms_step)
        line.update({
            'e2e': {'value': e2e_value, 'unit': 'img/s', 'h2d_bytes_per_step': input_bytes(host),
                    'd2h_bytes_per_step': 4, 'steps': n_e2e, 'host_cpus_bound': numa},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'kernel': 'ffl_kernel<256> (level-0 DSL spectrum loss, 128x256x256 maps per image)',
                         'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm'], 'unit': 'GB/s',
                         'frac': achieved / pk['hbm'], 'traffic': 15.48 * e_l0, 'peak_source': pk['src'],
                         'traffic_source': 'ncu --set full dram__bytes_read+write = 15.48 B/element (gpurun_out/prof_ffl256_r1i; part of the last gradient writes is still in L2 when the kernel ends) (profiles/ncu_r1_summary.md)',
                         'algorithmic_bytes_per_launch': 16.0 * e_l0, 'ms_per_launch': l0_ms},
            'roofline_vq': {'kernel': 'favae_vq_search_tc: tcgen05 cta_group::2 search + exact re-score + fallback, '
                                      '16384 x 256 codebook; algorithmic 2*N*K*D flops',
                            'bound': 'tensor', 'peak': pk['tf'], 'unit': 'TFLOP/s', 'peak_source': pk['src'],
                            'workload': vq_line(n_lat, vq_ms), 'microbench_large_n': vq_line(big_n, vq_big_ms),
                            'achieved': vq_line(big_n, vq_big_ms)['achieved'], 'frac': vq_line(big_n, vq_big_ms)['frac']},
            'roofline_blur': {'kernel': 'blur_fast_kernel<9,32,*> (forward, adjoint) + blur_sigma_kernel<9,32> on the level-0 maps', 'bound': 'hbm',
                              'peak': pk['hbm'], 'unit': 'GB/s',
                              'forward': {'ms': blur_ms['fwd'], 'achieved': 8.0 * e_l0 / (blur_ms['fwd'] * 1e-3) / 1e9,
                                          'frac': 8.0 * e_l0 / (blur_ms['fwd'] * 1e-3) / 1e9 / pk['hbm'],
                                          'algorithmic_bytes_per_element': 8},
                              'backward_with_sigma': {'ms': blur_ms['bwd'],
                                                      'achieved': 12.0 * e_l0 / (blur_ms['bwd'] * 1e-3) / 1e9,
                                                      'frac': 12.0 * e_l0 / (blur_ms['bwd'] * 1e-3) / 1e9 / pk['hbm'],
                                                      'algorithmic_bytes_per_element': 12}},
            'cpu_baseline': cpu,
        })
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_vq_search(n_lat, device, iters=5):
    """Device time of favae_vq_search_tc (tensor-core search + exact re-score + fallback) on n_lat
    synthetic latents against a 16384 x 256 codebook, inputs prepared once."""
    from favae_b200 import _lib
    x = torch.randn(n_lat, DIM, device=device)
    e = torch.nn.functional.normalize(torch.randn(K_CODES, DIM, device=device), dim=-1)
    xn = torch.empty(n_lat, DIM, device=device); xh = torch.empty(n_lat, DIM, device=device, dtype=torch.float16)
    en = torch.empty(K_CODES, DIM, device=device); eh = torch.empty(K_CODES, DIM, device=device, dtype=torch.float16)
    st = _lib.stream()
    _lib.call('favae_vq_prepare_rows', x.data_ptr(), n_lat, DIM, 1, 1, xn.data_ptr(), xh.data_ptr(), None, st)
    _lib.call('favae_vq_prepare_rows', e.data_ptr(), K_CODES, DIM, 1, 1, en.data_ptr(), eh.data_ptr(), None, st)
    nbytes = _lib.load().favae_vq_search_tc_workspace_bytes(n_lat, K_CODES, DIM)
    ws = torch.empty(nbytes, device=device, dtype=torch.uint8)
    idx = torch.empty(n_lat, device=device, dtype=torch.int64)
    keys = torch.empty(n_lat, device=device, dtype=torch.int64)

    def run():
        _lib.call('favae_vq_search_tc', xh.data_ptr(), eh.data_ptr(), xn.data_ptr(), en.data_ptr(), n_lat, K_CODES,
                  DIM, ws.data_ptr(), nbytes, keys.data_ptr(), idx.data_ptr(), _lib.stream())
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


def time_blur(batch, device, iters=20):
    """Device time of the level-0 blur forward and backward (+ sigma gradient) kernels."""
    from favae_b200 import _lib
    shape = (batch, 128, 256, 256)
    x = torch.randn(shape, device=device); g = torch.randn(shape, device=device)
    y = torch.empty_like(x); gs = torch.empty(1, device=device)
    sig = torch.tensor(SIGMA0, device=device)
    maps = batch * 128
    parts = torch.empty(int(_lib.load().favae_blur_partials(maps, 256, 256)), device=device)
    out = {}
    for name, fn in (
            ('fwd', lambda: _lib.call('favae_blur_forward', x.data_ptr(), maps, 256, 256, KSIZE, sig.data_ptr(),
                                      y.data_ptr(), _lib.stream())),
            ('bwd', lambda: _lib.call('favae_blur_backward', g.data_ptr(), x.data_ptr(), maps, 256, 256, KSIZE,
                                      sig.data_ptr(), y.data_ptr(), gs.data_ptr(), parts.data_ptr(), _lib.stream()))):
        for _ in range(20):                      # the e2e phase before this leaves the GPU mostly idle: let the clocks ramp up
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        out[name] = a.elapsed_time(b) / iters
    return out


if __name__ == '__main__':
    main()
