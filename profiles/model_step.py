"""The reference's own FA-VAE model (models/vqgan_fcm.py VQGANFCM from the git-ignored baseline/_ref,
unmodified; f=16, codebook 16384 x 256, FCM(Res) + non-pair-wise DSL, gaussian_kernel 9: BASELINE
configs[2]) through one full training step per iteration -- stage 0 forward, L1 + codebook + image
FFL + DSL losses, backward, Adam step; stage 1 discriminator forward / backward / Adam step
(favae_scripts/train_favae.py:75-116; LPIPS left out: its weights are absent offline) -- on one B200,
first as the reference runs it (its torch ops on the GPU, spectrum loss = the torch restatement of the
absent pip package) and then with favae_b200.patch_reference().  The conv backbone is out of this
repo's scope and identical in both arms: the difference between the two lines is the hot path.

    python profiles/model_step.py [--batch 8] [--steps 10] > profiles/model_step_r2.txt
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KW = dict(codebook_size=16384, n_embed=256, ch_mult=(1, 1, 2, 2, 4), attn_resolutions=[16], use_cosine_sim=True,
          use_l2_quantizer=True, kernel_size=9, dsl_init_sigma=3.0, use_gauss_resblock=True, commitment_weight=1.0)


def hinge_d(real, fake):
    return 0.5 * (torch.relu(1.0 - real).mean() + torch.relu(1.0 + fake).mean())


def run(fcm, vl, ffl, dsl, batch, steps, warmup, label):
    torch.manual_seed(0)
    model = fcm.VQGANFCM(device='cuda', **KW).cuda().train()
    g_params = list(model.encoder.parameters()) + list(model.decoder.parameters()) + list(model.quantizer.parameters())
    opt_g = torch.optim.Adam(g_params, lr=1e-5, betas=(0.5, 0.9))
    opt_d = torch.optim.Adam(model.discriminator.parameters(), lr=1e-5, betas=(0.5, 0.9))
    x = torch.rand(batch, 3, 256, 256, device='cuda', generator=torch.Generator('cuda').manual_seed(1)) * 2 - 1

    def step():
        opt_g.zero_grad(set_to_none=True)
        x_recon, loss_q, logits_fake, _, enc_feats, dec_feats = model(x, stage=0)
        loss = (x - x_recon).abs().mean() + loss_q.sum() - 0.1 * logits_fake.mean()
        loss = loss + vl.recon_ffl_loss(ffl, x, x_recon)
        l_dsl, _ = vl.recon_ffl_features_loss(dsl, enc_feats, dec_feats, 'cuda')
        loss = loss + l_dsl.sum()
        loss.backward()
        opt_g.step()
        opt_d.zero_grad(set_to_none=True)
        real, fake = model(x, stage=1)
        hinge_d(real, fake).backward()
        opt_d.step()
        return loss.detach()

    first = [float(step()) for _ in range(min(warmup, 3))]          # same weights / input in both variants
    for _ in range(warmup - len(first)):
        step()
    torch.cuda.synchronize()
    print(f'{label:<46} loss of the first steps: ' + ' '.join(f'{v:.6f}' for v in first), flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        last = step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    wall = (time.perf_counter() - t0) / steps * 1e3
    print(f'{label:<46} {ms:9.2f} ms/step (wall {wall:7.2f})  {batch / ms * 1e3:8.1f} img/s   loss {float(last):.5f}   '
          f'peak memory {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB', flush=True)
    del model, opt_g, opt_d
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    args = ap.parse_args()
    from oracle import ffl_oracle as fo
    from oracle import reference_tree
    mods = reference_tree.import_reference()
    if mods is None:
        raise SystemExit('baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists')
    l2q, vl, fcm = mods
    import favae_b200
    torch.backends.cudnn.benchmark = True
    print(f'# reference VQGANFCM training step on {torch.cuda.get_device_name(0)}, batch {args.batch}, 256^2, '
          f'K = 16384 x 256, FCM(Res) + DSL k = 9 (fp32, cuDNN TF32 convs as torch defaults)')
    ms_ref = run(fcm, vl, fo.FocalFrequencyLossOracle(loss_weight=1.0), fo.FocalFrequencyLossOracle(loss_weight=0.01),
                 args.batch, args.steps, args.warmup, 'reference modules on the GPU (unpatched)')
    done = favae_b200.patch_reference()
    ms_new = run(fcm, vl, favae_b200.FocalFrequencyLoss(loss_weight=1.0), favae_b200.FocalFrequencyLoss(loss_weight=0.01),
                 args.batch, args.steps, args.warmup, f'favae_b200.patch_reference() {done[1:]}')
    print(f'# whole-model speed-up from the drop-ins: {ms_ref / ms_new:.2f}x ({ms_ref - ms_new:.1f} ms of {ms_ref:.1f} ms '
          f'per step were the hot path\'s excess over the fused kernels)')


if __name__ == '__main__':
    main()
