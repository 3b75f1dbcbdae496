"""The UNMODIFIED reference model (models/vqgan_fcm.py VQGANFCM, shipped in the git-ignored
baseline/_ref by oracle/reference_tree.py) around the drop-ins, on a GPU: one stage-0 + stage-1 training
step (favae_scripts/train_favae.py:75-116, LPIPS left out: its weights are absent offline) with
favae_b200.patch_reference() against the same model, same weights, same batch, unpatched.

Two comparisons: (1) unpatched on the GPU vs patched on the GPU -- the backbone is the same cuDNN code
on both sides, so every difference is the drop-ins': held to the north_star tolerances; (2) unpatched
on the CPU vs the GPU runs -- a sanity bound only (conv rounding differs between CPU and cuDNN)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(codebook_size=1024, n_embed=256, ch_mult=(1, 1, 2, 2, 4), attn_resolutions=[16], use_cosine_sim=True,
          use_l2_quantizer=True, kernel_size=9, dsl_init_sigma=3.0, use_gauss_resblock=True, commitment_weight=1.0)


def _step(model, vl, ffl, dsl, x, device):
    """train_favae.py:75-116 without LPIPS / adaptive discriminator weight."""
    x_recon, loss_q, logits_fake, z, enc_feats, dec_feats = model(x, stage=0)
    lazy = [type(t).__name__ for t in enc_feats + dec_feats]
    loss_l1 = (x - x_recon).abs().mean()
    loss_ffl = vl.recon_ffl_loss(ffl, x, x_recon)
    loss_dsl, lst = vl.recon_ffl_features_loss(dsl, enc_feats, dec_feats, device)
    loss = loss_l1 + loss_q.sum() + loss_ffl + loss_dsl.sum() - 0.1 * logits_fake.mean()
    loss.backward()
    with torch.no_grad():
        _, _, indices, _ = model.encode(x)               # the quantizer call of stage 1 (second EMA update)
    out = dict(loss=float(loss), l1=float(loss_l1), q=float(loss_q.sum()), ffl=float(loss_ffl), dsl=float(loss_dsl.sum()),
               levels=[float(v) for v in lst], ind=indices.detach().cpu(), lazy=lazy,
               g_enc_in=model.encoder.conv_in.weight.grad.detach().cpu().clone(),
               g_sig_e=model.encoder.sigmas.grad.detach().cpu().clone(),
               g_sig_d=model.decoder.sigmas.grad.detach().cpu().clone(),
               embed=model.quantizer._codebook.embed.detach().cpu().clone(),
               cluster=model.quantizer._codebook.cluster_size.detach().cpu().clone())
    return out


def test_reference_model_with_dropins_one_training_step():
    from oracle import ffl_oracle as fo
    from oracle import reference_tree
    mods = reference_tree.import_reference()
    if mods is None:
        pytest.skip('baseline/_ref is not shipped (run __graft_entry__.build() where /root/reference exists)')
    import favae_b200
    l2q, vl, fcm = mods
    import models.codec as codec
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    saved = {(m, n): getattr(m, n) for m in (l2q,) for n in ('VectorQuantize', 'CosineSimCodebook', 'EuclideanCodebook')}
    saved.update({(vl, n): getattr(vl, n) for n in ('recon_ffl_loss', 'recon_ffl_features_loss',
                                                    'recon_sl_gaussian_features_loss')})
    blur_classes = [c for m in (codec, fcm) for c in vars(m).values() if isinstance(c, type) and '_gaussian_blur' in vars(c)]
    saved_blur = {c: c._gaussian_blur for c in blur_classes}
    try:
        torch.manual_seed(0)
        x = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(1)) * 2 - 1
        ref_cpu = fcm.VQGANFCM(device='cpu', **KW).train()
        sd = {k: v.clone() for k, v in ref_cpu.state_dict().items()}
        o_cpu = _step(ref_cpu, vl, fo.FocalFrequencyLossOracle(loss_weight=1.0),
                      fo.FocalFrequencyLossOracle(loss_weight=0.01), x, 'cpu')
        del ref_cpu
        ref_gpu = fcm.VQGANFCM(device='cuda', **KW).cuda().train()
        ref_gpu.load_state_dict(sd)
        o_ref = _step(ref_gpu, vl, fo.FocalFrequencyLossOracle(loss_weight=1.0),
                      fo.FocalFrequencyLossOracle(loss_weight=0.01), x.cuda(), 'cuda')
        del ref_gpu
        assert set(o_ref['lazy']) == {'Tensor'}

        done = favae_b200.patch_reference()
        assert 'models.l2_quantize' in done and 'losses.vqgan_losses' in done
        ours = fcm.VQGANFCM(device='cuda', **KW).cuda().train()
        assert type(ours.quantizer).__module__.startswith('favae_b200')
        ours.load_state_dict(sd)                             # same keys as the reference module
        o = _step(ours, vl, favae_b200.FocalFrequencyLoss(loss_weight=1.0),
                  favae_b200.FocalFrequencyLoss(loss_weight=0.01), x.cuda(), 'cuda')
        # the blurred features leave the model as deferred handles and the loss wrapper fused them
        assert set(o['lazy']) == {'LazyBlur'}
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)
        for c, f in saved_blur.items():
            c._gaussian_blur = f
        torch.backends.cudnn.allow_tf32 = old_tf32

    # (1) drop-ins vs the reference's own ops, same GPU backbone
    flips = int((o['ind'] != o_ref['ind']).sum())
    assert flips <= 2, f'{flips} of {o["ind"].numel()} code indices differ'
    for key, tol in (('l1', 1e-5), ('q', 1e-4), ('ffl', 1e-4), ('dsl', 1e-4), ('loss', 1e-4)):
        assert o[key] == pytest.approx(o_ref[key], rel=tol), key
    for a, b in zip(o['levels'], o_ref['levels']):
        assert a == pytest.approx(b, rel=1e-4)
    assert (o['g_enc_in'] - o_ref['g_enc_in']).abs().max() <= 2e-3 * o_ref['g_enc_in'].abs().max()
    # sigma gradients of a freshly initialised model are tiny cancelling sums (1e-6..1e-8): both sides
    # evaluate them in fp32, so they are compared on the scale of the largest one
    for key in ('g_sig_e', 'g_sig_d'):
        assert (o[key] - o_ref[key]).abs().max() <= 5e-3 * o_ref[key].abs().max(), (key, o[key], o_ref[key])
    if flips == 0:
        torch.testing.assert_close(o['embed'], o_ref['embed'], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(o['cluster'], o_ref['cluster'], rtol=1e-5, atol=1e-7)
    # (2) CPU reference: sanity bound across conv implementations
    assert int((o['ind'] != o_cpu['ind']).sum()) <= max(4, o['ind'].numel() // 50)
    for key in ('l1', 'q', 'ffl', 'dsl'):
        assert o[key] == pytest.approx(o_cpu[key], rel=2e-3), key
