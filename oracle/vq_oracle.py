"""CPU restatement of the reference vector quantizer (TEST INFRASTRUCTURE ONLY).

Follows ``/root/reference/models/l2_quantize.py``:

* ``cosine_codebook_forward``  -> ``CosineSimCodebook.forward``  (:392-444)
* ``euclid_codebook_forward``  -> ``EuclideanCodebook.forward``  (:265-306)
* ``vector_quantize_forward``  -> ``VectorQuantize.forward``     (:533-596)
* ``vector_quantize_backward`` -> autograd of :554-561 in closed form
* ``codebook_entry``           -> ``VectorQuantize.get_codebook_entry`` (:518-530)
* ``orthogonal_loss``          -> ``orthogonal_loss_fn`` (:174-179)

Written with plain torch CPU tensor ops in the same order as the reference so
that fp32 rounding matches the reference as closely as a restatement can.  State
(``embed``, ``cluster_size``, ``embed_avg``) is passed in and returned, never
held.  Pinned by ``tests/golden/vq_*.npz`` (see ``oracle/make_golden.py``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def l2norm(t: torch.Tensor) -> torch.Tensor:
    # l2_quantize.py:24-25  (F.normalize: x / max(||x||, 1e-12))
    return F.normalize(t, p=2, dim=-1)


def cosine_search(flat: torch.Tensor, embed: torch.Tensor):
    """flat (N,D), embed (K,D) un-normalised.  Returns (idx (N,) int64, xn, en).

    l2_quantize.py:403,408,410-411 -- first index wins ties (torch.argmax).
    """
    xn = l2norm(flat.float())
    en = l2norm(embed.float())
    # same contraction call as the reference (einsum -> bmm) so fp32 rounding, and with it
    # the winner among numerically tied codes, matches the reference on the same host
    sim = torch.einsum('h n d, h c d -> h n c', xn[None], en[None])[0]
    return sim.argmax(dim=-1), xn, en


def cosine_codebook_forward(flat, embed, cluster_size, *, training, decay=0.8,
                            all_reduce=None):
    """One call of CosineSimCodebook.forward on a flattened (N,D) input.

    Returns ``(quantize (N,D), idx (N,), new_embed (K,D), new_cluster_size (K,))``.
    ``all_reduce`` is an optional callable applied in place to ``bins`` and
    ``embed_sum`` (l2_quantize.py:419,427).
    """
    K = embed.shape[0]
    idx, xn, en = cosine_search(flat, embed)
    quantize = embed[idx]                                   # :415 pre-update, un-normalised
    if not training:
        return quantize, idx, embed, cluster_size
    bins = torch.bincount(idx, minlength=K).to(flat.dtype)  # :418
    if all_reduce is not None:
        all_reduce(bins)
    new_cluster = cluster_size * decay + bins * (1 - decay)  # :421
    zero = bins == 0                                        # :423
    bins_c = bins.masked_fill(zero, 1.0)                    # :424
    embed_sum = torch.zeros_like(embed).index_add_(0, idx, xn)  # :426
    if all_reduce is not None:
        all_reduce(embed_sum)
    en_new = l2norm(embed_sum / bins_c[:, None])            # :429-430
    en_new = torch.where(zero[:, None], en, en_new)         # :432-436 (normalised old code)
    new_embed = embed * decay + en_new * (1 - decay)        # :438
    return quantize, idx, new_embed, new_cluster


def euclid_codebook_forward(flat, embed, cluster_size, embed_avg, *, training,
                            decay=0.8, eps=1e-5, all_reduce=None):
    """EuclideanCodebook.forward (:265-306), including its quirk that ``embed_avg``
    is never updated (``embed_sum`` is computed at :294 and dropped)."""
    K = embed.shape[0]
    dist = -torch.cdist(flat.float()[None], embed.float()[None], p=2)[0]   # :280
    idx = dist.argmax(dim=-1)
    quantize = embed[idx]
    if not training:
        return quantize, idx, embed, cluster_size
    bins = torch.bincount(idx, minlength=K).to(flat.dtype)
    if all_reduce is not None:
        all_reduce(bins)
    new_cluster = cluster_size * decay + bins * (1 - decay)                 # :292
    smoothed = (new_cluster + eps) / (new_cluster.sum() + K * eps) * new_cluster.sum()  # :297
    new_embed = embed_avg / smoothed[:, None]                               # :299-300
    return quantize, idx, new_embed, new_cluster


def orthogonal_loss(codebook: torch.Tensor) -> torch.Tensor:
    # l2_quantize.py:174-179 with h == 1
    n = codebook.shape[0]
    c = l2norm(codebook)
    cs = c @ c.t()
    return ((cs - torch.eye(n)) ** 2).sum() / (n ** 2)


def vector_quantize_forward(x, embed, cluster_size, *, training=True,
                            commitment_weight=1.0, decay=0.8, use_cosine_sim=True,
                            embed_avg=None, eps=1e-5, all_reduce=None):
    """VectorQuantize.forward for ``accept_image_fmap=True``, heads=1, no projection.

    x: (B,C,h,w) fp32.  Returns dict with quantize (B,C,h,w), embed_ind (B,h,w),
    loss (1,), new_embed, new_cluster_size, and flat / q_flat for the backward.
    """
    B, C, h, w = x.shape
    flat = x.permute(0, 2, 3, 1).reshape(-1, C)             # :540
    if use_cosine_sim:
        q, idx, new_embed, new_cluster = cosine_codebook_forward(
            flat, embed, cluster_size, training=training, decay=decay, all_reduce=all_reduce)
    else:
        q, idx, new_embed, new_cluster = euclid_codebook_forward(
            flat, embed, cluster_size, embed_avg, training=training, decay=decay,
            eps=eps, all_reduce=all_reduce)
    if training:
        out = flat + (q - flat)                             # :554 forward value
        loss = torch.zeros(1)
        if commitment_weight > 0:
            loss = loss + F.mse_loss(out, flat) * commitment_weight  # :560-561 (on the ST value)
    else:
        out = q
        loss = torch.zeros(1)
    return dict(
        quantize=out.reshape(B, h, w, C).permute(0, 3, 1, 2).contiguous(),
        embed_ind=idx.reshape(B, h, w),
        loss=loss, new_embed=new_embed, new_cluster_size=new_cluster,
        flat=flat, q_flat=out,
    )


def vector_quantize_backward(flat, q_flat, grad_quantize_flat, grad_loss,
                             commitment_weight=1.0):
    """d/dx of (quantize, loss): g + 2*w*(x-q)/(N*D)*gl  (autograd of :554-561)."""
    n = flat.numel()
    return grad_quantize_flat + (2.0 * commitment_weight / n) * (flat - q_flat) * grad_loss


def codebook_entry(indices, embed, shape=None):
    # l2_quantize.py:518-530 -- one-hot matmul == row gather; NHWC view -> NCHW
    z = embed[indices.reshape(-1)]
    if shape is not None:
        z = z.view(shape).permute(0, 3, 1, 2).contiguous()
    return z
