"""Drop-in replacement for ``/root/reference/models/l2_quantize.py`` on B200.

Same module names, constructor arguments, ``forward`` return values and ``state_dict`` keys
(``_codebook.initted``, ``_codebook.cluster_size``, ``_codebook.embed`` [+ ``_codebook.embed_avg``,
``project_in.*``, ``project_out.*``]) as the reference, so ``models/vqgan_fcm.py:100-105,115`` and
``models/txt_cond_transformer.py:136,165`` keep working and published checkpoints load.

Everything on the hot path runs in hand-written sm_100a kernels behind the C ABI in
``include/favae_b200.h``:

  row l2norm + NCHW rearrange  -> favae_vq_prepare_rows      (l2_quantize.py:403,408,540)
  similarity GEMM + argmax     -> favae_vq_search_tc / _exact (:410-411, :280-282)
  gather + straight-through    -> favae_vq_gather_st          (:415, :554, :560)
  bincount + one-hot GEMM      -> favae_vq_code_stats         (:412, :418, :426)
  EMA / normalise / where      -> favae_vq_ema_update_*       (:421-438, :292-300)
  autograd backward            -> favae_vq_backward

The two all-reduces of the reference (:419, :427) become ONE all-reduce of the flat
``[bins | embed_sum]`` buffer, issued on a side stream; the EMA update that consumes it is
deferred to the next touch of the codebook (the next quantizer call, ``state_dict()``, any
attribute access to ``embed`` / ``cluster_size``), so the collective overlaps whatever the
training step does in between (the spectrum losses) instead of stalling the compute stream.
The forward of a call only ever uses the pre-update codebook (reference :415 gathers before
:421-438 update), so nothing observable changes.  There is no CPU path: CPU tensors raise.

Environment switches: ``FAVAE_VQ_SEARCH=auto|tc|exact`` (search kernel),
``FAVAE_VQ_DETERMINISTIC=1`` (bit-reproducible code statistics), ``FAVAE_VQ_CACHE=0`` (re-normalise
the codebook on every call instead of reusing the rows the EMA kernel emitted).
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch import nn

from . import _dist, _lib

__all__ = ['VectorQuantize', 'CosineSimCodebook', 'EuclideanCodebook', 'l2norm', 'orthogonal_loss_fn']


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


def l2norm(t):
    return F.normalize(t, p=2, dim=-1)


def uniform_init(*shape):
    t = torch.empty(shape)
    nn.init.kaiming_uniform_(t)
    return t


def orthogonal_loss_fn(t):
    """l2_quantize.py:174-179.  Cold option (no published FA-VAE config enables it): plain
    library GEMM on the normalised codes."""
    h, n = t.shape[:2]
    normed = l2norm(t)
    identity = torch.eye(n, device=t.device).expand(h, n, n)
    cosine_sim = torch.einsum('h i d, h j d -> h i j', normed, normed)
    return ((cosine_sim - identity) ** 2).sum() / (h * n ** 2)


def _search_mode():
    return os.environ.get('FAVAE_VQ_SEARCH', 'auto')


def _deterministic():
    return os.environ.get('FAVAE_VQ_DETERMINISTIC', '0') not in ('', '0')


def _cache_enabled():
    return os.environ.get('FAVAE_VQ_CACHE', '1') not in ('', '0')


_SIDE_STREAMS = {}


def _side_stream(device):
    """One extra stream per device for the statistics all-reduce."""
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


class _QuantizeFunction(torch.autograd.Function):
    """(x) -> (out, loss_sum) with d/dx = g_out + coef * (x - out) * g_loss."""

    @staticmethod
    def forward(ctx, x, codebook, hw, straight_through, want_loss):
        n = x.numel() // codebook.dim
        out, idx, loss_sum = codebook._search_and_gather(x, n, hw, straight_through, want_loss)
        ctx.save_for_backward(x, out)
        ctx.mark_non_differentiable(idx)
        ctx.codebook = codebook
        return out, idx, loss_sum

    @staticmethod
    def backward(ctx, g_out, _g_idx, g_loss):
        x, out = ctx.saved_tensors
        gx = torch.empty_like(x)
        g_out = g_out.contiguous() if g_out is not None else None
        g_loss = g_loss.contiguous() if g_loss is not None else None
        with _lib.on_device_of(x, g_out, g_loss):
            _lib.call('favae_vq_backward', _lib.ptr(x), _lib.ptr(out), _lib.ptr(g_out), _lib.ptr(g_loss),
                      x.numel(), 2.0, _lib.ptr(gx), _lib.stream())
            # The side-stream tail of the forward call (all-reduce + EMA) finished long ago: joining it here
            # costs nothing and orders everything the caller does after backward -- optimizer step, DDP's
            # buffer broadcast of the next forward, which reads the buffers without going through the module
            # attributes -- behind the codebook update.
            ctx.codebook._flush()
        return gx, None, None, None, None


class _CodebookBase(nn.Module):
    """Shared machinery of the two codebook classes (reference :183-306 and :308-444)."""

    cosine = True

    def __init__(self, dim, codebook_size, num_codebooks=1, kmeans_init=False, kmeans_iters=10,
                 decay=0.8, eps=1e-5, threshold_ema_dead_code=2, use_ddp=False,
                 learnable_codebook=False, sample_codebook_temp=0.):
        super().__init__()
        if num_codebooks != 1:
            raise NotImplementedError('favae_b200: separate codebooks per head are not built '
                                      '(never used by FA-VAE, models/vqgan_fcm.py:103-105)')
        if sample_codebook_temp != 0:
            raise NotImplementedError('favae_b200: gumbel sampling (sample_codebook_temp > 0) is not built')
        if use_ddp and (kmeans_init or threshold_ema_dead_code != 0):
            raise NotImplementedError('favae_b200: distributed k-means init / dead-code expiry '
                                      '(sample_vectors_distributed, :101-115) are not built')
        self.kmeans_init = kmeans_init
        self._init_done = not kmeans_init
        self.dim = dim
        self.decay = decay
        self.codebook_size = codebook_size
        self.num_codebooks = num_codebooks
        self.kmeans_iters = kmeans_iters
        self.eps = eps
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.sample_codebook_temp = sample_codebook_temp
        self.use_ddp = use_ddp
        self.learnable_codebook = learnable_codebook

        if kmeans_init:
            embed = torch.zeros(num_codebooks, codebook_size, dim)           # :200-201, :326-329
        else:
            embed = uniform_init(num_codebooks, codebook_size, dim)
            if self.cosine:
                embed = l2norm(embed)
        self.register_buffer('initted', torch.Tensor([not kmeans_init]))
        self.register_buffer('cluster_size', torch.zeros(num_codebooks, codebook_size))
        if not self.cosine:
            self.register_buffer('embed_avg', embed.clone())
        if learnable_codebook:
            self.embed = nn.Parameter(embed)
        else:
            self.register_buffer('embed', embed)

    # -- deferred EMA ---------------------------------------------------------------------
    _LAZY_NAMES = frozenset(('embed', 'cluster_size', 'embed_avg'))

    def __getattr__(self, name):
        # buffers / parameters live in _buffers / _parameters, so nn.Module resolves them here: a
        # pending EMA update (see _after_stats) is applied before anybody can look at the codebook
        if name in _CodebookBase._LAZY_NAMES and self.__dict__.get('_pending') is not None:
            self._flush()
        return super().__getattr__(name)

    def _raw(self, name):
        """Buffer / parameter without triggering the flush."""
        t = self._buffers.get(name)
        return t if t is not None else self._parameters[name]

    def _flush(self):
        """Apply the EMA update whose statistics were all-reduced on the side stream."""
        pending = self.__dict__.get('_pending')
        if pending is None:
            return
        self.__dict__['_pending'] = None
        stats, en, eh, done = pending
        with torch.cuda.device(stats.device):
            torch.cuda.current_stream(stats.device).wait_event(done)
            if en is not None:                   # (en is None: the side stream already applied the update)
                self._ema_update(en, eh, stats)

    def _save_to_state_dict(self, *args, **kwargs):
        self._flush()
        return super()._save_to_state_dict(*args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._flush()
        self.__dict__['_prep'] = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, *args, **kwargs):
        self._flush()
        self.__dict__['_prep'] = None
        self.__dict__['_scratch'] = {}
        return super()._apply(*args, **kwargs)

    def __getstate__(self):
        # pickling / deepcopy: finish the pending update, leave scratch, events and caches behind
        self._flush()
        state = self.__dict__.copy()
        for key in ('_pending', '_prep', '_scratch'):
            state.pop(key, None)
        return state

    def invalidate_cache(self):
        """Forget the cached normalised codebook rows.  Needed only after writing to ``embed`` through
        ``.data`` (which bypasses the tensor version counter that every other in-place write bumps)."""
        self._flush()
        self.__dict__['_prep'] = None

    def _distributed(self):
        import torch.distributed as dist
        return bool(self.use_ddp and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)

    def _ema_on_side_stream(self, t):
        """True when the tail of a training call -- the statistics all-reduce of a data-parallel run and the
        EMA update -- runs on the side stream (cosine codebook without dead-code expiry; FAVAE_EMA_STREAM=0
        disables it)."""
        return bool(t.is_cuda and self.threshold_ema_dead_code == 0 and self.cosine
                    and os.environ.get('FAVAE_EMA_STREAM', '1') not in ('', '0'))

    def _defers_ema(self, t):
        """True when the statistics all-reduce of this call runs on the side stream and the EMA itself waits
        for the next touch of the codebook (data-parallel runs of the codebooks the side-stream update does
        not cover)."""
        return bool(self._distributed() and t.is_cuda and self.threshold_ema_dead_code == 0
                    and not self._ema_on_side_stream(t))

    def _after_stats(self, en, eh, stats):
        """Statistics of this call are ready on the current stream: all-reduce + EMA."""
        if self._ema_on_side_stream(stats):
            # Nothing of this call reads the updated codebook (its outputs use the pre-update one, reference
            # :415 before :421-438), so the whole tail leaves the caller's stream: ONE all-reduce of
            # [bins | embed_sum] in a data-parallel run (reference: two blocking ones, :419/:427 and :291/:295)
            # and the EMA update run on the side stream, next to whatever the caller does after the quantizer
            # (decoder, spectrum losses, backward); the next touch of the codebook waits for the event.
            side = _side_stream(stats.device)
            side.wait_stream(torch.cuda.current_stream(stats.device))
            with torch.cuda.stream(side):
                if self.use_ddp:
                    _dist.all_reduce_stats(stats)      # (raises like the reference without a process group)
                self._ema_update(en, eh, stats)
                done = torch.cuda.Event()
                done.record(side)
            self.__dict__['_pending'] = (stats, None, None, done)
            return
        if self._defers_ema(stats):
            side = _side_stream(stats.device)
            side.wait_stream(torch.cuda.current_stream(stats.device))
            with torch.cuda.stream(side):
                _dist.all_reduce_stats(stats)
                done = torch.cuda.Event()
                done.record(side)
            self.__dict__['_pending'] = (stats, en, eh, done)
            return
        if self.use_ddp:
            _dist.all_reduce_stats(stats)      # raises like the reference without a process group
        self._ema_update(en, eh, stats)

    # -- kernels --------------------------------------------------------------------------
    def _buf(self, name, shape, dtype, device):
        """Per-shape scratch reused from call to call (workspaces, keys, statistics): stream-ordered
        reuse on the calling stream, never handed to the caller."""
        sc = self.__dict__.setdefault('_scratch', {})
        key = (name, tuple(shape), dtype, device)
        t = sc.get(key)
        if t is None:
            if len(sc) > 64:
                sc.clear()
            t = sc[key] = torch.empty(shape, device=device, dtype=dtype)
        return t

    def _prepare(self, x, n, hw, normalize, half, out=None):
        d = self.dim
        xn, xh = out if out is not None else (None, None)
        if xn is None:
            xn = torch.empty((n, d), device=x.device, dtype=torch.float32)
        if half and xh is None:
            xh = torch.empty((n, d), device=x.device, dtype=torch.float16)
        _lib.call('favae_vq_prepare_rows', _lib.ptr(x), n, d, hw, int(normalize), _lib.ptr(xn),
                  _lib.ptr(xh) if half else None, None, _lib.stream())
        return xn, (xh if half else None)

    def _prepared_codebook(self, use_tc):
        """(en, eh) = l2norm(embed) in fp32 and 16x that in fp16.  The reference re-normalises the
        whole codebook on every call (:408); here the EMA kernel emits the rows of the updated codebook
        and they are reused as long as ``embed`` has not been written by anybody else (tensor version
        counter + address)."""
        embed_t = self._raw('embed')
        embed = embed_t.detach()[0]
        key = (embed_t.data_ptr(), embed_t._version, embed_t.device)
        prep = self.__dict__.get('_prep')
        if prep is not None and prep['key'] == key and _cache_enabled() and (prep['eh'] is not None or not use_tc):
            return prep['en'], prep['eh']
        k = self.codebook_size
        out = (prep['en'], prep['eh']) if prep is not None and prep['en'].device == embed.device else None
        en, eh = self._prepare(embed, k, 1, True, use_tc, out)
        self.__dict__['_prep'] = {'key': key, 'en': en, 'eh': eh}
        return en, eh

    def _search_rows(self, xn, xh, codes, n):
        """Nearest code of every prepared latent row against ``codes`` (K, D), or against the module's
        own codebook when ``codes`` is None.  Returns (idx, en, eh)."""
        d, dev = self.dim, xn.device
        idx = torch.empty((n,), device=dev, dtype=torch.int64)
        keys = self._buf('keys', (n,), torch.int64, dev)
        eh = None
        if self.cosine:
            use_tc = xh is not None
            if codes is None:
                en, eh = self._prepared_codebook(use_tc)
            else:
                en, eh = self._prepare(codes, codes.shape[0], 1, True, use_tc)
            k = en.shape[0]
            if use_tc:
                ws_bytes = _lib.load().favae_vq_search_tc_workspace_bytes(n, k, d)
                ws = self._buf('ws', (ws_bytes,), torch.uint8, dev)
                _lib.call('favae_vq_search_tc', _lib.ptr(xh), _lib.ptr(eh), _lib.ptr(xn), _lib.ptr(en),
                          n, k, d, _lib.ptr(ws), ws_bytes, _lib.ptr(keys), _lib.ptr(idx), _lib.stream())
            else:
                _lib.call('favae_vq_search_exact', _lib.ptr(xn), _lib.ptr(en), None, n, k, d, 0,
                          _lib.ptr(keys), _lib.ptr(idx), _lib.stream())
        else:
            if codes is None:
                codes = self._raw('embed').detach()[0]
            k = codes.shape[0]
            esq = torch.empty((k,), device=dev, dtype=torch.float32)
            en = torch.empty((k, d), device=dev, dtype=torch.float32)
            _lib.call('favae_vq_prepare_rows', _lib.ptr(codes), k, d, 1, 0, _lib.ptr(en), None,
                      _lib.ptr(esq), _lib.stream())
            _lib.call('favae_vq_search_exact', _lib.ptr(xn), _lib.ptr(en), _lib.ptr(esq), n, k, d, 1,
                      _lib.ptr(keys), _lib.ptr(idx), _lib.stream())
        return idx, en, eh

    def _code_stats(self, rows, idx, n, out=None):
        k, d = self.codebook_size, self.dim
        stats = out if out is not None else torch.empty((k * (d + 1),), device=rows.device, dtype=torch.float32)
        _lib.call('favae_vq_code_stats', _lib.ptr(rows), _lib.ptr(idx), n, k, d, int(_deterministic()),
                  _lib.ptr(stats), _lib.stream())
        return stats

    @torch.no_grad()
    def _kmeans_init(self, rows, xh, n):
        """init_embed_ + kmeans (l2_quantize.py:124-164, :224-240, :351-367) on the prepared rows
        (normalised for the cosine codebook, raw for the Euclidean one), built from the search and
        statistics kernels.  Sampling uses the same torch RNG calls as the reference."""
        if self._init_done:
            return
        if bool(self.initted):                   # e.g. a checkpoint was loaded
            self._init_done = True
            return
        k, d = self.codebook_size, self.dim
        if n >= k:
            pick = torch.randperm(n, device=rows.device)[:k]
        else:
            pick = torch.randint(0, n, (k,), device=rows.device)
        means = rows[pick].contiguous()
        bins = torch.zeros(k, device=rows.device)
        for _ in range(self.kmeans_iters):
            idx, _, _ = self._search_rows(rows, xh, means, n)
            stats = self._code_stats(rows, idx, n)
            bins, esum = _dist.unpack_stats(stats, k, d)
            zero = bins == 0
            new_means = esum / bins.masked_fill(zero, 1.0)[:, None]
            if self.cosine:
                new_means = l2norm(new_means)
            means = torch.where(zero[:, None], means, new_means).contiguous()
        self.embed.data.copy_(means[None])
        if not self.cosine:
            self.embed_avg.data.copy_(means[None])
        self.cluster_size.data.copy_(bins[None])
        self.initted.data.fill_(1.0)
        self._init_done = True
        self.__dict__['_prep'] = None            # written through .data: the version counter did not move

    @torch.no_grad()
    def _expire_codes(self, rows):
        """expire_codes_ + replace (l2_quantize.py:242-262, :369-389): codes whose EMA cluster size
        fell under the threshold are replaced by l2-normalised latents sampled from the batch."""
        if self.threshold_ema_dead_code == 0:
            return
        expired = self.cluster_size[0] < self.threshold_ema_dead_code
        num = int(expired.sum())                 # host sync, as in the reference (:376 .item())
        if num == 0:
            return
        samples = rows if self.cosine else l2norm(rows)     # the reference normalises in both classes
        n = samples.shape[0]
        if n >= num:
            pick = torch.randperm(n, device=rows.device)[:num]
        else:
            pick = torch.randint(0, n, (num,), device=rows.device)
        self.embed.data[0][expired] = samples[pick]
        self.__dict__['_prep'] = None            # written through .data: the version counter did not move

    def _search_and_gather(self, x, n, hw, straight_through, want_loss):
        k, d, dev = self.codebook_size, self.dim, x.device
        self._flush()                            # a deferred EMA update of the previous call comes first
        mode = _search_mode()
        use_tc = (self.cosine and mode != 'exact' and n > 0
                  and _lib.load().favae_vq_search_tc_workspace_bytes(n, k, d) > 0)
        if mode == 'tc' and not use_tc:
            raise RuntimeError('favae_b200: FAVAE_VQ_SEARCH=tc but the tensor-core search does not '
                               f'support dim={d}, codebook_size={k}')
        # rows: l2-normalised latents (cosine) or the rearranged raw latents (Euclidean)
        xn, xh = self._prepare(x, n, hw, self.cosine, use_tc,
                               (self._buf('xn', (n, d), torch.float32, dev),
                                self._buf('xh', (n, d), torch.float16, dev) if use_tc else None))
        self._kmeans_init(xn, xh, n)
        embed = self._raw('embed').detach()[0]
        if not embed.is_contiguous():
            raise RuntimeError('favae_b200: codebook buffer must be contiguous')
        if embed.device != dev:
            raise RuntimeError(f'favae_b200: latents on {dev} but the codebook is on {embed.device}')
        idx, en, eh = self._search_rows(xn, xh, None, n)

        # When only the all-reduce leaves the caller's stream and the EMA is deferred to the next touch
        # (_defers_ema), the statistics come first so that the collective also overlaps the gather of this
        # call; otherwise the EMA (inline, or on the side stream right behind the collective) would overwrite
        # the codebook the gather still has to read (reference :415 gathers before :421-438 update), so the
        # order stays gather -> statistics -> [all-reduce ->] EMA.
        early = self.training and self._defers_ema(x)
        if early:
            stats = self._code_stats(xn, idx, n, self._buf('stats', (k * (d + 1),), torch.float32, dev))
            self._after_stats(en, eh, stats)

        out = torch.empty_like(x)
        # written in full by favae_vq_gather_st when the loss is wanted
        loss_sum = (torch.empty if want_loss else torch.zeros)((1,), device=dev, dtype=torch.float32)
        blocks = (n + 31) // 32 if hw == 1 else (n // hw) * ((hw + 31) // 32)
        partials = self._buf('partials', (max(blocks, 1),), torch.float32, dev) if want_loss else None
        _lib.call('favae_vq_gather_st', _lib.ptr(x), _lib.ptr(embed), _lib.ptr(idx), n, k, d, hw,
                  int(straight_through), _lib.ptr(out), _lib.ptr(partials),
                  _lib.ptr(loss_sum) if want_loss else None, _lib.stream())

        if self.training and not early:
            stats = self._code_stats(xn, idx, n, self._buf('stats', (k * (d + 1),), torch.float32, dev))
            self._after_stats(en, eh, stats)
            self._expire_codes(xn)
        return out, idx, loss_sum

    def _ema_update(self, en, eh, stats):
        raise NotImplementedError

    # -- reference-compatible codebook call: x (..., d) -> (quantize, embed_ind) --------------
    @torch.no_grad()
    def forward(self, x):
        _lib.require_cuda(x)
        x = x.float().contiguous()
        shape = x.shape
        n = x.numel() // self.dim
        with _lib.on_device_of(x, self._raw('embed')):
            out, idx, _ = self._search_and_gather(x, n, 1, False, False)
        return out.view(shape), idx.view(shape[:-1])


class CosineSimCodebook(_CodebookBase):
    """l2_quantize.py:308-444."""
    cosine = True

    def _ema_update(self, en, eh, stats):
        # the kernel also emits l2norm(updated codebook) into en / eh (in place), which stay valid for
        # the next search as long as nobody else writes embed (_prepared_codebook)
        embed_t = self._raw('embed')
        prep = self.__dict__.get('_prep')
        keep = _cache_enabled() and prep is not None and prep['en'] is en
        _lib.call('favae_vq_ema_update_cosine', _lib.ptr(embed_t), _lib.ptr(self._raw('cluster_size')),
                  _lib.ptr(en), _lib.ptr(stats), self.codebook_size, self.dim, float(self.decay),
                  _lib.ptr(en) if keep else None, _lib.ptr(eh) if keep and eh is not None else None,
                  _lib.stream())
        if keep:
            prep['key'] = (embed_t.data_ptr(), embed_t._version, embed_t.device)
            if eh is None:
                prep['eh'] = None


class EuclideanCodebook(_CodebookBase):
    """l2_quantize.py:183-306, including the quirk that ``embed_avg`` is never refreshed (:294-300)."""
    cosine = False

    def _ema_update(self, en, eh, stats):
        scratch = torch.empty((2,), device=stats.device, dtype=torch.float32)
        _lib.call('favae_vq_ema_update_euclid', _lib.ptr(self._raw('embed')), _lib.ptr(self._raw('cluster_size')),
                  _lib.ptr(self._raw('embed_avg')), _lib.ptr(stats), self.codebook_size, self.dim,
                  float(self.decay), float(self.eps), _lib.ptr(scratch), _lib.stream())


class VectorQuantize(nn.Module):
    """l2_quantize.py:448-596 -- same signature, same return triple ``(quantize, embed_ind, loss)``."""

    def __init__(self, dim, codebook_size, codebook_dim=None, heads=1, separate_codebook_per_head=False,
                 decay=0.8, eps=1e-5, kmeans_init=False, kmeans_iters=10, use_cosine_sim=False,
                 threshold_ema_dead_code=0, channel_last=True, accept_image_fmap=False,
                 commitment_weight=1., orthogonal_reg_weight=0., orthogonal_reg_active_codes_only=False,
                 orthogonal_reg_max_codes=None, sample_codebook_temp=0., sync_codebook=False):
        super().__init__()
        if separate_codebook_per_head and heads > 1:
            raise NotImplementedError('favae_b200: separate codebooks per head are not built '
                                      '(FA-VAE uses heads=1)')
        self.heads = heads
        self.separate_codebook_per_head = separate_codebook_per_head

        codebook_dim = default(codebook_dim, dim)
        codebook_input_dim = codebook_dim * heads
        requires_projection = codebook_input_dim != dim
        self.project_in = nn.Linear(dim, codebook_input_dim) if requires_projection else nn.Identity()
        self.project_out = nn.Linear(codebook_input_dim, dim) if requires_projection else nn.Identity()

        self.eps = eps
        self.commitment_weight = commitment_weight
        has_codebook_orthogonal_loss = orthogonal_reg_weight > 0
        self.orthogonal_reg_weight = orthogonal_reg_weight
        self.orthogonal_reg_active_codes_only = orthogonal_reg_active_codes_only
        self.orthogonal_reg_max_codes = orthogonal_reg_max_codes

        codebook_class = EuclideanCodebook if not use_cosine_sim else CosineSimCodebook
        self._codebook = codebook_class(
            dim=codebook_dim, num_codebooks=1, codebook_size=codebook_size, kmeans_init=kmeans_init,
            kmeans_iters=kmeans_iters, decay=decay, eps=eps,
            threshold_ema_dead_code=threshold_ema_dead_code, use_ddp=sync_codebook,
            learnable_codebook=has_codebook_orthogonal_loss, sample_codebook_temp=sample_codebook_temp)
        self.codebook_size = codebook_size
        self.accept_image_fmap = accept_image_fmap
        self.channel_last = channel_last

    @property
    def codebook(self):
        return self._codebook.embed[0]

    def get_codebook_entry(self, indices, shape):
        """l2_quantize.py:518-530 as a row gather (no one-hot GEMM); ``shape`` is (B, h, w, C)."""
        _lib.require_cuda(indices)
        idx = indices.reshape(-1).to(torch.int64).contiguous()
        embed = self._codebook.embed.detach()[0]
        n, d, k = idx.numel(), embed.shape[-1], embed.shape[0]
        if n and (int(idx.min()) < 0 or int(idx.max()) >= k):
            raise IndexError('favae_b200: code index out of range')
        if shape is not None:
            b, h, w, c = shape
            if c != d or b * h * w != n:
                raise RuntimeError(f'shape {tuple(shape)} does not match {n} indices of dim {d}')
            out = torch.empty((b, c, h, w), device=idx.device, dtype=torch.float32)
            hw = h * w
        else:
            out = torch.empty((n, d), device=idx.device, dtype=torch.float32)
            hw = 1
        with _lib.on_device_of(embed, idx):
            _lib.call('favae_vq_gather_rows', _lib.ptr(embed), _lib.ptr(idx), n, k, d, hw, _lib.ptr(out),
                      _lib.stream())
        return out

    def forward(self, x):
        _lib.require_cuda(x)
        with _lib.on_device_of(x, self._codebook._raw('embed')):
            return self._forward(x)

    def _forward(self, x):
        device = x.device
        need_transpose = not self.channel_last and not self.accept_image_fmap
        projected = not isinstance(self.project_in, nn.Identity)
        heads = self.heads
        cb = self._codebook
        training = self.training

        if self.accept_image_fmap:
            b, _, height, width = x.shape
            if projected or heads > 1:
                x = x.permute(0, 2, 3, 1).reshape(b, height * width, -1)      # :540
                hw = 1
            else:
                hw = height * width          # the kernels address NCHW directly
        else:
            hw = 1
            if need_transpose:
                x = x.transpose(1, 2)        # 'b d n -> b n d'
        x = self.project_in(x)
        if heads > 1:                        # shared codebook: 'b n (h d) -> 1 (b h) n d'  (:547-549)
            bb, nn_ = x.shape[0], x.shape[1]
            x = x.reshape(bb, nn_, heads, -1).permute(0, 2, 1, 3)
        x = x.float().contiguous()
        if x.shape[1 if hw > 1 else -1] != cb.dim:
            raise RuntimeError(f'expected {cb.dim} channels, got {tuple(x.shape)}')

        want_loss = training and self.commitment_weight > 0
        if training:
            quantize, idx, loss_sum = _QuantizeFunction.apply(x, cb, hw, True, want_loss)
        else:
            with torch.no_grad():
                quantize, idx, loss_sum = _QuantizeFunction.apply(x, cb, hw, False, False)

        if want_loss:
            loss = loss_sum * (self.commitment_weight / x.numel())               # :556, :560-561 (shape (1,))
        else:
            loss = torch.zeros(1, device=device, requires_grad=training)        # :556
        if training:
            if self.orthogonal_reg_weight > 0:                                  # :563-577
                # Reproduced as written in the reference, quirks included: the (1, K, D) codebook is
                # indexed and measured along dim 0 (the head axis, size 1).  `num_codes` is therefore 1,
                # orthogonal_reg_max_codes never subsamples and no randperm is drawn; and
                # orthogonal_reg_active_codes_only indexes dim 0 with code ids, which is an index error
                # for any id > 0 (a device-side assert in the reference; a clean IndexError here).
                codebook = cb.embed
                if self.orthogonal_reg_active_codes_only:
                    unique_code_ids = torch.unique(idx)
                    if int(unique_code_ids.max()) >= codebook.shape[0]:
                        raise IndexError(f'index {int(unique_code_ids.max())} is out of bounds for dimension 0 '
                                         f'with size {codebook.shape[0]} (orthogonal_reg_active_codes_only '
                                         'indexes the head axis of the codebook, l2_quantize.py:569)')
                    codebook = codebook[unique_code_ids]
                num_codes = codebook.shape[0]
                if exists(self.orthogonal_reg_max_codes) and num_codes > self.orthogonal_reg_max_codes:
                    rand_ids = torch.randperm(num_codes, device=device)[:self.orthogonal_reg_max_codes]
                    codebook = codebook[rand_ids]
                loss = loss + orthogonal_loss_fn(codebook) * self.orthogonal_reg_weight

        if heads > 1:                        # '1 (b h) n d -> b n (h d)', '1 (b h) n -> b n h'  (:579-585)
            quantize = quantize.permute(0, 2, 1, 3).reshape(bb, nn_, -1)
            idx = idx.view(bb, heads, nn_).permute(0, 2, 1)
        quantize = self.project_out(quantize)
        if self.accept_image_fmap:
            if projected or heads > 1:
                quantize = quantize.reshape(b, height, width, -1).permute(0, 3, 1, 2)
            embed_ind = idx.reshape(b, height, width, heads) if heads > 1 else idx.view(b, height, width)
        else:
            if need_transpose:
                quantize = quantize.transpose(1, 2)
            embed_ind = idx if heads > 1 else idx.view(x.shape[:-1])
        return quantize, embed_ind, loss
