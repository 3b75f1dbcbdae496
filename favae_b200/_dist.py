"""Data-parallel plumbing of the quantizer: ONE all-reduce of the flat ``[bins | embed_sum]``
statistics buffer per quantizer call (the reference issues two blocking all-reduces,
/root/reference/models/l2_quantize.py:419,427 and :291,295).  ``torch.distributed`` (NCCL over
NVLink on GPUs, gloo in the CPU tests) is used as plumbing only."""
from __future__ import annotations

import torch
import torch.distributed as distributed


def pack_stats(bins: torch.Tensor, embed_sum: torch.Tensor) -> torch.Tensor:
    """``[bins (K) | embed_sum (K*D)]`` -- the layout favae_vq_code_stats writes."""
    return torch.cat([bins.reshape(-1), embed_sum.reshape(-1)])


def unpack_stats(stats: torch.Tensor, k: int, d: int):
    return stats[:k], stats[k:].view(k, d)


def all_reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """SUM over ranks, in place.  Like the reference (``use_ddp`` -> ``distributed.all_reduce``)
    this raises if the process group has not been initialised."""
    distributed.all_reduce(stats, op=distributed.ReduceOp.SUM, group=group)
    return stats
