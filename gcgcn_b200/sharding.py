"""Document-sharded data parallelism.

Every document graph is independent in forward and backward, so ranks never exchange
activations: rank r runs the whole hot path on its own documents.  The only collective is one
all-reduce of the parameter gradients per optimizer step, issued as a single contiguous bucket
(2.2 MB fp32 for the GloVe configuration) so that it is one NVLink/NVSwitch message.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .batch import shard_documents  # noqa: F401  (re-exported)


class GradBucket:
    """One flat fp32 buffer holding every trainable parameter's gradient.

    ``pack()`` copies ``p.grad`` (zeros where a parameter got no gradient, e.g. the unused
    ``linears_k``) into the buffer, ``all_reduce()`` sums it over ranks in one call, ``unpack()``
    writes the reduced values back to ``p.grad``.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=dev)
        self.views = [v.view_as(p) for v, p in zip(torch.split(self.flat, self.sizes), self.params)]

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def pack(self):
        """p.grad -> bucket: one multi-tensor copy for all parameters that have a gradient."""
        vs = [v for v, p in zip(self.views, self.params) if p.grad is not None]
        gs = [p.grad for p in self.params if p.grad is not None]
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
        if vs:
            if hasattr(torch, "_foreach_copy_"):
                torch._foreach_copy_(vs, gs)
            else:
                for v, g in zip(vs, gs):
                    v.copy_(g)

    def all_reduce(self, group=None, average: bool = False, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average and not async_op:
            self.flat.div_(dist.get_world_size(group))
        return work

    def unpack(self, skip_none: bool = True):
        """bucket -> p.grad (one multi-tensor copy)."""
        vs, gs = [], []
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                if not skip_none:
                    p.grad = v.clone()
                continue
            vs.append(v)
            gs.append(p.grad)
        if gs:
            if hasattr(torch, "_foreach_copy_"):
                torch._foreach_copy_(gs, vs)
            else:
                for g, v in zip(gs, vs):
                    g.copy_(v)


def all_reduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = False,
                         bucket: Optional[GradBucket] = None) -> GradBucket:
    """Sum (or average) gradients over the ranks of ``group`` with one bucketed all-reduce."""
    bucket = bucket or GradBucket(params)
    bucket.pack()
    bucket.all_reduce(group=group, average=average)
    bucket.unpack()
    return bucket


def local_documents(sizes: Sequence[int], rank: int, world_size: int) -> List[int]:
    """Indices of the documents rank ``rank`` owns (balanced by sum n^2)."""
    return shard_documents(sizes, world_size)[rank]
