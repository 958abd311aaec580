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
        if average and async_op:
            raise ValueError("GradBucket.all_reduce: average=True needs the result; wait on an async_op=True "
                             "all-reduce yourself and divide, or fold 1/world_size into the optimiser's grad_scale")
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average:
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


class OverlappedBuckets:
    """Gradient all-reduce that overlaps the rest of the backward pass ("overlap with the tail
    of backward -- weight grads for MAGGC finish first since backward runs MAGGC -> CAGGC").

    The parameters are cut into buckets in the order their gradients become final (``groups``: lists of parameters,
    e.g. [MAGGC + MultiHeadAttention, CAGGC + GATAttention]).  A post-accumulate-grad hook per parameter counts a
    bucket down; when its last gradient lands, the bucket is packed and its all-reduce is issued ``async_op=True`` on
    a side stream while autograd keeps launching the remaining backward kernels on the main stream.  ``finish()``
    waits for the collectives and writes the reduced values back into ``p.grad``.  Parameters that never receive a
    gradient (``linears_k``, G:137) must be left out of the groups.
    """

    def __init__(self, groups: Sequence[Sequence[torch.nn.Parameter]], group=None):
        self.group = group
        self.buckets = [GradBucket(g) for g in groups]
        self._left = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._stream = None
        self._handles = []
        for bi, b in enumerate(self.buckets):
            for p in b.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))
        self.reset()

    @property
    def nbytes(self) -> int:
        return sum(b.nbytes for b in self.buckets)

    def reset(self):
        """Call before every backward pass."""
        self._left = [len(b.params) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def _make_hook(self, bi):
        def hook(_param):
            self._left[bi] -= 1
            if self._left[bi] == 0:
                self._launch(bi)
        return hook

    def _launch(self, bi):
        b = self.buckets[bi]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        if b.flat.is_cuda:
            main = torch.cuda.current_stream(b.flat.device)
            if self._stream is None:
                self._stream = torch.cuda.Stream(b.flat.device)
            self._stream.wait_stream(main)                 # the bucket's gradients are complete on the main stream
            with torch.cuda.stream(self._stream):
                b.pack()
                self._work[bi] = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            b.pack()
            self._work[bi] = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """After backward(): wait for every bucket's all-reduce and copy the sums into ``p.grad``."""
        for bi, b in enumerate(self.buckets):
            if self._left[bi] > 0:                          # a gradient never arrived: reduce what there is
                self._launch(bi)
            w = self._work[bi]
            if w is None:
                continue
            if b.flat.is_cuda:
                with torch.cuda.stream(self._stream):
                    w.wait()
                    b.unpack()
            else:
                w.wait()
                b.unpack()
        if self._stream is not None:
            torch.cuda.current_stream(self._stream.device).wait_stream(self._stream)

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


class FlatTrainer:
    """Config 5's optimiser state: parameters, gradients and both Adam moments each live in ONE flat fp32
    buffer, the module's parameters and their ``.grad`` are views into them.

    * autograd accumulates straight into the gradient bucket (``p.grad`` pre-exists as a view), so a step
      needs no pack / unpack copies: ``zero_grad()`` is one memset, ``all_reduce()`` one NCCL message,
      ``step()`` one fused kernel (``gcgcn_adam_step``, torch.optim.Adam semantics as at C:300);
    * parameters that never get a gradient (``linears_k``, G:137) stay out of the bucket when ``skip`` names
      them -- their ``.grad`` stays ``None`` like in the reference.

    No CPU path: ``step()`` on CPU tensors raises ``GcgcnError``.
    """

    def __init__(self, module: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, skip: Sequence[str] = ("linears_k",)):
        named = [(n, p) for n, p in module.named_parameters()
                 if p.requires_grad and not any(s in n for s in skip)]
        if not named:
            raise ValueError("no trainable parameters")
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        # every view starts on a 16-byte boundary so the float4 kernel and NCCL see aligned slices
        self.offsets, off = [], 0
        for s in self.sizes:
            self.offsets.append(off)
            off += (s + 3) // 4 * 4
        self.numel = off
        self.flat_params = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)          # gradient bucket
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.steps = 0
        with torch.no_grad():
            for p, o, s in zip(self.params, self.offsets, self.sizes):
                view = self.flat_params[o:o + s].view_as(p)
                view.copy_(p)
                p.data = view
                p.grad = self.flat[o:o + s].view_as(p)

    @property
    def nbytes(self) -> int:
        return self.numel * 4

    def zero_grad(self):
        self.flat.zero_()
        for p, o, s in zip(self.params, self.offsets, self.sizes):      # re-attach if someone set grad = None
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
                p.grad = self.flat[o:o + s].view_as(p)

    def _collect_detached(self):
        """``module.zero_grad()`` / ``optimizer.zero_grad(set_to_none=True)`` drop the views, after which autograd
        allocates fresh ``.grad`` tensors the bucket never sees.  Fold any such gradient into the bucket (and
        re-attach the view) so that all_reduce() / step() never run on a stale, zeroed bucket."""
        base = self.flat.data_ptr()
        for p, o, s in zip(self.params, self.offsets, self.sizes):
            g = p.grad
            if g is not None and g.data_ptr() == base + 4 * o:
                continue
            view = self.flat[o:o + s].view_as(p)
            if g is not None:
                view.add_(g)
            p.grad = view

    def all_reduce(self, group=None, async_op: bool = False):
        self._collect_detached()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def step(self, grad_scale: float = 1.0):
        from . import _lib
        self._collect_detached()
        if not self.flat.is_cuda:
            raise _lib.GcgcnError("FlatTrainer.step: gcgcn_b200 has no CPU path (parameters must live on a B200)")
        self.steps += 1
        stream = torch.cuda.current_stream(self.flat.device).cuda_stream
        _lib.call("gcgcn_adam_step", self.flat_params.data_ptr(), self.flat.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), self.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                  self.weight_decay, grad_scale, self.steps, stream)


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
