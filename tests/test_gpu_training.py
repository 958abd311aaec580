"""-m gpu: config 5's training step -- gradients accumulated straight into the flat bucket, fused Adam
(gcgcn_adam_step) against torch.optim.Adam as the reference's trainer builds it (config/Config.py:300)."""
import pytest
import torch

from helpers import maxdiff, upstream
from gcgcn_b200 import synthetic as S
from gcgcn_b200.batch import RaggedBatch
from gcgcn_b200.sharding import FlatTrainer
from gpu_common import DEV, cat_inputs, device_blocks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("count", [1, 3, 4, 1000, 545921])
def test_adam_kernel_matches_torch_adam(count):
    from gcgcn_b200 import _lib
    g = torch.Generator().manual_seed(count)
    p0 = torch.randn(count, generator=g)
    ref = torch.nn.Parameter(p0.clone().to(DEV))
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-2)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    stream = torch.cuda.current_stream().cuda_stream
    for step in range(1, 6):
        grad = torch.randn(count, generator=g).to(DEV)
        ref.grad = grad.clone() * 0.5
        opt.step()
        _lib.call("gcgcn_adam_step", p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), count,
                  1e-3, 0.9, 0.999, 1e-8, 1e-2, 0.5, step, stream)
    torch.cuda.synchronize()
    assert maxdiff(p, ref.detach()) <= 2e-6
    st = opt.state[ref]
    assert maxdiff(m, st["exp_avg"]) <= 1e-6 and maxdiff(v, st["exp_avg_sq"]) <= 1e-6


def test_flat_trainer_steps_like_torch_adam_on_the_hot_path():
    """Two optimiser steps over two micro-batches: the hot path's backward lands in the flat bucket in place
    (no pack step), and the parameters after FlatTrainer.step equal those of torch.optim.Adam fed the same
    gradients; linears_k never moves and keeps grad None (G:137)."""
    gb, _ = device_blocks(2, 8)
    twin, _ = device_blocks(2, 8)
    tw = dict(twin.named_parameters())
    opt = torch.optim.Adam([p for n, p in twin.named_parameters() if "linears_k" not in n], lr=1e-3)
    tr = FlatTrainer(gb, lr=1e-3)
    assert tr.numel >= 545921 and all(p.grad is not None for p in tr.params)
    k_before = {n: p.detach().clone() for n, p in gb.named_parameters() if "linears_k" in n}
    for docs in (S.make_batch((3, 7, 9)), S.make_batch((1, 10))):
        bt = RaggedBatch([d.n for d in docs], DEV)
        ups = [upstream(d.doc_id, (d.n, 128), (d.n, 128)) for d in docs]
        dy1 = torch.cat([u[0] for u in ups]).to(DEV)
        dy2 = torch.cat([u[1] for u in ups]).to(DEV)
        grads = {}
        for model in (gb, twin):
            tr.zero_grad() if model is gb else opt.zero_grad()
            x0, e0, e1, _ = cat_inputs(docs)
            out = model(x0, e0, e1, bt)
            torch.autograd.backward([out["y1"], out["y2"]], [dy1, dy2])
            grads[model] = {n: p.grad for n, p in model.named_parameters()}
        lo, hi = tr.flat.data_ptr(), tr.flat.data_ptr() + tr.nbytes
        for n, p in gb.named_parameters():
            if "linears_k" in n:
                assert p.grad is None and tw[n].grad is None
                continue
            assert lo <= p.grad.data_ptr() < hi, n                      # the bucket received it in place
            # same kernels, same inputs: the two passes agree (to fp32 noise where a gradient is analytically 0,
            # e.g. the GAT biases, which only move the softmax-invariant constant c)
            assert maxdiff(p.grad, tw[n].grad) <= 1e-4 * max(1.0, float(p.grad.abs().max())), n
            tw[n].grad = p.grad.detach().clone() * (1.0 / len(docs))   # feed torch's Adam the bucket's gradients
        opt.step()
        tr.step(grad_scale=1.0 / len(docs))
    torch.cuda.synchronize()
    for n, p in gb.named_parameters():
        if "linears_k" in n:
            assert torch.equal(p.detach(), k_before[n])
        else:
            assert maxdiff(p.detach(), tw[n].detach()) <= 5e-6, n
    assert tr.steps == 2
