"""Helpers shared by the -m gpu tests: run the CUDA path through the drop-in API."""
from __future__ import annotations

import torch

from helpers import upstream
from gcgcn_b200.batch import RaggedBatch
from gcgcn_b200.modules import GraphBlocks

DEV = torch.device("cuda:0")


def device_blocks(layers, heads, seed=0):
    """GraphBlocks on the GPU + its CPU state_dict (for the oracle)."""
    torch.manual_seed(seed)
    gb = GraphBlocks(layers, heads)
    state = {k: v.detach().clone() for k, v in gb.state_dict().items()}
    return gb.to(DEV).eval(), state


def cat_inputs(docs, edge_dtype=torch.float32, requires_grad=True):
    x0 = torch.cat([d.x0.float() for d in docs]).to(DEV).requires_grad_(requires_grad)
    e0 = torch.cat([d.e0.reshape(-1, 128) for d in docs]).to(DEV, edge_dtype).requires_grad_(requires_grad)
    e1 = torch.cat([d.e1.reshape(-1, 128) for d in docs]).to(DEV, edge_dtype).requires_grad_(requires_grad)
    adj = torch.cat([d.adj.reshape(-1) for d in docs]).to(DEV)
    return x0, e0, e1, adj


def run_blocks(gb, docs, edge_dtype=torch.float32, backward=True, with_adj=True):
    """Batched forward (+ backward with the shared upstream gradients) on the GPU."""
    bt = RaggedBatch([d.n for d in docs], DEV)
    x0, e0, e1, adj = cat_inputs(docs, edge_dtype, backward)
    gb.zero_grad()
    out = gb(x0, e0, e1, bt, adj if with_adj else None)
    res = {"bt": bt, "y1": out["y1"].detach(), "y2": out["y2"].detach(), "a0": out["a0"].detach(),
           "a1": out["a1"].detach(), "node_feats": out["node_feats"].detach()}
    if backward:
        ups = [upstream(d.doc_id, (d.n, 128), (d.n, 128)) for d in docs]
        dy1 = torch.cat([u[0] for u in ups]).to(DEV)
        dy2 = torch.cat([u[1] for u in ups]).to(DEV)
        ((out["y1"] * dy1).sum() + (out["y2"] * dy2).sum()).backward()
        res.update(dx0=x0.grad, de0=e0.grad, de1=e1.grad,
                   dparams={k: (None if p.grad is None else p.grad.detach().clone())
                            for k, p in gb.named_parameters()})
    torch.cuda.synchronize()
    return res
