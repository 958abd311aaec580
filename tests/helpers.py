"""Shared test helpers: oracle wiring, golden loading, comparison utilities."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from oracle import gcgcn_oracle as O
from gcgcn_b200 import synthetic as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = {"glove": (2, 8), "bert": (4, 4)}     # (layer_num, head_num): G:250-251, B:247-248
PREFIXES = ("get_weighted_adj_matrix", "graphcnn.0", "get_adj_matrix.0", "graphcnn.1")

FP32_TOL = 1e-4     # BASELINE.json north_star: <= 1e-4 abs on block outputs and gradients in fp32
BF16_TOL = 2e-2     # <= 2e-2 in bf16 (vs the fp32 oracle)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def sub(params, prefix):
    cut = len(prefix) + 1
    return {k[cut:]: v for k, v in params.items() if k.startswith(prefix + ".")}


def blocks_state(layers, heads, seed=0):
    """CPU state_dict of a GraphBlocks under manual_seed(seed) (== the reference's init)."""
    from gcgcn_b200.modules import GraphBlocks
    torch.manual_seed(seed)
    gb = GraphBlocks(layers, heads)
    return gb, {k: v.detach().clone() for k, v in gb.state_dict().items()}


def state_sha256(state) -> str:
    """Hash in the order make_golden.py used: gat, mha, cag, mag; keys relative to each module."""
    h = hashlib.sha256()
    for prefix in ("get_weighted_adj_matrix", "get_adj_matrix.0", "graphcnn.0", "graphcnn.1"):
        for k, v in sub(state, prefix).items():
            h.update(k.encode())
            h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def upstream(doc_id, shape1, shape2):
    gen = torch.Generator().manual_seed(99 + doc_id)
    return torch.randn(shape1, generator=gen), torch.randn(shape2, generator=gen)


def oracle_blocks(doc, state, layers, heads, keep=None, backward=True, apply_mask=False, device=None):
    """Run the oracle hop glue on one document; returns outputs and gradients (CPU; `device` runs the same
    ATen ops elsewhere -- bench.py's "reference PyTorch path on the B200" leg)."""
    dev = torch.device("cpu") if device is None else torch.device(device)
    ps = {k: v.detach().clone().to(dev).requires_grad_(backward) for k, v in state.items()}
    x0 = doc.x0.float().clone().to(dev).requires_grad_(backward)
    e0 = doc.e0.float().clone().to(dev).requires_grad_(backward)
    e1 = doc.e1.float().clone().to(dev).requires_grad_(backward)
    r = O.graph_blocks(x0, e0, e1, doc.adj.to(dev), sub(ps, PREFIXES[0]), sub(ps, PREFIXES[1]), sub(ps, PREFIXES[2]),
                       sub(ps, PREFIXES[3]), layers, heads, 1.0, keep, apply_mask)
    out = {k: (v.detach() if torch.is_tensor(v) else [t.detach() for t in v]) for k, v in r.items()}
    if backward:
        dy1, dy2 = upstream(doc.doc_id, r["y1"].shape, r["y2"].shape)
        ((r["y1"] * dy1.to(dev)).sum() + (r["y2"] * dy2.to(dev)).sum()).backward()
        out.update(dx0=x0.grad, de0=e0.grad, de1=e1.grad,
                   dparams={k: v.grad for k, v in ps.items()})
    return out


def relu_margin(doc, state, layers, heads) -> float:
    """Smallest |pre-activation| over every relu of the two blocks (oracle, eval mode)."""
    O.RELU_MARGIN_PROBE = []
    try:
        oracle_blocks(doc, state, layers, heads, backward=False)
        return min(O.RELU_MARGIN_PROBE)
    finally:
        O.RELU_MARGIN_PROBE = None


def maxdiff(a, b) -> float:
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) if a.numel() else 0.0


def assert_close(a, b, tol, what=""):
    d = maxdiff(a, b)
    assert d <= tol, f"{what}: max|diff| = {d:.3e} > {tol:.1e}"
