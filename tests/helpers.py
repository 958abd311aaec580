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


# ------------------------------------------------------------------------------- graph head
def head_shapes(layers=2, heads=8, hidden=128, dis_size=20, type_size=20, relations=97, hops=2, cls_dim=0):
    """state_dict keys -> shapes of everything the reference model owns after ``context_output`` (G:254-279)."""
    g = hidden // layers
    s = {}
    for nm in ("linear_node_h", "linear_node_t", "linear_edge_r"):
        s[f"get_weighted_adj_matrix.{nm}.weight"] = (hidden, hidden)
        s[f"get_weighted_adj_matrix.{nm}.bias"] = (hidden,)
    s["get_weighted_adj_matrix.wt.weight"] = (1, 3 * hidden)
    s["get_weighted_adj_matrix.wt.bias"] = (1,)
    for kind in ("linears_q", "linears_k"):
        for h in range(heads):
            s[f"get_adj_matrix.0.{kind}.{h}.weight"] = (hidden // heads, hidden)
            s[f"get_adj_matrix.0.{kind}.{h}.bias"] = (hidden // heads,)
    for l in range(layers):
        s[f"graphcnn.0.graphconv.{l}.weights_edge"] = (hidden, g)
        s[f"graphcnn.0.graphconv.{l}.weights_node"] = (hidden + g * l, g)
    s["graphcnn.0.linear_layer.weight"] = (hidden, hidden)
    s["graphcnn.0.linear_layer.bias"] = (hidden,)
    for h in range(heads):
        for l in range(layers):
            s[f"graphcnn.1.graphconv.{h * layers + l}.weights_edge"] = (hidden, g)
            s[f"graphcnn.1.graphconv.{h * layers + l}.weights_node"] = (hidden + g * l, g)
    s["graphcnn.1.linear_layer.weight"] = (hidden, hidden * heads)
    s["graphcnn.1.linear_layer.bias"] = (hidden,)
    for i in range(hops):
        s[f"word_attention.{i}.attention_sent.weight"] = (hidden, hidden)
        s[f"word_attention.{i}.attention_sent.bias"] = (hidden,)
        s[f"word_attention.{i}.attention_pos.weight"] = (hidden, dis_size)
        s[f"word_attention.{i}.attention_pos.bias"] = (hidden,)
        s[f"word_attention.{i}.attention_all.weight"] = (1, hidden)
        s[f"word_attention.{i}.attention_all.bias"] = (1,)
        s[f"sentence_attention.{i}.attention_sent.weight"] = (hidden, hidden)
        s[f"sentence_attention.{i}.attention_sent.bias"] = (hidden,)
        s[f"sentence_attention.{i}.attention_pos.weight"] = (hidden, hidden)
        s[f"sentence_attention.{i}.attention_pos.bias"] = (hidden,)
        s[f"sentence_attention.{i}.attention_all.weight"] = (1, hidden)
        s[f"sentence_attention.{i}.attention_all.bias"] = (1,)
        s[f"linear_word_att.{i}.weight"] = (hidden, 2 * hidden)
        s[f"linear_word_att.{i}.bias"] = (hidden,)
        s[f"linear_sentence_att.{i}.weight"] = (hidden, 2 * hidden)
        s[f"linear_sentence_att.{i}.bias"] = (hidden,)
    s["dense_layer.weight"] = (hidden, hidden * (hops + 1) + dis_size + type_size)
    s["dense_layer.bias"] = (hidden,)
    s["bili_layer_01.weight"] = (relations, hidden, hidden)
    s["bili_layer_01.bias"] = (relations,)
    s["classification_layer_01.weight"] = (relations, 2 * hidden)
    s["classification_layer_01.bias"] = (relations,)
    if cls_dim:                                     # BERT variant only (B:265)
        s["linear_cls.weight"] = (relations, cls_dim)
        s["linear_cls.bias"] = (relations,)
    s["dis_embed.weight"] = (21, dis_size)
    s["ner_emb.weight"] = (7, type_size)
    return s


def head_state(seed=0, layers=2, heads=8, cls_dim=0):
    """Deterministic weights for every head parameter (uniform, fan-in scaled), independent of any module's init
    order: tests/golden/make_golden_edge.py loads the same values into the UNMODIFIED reference model, the GPU tests
    load them into the drop-in modules.  Embedding row 0 of ner_emb is the padding row (G:242)."""
    gen = torch.Generator().manual_seed(4321 + seed)
    out = {}
    for name, shape in head_shapes(layers, heads, cls_dim=cls_dim).items():
        if len(shape) == 1:
            bound = 0.1
        elif name.endswith("weights_node") or name.endswith("weights_edge"):
            bound = (6.0 / (shape[0] + shape[1])) ** 0.5
        elif name in ("dis_embed.weight", "ner_emb.weight"):
            bound = 0.5
        else:
            bound = 1.0 / shape[-1] ** 0.5
        out[name] = (torch.rand(shape, generator=gen) * 2 - 1) * bound
    out["ner_emb.weight"][0] = 0.0
    return out


def head_labels(seed, n, relations=97):
    """Sparse synthetic multi-hot labels [n, n, R] float32 (about 2 % positives), zero diagonal."""
    gen = torch.Generator().manual_seed(555 + seed)
    lab = (torch.rand(n, n, relations, generator=gen) < 0.02).float()
    lab[torch.arange(n), torch.arange(n)] = 0.0
    return lab
