"""CPU oracle for the GCGCN entity-graph convolution hot path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU baseline.  Nothing
under ``gcgcn_b200/`` imports it; the product path has no CPU fallback.

It is a functional restatement, in stock fp32 PyTorch on the CPU, of what the
reference computes on this path.  Every function cites the reference lines it
follows (G = /root/reference/models/GCGCN_glove.py, B =
models/GraphCNN_multihead_bert_gate_cls.py -- graph classes byte-identical to G,
C = config/Config.py).  The operations are issued "as written" (expanded n*n
linears, einsum over the full edge tensor, chained matmul) so that low-order bits
track the reference, not the algebraically collapsed forms the CUDA kernels use.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against the reference's own modules executed in the build
container -- ``oracle/pin_against_reference.py`` (bit-exact on forward outputs and
gradients for every function below) -- and against the committed fixtures in
``tests/golden/`` that were generated from the reference by
``tests/golden/make_golden.py``.

Parameters are passed as a flat ``dict`` that uses the reference's own
``state_dict`` key names, e.g. ``linear_node_h.weight`` or
``graphconv.3.weights_node``.

Dropout: torch's RNG stream cannot be reproduced by a CUDA kernel, so every
function takes optional *keep-scale masks* (entries are 0 or 1/(1-p)); ``None``
means eval mode.  They multiply exactly where the reference's ``nn.Dropout``
instances sit.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

HIDDEN = 128  # G:234

# Test hook: when set to a list, every relu of the dense stack appends the smallest |pre-activation|
# it saw.  relu'(x) jumps at 0, so a document whose margin is within float rounding of 0 has
# ill-conditioned gradients (any re-implementation may land on the other side of the kink); the
# tests use this to pick well-conditioned documents for gradient comparisons.
RELU_MARGIN_PROBE = None


def _sub(params: Params, prefix: str) -> Params:
    """Select ``prefix.*`` entries and strip the prefix."""
    cut = len(prefix) + 1
    return {k[cut:]: v for k, v in params.items() if k.startswith(prefix + ".")}


def _affine(x: Tensor, p: Params, name: str) -> Tensor:
    return torch.nn.functional.linear(x, p[name + ".weight"], p[name + ".bias"])


def _chain3(a: Tensor, b: Tensor, c: Tensor) -> Tensor:
    # G:42 calls torch.chain_matmul, which torch implements as linalg.multi_dot
    # (it picks the cheaper association order).
    return torch.linalg.multi_dot([a, b, c])


# --------------------------------------------------------------------------- a3
def graph_conv(x_in: Tensor, edge: Tensor, att: Tensor, w_edge: Tensor, w_node: Tensor,
               bias: Optional[Tensor] = None) -> Tensor:
    """GraphConv.forward, G:36-50.

    out_i = ( mean_j(e_ij W_e) + (A X W_n)_i ) / r_i,  r_i = sum_j A_ij + [sum_j A_ij == 0].
    Both the edge and the node term are divided by r (quirk 4).
    """
    edge_term = torch.einsum("ijk,kp->ijp", edge, w_edge).mean(dim=1)      # G:40-41
    node_term = _chain3(att, x_in, w_node)                                 # G:42
    total = edge_term + node_term                                          # G:43
    if bias is not None:                                                   # G:45-46
        total = total + bias
    r = att.sum(1)                                                         # G:47
    r = r + torch.eq(r, 0).float()                                         # G:48-49
    return total / r.unsqueeze(1).expand_as(total)                         # G:49-50


def _dense_stack(x: Tensor, edge: Tensor, att: Tensor, p: Params, first: int, layers: int,
                 keep: Optional[Sequence[Optional[Tensor]]]) -> Tensor:
    """The densely connected sub-layer stack shared by G:67-76 and G:103-113.

    Returns cat_l(drop(g_l)) + x.  Dropout hits only the copy that is
    concatenated for the output, never the dense-connect cache (G:72-74).
    """
    cache = [x]
    outs = []
    cur = x
    for l in range(layers):
        k = first + l
        pre = graph_conv(cur, edge, att, p[f"graphconv.{k}.weights_edge"],
                         p[f"graphconv.{k}.weights_node"])
        if RELU_MARGIN_PROBE is not None:
            RELU_MARGIN_PROBE.append(float(pre.detach().abs().min()))
        g = torch.relu(pre)                                                # G:71 / G:108
        cache.append(g)                                                    # G:72
        cur = torch.cat(cache, dim=-1)                                     # G:73
        m = None if keep is None else keep[l]
        outs.append(g if m is None else g * m)                             # G:74
    return torch.cat(outs, dim=-1) + x                                     # G:75-76


# --------------------------------------------------------------------------- a4
def caggc_conv(x: Tensor, edge: Tensor, att: Tensor, p: Params, layers: int,
               keep: Optional[Sequence[Optional[Tensor]]] = None) -> Tensor:
    """GraphConvolution.forward (CAGGC convolution), G:63-80."""
    f = _dense_stack(x, edge, att, p, 0, layers, keep)
    return _affine(f, p, "linear_layer")                                   # G:78


# --------------------------------------------------------------------------- a6
def maggc_conv(x: Tensor, edge: Tensor, atts: Sequence[Tensor], p: Params, layers: int,
               heads: int, keep: Optional[Sequence[Sequence[Optional[Tensor]]]] = None) -> Tensor:
    """MultiGraphConvolution.forward (MAGGC convolution), G:97-120."""
    per_head = []
    for h in range(heads):                                                 # G:102
        kh = None if keep is None else keep[h]
        per_head.append(_dense_stack(x, edge, atts[h], p, h * layers, layers, kh))
    return _affine(torch.cat(per_head, -1), p, "linear_layer")             # G:117-118


# --------------------------------------------------------------------------- a5
def mha_attention(x: Tensor, p: Params, heads: int,
                  keep: Optional[Sequence[Optional[Tensor]]] = None) -> List[Tensor]:
    """MultiHeadAttention.forward, G:133-142.

    The key is produced by ``linears_q`` as well (quirk 3): scores are q q^T / sqrt(d_h);
    ``linears_k`` is never used.  The ``mask`` argument is ignored by the reference.
    """
    d_h = p["linears_q.0.weight"].shape[0]
    out = []
    for h in range(heads):
        q = _affine(x, p, f"linears_q.{h}")                                # G:136
        kt = _affine(x, p, f"linears_q.{h}").transpose(0, 1)               # G:137
        a = torch.softmax(torch.mm(q, kt) / math.sqrt(d_h), dim=-1)        # G:138
        if keep is not None and keep[h] is not None:                       # G:139-140
            a = a * keep[h]
        out.append(a)
    return out


# --------------------------------------------------------------------------- a2
def gat_attention(x: Tensor, edge: Tensor, p: Params, mask: Optional[Tensor] = None,
                  keep: Optional[Tensor] = None, apply_mask: bool = False) -> Tensor:
    """GATAttention.forward, G:154-168.

    Both "head" and "tail" node features index the column entity j (quirk 2).
    ``mask`` is accepted but has no effect (quirk 1: the reference calls the
    out-of-place ``masked_fill`` and drops its result, G:164) unless
    ``apply_mask=True``, which restates the evidently intended in-place fill and
    is the oracle for the product's opt-in ``apply_mask`` flag.
    """
    n = x.size(0)
    xh = x.unsqueeze(0).expand(n, n, -1)                                   # G:156
    xt = x.unsqueeze(0).expand(n, n, -1)                                   # G:157
    feat = torch.cat([_affine(xh, p, "linear_node_h"),
                      _affine(xt, p, "linear_node_t"),
                      _affine(edge, p, "linear_edge_r")], -1)              # G:159-162
    energy = _affine(feat, p, "wt").squeeze(-1)                            # G:162
    if mask is not None and apply_mask:
        energy = energy.masked_fill(mask, -100000.0)
    a = torch.softmax(energy, dim=-1)                                      # G:165
    if keep is not None:                                                   # G:166-167
        a = a * keep
    return a


# --------------------------------------------------------------------------- a7
def graph_blocks(x0: Tensor, e0: Tensor, e1: Tensor, adj: Optional[Tensor],
                 gat_p: Params, cag_p: Params, mha_p: Params, mag_p: Params,
                 layers: int, heads: int, alpha: float = 1.0,
                 keep: Optional[dict] = None, apply_mask: bool = False) -> dict:
    """Hop glue, G:329-341 with graph_hop = 2 (C:71).

    hop 0: A = GAT(x0, e0, adj == 0); new = CAGGC(x0, e0, A)
    hop 1: A_h = MHA(y1);            new = MAGGC(y1, e1, A_h)
    after each hop: node_feats.append(old) *before* the update (quirk 5),
    node = alpha*new + (1-alpha)*old, then dropout(0.2).
    Returns y1, y2 and node_feats = cat[x0, x0, y1] (what the classifier sees).
    ``keep`` (train-mode parity) may hold: 'gat' [n,n], 'cag' [L][n,g], 'out0' [n,128],
    'mha' [H][n,n], 'mag' [H][L][n,g], 'out1' [n,128].
    """
    keep = keep or {}
    mask = None if adj is None else torch.eq(adj, 0)                       # G:330
    a0 = gat_attention(x0, e0, gat_p, mask, keep.get("gat"), apply_mask)   # G:332
    new = caggc_conv(x0, e0, a0, cag_p, layers, keep.get("cag"))           # G:333
    feats = [x0, x0]                                                       # G:310, G:338
    y1 = alpha * new + (1 - alpha) * x0                                    # G:339
    if keep.get("out0") is not None:                                       # G:341
        y1 = y1 * keep["out0"]
    a1 = mha_attention(y1, mha_p, heads, keep.get("mha"))                  # G:336
    new = maggc_conv(y1, e1, a1, mag_p, layers, heads, keep.get("mag"))    # G:337
    feats.append(y1)                                                       # G:338
    y2 = alpha * new + (1 - alpha) * y1                                    # G:339
    if keep.get("out1") is not None:
        y2 = y2 * keep["out1"]
    return {"y1": y1, "y2": y2, "a0": a0, "a1": a1,
            "node_feats": torch.cat(feats, 1)}                             # G:344


# --------------------------------------------------------------------------- a1
def build_node_pos(spans: Sequence[Sequence[Sequence[int]]], doc_len: int,
                   max_length: int = 512) -> Tensor:
    """Mention->entity pooling weights, C:169-176 + truncation C:223.

    ``spans[e]`` lists the [start, end) token spans of entity e.  Each span is
    *assigned* 1/(end-start) (later spans overwrite earlier ones where they
    overlap, C:174), then the whole row is scaled by 1/#mentions (C:175).
    Computed in float64 and cast to float32 like the reference.
    """
    w = np.zeros((len(spans), doc_len))
    for e, ms in enumerate(spans):
        for s, t in ms:
            w[e, s:t] = 1.0 / (t - s)
        w[e, :] *= 1.0 / len(ms)
    return torch.FloatTensor(w[:, :max_length])


def pool_nodes(node_pos: Tensor, ctx: Tensor) -> Tensor:
    """node_feat = node_pos x context_output as broadcast-multiply + sum, G:297-298.

    node_pos [n, L], ctx [1, L, d] -> [n, d].
    """
    n, d = node_pos.size(0), ctx.size(-1)
    prod = node_pos.unsqueeze(2).expand(-1, -1, d) * ctx.expand(n, -1, -1)
    return prod.sum(dim=1)


# --------------------------------------------------------------------------- a8
def make_dis2idx() -> np.ndarray:
    """Log-bucket distance table, C:105-116."""
    t = np.zeros(1024, dtype="int64")
    t[1] = 1
    for lo, v in ((2, 2), (4, 3), (8, 4), (16, 5), (32, 6), (64, 7), (128, 8), (256, 9), (512, 10)):
        t[lo:] = v
    return t


def build_node_relative_pos(first_pos: Sequence[int]) -> Tensor:
    """node_relative_pos, C:207-217 + C:225: signed bucket of first-mention offsets."""
    tab = make_dis2idx()
    n = len(first_pos)
    rp = np.zeros((n, n))
    for h in range(n):
        for t in range(n):
            if h == t:
                continue
            d = first_pos[h] - first_pos[t]
            rp[h, t] = -tab[-d] if d < 0 else tab[d]
    return torch.LongTensor(rp)


def pair_gather_classifier(node_feats_with_type: Tensor, rel_pos: Tensor, dis_table: Tensor,
                           dis_plus: int = 10):
    """Classifier-side pair tensors, G:306-307 + G:351-352.

    P_h[i,j] = cat(F[j], dis[dis_plus + rp[i,j]]),  P_t[i,j] = cat(F[i], dis[dis_plus - rp[i,j]])
    (quirk 6: the "h" tensor gathers entity j, the "t" tensor entity i).
    """
    n = node_feats_with_type.size(0)
    rel_h = torch.nn.functional.embedding(dis_plus + rel_pos, dis_table)   # G:306
    rel_t = torch.nn.functional.embedding(dis_plus - rel_pos, dis_table)   # G:307
    p_h = torch.cat([node_feats_with_type.unsqueeze(0).expand(n, -1, -1), rel_h], -1)   # G:351
    p_t = torch.cat([node_feats_with_type.unsqueeze(1).expand(-1, n, -1), rel_t], -1)   # G:352
    return p_h, p_t


def pair_gather_inloop(node_feat: Tensor, slots: int):
    """In-loop pair views fed to SentenceAttention, G:321-322: h[i,j,s]=x[j], t[i,j,s]=x[i]."""
    n, d = node_feat.shape
    shape = (n, n, slots, d)
    emb_h = node_feat.unsqueeze(0).unsqueeze(2).expand(shape)              # G:321
    emb_t = node_feat.unsqueeze(1).unsqueeze(2).expand(shape)              # G:322
    return emb_h, emb_t


def pair_index_tables(rel_pos: Tensor, dis_plus: int = 10):
    """The pure-integer content of the pair gathers: h_idx[i,j]=j, t_idx[i,j]=i,
    dis_h = dis_plus + rp, dis_t = dis_plus - rp (G:306-307, 351-352)."""
    n = rel_pos.size(0)
    ar = torch.arange(n, dtype=torch.int64)
    h_idx = ar.unsqueeze(0).expand(n, n).contiguous()
    t_idx = ar.unsqueeze(1).expand(n, n).contiguous()
    return h_idx, t_idx, dis_plus + rel_pos, dis_plus - rel_pos


def node_feats_with_type(node_feats: Tensor, node_type: Tensor, ner_table: Tensor) -> Tensor:
    """cat[node_feats, ner_emb(node_type)], G:345-347 (ner_emb has padding_idx=0, G:242)."""
    return torch.cat([node_feats, torch.nn.functional.embedding(node_type, ner_table)], 1)


# ----------------------------------------------------------------- whole hot path
def hot_path(ctx: Tensor, node_pos: Tensor, e0: Tensor, e1: Tensor, adj: Tensor,
             node_type: Tensor, rel_pos: Tensor, params: Params, layers: int, heads: int,
             alpha: float = 1.0, keep: Optional[dict] = None) -> dict:
    """pooling -> CAGGC -> MAGGC -> classifier-side pair gathers.

    ``params`` uses the top-level model's key names: get_weighted_adj_matrix.*,
    graphcnn.0.*, get_adj_matrix.0.*, graphcnn.1.*, ner_emb.weight, dis_embed.weight.
    """
    x0 = pool_nodes(node_pos, ctx.unsqueeze(0) if ctx.dim() == 2 else ctx)
    r = graph_blocks(x0, e0, e1, adj,
                     _sub(params, "get_weighted_adj_matrix"), _sub(params, "graphcnn.0"),
                     _sub(params, "get_adj_matrix.0"), _sub(params, "graphcnn.1"),
                     layers, heads, alpha, keep)
    f = node_feats_with_type(r["node_feats"], node_type, params["ner_emb.weight"])
    p_h, p_t = pair_gather_classifier(f, rel_pos, params["dis_embed.weight"])
    r.update(x0=x0, pair_h=p_h, pair_t=p_t)
    return r
