"""-m gpu: the edge-feature producer (WordAttention + SentenceAttention + their linears,
G:171-214 as used at G:299-327) on the GPU from the wire format, against

  * the goldens written by the UNMODIFIED reference model (tests/golden/edge_head.npz, make_golden_edge.py), and
  * oracle/edge_oracle.py (the dense restatement, pinned to the reference) on the same inputs, values and every
    gradient.  Tolerance: 1e-4 absolute on context_sent_att, and 1e-4 relative to the largest entry for gradients
    (they sum hundreds of terms).
"""
import numpy as np
import pytest
import torch

from helpers import FP32_TOL, golden, head_state, sub
from oracle import edge_oracle as EO
from oracle import featurize_oracle as FO
from gcgcn_b200 import synthetic as S
from gcgcn_b200.batch import RaggedBatch
from gcgcn_b200.edgefeat import EdgeFeatures, EdgeTables
from gcgcn_b200.featurize import wire_from_record

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
EDGE_KEYS = ("word_attention", "sentence_attention", "linear_word_att", "linear_sentence_att")


def producer(state):
    ef = EdgeFeatures()
    own = {k: v for k, v in state.items() if k.split(".")[0] in EDGE_KEYS}
    ef.load_state_dict(own, strict=True)              # the reference's key names load unchanged
    return ef.to(DEV)


def rel_close(a, b, tol, what):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(float(b.abs().max()), 1e-2)      # (softmax-shift biases have a true gradient of 0: rounding noise only)
    d = float((a - b).abs().max()) / scale
    assert d <= tol, f"{what}: max|diff| / max|ref| = {d:.3e} > {tol:.1e}"


def oracle_edges(hop, ctx, node_feat, dense, state, upstream):
    """oracle value + gradients of sum(e * upstream) for one document (CPU)."""
    ps = {k: v.clone().requires_grad_(True) for k, v in state.items()
          if k.split(".")[0] in EDGE_KEYS or k == "dis_embed.weight"}
    c = ctx.clone().requires_grad_(True)
    x = node_feat.clone().requires_grad_(True)
    tt = lambda k, dt: torch.from_numpy(np.asarray(dense[k])).to(dt)
    e = EO.edge_features(c, x, tt("sen_matrix", torch.bool), tt("pos_matrix_h", torch.int64),
                         tt("pos_matrix_t", torch.int64), ps["dis_embed.weight"], sub(ps, f"word_attention.{hop}"),
                         sub(ps, f"sentence_attention.{hop}"), sub(ps, f"linear_word_att.{hop}"),
                         sub(ps, f"linear_sentence_att.{hop}"))
    (e * upstream).sum().backward()
    return e.detach(), c.grad, x.grad, {k: v.grad for k, v in ps.items() if v.grad is not None}


def gpu_edges(ef, hop, ctx, node_feat, dis, tabs, upstream):
    c = ctx.to(DEV).requires_grad_(True)
    x = node_feat.to(DEV).requires_grad_(True)
    d = dis.to(DEV).requires_grad_(True)
    ef.zero_grad()
    e = ef(hop, c, x, d, tabs)
    (e * upstream.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in ef.named_parameters() if p.grad is not None}
    grads["dis_embed.weight"] = d.grad
    return e.detach(), c.grad, x.grad, grads


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_edge_features_match_the_reference_goldens(idx):
    g = golden("edge_head.npz")
    seed, n, L, Sx, active = g[f"d{idx}_meta"].tolist()
    item = S.make_record(seed, n=n, L=L, S=None if Sx < 0 else Sx)
    w = wire_from_record(item)
    state = head_state(0)
    ef = producer(state)
    bt = RaggedBatch([n], DEV)
    tabs = EdgeTables([w], bt, DEV)
    assert tabs.num_slots == active
    ctx = torch.from_numpy(g[f"d{idx}_ctx"])
    dense = FO.from_list_to_tensor(item)
    x0 = torch.from_numpy(dense["node_pos"]) @ ctx                      # hop-0 node features (G:297-298)
    with torch.no_grad():
        e0 = ef(0, ctx.to(DEV), x0.to(DEV), state["dis_embed.weight"].to(DEV), tabs)
    assert float((e0.cpu().view(n, n, 128) - torch.from_numpy(g[f"d{idx}_e0"])).abs().max()) <= FP32_TOL
    # hop 1 consumes y1; the oracle head (pinned to the reference) supplies it
    tt = lambda k, dt: torch.from_numpy(np.asarray(dense[k])).to(dt)
    with torch.no_grad():
        r = EO.graph_head(ctx, tt("node_pos", torch.float32), tt("sen_matrix", torch.bool), tt("pos_matrix_h", torch.int64),
                          tt("pos_matrix_t", torch.int64), tt("adj_matrix", torch.float32), tt("node_type", torch.int64),
                          tt("node_relative_pos", torch.int64), state, 2, 8)
        e1 = ef(1, ctx.to(DEV), r["y1"].to(DEV), state["dis_embed.weight"].to(DEV), tabs)
    assert float((e1.cpu().view(n, n, 128) - torch.from_numpy(g[f"d{idx}_e1"])).abs().max()) <= FP32_TOL


@pytest.mark.parametrize("seed,n,L,Sx,hop", [(8, 7, 150, None, 0), (2, 6, 90, 3, 1), (10, 9, 230, None, 0), (4, 9, 230, None, 1)])
def test_edge_features_values_and_gradients_against_the_oracle(seed, n, L, Sx, hop):
    item = S.make_record(seed, n=n, L=L, S=Sx)
    w = wire_from_record(item)
    dense = FO.from_list_to_tensor(item)
    state = head_state(1)
    gen = torch.Generator().manual_seed(seed)
    ctx = torch.tanh(torch.randn(w.length, 128, generator=gen))
    node_feat = torch.tanh(torch.randn(n, 128, generator=gen))
    up = torch.randn(n, n, 128, generator=gen)
    eo, dco, dxo, go = oracle_edges(hop, ctx, node_feat, dense, state, up)
    ef = producer(state)
    bt = RaggedBatch([n], DEV)
    tabs = EdgeTables([w], bt, DEV)
    assert tabs.num_slots > 0
    e, dc, dx, gg = gpu_edges(ef, hop, ctx, node_feat, state["dis_embed.weight"], tabs, up.view(-1, 128))
    assert float((e.cpu().view(n, n, 128) - eo).abs().max()) <= FP32_TOL
    rel_close(dc, dco, 1e-4, "dctx")
    rel_close(dx, dxo, 1e-4, "dnode_feat")
    assert len(go) == 17                                  # 16 parameters of the hop + dis_embed.weight
    for k, v in go.items():
        rel_close(gg[k], v, 1e-4, k)
    # the other hop's parameters are untouched
    other = 1 - hop
    assert all(p.grad is None for k, p in ef.named_parameters() if f".{other}." in k)


def test_ragged_batch_equals_per_document_runs():
    recs = [(8, 7, 150, None), (3, 5, 60, None), (2, 6, 90, 3), (18, 9, 230, None)]      # (3, 5, 60): no active slot
    items = [S.make_record(s, n=n, L=L, S=Sx) for s, n, L, Sx in recs]
    wires = [wire_from_record(it) for it in items]
    state = head_state(2)
    ef = producer(state)
    gen = torch.Generator().manual_seed(5)
    ctxs = [torch.tanh(torch.randn(w.length, 128, generator=gen)) for w in wires]
    xs = [torch.tanh(torch.randn(w.n, 128, generator=gen)) for w in wires]
    ups = [torch.randn(w.n * w.n, 128, generator=gen) for w in wires]
    bt = RaggedBatch([w.n for w in wires], DEV)
    tabs = EdgeTables(wires, bt, DEV)
    e, dc, dx, gg = gpu_edges(ef, 0, torch.cat(ctxs), torch.cat(xs), state["dis_embed.weight"], tabs, torch.cat(ups))
    gsum = None
    eo, dco, dxo = [], [], []
    for w, it, c, x, u in zip(wires, items, ctxs, xs, ups):
        o = oracle_edges(0, c, x, FO.from_list_to_tensor(it), state, u.view(w.n, w.n, 128))
        eo.append(o[0].reshape(-1, 128)); dco.append(o[1]); dxo.append(o[2])
        gsum = o[3] if gsum is None else {k: gsum[k] + v for k, v in o[3].items()}
    assert float((e.cpu() - torch.cat(eo)).abs().max()) <= FP32_TOL
    rel_close(dc, torch.cat(dco), 1e-4, "dctx")
    rel_close(dx, torch.cat(dxo), 1e-4, "dnode_feat")
    for k, v in gsum.items():
        if ".1." in k:
            continue
        rel_close(gg[k], v, 1e-4, k)


def test_document_without_active_slots_gets_the_bias_and_bf16_storage():
    # record whose first sentence holds no entity pair: context_sent_att = linear_sentence_att.bias for every pair
    for seed in range(40):
        item = S.make_record(seed, n=5, L=80)
        w = wire_from_record(item)
        if not any(r[3] <= 0 < min(r[4], w.length) and r[2] < w.max_num for r in w.slots.tolist()):
            break
    else:
        pytest.skip("no such record")
    state = head_state(0)
    ef = producer(state)
    bt = RaggedBatch([w.n], DEV)
    tabs = EdgeTables([w], bt, DEV)
    assert tabs.num_slots == 0
    ctx = torch.randn(w.length, 128).to(DEV).requires_grad_(True)
    x = torch.randn(w.n, 128).to(DEV)
    e = ef(0, ctx, x, state["dis_embed.weight"].to(DEV), tabs)
    assert torch.equal(e.cpu(), state["linear_sentence_att.0.bias"].expand(w.n * w.n, 128))
    e.sum().backward()
    assert torch.allclose(ef.linear_sentence_att[0].bias.grad.cpu(), torch.full((128,), float(w.n * w.n)))
    eb = ef(0, ctx, x, state["dis_embed.weight"].to(DEV), tabs, edge_dtype=torch.bfloat16)
    assert eb.dtype == torch.bfloat16
    assert float((eb.detach().float().cpu() - e.detach().cpu()).abs().max()) <= 2e-2


def test_pair_with_every_slot_active_divides_by_1e_minus_10():
    """G:213 divides by (sent_num + 1e-10) where sent_num counts the PADDED slots: a pair whose slots all contain
    token 0 is scaled by 1e10, in the reference and here."""
    import networkx as nx
    g = nx.DiGraph()
    g.add_node(0, exist_pos=[(1, 3)], type=[1])
    g.add_node(1, exist_pos=[(5, 6)], type=[2])
    g.add_node(2, exist_pos=[(12, 14)], type=[3])
    g.add_edge(0, 1, sentences=[(0, 9)], position=[(1, 3, 5, 6)])
    g.add_edge(1, 0, sentences=[(0, 9)], position=[(5, 6, 1, 3)])
    g.graph["max_sentence_num"] = 1
    item = {"document": list(range(20)), "graph": g}
    w = wire_from_record(item)
    dense = FO.from_list_to_tensor(item)
    state = head_state(3)
    gen = torch.Generator().manual_seed(0)
    ctx, x = torch.tanh(torch.randn(20, 128, generator=gen)), torch.tanh(torch.randn(3, 128, generator=gen))
    up = torch.randn(3, 3, 128, generator=gen)
    eo, dco, dxo, go = oracle_edges(0, ctx, x, dense, state, up)
    ef = producer(state)
    tabs = EdgeTables([w], RaggedBatch([3], DEV), DEV)
    assert tabs.num_pairs == 2 and float(tabs.pair_denom_host[0]) == np.float32(1e-10)
    e, dc, dx, gg = gpu_edges(ef, 0, ctx, x, state["dis_embed.weight"], tabs, up.view(-1, 128))
    rel_close(e.cpu().view(3, 3, 128), eo, 1e-4, "e")
    rel_close(dc, dco, 1e-4, "dctx")
