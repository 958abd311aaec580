"""CPU: oracle/edge_oracle.py against the goldens written by the UNMODIFIED reference model
(tests/golden/edge_head.npz, make_golden_edge.py), and the host-side table builder of the edge-feature producer
(gcgcn_b200.edgefeat.EdgeTables) against the dense tensors of the reference's featurisation."""
import numpy as np
import pytest
import torch

from helpers import golden, head_labels, head_shapes, head_state
from oracle import edge_oracle as EO
from oracle import featurize_oracle as FO
from gcgcn_b200 import synthetic as S
from gcgcn_b200.featurize import wire_from_record


def _dense(item):
    t = FO.from_list_to_tensor(item)
    tt = lambda k, dt: torch.from_numpy(np.asarray(t[k])).to(dt)
    return (tt("node_pos", torch.float32), tt("sen_matrix", torch.bool), tt("pos_matrix_h", torch.int64),
            tt("pos_matrix_t", torch.int64), tt("adj_matrix", torch.float32), tt("node_type", torch.int64),
            tt("node_relative_pos", torch.int64))


@pytest.mark.parametrize("idx", [0, 2])
def test_oracle_head_reproduces_the_reference_goldens(idx):
    g = golden("edge_head.npz")
    seed, n, L, Sx, _ = g[f"d{idx}_meta"].tolist()
    item = S.make_record(seed, n=n, L=L, S=None if Sx < 0 else Sx)
    state = {k: v.clone().requires_grad_(True) for k, v in head_state(0).items()}
    ctx = torch.from_numpy(g[f"d{idx}_ctx"]).requires_grad_(True)
    r = EO.graph_head(ctx, *_dense(item), state, 2, 8)
    loss = EO.loss_as_written(r["logits"], head_labels(seed, n))
    loss.backward()
    assert float((r["logits"] - torch.from_numpy(g[f"d{idx}_logits"])).abs().max()) <= 2e-5
    assert float((r["e0"] - torch.from_numpy(g[f"d{idx}_e0"])).abs().max()) <= 2e-6
    assert float((r["e1"] - torch.from_numpy(g[f"d{idx}_e1"])).abs().max()) <= 2e-6
    assert abs(float(loss) - float(g[f"d{idx}_loss"].reshape(-1)[0])) <= 1e-6
    ref = torch.from_numpy(g[f"d{idx}_dctx"])
    assert float((ctx.grad - ref).abs().max()) <= 2e-5 * max(float(ref.abs().max()), 1e-3)
    k = "bili_layer_01.weight"
    got = state[k].grad.reshape(-1)[::997]
    assert float((got - torch.from_numpy(g[f"d{idx}_g_{k}"])).abs().max()) <= 2e-5 * float(np.abs(g[f"d{idx}_g_{k}"]).max())


def test_head_state_covers_exactly_the_reference_head_keys():
    st = head_state(0)
    assert list(st) == list(head_shapes())
    from gcgcn_b200.head import GraphHead
    m = GraphHead()
    res = m.load_state_dict(st, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert float(st["ner_emb.weight"][0].abs().max()) == 0.0


@pytest.mark.parametrize("seed,n,L,Sx", [(8, 7, 150, None), (2, 6, 90, 3), (5, 5, 60, None), (10, 10, 600, None), (3, 5, 60, None)])
def test_edge_tables_list_exactly_the_slots_that_survive_the_token0_mask(seed, n, L, Sx):
    """EdgeTables (host) against the dense tensors: active slots = sen_matrix[:, :, :, 0] (G:302), their lengths and
    position buckets = the rows of pos_matrix_h / _t, denominators = padded-slot counts + 1e-10 (G:206, 213)."""
    from gcgcn_b200.batch import RaggedBatch
    from gcgcn_b200.edgefeat import EdgeTables
    item = S.make_record(seed, n=n, L=L, S=Sx)
    w = wire_from_record(item)
    dense = FO.from_list_to_tensor(item)
    sen, ph, pt = dense["sen_matrix"], dense["pos_matrix_h"], dense["pos_matrix_t"]
    tabs = EdgeTables([w], RaggedBatch([n], "cpu"), None)
    act = np.argwhere(sen[:, :, :, 0])
    assert tabs.num_slots == len(act)
    h = tabs.host
    if len(act) == 0:
        assert tabs.num_tokens == 0 and tabs.num_pairs == 0
        return
    assert tabs.num_tokens == int(h["slot_len"].max())
    for k, (u, v, s) in enumerate(act.tolist()):             # argwhere order == (pair, slot) order of the tables
        assert h["slot_rowi"][k] == u and h["slot_rowj"][k] == v
        ln = int(h["slot_len"][k])
        assert sen[u, v, s, :ln].all() and not sen[u, v, s, ln:].any()
        h0, h1, t0, t1 = h["slot_span"][4 * k:4 * k + 4].tolist()
        tok = np.arange(ln)
        bucket = lambda a, b: np.where(tok < a, -FO.make_dis2idx()[np.maximum(a - tok, 0)],
                                       np.where(tok > b, FO.make_dis2idx()[np.maximum(tok - b, 0)], 0)) + 10
        assert np.array_equal(bucket(h0, h1), ph[u, v, s, :ln]) and np.array_equal(bucket(t0, t1), pt[u, v, s, :ln])
    sent_num = (~sen[:, :, :, 0]).sum(-1)
    pairs = np.argwhere(sen[:, :, :, 0].any(-1))
    assert tabs.num_pairs == len(pairs)
    for k, (u, v) in enumerate(pairs.tolist()):
        assert tabs.pair_idx_host[k] == u * n + v
        assert tabs.pair_denom_host[k] == np.float32(sent_num[u, v]) + np.float32(1e-10)
    # every (slot, side) appears once in the node -> entries table, under the node whose embedding it uses
    ptr, ctr = h["node_ctr_ptr"], h["node_ctr"]
    assert ptr[-1] == 2 * tabs.num_slots and sorted(ctr.tolist()) == list(range(2 * tabs.num_slots))
    for r in range(n):
        for e in ctr[ptr[r]:ptr[r + 1]].tolist():
            s, side = divmod(e, 2)
            assert (h["slot_rowj"][s] if side == 0 else h["slot_rowi"][s]) == r


def test_oracle_bert_variant_reproduces_the_reference_golden():
    g = golden("edge_head_bert.npz")
    seed, n, L, Sx, _ = g["d1_meta"].tolist()
    item = S.make_record(seed, n=n, L=L, S=None if Sx < 0 else Sx)
    state = head_state(0, 4, 4, cls_dim=768)
    with torch.no_grad():
        r = EO.graph_head(torch.from_numpy(g["d1_ctx"]), *_dense(item), state, 4, 4, cls_feat=torch.from_numpy(g["d1_cls_feat"]))
    assert float((r["logits"] - torch.from_numpy(g["d1_logits"])).abs().max()) <= 2e-5
    from gcgcn_b200.head import GraphHead
    res = GraphHead(4, 4, cls_dim=768).load_state_dict(state, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
