"""-m gpu: pooling -> CAGGC -> MAGGC -> pair gathers end to end against
oracle.hot_path, including the relation-argmax criterion of the north star (>= 99.9 % of pairs)."""
import pytest
import torch

from helpers import FP32_TOL, assert_close
from gpu_common import DEV, device_blocks
from oracle import gcgcn_oracle as O
from gcgcn_b200 import synthetic as S
from gcgcn_b200.batch import PairTables, PoolTable, RaggedBatch, node_relative_pos
from gcgcn_b200.modules import pair_gather, pool_nodes

pytestmark = pytest.mark.gpu


def _classifier(gen):
    """Random stand-in for dense_layer + Bilinear + Linear (G:272-276, 354-358)."""
    return {"dense.w": torch.randn(128, 424, generator=gen) * 0.05, "dense.b": torch.randn(128, generator=gen) * 0.05,
            "bili.w": torch.randn(97, 128, 128, generator=gen) * 0.05, "bili.b": torch.randn(97, generator=gen) * 0.05,
            "cls.w": torch.randn(97, 256, generator=gen) * 0.05, "cls.b": torch.randn(97, generator=gen) * 0.05}


def _logits(ph, pt, c):
    fh = torch.tanh(torch.nn.functional.linear(ph, c["dense.w"], c["dense.b"]))
    ft = torch.tanh(torch.nn.functional.linear(pt, c["dense.w"], c["dense.b"]))
    return (torch.nn.functional.bilinear(fh, ft, c["bili.w"], c["bili.b"])
            + torch.nn.functional.linear(torch.cat([fh, ft], -1), c["cls.w"], c["cls.b"]))


def test_whole_hot_path_and_relation_argmax():
    gb, state = device_blocks(2, 8)
    gen = torch.Generator().manual_seed(11)
    ner = torch.randn(7, 20, generator=gen)
    ner[0] = 0
    dis = torch.randn(21, 20, generator=gen)
    cls = _classifier(gen)
    docs = S.make_batch()
    bt = RaggedBatch([d.n for d in docs], DEV)
    rps = [node_relative_pos(d.first_pos) for d in docs]
    pool = PoolTable.from_spans([d.spans for d in docs], [d.L for d in docs], device=DEV)
    pairs = PairTables(bt, rps, device=DEV)
    ctx = torch.cat([d.ctx for d in docs]).to(DEV)
    e0 = torch.cat([d.e0.reshape(-1, 128) for d in docs]).to(DEV)
    e1 = torch.cat([d.e1.reshape(-1, 128) for d in docs]).to(DEV)
    node_type = torch.cat([d.node_type for d in docs]).to(DEV)
    with torch.no_grad():
        x0 = pool_nodes(ctx, pool)
        out = gb(x0, e0, e1, bt)
        feats = torch.cat([out["node_feats"], ner.to(DEV)[node_type]], 1)       # G:345-347
        ph, pt = pair_gather(feats, dis.to(DEV), pairs, bt)
    params = dict(state)
    params["ner_emb.weight"], params["dis_embed.weight"] = ner, dis
    agree = total = 0
    for b, d in enumerate(docs):
        with torch.no_grad():
            r = O.hot_path(d.ctx, O.build_node_pos(d.spans, d.L), d.e0, d.e1, d.adj, d.node_type,
                           torch.from_numpy(rps[b]), params, 2, 8)
        lo, hi = int(bt.pair_ptr_host[b]), int(bt.pair_ptr_host[b + 1])
        n0 = int(bt.node_ptr_host[b])
        assert_close(x0[n0:n0 + d.n], r["x0"], 2e-6, "pooled x0")
        assert_close(out["y1"][n0:n0 + d.n], r["y1"], FP32_TOL, "y1")
        assert_close(out["y2"][n0:n0 + d.n], r["y2"], FP32_TOL, "y2")
        assert_close(ph[lo:hi], r["pair_h"].reshape(-1, 424), FP32_TOL, "pair_h")
        assert_close(pt[lo:hi], r["pair_t"].reshape(-1, 424), FP32_TOL, "pair_t")
        mine = _logits(ph[lo:hi].cpu(), pt[lo:hi].cpu(), cls).argmax(-1)
        ref = _logits(r["pair_h"].reshape(-1, 424), r["pair_t"].reshape(-1, 424), cls).argmax(-1)
        agree += int((mine == ref).sum())
        total += mine.numel()
    assert agree / total >= 0.999, f"relation argmax agreement {agree}/{total}"
