"""-m gpu: pair_dense (tanh(U[j] + Vd[10 + rp]) / tanh(U[i] + Vd[10 - rp]), gcgcn_pair_dense_fwd/bwd) against the
reference lines G:351-355 executed literally on the CPU (expand + cat + dense_layer + tanh), values and every
gradient; and against the product's own gather route tanh(dense_layer(pair_gather(...)))."""
import pytest
import torch
import torch.nn as nn

from helpers import FP32_TOL, assert_close
from gcgcn_b200 import synthetic as S
from gcgcn_b200.batch import PairTables, RaggedBatch, node_relative_pos
from gcgcn_b200.modules import pair_dense, pair_gather

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference_entity_features(feats, rp, dis_embed, dense):
    """G:306-307 + G:351-355, one document."""
    n = feats.size(0)
    rel_h = dis_embed[10 + rp]
    rel_t = dis_embed[10 - rp]
    ph = torch.cat([feats.unsqueeze(0).expand(n, -1, -1), rel_h], -1)
    pt = torch.cat([feats.unsqueeze(1).expand(-1, n, -1), rel_t], -1)
    return torch.tanh(dense(ph)), torch.tanh(dense(pt))


@pytest.mark.parametrize("doc_ids", [(2, 7, 11), (0,), (4, 4, 9, 10)])
def test_pair_dense_matches_the_reference_lines(doc_ids):
    docs = S.make_batch(doc_ids)
    bt = RaggedBatch([d.n for d in docs], DEV)
    gen = torch.Generator().manual_seed(17)
    dense = nn.Linear(424, 128)
    with torch.no_grad():
        dense.weight.copy_(torch.randn(128, 424, generator=gen) * 0.05)
        dense.bias.copy_(torch.randn(128, generator=gen) * 0.05)
    dis = (torch.randn(21, 20, generator=gen) * 0.5).requires_grad_(True)
    feats = [torch.randn(d.n, 404, generator=gen).requires_grad_(True) for d in docs]
    rps = [node_relative_pos(d.first_pos) for d in docs]
    gh = [torch.randn(d.n, d.n, 128, generator=gen) for d in docs]
    gt = [torch.randn(d.n, d.n, 128, generator=gen) for d in docs]
    refs = []
    for f, rp, a, b in zip(feats, rps, gh, gt):
        eh, et = reference_entity_features(f, torch.from_numpy(rp), dis, dense)
        ((eh * a).sum() + (et * b).sum()).backward()
        refs.append((eh.detach(), et.detach()))

    dense_dev = nn.Linear(424, 128).to(DEV)
    dense_dev.load_state_dict(dense.state_dict())
    dis_dev = dis.detach().clone().to(DEV).requires_grad_(True)
    feat_dev = torch.cat([f.detach() for f in feats]).to(DEV).requires_grad_(True)
    tabs = PairTables(bt, rps, device=DEV)
    eh, et = pair_dense(feat_dev, dense_dev, dis_dev, tabs, bt)
    up_h = torch.cat([a.reshape(-1, 128) for a in gh]).to(DEV)
    up_t = torch.cat([b.reshape(-1, 128) for b in gt]).to(DEV)
    ((eh * up_h).sum() + (et * up_t).sum()).backward()
    for b, (d, (rh, rt)) in enumerate(zip(docs, refs)):
        lo, hi = int(bt.pair_ptr_host[b]), int(bt.pair_ptr_host[b + 1])
        assert_close(eh[lo:hi], rh.reshape(-1, 128), FP32_TOL, f"entity_feature_h doc{b}")
        assert_close(et[lo:hi], rt.reshape(-1, 128), FP32_TOL, f"entity_feature_t doc{b}")
        n0 = int(bt.node_ptr_host[b])
        assert_close(feat_dev.grad[n0:n0 + d.n], feats[b].grad, 2 * FP32_TOL, f"dF doc{b}")
    assert_close(dis_dev.grad, dis.grad, 2e-3, "d dis_embed (sums over ~1e3 pairs per row)")
    assert_close(dense_dev.weight.grad, dense.weight.grad, 2e-3, "d dense_layer.weight")
    assert_close(dense_dev.bias.grad, dense.bias.grad, 2e-3, "d dense_layer.bias")

    # same values as the gather route of this package (the tensors pair_dense never forms)
    ph, pt = pair_gather(feat_dev.detach(), dis_dev.detach(), tabs, bt)
    assert_close(eh, torch.tanh(dense_dev(ph)), FP32_TOL, "vs tanh(dense_layer(pair_gather)) h")
    assert_close(et, torch.tanh(dense_dev(pt)), FP32_TOL, "vs tanh(dense_layer(pair_gather)) t")


def test_pair_dense_rejects_cpu_tensors():
    from gcgcn_b200 import _lib
    docs = S.make_batch((11,))
    bt = RaggedBatch([d.n for d in docs], DEV)
    tabs = PairTables(bt, [node_relative_pos(d.first_pos) for d in docs], device=DEV)
    with pytest.raises(_lib.GcgcnError):
        pair_dense(torch.randn(docs[0].n, 404), nn.Linear(424, 128), torch.randn(21, 20), tabs, bt)
