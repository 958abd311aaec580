"""-m gpu: each drop-in module, called per document with the reference's own forward
signature (G:36, 63, 97, 133, 154), against the matching oracle function."""
import pytest
import torch

from helpers import FP32_TOL, PREFIXES, VARIANTS, assert_close, maxdiff, sub
from gpu_common import DEV, device_blocks
from oracle import gcgcn_oracle as O
from gcgcn_b200 import synthetic as S
from gcgcn_b200 import modules as M

pytestmark = pytest.mark.gpu
DOCS = (0, 6, 10, 11)


def _grads(tensors):
    return [None if t.grad is None else t.grad.detach().cpu() for t in tensors]


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_gat_attention_forward_backward(variant):
    layers, heads = VARIANTS[variant]
    gb, state = device_blocks(layers, heads)
    gat = gb.get_weighted_adj_matrix
    for i in DOCS:
        d = S.make_doc(i)
        x, e = d.x0.to(DEV).requires_grad_(True), d.e0.to(DEV).requires_grad_(True)
        gat.zero_grad()
        a = gat(x, e, torch.eq(d.adj, 0).to(DEV))
        assert a.shape == (d.n, d.n)
        g = torch.randn(d.n, d.n, generator=torch.Generator().manual_seed(i))
        (a * g.to(DEV)).sum().backward()
        ps = {k: v.clone().requires_grad_(True) for k, v in sub(state, PREFIXES[0]).items()}
        xo, eo = d.x0.clone().requires_grad_(True), d.e0.clone().requires_grad_(True)
        ao = O.gat_attention(xo, eo, ps, torch.eq(d.adj, 0))
        (ao * g).sum().backward()
        assert_close(a, ao, 1e-5, "A")
        assert_close(x.grad, xo.grad, FP32_TOL, "dx")
        assert_close(e.grad, eo.grad, FP32_TOL, "de")
        for k, p in gat.named_parameters():
            assert_close(p.grad, ps[k].grad, FP32_TOL, "d" + k)


def test_gat_mask_is_ignored_by_default_and_applied_on_request():
    gb, state = device_blocks(2, 8)
    gat = gb.get_weighted_adj_matrix
    d = S.make_doc(3)
    x, e = d.x0.to(DEV), d.e0.to(DEV)
    mask = torch.eq(d.adj, 0)
    a_masked_arg = gat(x, e, mask.to(DEV))
    a_no_mask = gat(x, e, None)
    assert torch.equal(a_masked_arg, a_no_mask)                      # quirk 1 (G:163-164)
    gat.apply_mask = True
    x.requires_grad_(True), e.requires_grad_(True)
    a = gat(x, e, mask.to(DEV))
    g = torch.randn(d.n, d.n, generator=torch.Generator().manual_seed(1))
    (a * g.to(DEV)).sum().backward()
    ps = {k: v.clone().requires_grad_(True) for k, v in sub(state, PREFIXES[0]).items()}
    xo, eo = d.x0.clone().requires_grad_(True), d.e0.clone().requires_grad_(True)
    ao = O.gat_attention(xo, eo, ps, mask, apply_mask=True)
    (ao * g).sum().backward()
    assert_close(a, ao, 1e-5, "masked A")
    assert float(a[mask.to(DEV)].max()) < 1e-30 or bool(mask.all(1).any())
    assert_close(x.grad, xo.grad, FP32_TOL, "dx")
    assert_close(e.grad, eo.grad, FP32_TOL, "de")
    gat.apply_mask = False


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_multi_head_attention(variant):
    layers, heads = VARIANTS[variant]
    gb, state = device_blocks(layers, heads)
    mha = gb.get_adj_matrix[0]
    for i in DOCS:
        d = S.make_doc(i)
        x = d.x0.to(DEV).requires_grad_(True)
        mha.zero_grad()
        atts = mha(x, d.e1.to(DEV))          # second positional arg lands in `mask` and is ignored (G:336)
        assert isinstance(atts, list) and len(atts) == heads and atts[0].shape == (d.n, d.n)
        gen = torch.Generator().manual_seed(i)
        gs = [torch.randn(d.n, d.n, generator=gen) for _ in range(heads)]
        sum((a * g.to(DEV)).sum() for a, g in zip(atts, gs)).backward()
        ps = {k: v.clone().requires_grad_(True) for k, v in sub(state, PREFIXES[2]).items()}
        xo = d.x0.clone().requires_grad_(True)
        ao = O.mha_attention(xo, ps, heads)
        sum((a * g).sum() for a, g in zip(ao, gs)).backward()
        for h in range(heads):
            assert_close(atts[h], ao[h], 1e-5, f"A[{h}]")
        assert_close(x.grad, xo.grad, FP32_TOL, "dx")
        for k, p in mha.named_parameters():
            if "linears_k" in k:
                assert p.grad is None and ps[k].grad is None          # quirk 3
            else:
                assert_close(p.grad, ps[k].grad, FP32_TOL, "d" + k)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_graph_convolution_and_multi_graph_convolution(variant):
    layers, heads = VARIANTS[variant]
    gb, state = device_blocks(layers, heads)
    cag, mag = gb.graphcnn[0], gb.graphcnn[1]
    for i in DOCS:
        d = S.make_doc(i)
        gen = torch.Generator().manual_seed(100 + i)
        # arbitrary non-negative maps incl. an all-zero row: exercises r = rowsum + [rowsum == 0]
        atts = [torch.rand(d.n, d.n, generator=gen) * (torch.rand(d.n, d.n, generator=gen) < 0.7) for _ in range(heads)]
        atts[0][0].zero_()
        g = torch.randn(d.n, 128, generator=gen)
        for mod, prefix, att_in in ((cag, PREFIXES[1], atts[0]), (mag, PREFIXES[3], atts)):
            x, e = d.x0.to(DEV).requires_grad_(True), d.e1.to(DEV).requires_grad_(True)
            multi = isinstance(att_in, list)
            a_dev = [a.to(DEV).requires_grad_(True) for a in att_in] if multi else att_in.to(DEV).requires_grad_(True)
            mod.zero_grad()
            y = mod(x, e, a_dev)
            (y * g.to(DEV)).sum().backward()
            ps = {k: v.clone().requires_grad_(True) for k, v in sub(state, prefix).items()}
            xo, eo = d.x0.clone().requires_grad_(True), d.e1.clone().requires_grad_(True)
            if multi:
                ao = [a.clone().requires_grad_(True) for a in att_in]
                yo = O.maggc_conv(xo, eo, ao, ps, layers, heads)
            else:
                ao = att_in.clone().requires_grad_(True)
                yo = O.caggc_conv(xo, eo, ao, ps, layers)
            (yo * g).sum().backward()
            assert_close(y, yo, FP32_TOL, "y")
            assert_close(x.grad, xo.grad, FP32_TOL, "dx")
            assert_close(e.grad, eo.grad, FP32_TOL, "de")
            if multi:
                for h in range(heads):
                    assert_close(a_dev[h].grad, ao[h].grad, FP32_TOL, f"dA[{h}]")
            else:
                assert_close(a_dev.grad, ao.grad, FP32_TOL, "dA")
            for k, p in mod.named_parameters():
                assert_close(p.grad, ps[k].grad, 2 * FP32_TOL, "d" + k)


@pytest.mark.parametrize("in_dim,out_dim", [(128, 64), (192, 64), (224, 32), (128, 128), (160, 16)])
def test_stand_alone_graph_conv(in_dim, out_dim):
    torch.manual_seed(4)
    conv = M.GraphConv(in_dim, 128, out_dim)
    w_e, w_n = conv.weights_edge.detach().clone(), conv.weights_node.detach().clone()
    conv = conv.to(DEV)
    d = S.make_doc(5)
    gen = torch.Generator().manual_seed(9)
    x_cpu = torch.randn(d.n, in_dim, generator=gen)
    att = torch.softmax(torch.randn(d.n, d.n, generator=gen), -1)
    att[2].zero_()
    g = torch.randn(d.n, out_dim, generator=gen)
    x, e, a = (t.to(DEV).requires_grad_(True) for t in (x_cpu, d.e0, att))
    y = conv(x, e, a)
    (y * g.to(DEV)).sum().backward()
    xo, eo, ao = (t.clone().requires_grad_(True) for t in (x_cpu, d.e0, att))
    we, wn = w_e.clone().requires_grad_(True), w_n.clone().requires_grad_(True)
    yo = O.graph_conv(xo, eo, ao, we, wn)
    (yo * g).sum().backward()
    assert_close(y, yo, FP32_TOL, "out")
    assert_close(x.grad, xo.grad, FP32_TOL, "dx")
    assert_close(e.grad, eo.grad, FP32_TOL, "de")
    assert_close(a.grad, ao.grad, FP32_TOL, "dA")
    assert_close(conv.weights_edge.grad, we.grad, FP32_TOL, "dWe")
    assert_close(conv.weights_node.grad, wn.grad, FP32_TOL, "dWn")


def test_reference_call_sequence_shares_the_edge_pass():
    """G:330-341 written exactly as the reference writes it, with the drop-in modules."""
    gb, state = device_blocks(2, 8)
    gat, mha, cag, mag = gb.get_weighted_adj_matrix, gb.get_adj_matrix[0], gb.graphcnn[0], gb.graphcnn[1]
    d = S.make_doc(4)
    from helpers import oracle_blocks, upstream
    node_feat = d.x0.to(DEV).requires_grad_(True)
    e0, e1 = d.e0.to(DEV).requires_grad_(True), d.e1.to(DEV).requires_grad_(True)
    mask = torch.eq(d.adj.to(DEV), 0)
    weight_adj_matrix = gat(node_feat, e0, mask)
    assert getattr(M._EBAR_TLS, "slot", None) is not None
    y1 = cag(node_feat, e0, weight_adj_matrix)
    assert getattr(M._EBAR_TLS, "slot", None) is None     # consumed: e0 was streamed once
    adj_matrix_list = mha(y1, e1)
    y2 = mag(y1, e1, adj_matrix_list)
    dy1, dy2 = upstream(d.doc_id, y1.shape, y2.shape)
    ((y1 * dy1.to(DEV)).sum() + (y2 * dy2.to(DEV)).sum()).backward()
    r = oracle_blocks(d, state, 2, 8)
    assert_close(y1, r["y1"], FP32_TOL, "y1")
    assert_close(y2, r["y2"], FP32_TOL, "y2")
    assert_close(node_feat.grad, r["dx0"], FP32_TOL, "dx0")
    assert_close(e0.grad, r["de0"], FP32_TOL, "de0")
    assert_close(e1.grad, r["de1"], FP32_TOL, "de1")
    # a plain python list of H separate tensors works too (the reference's type)
    y2b = mag(y1.detach(), e1.detach(), [a.detach().clone() for a in adj_matrix_list])
    assert_close(y2b, y2, 1e-6, "list input")


def test_ebar_stashed_under_no_grad_is_not_reused_in_grad_mode():
    """a GATAttention.forward under no_grad must not hand its history-free edge mean to a grad-mode
    GraphConvolution.forward on the same tensor (de0 would silently lose the GraphConv edge-mean term)."""
    gb, state = device_blocks(2, 8)
    gat, cag = gb.get_weighted_adj_matrix, gb.graphcnn[0]
    d = S.make_doc(7)
    x = d.x0.to(DEV).requires_grad_(True)
    e0 = d.e0.to(DEV).requires_grad_(True)
    with torch.no_grad():
        att = gat(x, e0, None)
    y = cag(x, e0, att)                      # recomputes the mean with history
    y.sum().backward()
    assert e0.grad is not None and float(e0.grad.abs().max()) > 0.0
    xo, eo = d.x0.clone().requires_grad_(True), d.e0.clone().requires_grad_(True)
    yo = O.caggc_conv(xo, eo, att.cpu(), sub(state, "graphcnn.0"), 2)
    yo.sum().backward()
    assert_close(e0.grad, eo.grad, FP32_TOL, "de0")
