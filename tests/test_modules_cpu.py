"""Drop-in modules on the CPU: structure, parameter naming, and the no-fallback rule."""
import pytest
import torch

from gcgcn_b200 import _lib
from gcgcn_b200 import modules as M
from gcgcn_b200.batch import RaggedBatch


def test_constructor_signatures_and_parameter_shapes():
    gc = M.GraphConv(192, 128, 64)
    assert gc.weights_edge.shape == (128, 64) and gc.weights_node.shape == (192, 64) and gc.bias is None
    cag = M.GraphConvolution(2, 128, 128)
    assert [tuple(p.shape) for p in cag.parameters()] == [(128, 64), (128, 64), (128, 64), (192, 64),
                                                          (128, 128), (128,)]
    mag = M.MultiGraphConvolution(4, 4, 128, 128)
    assert len(mag.graphconv) == 16 and mag.linear_layer.weight.shape == (128, 512)
    assert mag.graphconv[7].weights_node.shape == (128 + 32 * 3, 32)
    mha = M.MultiHeadAttention(8, 128)
    assert len(mha.linears_q) == 8 and len(mha.linears_k) == 8 and mha.linears_q[0].weight.shape == (16, 128)
    gat = M.GATAttention(128, 128)
    assert gat.wt.weight.shape == (1, 384)
    # used-parameter totals of SURVEY.md section 8a
    gb = M.GraphBlocks(2, 8)
    used = sum(p.numel() for n, p in gb.named_parameters() if "linears_k" not in n)
    dead = sum(p.numel() for n, p in gb.named_parameters() if "linears_k" in n)
    assert (used, dead) == (545921, 16512)
    gb = M.GraphBlocks(4, 4)
    used = sum(p.numel() for n, p in gb.named_parameters() if "linears_k" not in n)
    assert used == 343169


def test_collapse_is_exact_algebra():
    torch.manual_seed(1)
    gat = M.GATAttention(128, 128)
    n = 9
    x, e = torch.randn(n, 128), torch.randn(n, n, 128)
    u, v, c = gat.collapse()
    energy = (x @ u).unsqueeze(0) + e @ v + c
    xh = x.unsqueeze(0).expand(n, n, -1)
    ref = gat.wt(torch.cat([gat.linear_node_h(xh), gat.linear_node_t(xh), gat.linear_edge_r(e)], -1)).squeeze(-1)
    assert float((energy - ref).abs().max()) < 1e-5


def test_cpu_tensors_are_rejected_not_computed():
    gb = M.GraphBlocks(2, 8)
    x, e = torch.randn(5, 128), torch.randn(5, 5, 128)
    with pytest.raises(_lib.GcgcnError, match="CPU"):
        gb.get_weighted_adj_matrix(x, e, None)
    with pytest.raises(_lib.GcgcnError, match="CPU"):
        gb.graphcnn[0](x, e, torch.eye(5))
    with pytest.raises(_lib.GcgcnError, match="CPU"):
        gb.get_adj_matrix[0](x)


def test_unsupported_shapes_fail_loudly():
    with pytest.raises(_lib.GcgcnError):
        M.GraphConvolution(3, 128, 128)        # 128 / 3 is not an integer width
    with pytest.raises(_lib.GcgcnError):
        M.GATAttention(64, 64)
    with pytest.raises(_lib.GcgcnError):
        M.GraphConv(128, 128, 64, bias=True)
