"""-m gpu: gcgcn_expand_pair_context (wire format -> the reference's dense [n, n, S, L] tensors, on the device)
against the oracle restatement of config/Config.py:180-205, 219-222 and the reference's own outputs: integer /
boolean work, bit-exact."""
import numpy as np
import pytest
import torch

from helpers import golden
from oracle import featurize_oracle as FO
from gcgcn_b200 import synthetic as S
from gcgcn_b200.featurize import wire_from_record

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("seed,max_length,max_num", [(0, 512, 5), (1, 512, 5), (2, 40, 2), (3, 512, 5), (4, 64, 3),
                                                     (5, 512, 5), (6, 512, 1), (8, 512, 5), (9, 100, 4)])
def test_expand_pair_context_is_bit_exact(seed, max_length, max_num):
    item = S.make_record(seed)
    ref = FO.from_list_to_tensor(item, max_length, max_num)
    w = wire_from_record(item, max_length, max_num)
    sen, ph, pt = w.expand_pair_context(DEV)
    assert sen.dtype == torch.bool and ph.dtype == torch.int64
    assert np.array_equal(sen.cpu().numpy(), ref["sen_matrix"])
    assert np.array_equal(ph.cpu().numpy(), ref["pos_matrix_h"])
    assert np.array_equal(pt.cpu().numpy(), ref["pos_matrix_t"])
    assert np.array_equal(w.adjacency(DEV).cpu().numpy(), ref["adj_matrix"])


def test_expand_matches_the_reference_golden_and_docred_size():
    g = golden("featurize.npz")
    w = wire_from_record(S.make_record(7, n=6, L=60, S=3), max_length=48, max_num=2)
    sen, ph, pt = w.expand_pair_context(DEV)
    assert np.array_equal(sen.cpu().numpy(), g["sen_matrix"])
    assert np.array_equal(ph.cpu().numpy(), g["pos_matrix_h"]) and np.array_equal(pt.cpu().numpy(), g["pos_matrix_t"])
    # DocRED maximum: 42 entities, 512 tokens, 5 slots (77 MB of dense tensors from a few KB)
    item = S.make_record(11, n=42, L=600, S=5)
    ref = FO.from_list_to_tensor(item, 512, 5)
    w = wire_from_record(item)
    sen, ph, pt = w.expand_pair_context(DEV)
    assert np.array_equal(ph.cpu().numpy(), ref["pos_matrix_h"]) and np.array_equal(sen.cpu().numpy(), ref["sen_matrix"])
    assert w.dense_nbytes() > 70e6 and w.nbytes < 200e3


def test_graph_without_edges_and_single_entity():
    import networkx as nx
    g = nx.DiGraph()
    g.add_node(0, exist_pos=[(3, 5)], type=[2])
    g.graph["max_sentence_num"] = 1
    w = wire_from_record({"document": list(range(20)), "graph": g})
    sen, ph, pt = w.expand_pair_context(DEV)
    assert sen.shape == (1, 1, 1, 20) and not sen.any() and not ph.any() and not pt.any()
