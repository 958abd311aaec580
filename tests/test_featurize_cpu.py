"""Host featurisation: the oracle restatement of config/Config.py:162-233 against the reference's
own outputs (tests/golden/featurize.npz, written by oracle/pin_featurize.py --write), and the wire format's host-side
content -- spans, types, relative positions, edge list, slot rows -- against the oracle.  No GPU, no compute calls
into the library; the dense expansion used here to read the slot rows back is test code."""
import numpy as np
import pytest
import torch

from helpers import golden
from oracle import featurize_oracle as FO
from gcgcn_b200 import synthetic as S
from gcgcn_b200.featurize import wire_from_record

CASES = [(0, 512, 5), (1, 512, 5), (2, 40, 2), (3, 512, 5), (4, 64, 3), (5, 512, 5), (6, 512, 1)]


def expand_slots_numpy(w):
    """Test-side reading of the slot rows (the product does this on the GPU: gcgcn_expand_pair_context)."""
    tab = FO.make_dis2idx()
    shape = (w.n, w.n, w.max_num, w.length)
    sen, ph, pt = np.zeros(shape, bool), np.zeros(shape, np.int64), np.zeros(shape, np.int64)
    for u, v, j, s0, s1, h0, h1, t0, t1 in w.slots:
        if j >= w.max_num:
            continue
        for k in range(s0, min(s1, w.length)):
            sen[u, v, j, k] = True
            ph[u, v, j, k] = 10 + (-tab[h0 - k] if k < h0 else tab[k - h1] if k > h1 else 0)
            pt[u, v, j, k] = 10 + (-tab[t0 - k] if k < t0 else tab[k - t1] if k > t1 else 0)
    return sen, ph, pt


def test_oracle_matches_the_reference_golden():
    g = golden("featurize.npz")
    got = FO.from_list_to_tensor(S.make_record(7, n=6, L=60, S=3), max_length=48, max_num=2)
    for k in ("adj_matrix", "sen_matrix", "pos_matrix_h", "pos_matrix_t", "node_pos", "node_type", "node_relative_pos"):
        assert got[k].dtype == g[k].dtype and np.array_equal(got[k], g[k]), k
    assert g["sen_matrix"].any() and (g["pos_matrix_h"] != 0).any()


@pytest.mark.parametrize("seed,max_length,max_num", CASES)
def test_wire_format_carries_everything_the_dense_tensors_hold(seed, max_length, max_num):
    item = S.make_record(seed)
    ref = FO.from_list_to_tensor(item, max_length, max_num)
    w = wire_from_record(item, max_length, max_num)
    assert ref["sen_matrix"].shape == (w.n, w.n, w.max_num, w.length)
    # pooling weights: bit-exact float32 (C:169-176, 223)
    tab = w.pool_table(max_length=max_length)
    assert torch.equal(tab.dense(slice(0, w.n), 0, w.length), torch.from_numpy(ref["node_pos"]))
    assert np.array_equal(w.node_type, ref["node_type"])
    assert np.array_equal(w.relative_pos(), ref["node_relative_pos"])
    assert np.array_equal(w.adjacency("cpu").numpy(), ref["adj_matrix"])
    sen, ph, pt = expand_slots_numpy(w)
    assert np.array_equal(sen, ref["sen_matrix"])
    assert np.array_equal(ph, ref["pos_matrix_h"]) and np.array_equal(pt, ref["pos_matrix_t"])
    assert w.nbytes * 20 < w.dense_nbytes() or w.n < 3


def test_expansion_has_no_cpu_path():
    from gcgcn_b200 import _lib
    w = wire_from_record(S.make_record(0))
    with pytest.raises(_lib.GcgcnError):
        w.expand_pair_context("cpu")
