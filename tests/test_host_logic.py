"""Host-side index builders and batch descriptors vs the oracle's literal restatement of
config/Config.py (C:169-176, 207-217, 223) and G:351-352.  Integer work: bit-exact."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import gcgcn_oracle as O
from gcgcn_b200 import synthetic as S
from gcgcn_b200.batch import (PairTables, PoolTable, RaggedBatch, make_dis2idx, node_relative_pos,
                              shard_documents)


def test_ragged_batch_offsets():
    bt = RaggedBatch([3, 1, 5, 0, 2], "cpu")
    assert bt.node_ptr_host.tolist() == [0, 3, 4, 9, 9, 11]
    assert bt.pair_ptr_host.tolist() == [0, 9, 10, 35, 35, 39]
    assert bt.row_doc_host.tolist() == [0, 0, 0, 1, 2, 2, 2, 2, 2, 4, 4]
    assert (bt.total_nodes, bt.total_pairs, bt.max_nodes, bt.num_docs) == (11, 39, 5, 5)
    # algorithmic bytes formula of SURVEY.md 8d: s*d*(5 n^2 + 8 n)
    one = RaggedBatch([42], "cpu")
    assert one.algorithmic_bytes(4, backward=True) == 4 * 128 * (5 * 42 * 42 + 8 * 42)
    assert one.algorithmic_bytes(4, backward=False) == 4 * 128 * (2 * 42 * 42 + 3 * 42)
    assert round(one.algorithmic_bytes() / 1e6, 2) == 4.69


def test_empty_batch():
    bt = RaggedBatch([], "cpu")
    assert (bt.total_nodes, bt.total_pairs, bt.max_nodes) == (0, 0, 0)


def _check_pool_table(spans_list, lens):
    tab = PoolTable.from_spans(spans_list, lens)
    node0, tok0 = 0, 0
    for spans, L in zip(spans_list, lens):
        dense_ref = O.build_node_pos(spans, L)                      # [n, min(L,512)] float32
        Lt = dense_ref.shape[1]
        got = tab.dense(slice(node0, node0 + len(spans)), tok0, Lt)
        assert torch.equal(got, dense_ref)                          # bit-exact weights and positions
        node0 += len(spans)
        tok0 += Lt
    assert tab.total_tokens == tok0 and tab.total_nodes == node0
    # the table from the dense matrices is the same table
    tab2 = PoolTable.from_node_pos([O.build_node_pos(s, L) for s, L in zip(spans_list, lens)])
    for name in ("ent_ptr_host", "tok_idx_host", "w_host", "tok_ptr_host", "ent_idx_host", "w_t_host"):
        assert np.array_equal(getattr(tab, name), getattr(tab2, name)), name
    # transpose really is the transpose
    dense = np.zeros((tab.total_nodes, tab.total_tokens), np.float32)
    for e in range(tab.total_nodes):
        for k in range(tab.ent_ptr_host[e], tab.ent_ptr_host[e + 1]):
            dense[e, tab.tok_idx_host[k]] = tab.w_host[k]
    dense_t = np.zeros_like(dense)
    for t in range(tab.total_tokens):
        for k in range(tab.tok_ptr_host[t], tab.tok_ptr_host[t + 1]):
            dense_t[tab.ent_idx_host[k], t] = tab.w_t_host[k]
    assert np.array_equal(dense, dense_t)


def test_pool_table_matches_reference_weights_on_synthetic_batch():
    docs = S.make_batch()
    _check_pool_table([d.spans for d in docs], [d.L for d in docs])


def test_pool_table_overlap_overwrite_and_truncation():
    # overlapping mentions: the later span overwrites (C:174); tokens >= 512 are cut (C:223)
    spans = [[[0, 4], [2, 5]], [[510, 514]], [[600, 603], [7, 8]], [[3, 4], [3, 4], [3, 4]]]
    _check_pool_table([spans], [700])
    tab = PoolTable.from_spans([spans], [700])
    assert tab.total_tokens == 512
    # entity 1 keeps only tokens 510, 511 with weight 1/4 each; entity 2 loses its first mention
    assert tab.tok_idx_host[tab.ent_ptr_host[1]:tab.ent_ptr_host[2]].tolist() == [510, 511]
    assert tab.tok_idx_host[tab.ent_ptr_host[2]:tab.ent_ptr_host[3]].tolist() == [7]
    assert tab.w_host[tab.ent_ptr_host[2]] == np.float32(0.5)


@settings(max_examples=30, deadline=None)
@given(st.lists(st.lists(st.tuples(st.integers(0, 560), st.integers(1, 6)), min_size=1, max_size=4),
                min_size=1, max_size=9), st.integers(520, 640))
def test_pool_table_property(entities, L):
    spans = [[[s, min(s + ln, L)] for s, ln in ms if s < L] or [[0, 1]] for ms in entities]
    spans = [[m for m in ms if m[1] > m[0]] or [[0, 1]] for ms in spans]
    _check_pool_table([spans], [L])


def test_dis2idx_and_relative_pos():
    assert np.array_equal(make_dis2idx(), O.make_dis2idx())
    for d in S.make_batch():
        assert np.array_equal(node_relative_pos(d.first_pos), O.build_node_relative_pos(d.first_pos).numpy())
    rp = node_relative_pos([0, 1, 3, 700, 90])
    assert rp[0, 3] == -10 and rp[3, 0] == 10 and rp[1, 0] == 1 and rp[0, 2] == -2 and rp[4, 0] == 7


def test_pair_tables_orientation_and_distance_rows():
    docs = S.make_batch((0, 5, 11))
    bt = RaggedBatch([d.n for d in docs], "cpu")
    rps = [node_relative_pos(d.first_pos) for d in docs]
    tabs = PairTables(bt, rps)
    for b, d in enumerate(docs):
        lo, hi = bt.pair_ptr_host[b], bt.pair_ptr_host[b + 1]
        h_ref, t_ref, dh_ref, dt_ref = O.pair_index_tables(torch.from_numpy(rps[b]))
        base = bt.node_ptr_host[b]
        assert np.array_equal(tabs.h_idx_host[lo:hi] - base, h_ref.reshape(-1).numpy())   # h_idx[i,j] = j
        assert np.array_equal(tabs.t_idx_host[lo:hi] - base, t_ref.reshape(-1).numpy())   # t_idx[i,j] = i
        assert np.array_equal(tabs.dis_h_host[lo:hi], dh_ref.reshape(-1).numpy())
        assert np.array_equal(tabs.dis_t_host[lo:hi], dt_ref.reshape(-1).numpy())
        assert tabs.dis_h_host[lo:hi].min() >= 0 and tabs.dis_h_host[lo:hi].max() <= 20
    assert tabs.h_idx_host.dtype == np.int32


def test_shard_documents_balances_pair_counts():
    sizes = S.shard_doc_sizes(1200)
    for world in (1, 2, 4, 8):
        shards = shard_documents(sizes, world)
        assert sorted(i for s in shards for i in s) == list(range(1200))
        loads = [int((sizes[s] ** 2).sum()) for s in shards]
        assert max(loads) - min(loads) <= 42 * 42


def test_flat_trainer_views_and_cpu_step_is_an_error():
    """Host logic of config 5's trainer: parameters and gradients become views of two flat buffers (16-byte
    aligned slices), autograd accumulates into the bucket in place, and step() on CPU raises (no CPU path)."""
    import pytest
    from gcgcn_b200 import _lib
    from gcgcn_b200.modules import GraphBlocks
    from gcgcn_b200.sharding import FlatTrainer
    torch.manual_seed(0)
    gb = GraphBlocks(2, 8)
    before = {n: p.detach().clone() for n, p in gb.named_parameters()}
    tr = FlatTrainer(gb)
    assert sum(tr.sizes) == 545921 and tr.numel % 4 == 0 and all(o % 4 == 0 for o in tr.offsets)
    for n, p in gb.named_parameters():
        assert torch.equal(p.detach(), before[n])
        if "linears_k" in n:
            assert p.grad is None
        else:
            assert p.data_ptr() - tr.flat_params.data_ptr() == p.grad.data_ptr() - tr.flat.data_ptr()
    w = gb.graphcnn[0].linear_layer.weight
    (w.sum() * 3.0).backward()
    assert float(tr.flat.sum()) == 3.0 * w.numel()          # landed in the bucket, no pack step
    tr.zero_grad()
    assert float(tr.flat.abs().sum()) == 0.0 and w.grad.data_ptr() >= tr.flat.data_ptr()
    with pytest.raises(_lib.GcgcnError):
        tr.step()


def test_pool_table_concat_appends_rows_over_the_same_tokens():
    """PoolTable.concat (one gather for the entity pooling and the producer's active context rows)."""
    import numpy as np
    from gcgcn_b200.batch import PoolTable
    a = PoolTable.from_spans([[[[0, 2]], [[3, 4], [5, 7]]]], [10])
    b = PoolTable(np.arange(4), np.asarray([0, 1, 2]), np.ones(3, np.float32), 10)
    c = PoolTable.concat(a, b)
    assert c.total_nodes == a.total_nodes + 3 and c.total_tokens == 10
    dense = np.zeros((c.total_nodes, 10), np.float32)
    for e in range(c.total_nodes):
        k0, k1 = c.ent_ptr_host[e], c.ent_ptr_host[e + 1]
        dense[e, c.tok_idx_host[k0:k1]] = c.w_host[k0:k1]
    assert np.allclose(dense[0, :2], 0.5) and np.allclose(dense[1, [3, 5, 6]], [0.5, 0.25, 0.25])
    assert np.array_equal(dense[2:], np.eye(10, dtype=np.float32)[:3])
    # the transposed table (backward) lists, per token, the rows that read it
    assert c.tok_ptr_host[-1] == c.tok_idx_host.size
    rows_of_tok0 = c.ent_idx_host[c.tok_ptr_host[0]:c.tok_ptr_host[1]].tolist()
    assert rows_of_tok0 == [0, 2]


def test_synthetic_wire_documents_are_consistent():
    from gcgcn_b200 import synthetic as S
    for d in S.make_batch()[:4]:
        w = S.make_wire(d)
        assert w.n == d.n and w.length == d.L and 1 <= w.max_num <= 5
        sl = w.slots
        assert ((sl[:, 0] != sl[:, 1]) & (sl[:, 3] < sl[:, 4])).all()
        # an edge exists iff it has a slot; slot numbers per edge start at 0 and are consecutive
        edges = {tuple(e) for e in w.edges.tolist()}
        assert edges == {(u, v) for u, v in sl[:, :2].tolist()}
        for (u, v) in list(edges)[:10]:
            js = sorted(sl[(sl[:, 0] == u) & (sl[:, 1] == v), 2].tolist())
            assert js == list(range(len(js)))
        # mentions lie inside the sentence of their slot (head mention starts in it)
        assert ((sl[:, 5] >= sl[:, 3]) & (sl[:, 5] < sl[:, 4])).all()


def test_pack_tiles_is_a_greedy_partition_of_consecutive_documents():
    """gcgcn_batch.tile_doc (the packed-tile MAGGC kernel's work list): every tile holds consecutive documents with
    at most TILE_ROWS rows in total, no tile could take the next document, and the hint is off when one document alone
    is too large or the batch is empty."""
    from gcgcn_b200.batch import TILE_ROWS, pack_tiles
    rng = np.random.default_rng(0)
    for _ in range(200):
        ns = rng.integers(0, 65, size=int(rng.integers(1, 50)))
        td = pack_tiles(ns)
        assert td.dtype == np.int32 and td[0] == 0 and td[-1] == ns.size
        for a, b in zip(td[:-1], td[1:]):
            assert b > a and ns[a:b].sum() <= TILE_ROWS
            if b < ns.size:
                assert ns[a:b + 1].sum() > TILE_ROWS
    assert pack_tiles(np.array([TILE_ROWS, 1, TILE_ROWS - 1, 0, 0, 3])).tolist() == [0, 1, 5, 6]
    assert pack_tiles(np.array([0, 0])).tolist() == [0, 2]
    assert pack_tiles(np.array([TILE_ROWS + 1, 2])).tolist() == [0]
    assert pack_tiles(np.array([], dtype=np.int64)).tolist() == [0]
    bt = RaggedBatch([5, 92, 7], "cpu")
    assert bt.num_tiles == 3 and bt.c_struct.num_tiles == 3 and bt.c_struct.tile_rows == TILE_ROWS
    assert RaggedBatch([200], "cpu").num_tiles == 0
