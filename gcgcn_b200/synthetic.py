"""Synthetic DocRED-shaped documents for the graph hot path.

There is no network and no DocRED in this environment, so every test and bench
input is synthetic with the shapes the reference sees: entities n <= 42, tokens
L <= 512, hidden d = 128, a full n x n pair grid of 128-d edge features per hop.

A document is generated from ``torch.Generator().manual_seed(1337 + doc_id)`` on the
CPU, so the same ids give bit-identical inputs here and on the GPU box.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

HIDDEN = 128
# 12-document batch of configs 1/2/3; n cycles through this list in config 5.
DOC_N = [42, 35, 28, 24, 21, 19, 19, 17, 14, 11, 8, 5]
DOC_L = [512, 480, 400, 330, 260, 220, 198, 197, 180, 150, 120, 90]
DOC_S = [5, 4, 3, 3, 3, 2, 3, 3, 2, 2, 1, 1]
SEED0 = 1337


@dataclass
class Doc:
    doc_id: int
    n: int
    L: int
    S: int
    x0: torch.Tensor          # [n, 128]   tanh(N(0,1))  -- standalone graph-block input
    e0: torch.Tensor          # [n, n, 128] hop-0 edge features
    e1: torch.Tensor          # [n, n, 128] hop-1 edge features
    adj: torch.Tensor         # [n, n] f32 0/1, zero diagonal
    ctx: torch.Tensor         # [L, 128]   tanh(N(0,1))  -- encoder output stand-in
    spans: List[List[List[int]]]   # per entity: [[start, end), ...]
    node_type: torch.Tensor   # [n] int64 in 1..6
    first_pos: List[int]      # first listed mention start per entity


def _rand_spans(gen: torch.Generator, n: int, L: int, overlap_case: bool):
    """1-3 mentions per entity, 1-4 tokens each, non-overlapping by construction:
    the document is cut into one slot per mention and each mention lives in its slot."""
    counts = torch.randint(1, 4, (n,), generator=gen).tolist()
    total = sum(counts)
    width = max(1, L // total)
    perm = torch.randperm(total, generator=gen).tolist()
    lens = torch.randint(1, 5, (total,), generator=gen).tolist()
    offs = torch.randint(0, 1 << 20, (total,), generator=gen).tolist()
    spans, k = [], 0
    for e in range(n):
        ms = []
        for _ in range(counts[e]):
            slot = perm[k]
            ln = min(lens[k], width)
            st = slot * width + offs[k] % (width - ln + 1)
            ms.append([st, st + ln])
            k += 1
        spans.append(ms)
    if overlap_case and L >= 8:
        # pins the "later span overwrites" semantics of C:174
        spans[0] = [[0, 4], [2, 5]]
    return spans


def make_doc(doc_id: int, n: Optional[int] = None, L: Optional[int] = None,
             S: Optional[int] = None, dtype: torch.dtype = torch.float32) -> Doc:
    k = doc_id % len(DOC_N)
    n = DOC_N[k] if n is None else n
    L = DOC_L[k] if L is None else L
    S = DOC_S[k] if S is None else S
    g = torch.Generator().manual_seed(SEED0 + doc_id)
    x0 = torch.tanh(torch.randn(n, HIDDEN, generator=g))
    e0 = torch.randn(n, n, HIDDEN, generator=g)
    e1 = torch.randn(n, n, HIDDEN, generator=g)
    adj = (torch.rand(n, n, generator=g) < 0.3).float()
    adj.fill_diagonal_(0.0)
    ctx = torch.tanh(torch.randn(L, HIDDEN, generator=g))
    spans = _rand_spans(g, n, L, overlap_case=(doc_id == 0))
    node_type = torch.randint(1, 7, (n,), generator=g)
    first_pos = [ms[0][0] for ms in spans]
    return Doc(doc_id, n, L, S, x0.to(dtype), e0.to(dtype), e1.to(dtype), adj, ctx.to(dtype),
               spans, node_type, first_pos)


def make_batch(doc_ids: Sequence[int] = tuple(range(12)), **kw) -> List[Doc]:
    return [make_doc(i, **kw) for i in doc_ids]


def make_keep_masks(doc_id: int, n: int, layers: int, heads: int, hidden: int = HIDDEN,
                    p_att: float = 0.1, p_gcn: float = 0.2, p_out: float = 0.2) -> dict:
    """Keep-scale masks (0 or 1/(1-p)) for every dropout site on the path:
    GAT attention (G:152), CAGGC sub-layer outputs (G:59), hop output (G:232, 341),
    MHA attention (G:131), MAGGC sub-layer outputs (G:90)."""
    gen = torch.Generator().manual_seed(7331 + doc_id)
    gsz = hidden // layers

    def m(shape, p):
        return (torch.rand(shape, generator=gen) >= p).float() / (1.0 - p)

    return {
        "gat": m((n, n), p_att),
        "cag": [m((n, gsz), p_gcn) for _ in range(layers)],
        "out0": m((n, hidden), p_out),
        "mha": [m((n, n), p_att) for _ in range(heads)],
        "mag": [[m((n, gsz), p_gcn) for _ in range(layers)] for _ in range(heads)],
        "out1": m((n, hidden), p_out),
    }


def shard_doc_sizes(num_docs: int) -> np.ndarray:
    """Entity counts of a config-5 style shard: n cycles through DOC_N (mean 20.25)."""
    return np.asarray([DOC_N[i % len(DOC_N)] for i in range(num_docs)], dtype=np.int64)


# ------------------------------------------------------------------------------- pickle-record stand-ins
def make_record(seed: int, n: Optional[int] = None, L: Optional[int] = None, S: Optional[int] = None) -> dict:
    """A synthetic record with the schema of the reference's pickled documents (gen_data_extend_graph.py:115-290):
    `document` token ids, and `graph`, an nx.DiGraph whose nodes carry `exist_pos` [(start, end), ...] and `type`,
    whose edges carry `sentences` [(s0, s1), ...] and `position` [(h0, h1, t0, t1), ...] (one slot per sentence the
    two entities share), and whose graph attribute `max_sentence_num` bounds the slots.  Sentences partition the
    document; an edge exists iff two entities have mentions in a common sentence (gen_data_extend_graph.py:209-261).
    Documents may be longer than max_length and edges may have more slots than max_num (both get truncated)."""
    import networkx as nx
    rng = np.random.RandomState(4242 + seed)
    n = int(rng.randint(2, 12)) if n is None else n
    L = int(rng.randint(40, 700)) if L is None else L
    bounds = [0]
    while bounds[-1] < L:
        bounds.append(min(L, bounds[-1] + int(rng.randint(6, 40))))
    sents = list(zip(bounds[:-1], bounds[1:]))
    g = nx.DiGraph()
    for e in range(n):
        poses, where = [], []
        for _ in range(int(rng.randint(1, 4))):
            si = int(rng.randint(len(sents)))
            s0, s1 = sents[si]
            a = int(rng.randint(s0, s1))
            b = min(s1, a + int(rng.randint(1, 5)))
            poses.append((a, b))
            where.append(sents[si])
        g.add_node(e, exist_pos=poses, exist_sentence=where, type=[int(rng.randint(1, 7))])
    most = 0
    for a in range(n):
        for b in range(n):
            if a == b:
                continue
            common, cpos = [], []
            for pa, sa in zip(g.nodes[a]["exist_pos"], g.nodes[a]["exist_sentence"]):
                for pb, sb in zip(g.nodes[b]["exist_pos"], g.nodes[b]["exist_sentence"]):
                    if sa == sb:
                        common.append(sa)
                        cpos.append(pa + pb)
            if common:
                g.add_edge(a, b, sentences=common, position=cpos)
                most = max(most, len(common))
    g.graph["max_sentence_num"] = max(most, 1) if S is None else max(S, most, 1)
    return {"document": rng.randint(1, 1000, size=L).tolist(), "document_pos": rng.randint(0, n + 1, size=L).tolist(),
            "document_ner": rng.randint(0, 7, size=L).tolist(), "graph": g, "title": f"synthetic-{seed}",
            "label_matrix": np.zeros((n, n, 97), dtype=np.float32), "label_mask": []}


# ------------------------------------------------------------------------------- wire format of the DocRED-shaped batch
def make_wire(doc: Doc, max_num: int = 5):
    """The wire format (featurize.WireDoc) of a synthetic ``Doc``: its mention spans, types and first positions, plus
    sentences, edges and (edge, sentence slot) rows generated the way the reference's preprocessing defines them
    (gen_data_extend_graph.py:209-261): sentences of 12-40 tokens partition the document, an edge (u, v) exists iff
    the two entities have mentions in a common sentence, with one slot per such mention pair."""
    from .featurize import WireDoc
    rng = np.random.RandomState(977 + doc.doc_id)
    bounds = [0]
    while bounds[-1] < doc.L:
        bounds.append(min(doc.L, bounds[-1] + int(rng.randint(12, 41))))
    bounds = np.asarray(bounds)
    sent_of = lambda a: int(np.searchsorted(bounds, a, side="right") - 1)
    ment = [[(a, b, sent_of(a)) for a, b in ms] for ms in doc.spans]
    edges, slots, most = [], [], 0
    for u in range(doc.n):
        for v in range(doc.n):
            if u == v:
                continue
            j = 0
            for a0, a1, sa in ment[u]:
                for b0, b1, sb in ment[v]:
                    if sa == sb:
                        slots.append((u, v, j, int(bounds[sa]), int(bounds[sa + 1]), a0, a1, b0, b1))
                        j += 1
            if j:
                edges.append((u, v))
                most = max(most, j)
    return WireDoc(n=doc.n, length=doc.L, max_num=min(max(most, 1), max_num), spans=doc.spans,
                   node_type=doc.node_type.numpy().astype(np.int64), first_pos=np.asarray(doc.first_pos, dtype=np.int64),
                   edges=np.asarray(edges, dtype=np.int32).reshape(-1, 2), slots=np.asarray(slots, dtype=np.int32).reshape(-1, 9))
