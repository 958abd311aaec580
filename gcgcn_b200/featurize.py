"""Host featurisation as a compact wire format.

The reference's ``Config.from_list_to_tensor`` (config/Config.py:162-233) turns one pickled document into dense
tensors on the host -- among them three ``[n, n, S, L]`` matrices filled by a triple Python loop (C:183-205) -- and
ships all of them to the GPU for every document of every epoch (C:343-353; 77 MB at n = 42, S = 5, L = 512).
``wire_from_record`` keeps what those tensors are made of instead: mention spans, entity types, first-mention
positions, the edge list and one 9-int row per (edge, sentence slot).  The graph hot path consumes the spans and
positions directly (``PoolTable``, ``PairTables``); the dense pair-context tensors of the edge-feature producer
(out of scope here) are re-materialised ON the GPU by ``gcgcn_expand_pair_context`` with the reference's dtypes and
truncation rules, bit for bit (tests/test_gpu_featurize.py against the oracle and the reference's own outputs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from . import _lib
from .batch import PairTables, PoolTable, RaggedBatch, node_relative_pos

DIS_PLUS = 10      # config/Config.py:118


@dataclass
class WireDoc:
    n: int                       # entities (graph nodes)
    length: int                  # min(len(document), max_length)                        (C:164, 223)
    max_num: int                 # min(graph max_sentence_num, max_num) sentence slots   (C:220-222)
    spans: List[List[List[int]]]   # per entity [[start, end), ...] -- the whole of node_pos (C:169-176)
    node_type: np.ndarray        # [n] int64                                             (C:176)
    first_pos: np.ndarray        # [n] int64 first listed mention start -> node_relative_pos (C:207-217)
    edges: np.ndarray            # [e, 2] int32 (u, v) -> adj_matrix                     (C:184)
    slots: np.ndarray            # [m, 9] int32 (u, v, slot, s0, s1, h0, h1, t0, t1)     (C:185-205)

    @property
    def nbytes(self) -> int:
        nspans = sum(len(s) for s in self.spans)
        return 8 * nspans + 8 * self.n + 8 * self.n + self.edges.nbytes + self.slots.nbytes

    def dense_nbytes(self) -> int:
        """What the reference copies host -> device for the same document (C:343-353), graph inputs only."""
        cells = self.n * self.n * self.max_num * self.length
        return cells * (1 + 8 + 8) + self.n * self.length * 4 + self.n * self.n * (4 + 8) + self.n * 8

    # ---- what the graph hot path consumes ----------------------------------------------------------
    def pool_table(self, device=None, max_length: int = 512) -> PoolTable:
        return PoolTable.from_spans([self.spans], [self.length], max_length=max_length, device=device)

    def relative_pos(self) -> np.ndarray:
        return node_relative_pos(self.first_pos)

    def pair_tables(self, batch: RaggedBatch, device=None) -> PairTables:
        return PairTables(batch, [self.relative_pos()], device=device)

    def adjacency(self, device) -> torch.Tensor:
        """adj_matrix [n, n] float32 0/1 (C:178, 184, 219), scattered on the device from the edge list."""
        adj = torch.zeros(self.n, self.n, dtype=torch.float32, device=device)
        if len(self.edges):
            e = torch.from_numpy(self.edges.astype(np.int64)).to(device)
            adj[e[:, 0], e[:, 1]] = 1.0
        return adj

    # ---- what the (out-of-scope) edge-feature producer consumes ------------------------------------
    def expand_pair_context(self, device):
        """(sen_matrix bool, pos_matrix_h int64, pos_matrix_t int64), each [n, n, max_num, length], written by
        ``gcgcn_expand_pair_context`` on ``device`` (CUDA only: there is no host fallback)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.GcgcnError("expand_pair_context: gcgcn_b200 has no CPU path")
        shape = (self.n, self.n, self.max_num, self.length)
        sen = torch.empty(shape, dtype=torch.bool, device=device)
        pos_h = torch.empty(shape, dtype=torch.int64, device=device)
        pos_t = torch.empty(shape, dtype=torch.int64, device=device)
        slots = torch.from_numpy(self.slots).to(device)
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.call("gcgcn_expand_pair_context", slots.data_ptr() if len(self.slots) else None, len(self.slots), self.n,
                  self.max_num, self.length, DIS_PLUS, sen.data_ptr(), pos_h.data_ptr(), pos_t.data_ptr(), stream)
        return sen, pos_h, pos_t


def wire_from_record(item: dict, max_length: int = 512, max_num: int = 5) -> WireDoc:
    """``item`` is a record of the reference's pickles (gen_data_extend_graph.py:115-290): ``item['document']`` and
    ``item['graph']`` (an ``nx.DiGraph``; node attributes ``exist_pos``/``type``, edge attributes
    ``sentences``/``position``, graph attribute ``max_sentence_num``)."""
    graph = item["graph"]
    n = len(graph.nodes())
    spans = [[[int(a), int(b)] for a, b in graph.nodes[e]["exist_pos"]] for e in range(n)]
    node_type = np.asarray([graph.nodes[e]["type"][0] for e in range(n)], dtype=np.int64)
    first_pos = np.asarray([graph.nodes[e]["exist_pos"][0][0] for e in range(n)], dtype=np.int64)
    edges, slots = [], []
    for u, v, edge in graph.edges(data=True):
        edges.append((u, v))
        for j, (sent, pos) in enumerate(zip(edge["sentences"], edge["position"])):
            slots.append((u, v, j, sent[0], sent[1], pos[0], pos[1], pos[2], pos[3]))
    return WireDoc(
        n=n, length=min(len(item["document"]), max_length),
        max_num=min(int(graph.graph["max_sentence_num"]), max_num), spans=spans, node_type=node_type,
        first_pos=first_pos, edges=np.asarray(edges, dtype=np.int32).reshape(-1, 2),
        slots=np.asarray(slots, dtype=np.int32).reshape(-1, 9))
