"""Host-side descriptors and index builders for ragged batches of document graphs.

Pure host logic (numpy / CPU torch) plus the device copies of the tables; no kernels here.

* ``RaggedBatch``  -- the ``gcgcn_batch`` descriptor: node and pair offsets of B concatenated
  documents.  B = 1 is the reference's un-batched call (G:281 takes one document).
* ``PoolTable``    -- mention->entity pooling weights as CSR, derived bit-exactly from the
  reference's dense ``node_pos`` construction (config/Config.py:169-176, truncation C:223).
* ``PairTables``   -- the integer content of the h/t pair gathers (G:306-307, G:351-352).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib

MAX_LENGTH = 512   # config/Config.py:67
DIS_PLUS = 10      # config/Config.py:119


TILE_ROWS = 96          # GCGCN_TILE_ROWS of include/gcgcn_b200.h


def pack_tiles(sizes: np.ndarray) -> np.ndarray:
    """Greedy packing of consecutive documents into tiles of at most TILE_ROWS node rows: returns ``tile_doc``
    [num_tiles + 1] (int32), tile t = documents [tile_doc[t], tile_doc[t+1]).  Empty (size 1, no tiles) when there
    are no documents or one of them alone exceeds a tile."""
    ns = np.asarray(sizes, dtype=np.int64)
    if ns.size == 0 or int(ns.max()) > TILE_ROWS:
        return np.zeros(1, dtype=np.int32)
    # vectorised greedy: a tile starting at document s ends before the first document whose cumulative row count
    # from s exceeds TILE_ROWS -- found with searchsorted on the prefix sums
    ptr = np.zeros(ns.size + 1, dtype=np.int64)
    np.cumsum(ns, out=ptr[1:])
    starts = [0]
    s = 0
    while s < ns.size:
        s = int(np.searchsorted(ptr, ptr[s] + TILE_ROWS, side="right")) - 1
        starts.append(s)
    return np.asarray(starts, dtype=np.int32)


class RaggedBatch:
    """Offsets of B document graphs laid out back to back (include/gcgcn_b200.h)."""

    def __init__(self, sizes: Sequence[int], device: torch.device | str):
        ns = np.asarray(list(sizes), dtype=np.int64)
        if ns.ndim != 1 or (ns < 0).any():
            raise ValueError("sizes must be a 1-D list of non-negative entity counts")
        self.sizes = ns
        self.device = torch.device(device)
        self.num_docs = int(ns.size)
        node_ptr = np.zeros(ns.size + 1, dtype=np.int64)
        np.cumsum(ns, out=node_ptr[1:])
        pair_ptr = np.zeros(ns.size + 1, dtype=np.int64)
        np.cumsum(ns * ns, out=pair_ptr[1:])
        if node_ptr[-1] >= 2 ** 31:
            raise ValueError("too many node rows for int32 offsets")
        self.total_nodes = int(node_ptr[-1])
        self.total_pairs = int(pair_ptr[-1])
        self.max_nodes = int(ns.max()) if ns.size else 0
        self.node_ptr_host = node_ptr.astype(np.int32)
        self.pair_ptr_host = pair_ptr
        self.row_doc_host = np.repeat(np.arange(ns.size, dtype=np.int32), ns)
        self.node_ptr = torch.from_numpy(self.node_ptr_host).to(self.device)
        self.pair_ptr = torch.from_numpy(self.pair_ptr_host).to(self.device)
        self.row_doc = torch.from_numpy(self.row_doc_host).to(self.device)
        # scheduling hint: documents by descending size + size-class boundaries (n > 48, 32, 16, 0)
        self.doc_order_host = np.argsort(-ns, kind="stable").astype(np.int32)
        self.class_end_host = [int((ns > t).sum()) for t in (48, 32, 16, 0)]
        self.doc_order = torch.from_numpy(self.doc_order_host).to(self.device)
        # packing hint: consecutive documents in tiles of <= 128 node rows (tensor-core block kernels)
        self.tile_doc_host = pack_tiles(ns)
        self.num_tiles = max(int(self.tile_doc_host.size) - 1, 0)
        self.tile_doc = torch.from_numpy(self.tile_doc_host).to(self.device)
        self.c_struct = _lib.Batch(
            self.num_docs, self.total_nodes, self.total_pairs, self.max_nodes, 0,
            self.node_ptr.data_ptr(), self.pair_ptr.data_ptr(),
            self.row_doc.data_ptr() if self.total_nodes else None,
            self.doc_order.data_ptr() if self.num_docs else None,
            (ctypes.c_int32 * 4)(*self.class_end_host),
            self.tile_doc.data_ptr() if self.num_tiles else None, self.num_tiles, TILE_ROWS)

    @property
    def ref(self):
        return ctypes.byref(self.c_struct)

    def split_nodes(self, t: torch.Tensor) -> List[torch.Tensor]:
        """[total_nodes, *] -> per-document views."""
        return list(torch.split(t, self.sizes.tolist(), dim=0))

    def split_pairs(self, t: torch.Tensor) -> List[torch.Tensor]:
        """[total_pairs, *] -> per-document [n, n, *] views."""
        out = []
        for b, n in enumerate(self.sizes.tolist()):
            lo, hi = int(self.pair_ptr_host[b]), int(self.pair_ptr_host[b + 1])
            out.append(t[lo:hi].view(n, n, *t.shape[1:]))
        return out

    # algorithmic HBM bytes per SURVEY.md section 8d (weights excluded)
    def algorithmic_bytes(self, elem_bytes: int = 4, backward: bool = True) -> int:
        d = 128
        n2, n1 = self.total_pairs, self.total_nodes
        fwd = elem_bytes * d * (2 * n2 + 3 * n1)
        bwd = elem_bytes * d * (3 * n2 + 5 * n1)
        return fwd + (bwd if backward else 0)


_BATCH_CACHE: dict = {}


def single_doc_batch(n: int, device) -> RaggedBatch:
    """Cached B = 1 descriptor used by the per-document drop-in modules."""
    key = (int(n), str(device))
    bt = _BATCH_CACHE.get(key)
    if bt is None:
        if len(_BATCH_CACHE) > 512:
            _BATCH_CACHE.clear()
        bt = _BATCH_CACHE[key] = RaggedBatch([n], device)
    return bt


# ------------------------------------------------------------------------------- pooling
def _entity_weights(spans: Sequence[Sequence[int]], max_length: int):
    """Sparse restatement of C:169-176 for one entity: each span *assigns* 1/len to its tokens
    (a later span overwrites an earlier one where they overlap, C:174), then the row is scaled
    by 1/#mentions (C:175) -- all in float64 -- and cast to float32 (C:223).  Tokens at or past
    ``max_length`` are truncated away (C:223); exact zeros are dropped like ``node_pos != 0``."""
    vals = {}
    for s, t in spans:
        v = 1.0 / (t - s)
        for k in range(s, t):
            vals[k] = v
    scale = 1.0 / len(spans)
    toks = sorted(k for k in vals if k < max_length)
    w = np.asarray([vals[k] * scale for k in toks], dtype=np.float64).astype(np.float32)
    keep = w != 0
    return np.asarray(toks, dtype=np.int64)[keep], w[keep]


class PoolTable:
    """CSR (entity -> tokens) and its transpose (token -> entities) for a ragged batch.

    Token indices are global rows of the concatenated context ``[sum_b L_b, 128]``.
    """

    def __init__(self, ent_ptr, tok_idx, w, total_tokens: int, device=None):
        self.ent_ptr_host = np.asarray(ent_ptr, dtype=np.int32)
        self.tok_idx_host = np.asarray(tok_idx, dtype=np.int32)
        self.w_host = np.asarray(w, dtype=np.float32)
        self.total_tokens = int(total_tokens)
        self.total_nodes = int(self.ent_ptr_host.size - 1)
        # transpose: stable sort by token keeps entities ascending inside a token's list
        ent_of = np.repeat(np.arange(self.total_nodes, dtype=np.int32), np.diff(self.ent_ptr_host))
        order = np.argsort(self.tok_idx_host, kind="stable")
        self.ent_idx_host = ent_of[order]
        self.w_t_host = self.w_host[order]
        counts = np.bincount(self.tok_idx_host, minlength=self.total_tokens)
        self.tok_ptr_host = np.zeros(self.total_tokens + 1, dtype=np.int32)
        np.cumsum(counts, out=self.tok_ptr_host[1:])
        self.device = None
        if device is not None:
            self.to(device)

    def to(self, device):
        self.device = torch.device(device)
        for name in ("ent_ptr", "tok_idx", "w", "tok_ptr", "ent_idx", "w_t"):
            setattr(self, name, torch.from_numpy(getattr(self, name + "_host")).to(self.device))
        return self

    @classmethod
    def concat(cls, a: "PoolTable", b: "PoolTable", device=None):
        """Rows of ``a`` followed by rows of ``b`` over the same token space: one gather launch (and one transposed
        gather in the backward) serves both -- the graph head pools the entities and fetches the producer's active
        context rows in a single pass over ``context_output``."""
        if a.total_tokens != b.total_tokens:
            raise ValueError("both tables must index the same token rows")
        ent_ptr = np.concatenate([a.ent_ptr_host, a.ent_ptr_host[-1] + b.ent_ptr_host[1:]])
        return cls(ent_ptr, np.concatenate([a.tok_idx_host, b.tok_idx_host]), np.concatenate([a.w_host, b.w_host]),
                   a.total_tokens, device)

    @classmethod
    def from_spans(cls, docs_spans: Sequence[Sequence[Sequence[Sequence[int]]]],
                   doc_lens: Sequence[int], max_length: int = MAX_LENGTH, device=None):
        """docs_spans[b][e] = [[start, end), ...] in document-local token positions."""
        ent_ptr, tok, wts, base = [0], [], [], 0
        for spans, L in zip(docs_spans, doc_lens):
            Lt = min(int(L), max_length)
            for ms in spans:
                t, w = _entity_weights(ms, Lt)
                tok.append(t + base)
                wts.append(w)
                ent_ptr.append(ent_ptr[-1] + t.size)
            base += Lt
        tok = np.concatenate(tok) if tok else np.zeros(0, np.int64)
        wts = np.concatenate(wts) if wts else np.zeros(0, np.float32)
        return cls(ent_ptr, tok, wts, base, device)

    @classmethod
    def from_node_pos(cls, node_pos_list: Sequence[torch.Tensor], device=None):
        """From the reference's dense [n, L] float32 weight matrices (already truncated)."""
        ent_ptr, tok, wts, base = [0], [], [], 0
        for npos in node_pos_list:
            a = npos.detach().cpu().numpy()
            for row in a:
                nz = np.nonzero(row)[0]
                tok.append(nz + base)
                wts.append(row[nz])
                ent_ptr.append(ent_ptr[-1] + nz.size)
            base += a.shape[1]
        tok = np.concatenate(tok) if tok else np.zeros(0, np.int64)
        wts = np.concatenate(wts) if wts else np.zeros(0, np.float32)
        return cls(ent_ptr, tok, wts, base, device)

    def dense(self, b_nodes: slice, tok_base: int, L: int) -> torch.Tensor:
        """Dense [n, L] float32 block of one document (for tests)."""
        lo, hi = b_nodes.start, b_nodes.stop
        out = np.zeros((hi - lo, L), dtype=np.float32)
        for e in range(lo, hi):
            k0, k1 = self.ent_ptr_host[e], self.ent_ptr_host[e + 1]
            out[e - lo, self.tok_idx_host[k0:k1] - tok_base] = self.w_host[k0:k1]
        return torch.from_numpy(out)


# ------------------------------------------------------------------------------- pair gathers
def make_dis2idx() -> np.ndarray:
    """Log-bucket distance table of config/Config.py:105-116."""
    t = np.zeros(1024, dtype=np.int64)
    t[1:2] = 1
    edge = 2
    for v in range(2, 11):
        t[edge:] = v
        edge *= 2
    return t


def node_relative_pos(first_pos: Sequence[int]) -> np.ndarray:
    """Signed distance bucket between first mentions, config/Config.py:207-217 (zero diagonal)."""
    tab = make_dis2idx()
    p = np.asarray(first_pos, dtype=np.int64)
    d = p[:, None] - p[None, :]
    rp = np.where(d < 0, -tab[np.abs(d)], tab[np.abs(d)])
    np.fill_diagonal(rp, 0)
    return rp.astype(np.int64)


class PairTables:
    """int32 gather tables over every pair of the batch (pair-major, [total_pairs]):

    h_idx[(b,i,j)] = node_ptr[b] + j     -- the "h" tensor gathers the column entity (G:351)
    t_idx[(b,i,j)] = node_ptr[b] + i     -- the "t" tensor gathers the row entity    (G:352)
    dis_h = dis_plus + rp[i,j], dis_t = dis_plus - rp[i,j]                     (G:306-307)
    """

    def __init__(self, batch: RaggedBatch, rel_pos: Optional[Sequence[np.ndarray]] = None,
                 dis_plus: int = DIS_PLUS, device=None):
        h, t, dh, dt = [], [], [], []
        for b, n in enumerate(batch.sizes.tolist()):
            base = int(batch.node_ptr_host[b])
            ar = np.arange(n, dtype=np.int64)
            h.append(np.broadcast_to(base + ar[None, :], (n, n)).reshape(-1))
            t.append(np.broadcast_to(base + ar[:, None], (n, n)).reshape(-1))
            if rel_pos is not None:
                rp = np.asarray(rel_pos[b], dtype=np.int64).reshape(n, n)
                dh.append((dis_plus + rp).reshape(-1))
                dt.append((dis_plus - rp).reshape(-1))
        cat = lambda xs: (np.concatenate(xs) if xs else np.zeros(0, np.int64)).astype(np.int32)
        self.h_idx_host, self.t_idx_host = cat(h), cat(t)
        self.has_dis = rel_pos is not None
        self.dis_h_host = cat(dh) if self.has_dis else None
        self.dis_t_host = cat(dt) if self.has_dis else None
        # range of the distance rows the tables address; checked against the embedding's row count by the callers
        # (an out-of-range row would read, and in the backward write, out of bounds on the device)
        self.dis_min = int(min(self.dis_h_host.min(), self.dis_t_host.min())) if self.has_dis and self.dis_h_host.size else 0
        self.dis_max = int(max(self.dis_h_host.max(), self.dis_t_host.max())) if self.has_dis and self.dis_h_host.size else 0
        self.device = None
        if device is not None:
            self.to(device)

    def check_dis_rows(self, rows: int):
        if self.has_dis and (self.dis_min < 0 or self.dis_max >= rows):
            raise _lib.GcgcnError(f"PairTables address distance rows {self.dis_min}..{self.dis_max} but the table has "
                                  f"{rows} rows (dis_plus +/- node_relative_pos must stay inside the embedding, G:306-307)")

    def to(self, device):
        self.device = torch.device(device)
        self.h_idx = torch.from_numpy(self.h_idx_host).to(self.device)
        self.t_idx = torch.from_numpy(self.t_idx_host).to(self.device)
        self.dis_h = torch.from_numpy(self.dis_h_host).to(self.device) if self.has_dis else None
        self.dis_t = torch.from_numpy(self.dis_t_host).to(self.device) if self.has_dis else None
        return self


# ------------------------------------------------------------------------------- sharding
def shard_documents(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Document -> rank assignment balanced by sum n^2 (the cost driver):
    documents sorted by n^2 descending are dealt to the currently lightest rank."""
    ns = np.asarray(list(sizes), dtype=np.int64)
    order = np.argsort(-(ns * ns), kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order.tolist():
        r = int(np.argmin(load))
        shards[r].append(i)
        load[r] += int(ns[i]) ** 2
    for s in shards:
        s.sort()
    return shards
