"""Edge-feature producer of GCGCN on the GPU: ``WordAttention`` + ``SentenceAttention``
(/root/reference/models/GCGCN_glove.py:171-214) as the in-loop code uses them (G:299-327), for ragged batches of
documents, fed by the wire format of ``featurize.py`` instead of the dense ``[n, n, S, L]`` host tensors.

The module classes keep the reference's class names, constructor signatures and parameter names, so the
``word_attention.i`` / ``sentence_attention.i`` / ``linear_word_att.i`` / ``linear_sentence_att.i`` slices of a
reference ``state_dict`` load unchanged.  The arithmetic is this package's CUDA (csrc/edgefeat.cu + gcgcn_gemm);
there is no CPU path.  What makes it cheap (each step is exact algebra on the reference's formulas, see edgefeat.cu):
the word scores are a 21 x L table per document, and only slots whose sentence contains token 0 survive the
``~sen_matrix[:, :, :, 0:1]`` mask of G:302 -- all other slots contribute exactly 0 and get exactly zero gradient.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn
from torch.autograd import Function

from . import _lib
from .batch import DIS_PLUS, PoolTable, RaggedBatch
from .functional import D, LinearFn, PoolFn, _cuda, _edge, _p, _stream, linear, workspace

BUCKETS = 21           # config/Config.py:118 dis_num


# ------------------------------------------------------------------------------- host tables
class EdgeTables:
    """Index tables of the active slots of a batch (``struct gcgcn_edge_tables``), built once per batch on the host.

    docs: objects with ``n``, ``length``, ``max_num`` and ``slots`` ([m, 9] int rows
    (u, v, slot, s0, s1, h0, h1, t0, t1)) -- ``featurize.WireDoc``.  Token rows refer to the concatenated context
    ``[sum_b length_b, 128]`` (the layout ``PoolTable.from_spans`` uses).
    """

    def __init__(self, docs: Sequence, batch: RaggedBatch, device=None, dis_plus: int = DIS_PLUS):
        if len(docs) != batch.num_docs:
            raise ValueError("one wire document per document of the batch")
        tok_base = np.zeros(len(docs) + 1, dtype=np.int64)
        np.cumsum([int(d.length) for d in docs], out=tok_base[1:])
        self.total_tokens = int(tok_base[-1])
        self.total_nodes, self.total_pairs = batch.total_nodes, batch.total_pairs
        rows = []          # (doc, pair_local, slot, len, h0, h1, t0, t1, u, v)
        for b, d in enumerate(docs):
            if int(d.n) != int(batch.sizes[b]):
                raise ValueError(f"document {b}: wire format has {d.n} entities, the batch {batch.sizes[b]}")
            sl = np.asarray(d.slots, dtype=np.int64).reshape(-1, 9)
            for u, v, slot, s0, s1, h0, h1, t0, t1 in sl.tolist():
                if slot >= d.max_num or u >= d.n or v >= d.n:          # truncated away (C:220-222)
                    continue
                ln = min(s1, d.length)
                if s0 <= 0 < ln:                                        # sen_matrix[u, v, slot, 0] is set (G:302)
                    rows.append((b, u * d.n + v, slot, ln, h0, h1, t0, t1, u, v))
        rows.sort(key=lambda r: (r[0], r[1], r[2]))
        A = len(rows)
        r = np.asarray(rows, dtype=np.int64).reshape(A, 10)
        doc_of = r[:, 0]
        # active tokens: [0, Lact_b) of every document with an active slot
        lact = np.zeros(len(docs), dtype=np.int64)
        if A:
            np.maximum.at(lact, doc_of, r[:, 3])
        act_first = np.zeros(len(docs) + 1, dtype=np.int64)
        np.cumsum(lact, out=act_first[1:])
        Na = int(act_first[-1])
        tok_doc = np.repeat(np.arange(len(docs)), lact)
        self.act_tok_host = (tok_base[tok_doc] + (np.arange(Na) - act_first[tok_doc])).astype(np.int32)
        doc_slot_lo = np.searchsorted(doc_of, np.arange(len(docs)), side="left")
        doc_slot_hi = np.searchsorted(doc_of, np.arange(len(docs)), side="right")
        host = {
            "tok_first": act_first[tok_doc], "tok_slot_lo": doc_slot_lo[tok_doc], "tok_slot_hi": doc_slot_hi[tok_doc],
            "slot_tok0": act_first[doc_of], "slot_len": r[:, 3], "slot_span": r[:, 4:8].reshape(-1),
            "slot_rowi": batch.node_ptr_host[doc_of] + r[:, 8], "slot_rowj": batch.node_ptr_host[doc_of] + r[:, 9],
        }
        # documents with active slots (one CTA each in the word-attention kernels)
        adocs = np.nonzero(lact > 0)[0]
        host.update(adoc_tok0=act_first[adocs], adoc_len=lact[adocs], adoc_slot_lo=doc_slot_lo[adocs],
                    adoc_slot_hi=doc_slot_hi[adocs])
        self.num_active_docs, self.max_active_len = int(adocs.size), int(lact.max()) if len(docs) else 0
        att_off = np.zeros(A + 1, dtype=np.int64)
        np.cumsum(2 * r[:, 3], out=att_off[1:])
        host["slot_att"] = att_off[:-1]
        self.att_total = int(att_off[-1])
        # active pairs (slots are sorted by pair already)
        key = doc_of * (1 << 32) + r[:, 1]
        first = np.ones(A, dtype=bool)
        first[1:] = key[1:] != key[:-1]
        starts = np.nonzero(first)[0]
        AP = int(starts.size)
        pair_ptr = np.append(starts, A).astype(np.int64)
        pdoc = doc_of[starts] if AP else np.zeros(0, np.int64)
        counts = np.diff(pair_ptr)
        S_of = np.asarray([int(d.max_num) for d in docs], dtype=np.int64)
        sent_num = S_of[pdoc] - counts                                              # padded slots (G:206)
        self.pair_denom_host = sent_num.astype(np.float32) + np.float32(1e-10)      # G:213, in float32 like torch
        self.pair_idx_host = (batch.pair_ptr_host[pdoc] + r[starts, 1]).astype(np.int64) if AP else np.zeros(0, np.int64)
        host["pair_slot_ptr"] = pair_ptr
        # node row -> (slot, side) entries whose node embedding it is: side h embeds the column entity (G:320)
        ent_rows = np.concatenate([host["slot_rowj"], host["slot_rowi"]]) if A else np.zeros(0, np.int64)
        ent_ids = np.concatenate([2 * np.arange(A), 2 * np.arange(A) + 1]) if A else np.zeros(0, np.int64)
        order = np.lexsort((ent_ids, ent_rows))
        ctr_ptr = np.zeros(self.total_nodes + 1, dtype=np.int64)
        np.cumsum(np.bincount(ent_rows.astype(np.int64), minlength=self.total_nodes), out=ctr_ptr[1:])
        host["node_ctr_ptr"], host["node_ctr"] = ctr_ptr, ent_ids[order]
        self.num_tokens, self.num_slots, self.num_pairs, self.dis_plus = Na, A, AP, int(dis_plus)
        self.host = {k: np.ascontiguousarray(v, dtype=np.int32) for k, v in host.items()}
        self.device = None
        if device is not None:
            self.to(device)

    def to(self, device):
        self.device = torch.device(device)
        self.dev = {k: torch.from_numpy(v).to(self.device) for k, v in self.host.items()}
        self.pair_idx = torch.from_numpy(self.pair_idx_host).to(self.device)
        self.pair_denom = torch.from_numpy(self.pair_denom_host).to(self.device)
        # the active token rows as a unit-weight gather table: ctx_act = ctx[act_tok] (backward = its transpose)
        self.gather = PoolTable(np.arange(self.num_tokens + 1), self.act_tok_host, np.ones(self.num_tokens, np.float32),
                                self.total_tokens, self.device)
        ptr = lambda t: t.data_ptr() if t.numel() else None
        d = self.dev
        self.c_struct = _lib.EdgeTablesC(
            self.num_tokens, self.num_slots, self.num_pairs, self.att_total, self.dis_plus, 0,
            ptr(d["tok_first"]), ptr(d["tok_slot_lo"]), ptr(d["tok_slot_hi"]), ptr(d["slot_tok0"]), ptr(d["slot_len"]),
            ptr(d["slot_span"]), ptr(d["slot_att"]), ptr(d["slot_rowi"]), ptr(d["slot_rowj"]), ptr(self.pair_idx),
            ptr(d["pair_slot_ptr"]), ptr(self.pair_denom), ptr(d["node_ctr_ptr"]), ptr(d["node_ctr"]),
            self.num_active_docs, self.max_active_len, ptr(d["adoc_tok0"]), ptr(d["adoc_len"]), ptr(d["adoc_slot_lo"]),
            ptr(d["adoc_slot_hi"]))
        return self

    @property
    def ref(self):
        return ctypes.byref(self.c_struct)

    def ws(self, device):
        n = _lib.load().gcgcn_edgefeat_ws_bytes(self.num_tokens, self.att_total, self.num_slots, self.num_pairs,
                                                self.total_pairs)
        t = workspace(device, n)
        return t.data_ptr(), t.numel()


# ------------------------------------------------------------------------------- autograd bindings
class WordTableFn(Function):
    """T[a][k] = wa . tanh(SF[a] + DF[k]) + ba  (G:183 over the 21 x L distinct arguments)."""

    @staticmethod
    def forward(ctx, SF, DF, wa, ba):
        SF, DF, wa, ba = _cuda(SF, "SF"), _cuda(DF, "DF"), _cuda(wa, "wa").reshape(-1), _cuda(ba, "ba").reshape(-1)
        T = torch.empty(SF.shape[0], BUCKETS, device=SF.device)
        _lib.call("gcgcn_word_table_fwd", _p(SF), _p(DF), _p(wa), _p(ba), SF.shape[0], _p(T), _stream(SF.device))
        ctx.save_for_backward(SF, DF, wa)
        return T

    @staticmethod
    def backward(ctx, dT):
        SF, DF, wa = ctx.saved_tensors
        dev = SF.device
        dT = _cuda(dT, "dT")
        dSF = torch.empty_like(SF)
        dpar = torch.empty(BUCKETS * D + D + 4, device=dev)
        n = _lib.load().gcgcn_edgefeat_ws_bytes(SF.shape[0], 0, 0, 0, 0)
        ws = workspace(dev, n)
        _lib.call("gcgcn_word_table_bwd", _p(SF), _p(DF), _p(wa), _p(dT), SF.shape[0], _p(dSF), _p(dpar), ws.data_ptr(),
                  ws.numel(), _stream(dev))
        return dSF, dpar[:BUCKETS * D].view(BUCKETS, D), dpar[BUCKETS * D:BUCKETS * D + D].view(1, D), dpar[BUCKETS * D + D:BUCKETS * D + D + 1]


class WordPoolFn(Function):
    """cwa [slots, 256]: per active slot and side, softmax over the sentence tokens + weighted context sum (G:186-188)."""

    @staticmethod
    def forward(ctx, T, ctx_act, tabs: EdgeTables):
        T, ctx_act = _cuda(T, "T"), _cuda(ctx_act, "ctx")
        dev = T.device
        att = torch.empty(max(tabs.att_total, 1), device=dev)
        cwa = torch.empty(tabs.num_slots, 2 * D, device=dev)
        _lib.call("gcgcn_word_pool_fwd", tabs.ref, _p(T), _p(ctx_act), _p(att), _p(cwa), _stream(dev))
        ctx.save_for_backward(ctx_act, att)
        ctx.tabs = tabs
        return cwa

    @staticmethod
    def backward(ctx, dcwa):
        ctx_act, att = ctx.saved_tensors
        tabs, dev = ctx.tabs, ctx_act.device
        dcwa = _cuda(dcwa, "dcwa")
        dctx = torch.zeros_like(ctx_act) if tabs.num_slots == 0 else torch.empty_like(ctx_act)
        dT = torch.zeros(ctx_act.shape[0], BUCKETS, device=dev) if tabs.num_slots == 0 else \
            torch.empty(ctx_act.shape[0], BUCKETS, device=dev)
        ws, wsb = tabs.ws(dev)
        _lib.call("gcgcn_word_pool_bwd", tabs.ref, _p(ctx_act), _p(att), _p(dcwa), _p(dctx), _p(dT), ws, wsb, _stream(dev))
        return dT, dctx, None


class SentPoolFn(Function):
    """csa [pairs, 256]: relu-weighted sum of a pair's active slots / (sent_num + 1e-10)  (G:202-213)."""

    @staticmethod
    def forward(ctx, cw, sfeat, nfeat, va, ca, tabs: EdgeTables):
        cw, sfeat, nfeat = _cuda(cw, "cw"), _cuda(sfeat, "sfeat"), _cuda(nfeat, "nfeat")
        va, ca = _cuda(va, "va").reshape(-1), _cuda(ca, "ca").reshape(-1)
        dev = cw.device
        score = torch.empty(max(tabs.num_slots, 1), 2, device=dev)
        csa = torch.empty(tabs.num_pairs, 2 * D, device=dev)
        _lib.call("gcgcn_sent_pool_fwd", tabs.ref, _p(cw), _p(sfeat), _p(nfeat), _p(va), _p(ca), _p(score), _p(csa),
                  _stream(dev))
        ctx.save_for_backward(cw, sfeat, nfeat, va, score)
        ctx.tabs = tabs
        return csa

    @staticmethod
    def backward(ctx, dcsa):
        cw, sfeat, nfeat, va, score = ctx.saved_tensors
        tabs, dev = ctx.tabs, cw.device
        dcsa = _cuda(dcsa, "dcsa")
        dcw, dsfeat, dnfeat = torch.empty_like(cw), torch.empty_like(sfeat), torch.empty_like(nfeat)
        dpar = torch.empty(D + 4, device=dev)
        ws, wsb = tabs.ws(dev)
        _lib.call("gcgcn_sent_pool_bwd", tabs.ref, nfeat.shape[0], _p(cw), _p(sfeat), _p(nfeat), _p(va), _p(score),
                  _p(dcsa), _p(dcw), _p(dsfeat), _p(dnfeat), _p(dpar), ws, wsb, _stream(dev))
        return dcw, dsfeat, dnfeat, dpar[:D].view(1, D), dpar[D:D + 1], None


class EdgeFillFn(Function):
    """e [total_pairs, 128]: bias everywhere, bias + rows[k] at the active pairs (G:326 on a mostly-zero input)."""

    @staticmethod
    def forward(ctx, rows, bias, tabs: EdgeTables, edge_dtype):
        rows, bias = _cuda(rows, "rows"), _cuda(bias, "bias")
        dev = bias.device
        e = torch.empty(tabs.total_pairs, D, device=dev, dtype=edge_dtype)
        dt = _lib.F32 if edge_dtype == torch.float32 else _lib.BF16
        _lib.call("gcgcn_edge_fill_fwd", _p(bias), _p(rows), _p(tabs.pair_idx) if tabs.num_pairs else None, tabs.num_pairs,
                  tabs.total_pairs, dt, _p(e), _stream(dev))
        ctx.tabs, ctx.dt = tabs, dt
        return e

    @staticmethod
    def backward(ctx, de):
        tabs = ctx.tabs
        de, dt = _edge(de, "de")
        dev = de.device
        drows = torch.empty(tabs.num_pairs, D, device=dev)
        dbias = torch.empty(D, device=dev)
        ws, wsb = tabs.ws(dev)
        _lib.call("gcgcn_edge_fill_bwd", _p(de), _p(tabs.pair_idx) if tabs.num_pairs else None, tabs.num_pairs,
                  tabs.total_pairs, dt, _p(drows), _p(dbias), ws, wsb, _stream(dev))
        return drows, dbias, None, None


# ------------------------------------------------------------------------------- modules
class WordAttention(nn.Module):
    """WordAttention(input_dim, hidden_dim, position_dim) -- G:171-192 (parameters only; the batched evaluation
    lives in ``EdgeFeatures.forward``, which never forms the ``[n, n, S, L, *]`` tensors the reference method takes)."""

    def __init__(self, input_dim, hidden_dim, position_dim):
        super().__init__()
        self.attention_sent = nn.Linear(input_dim, hidden_dim)
        self.attention_pos = nn.Linear(position_dim, hidden_dim)
        self.attention_all = nn.Linear(hidden_dim, 1)
        if input_dim != D or hidden_dim != D:
            raise _lib.GcgcnError("gcgcn_b200 supports hidden size 128 (G:234)")


class SentenceAttention(nn.Module):
    """SentenceAttention(input_dim, hidden_dim) -- G:195-214 (parameters only, see ``EdgeFeatures``)."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.attention_sent = nn.Linear(input_dim, hidden_dim)
        self.attention_pos = nn.Linear(input_dim, hidden_dim)
        self.attention_all = nn.Linear(hidden_dim, 1)
        if input_dim != D or hidden_dim != D:
            raise _lib.GcgcnError("gcgcn_b200 supports hidden size 128 (G:234)")


class EdgeFeatures(nn.Module):
    """The per-hop producers of ``context_sent_att`` with the reference model's attribute names (G:266-269):
    ``word_attention.i``, ``sentence_attention.i``, ``linear_word_att.i``, ``linear_sentence_att.i``."""

    def __init__(self, hidden_size=D, dis_size=20, graph_hop=2):
        super().__init__()
        self.word_attention = nn.ModuleList([WordAttention(hidden_size, hidden_size, dis_size) for _ in range(graph_hop)])
        self.sentence_attention = nn.ModuleList([SentenceAttention(hidden_size, hidden_size) for _ in range(graph_hop)])
        self.linear_word_att = nn.ModuleList([nn.Linear(hidden_size * 2, hidden_size) for _ in range(graph_hop)])
        self.linear_sentence_att = nn.ModuleList([nn.Linear(hidden_size * 2, hidden_size) for _ in range(graph_hop)])

    def forward(self, hop: int, context: Optional[torch.Tensor], node_feat: torch.Tensor, dis_embed: torch.Tensor,
                tabs: EdgeTables, edge_dtype=torch.float32, ctx_act: Optional[torch.Tensor] = None) -> torch.Tensor:
        """context [total_tokens, 128] (``context_output`` of every document back to back), node_feat
        [total_nodes, 128] (the hop's node features), dis_embed [21, 20] (``dis_embed.weight``) ->
        context_sent_att of every pair of the batch, [total_pairs, 128]  (G:313-326 for hop ``hop``).
        ``ctx_act`` = the active context rows ``context[tabs.act_tok]`` when the caller has gathered them already
        (both hops read the same rows; ``context`` may then be None)."""
        if context is None:
            context = ctx_act
        if not context.is_cuda:
            raise _lib.GcgcnError("EdgeFeatures: gcgcn_b200 runs on CUDA only (no CPU fallback)")
        wa, sa = self.word_attention[hop], self.sentence_attention[hop]
        lw, ls = self.linear_word_att[hop], self.linear_sentence_att[hop]
        if tabs.num_slots == 0:           # no sentence contains token 0: every pair gets the bias (G:326 on zeros)
            return EdgeFillFn.apply(context.new_zeros(0, D), ls.bias, tabs, edge_dtype)
        if ctx_act is None:
            ctx_act = PoolFn.apply(context, tabs.gather)                              # rows of the first sentences
        SF = linear(ctx_act, wa.attention_sent)                                       # G:179
        DF = linear(dis_embed, wa.attention_pos)                                      # G:180 on the 21 table rows
        T = WordTableFn.apply(SF, DF, wa.attention_all.weight, wa.attention_all.bias)  # G:183
        cwa = WordPoolFn.apply(T, ctx_act, tabs)                                      # G:186-188, h | t (G:316)
        cw = linear(cwa, lw)                                                          # G:317
        sfeat = linear(cw, sa.attention_sent)                                         # G:202
        nfeat = linear(node_feat, sa.attention_pos)                                   # G:203 at node level
        csa = SentPoolFn.apply(cw, sfeat, nfeat, sa.attention_all.weight, sa.attention_all.bias, tabs)   # G:204-213, G:325
        rows = LinearFn.apply(csa, ls.weight, None)                                   # G:326 without the bias
        return EdgeFillFn.apply(rows, ls.bias, tabs, edge_dtype)
