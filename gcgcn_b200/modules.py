"""Drop-in replacements for the reference's graph-block modules.

Same class names, constructor signatures, parameter names/shapes (so ``state_dict`` round-trips
with the reference) and ``forward`` signatures as
/root/reference/models/GCGCN_glove.py:18-168 (identical in
models/GraphCNN_multihead_bert_gate_cls.py:18-172).  ``forward`` takes one un-batched document
exactly like the reference; ``forward_batched`` runs a ragged batch of documents in one set of
launches.  All arithmetic happens in this package's CUDA kernels via ``functional``.

Reference quirks that are reproduced on purpose:
  1. GATAttention ignores its mask unless ``apply_mask`` is set (G:163-164);
  2. GAT "head" and "tail" node features both index the column entity (G:156-157);
  3. MultiHeadAttention's key is produced by ``linears_q``; ``linears_k`` is allocated but unused
     and never receives a gradient (G:136-137);
  4. GraphConv divides node *and* edge term by the attention row sum (G:43-50);
  5./6. hop glue and pair orientation: see ``GraphBlocks`` and ``batch.PairTables``.
"""
from __future__ import annotations

import threading
import weakref
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _lib
from .batch import PairTables, PoolTable, RaggedBatch, single_doc_batch
from .functional import (CaggcFn, EdgeMeanFn, GatCollapseFn, GatFn, LinearFn, PackRowsFn, MhaFn, MhaStackFn, PackStackFn, PairDenseFn, PairGatherFn, PoolFn, StackFn,
                         block_supported)

HIDDEN = 128


class _KeepMixin:
    """Train-mode dropout = multiplication by a keep-scale mask (0 or 1/(1-p)) that the kernels
    take as an input.  Masks come from torch's generator unless a test queued explicit ones with
    ``inject_keep`` (that is how train-mode parity with the oracle is checked)."""

    def inject_keep(self, masks: Optional[Sequence[torch.Tensor]]):
        self._injected = None if masks is None else list(masks)

    def _keep(self, shape, p: float, device) -> Optional[torch.Tensor]:
        inj = getattr(self, "_injected", None)
        if inj:
            return inj.pop(0).to(device=device, dtype=torch.float32).reshape(shape)
        if not self.training or p <= 0.0:
            return None
        return (torch.rand(shape, device=device) >= p).to(torch.float32) / (1.0 - p)

    def _dropping(self) -> bool:
        return bool(getattr(self, "_injected", None)) or self.training


# edge-mean hand-off between GATAttention.forward and the GraphConvolution.forward that follows it
# on the same edge tensor (G:332-333): one pass over e serves both.
# The slot is per thread and remembers the autograd mode it was filled under: an ebar stashed under no_grad (or
# by another thread's model) has no history, and handing it to a grad-mode GraphConvolution would silently drop
# the edge-mean term of de0 -- such a stash is ignored (and dropped), the mean is simply recomputed.
_EBAR_TLS = threading.local()


def _stash_ebar(edge: torch.Tensor, ebar: torch.Tensor):
    _EBAR_TLS.slot = (weakref.ref(edge), edge._version, ebar, torch.is_grad_enabled(), ebar.requires_grad)


def _take_ebar(edge: torch.Tensor) -> Optional[torch.Tensor]:
    slot = getattr(_EBAR_TLS, "slot", None)
    _EBAR_TLS.slot = None                     # one consumer; never keeps a tensor (and its graph) alive longer
    if slot is None:
        return None
    ref, version, ebar, grad_mode, had_grad = slot
    if ref() is not edge or version != edge._version:
        return None
    if grad_mode != torch.is_grad_enabled() or (grad_mode and edge.requires_grad and not had_grad):
        return None
    return ebar


def _pairs2d(edge: torch.Tensor) -> torch.Tensor:
    return edge.reshape(-1, edge.shape[-1])


class AttentionList(list):
    """List of H [n,n] views that remembers the head-major tensor they were cut from."""
    stacked: Optional[torch.Tensor] = None


# ------------------------------------------------------------------------------------- a3
class GraphConv(nn.Module, _KeepMixin):
    """GraphConv(input_dim, edge_dim, output_dim, bias=False) -- G:18-50."""

    def __init__(self, input_dim, edge_dim, output_dim, bias=False):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.weights_edge = nn.Parameter(torch.empty(edge_dim, output_dim))
        self.weights_node = nn.Parameter(torch.empty(input_dim, output_dim))
        if bias:
            # the reference leaves this parameter uninitialised (G:27) and never enables it
            raise _lib.GcgcnError("GraphConv(bias=True) is never used by the reference and is not supported")
        self.register_parameter("bias", None)
        self.init()

    def init(self):
        nn.init.xavier_uniform_(self.weights_edge.data)   # G:33
        nn.init.xavier_uniform_(self.weights_node.data)   # G:34

    def forward(self, inputs, edge_inputs, adjacency_matrix):
        n = inputs.size(0)
        bt = single_doc_batch(n, inputs.device)
        return self.forward_batched(inputs, _pairs2d(edge_inputs), adjacency_matrix.reshape(1, -1), bt)

    def forward_batched(self, x, edge, att, batch: RaggedBatch, ebar=None):
        if self.weights_edge.shape[0] != HIDDEN:
            raise _lib.GcgcnError(f"edge_dim must be {HIDDEN}, got {self.weights_edge.shape[0]}")
        if ebar is None:
            ebar = EdgeMeanFn.apply(edge, batch)
        return StackFn.apply(x, ebar, att, self.weights_node, self.weights_edge, None, None, None, batch,
                             1, 1, self.input_dim, self.output_dim, 0, None)


def _pack_stack(convs: Sequence[GraphConv], heads: int, layers: int, g: int):
    """Pack per-GraphConv parameters into the layouts of gcgcn_graphconv_stack_* (autograd-visible,
    so gradients land on the reference-named parameters)."""
    if convs[0].weights_node.is_cuda:
        flat = []
        for c in convs:
            flat += [c.weights_node, c.weights_edge]
        packed = PackStackFn.apply(heads, layers, *flat)
        return (packed[0], packed[1], packed[2] if layers > 1 else None)
    wn_x = torch.cat([c.weights_node[:HIDDEN] for c in convs], dim=1)          # [128, H*128]
    w_e = torch.cat([c.weights_edge for c in convs], dim=1)                    # [128, H*128]
    if layers == 1:
        return wn_x, w_e, None
    blocks = []
    for k, c in enumerate(convs):
        l = k % layers
        inner = c.weights_node[HIDDEN:]                                        # [l*g, g]
        pad = inner.new_zeros(HIDDEN - l * g, g)
        blocks.append(torch.cat([inner, pad], dim=0))
    winner = torch.stack(blocks, 0).view(heads, layers, HIDDEN, g)
    return wn_x, w_e, winner


_FULL = _lib.STACK_RELU | _lib.STACK_RESIDUAL | _lib.STACK_LINEAR


# ------------------------------------------------------------------------------------- a4
class GraphConvolution(nn.Module, _KeepMixin):
    """CAGGC convolution: GraphConvolution(layer_num, input_dim, output_dim, bias=False) -- G:52-80."""

    def __init__(self, layer_num, input_dim, output_dim, bias=False):
        super().__init__()
        self.input_dim = input_dim
        self.layer_num = layer_num
        hidden_dim = output_dim
        graph_hidden_dim = int(hidden_dim / layer_num)
        self.gcn_dropout = nn.Dropout(0.2)
        self.graphconv = nn.ModuleList(
            [GraphConv(input_dim + graph_hidden_dim * i, input_dim, graph_hidden_dim) for i in range(layer_num)])
        self.linear_layer = nn.Linear(hidden_dim, output_dim)
        self._g = graph_hidden_dim
        if input_dim != HIDDEN or output_dim != HIDDEN or graph_hidden_dim * layer_num != HIDDEN:
            raise _lib.GcgcnError("gcgcn_b200 supports hidden size 128 with layer_num dividing 128 (G:234)")

    def forward(self, node_feat, edge_feat, adj_matrix):
        n = node_feat.size(0)
        bt = single_doc_batch(n, node_feat.device)
        ebar = _take_ebar(edge_feat)
        if ebar is None:
            ebar = EdgeMeanFn.apply(_pairs2d(edge_feat), bt)
        return self.forward_batched(node_feat, ebar, adj_matrix.reshape(1, -1), bt)

    def forward_batched(self, x, ebar, att, batch: RaggedBatch):
        wn_x, w_e, winner = _pack_stack(self.graphconv, 1, self.layer_num, self._g)
        keep = None
        if self._dropping():
            ks = [self._keep((batch.total_nodes, self._g), self.gcn_dropout.p, x.device)
                  for _ in range(self.layer_num)]
            keep = None if ks[0] is None else torch.cat(ks, dim=1)
        return StackFn.apply(x, ebar, att, wn_x, w_e, winner, self.linear_layer.weight,
                             self.linear_layer.bias, batch, 1, self.layer_num, HIDDEN, HIDDEN, _FULL, keep)


# ------------------------------------------------------------------------------------- a6
class MultiGraphConvolution(nn.Module, _KeepMixin):
    """MAGGC convolution: MultiGraphConvolution(layer_num, head_num, input_dim, output_dim) -- G:82-120."""

    def __init__(self, layer_num, head_num, input_dim, output_dim, bias=False):
        super().__init__()
        self.input_dim = input_dim
        self.layer_num = layer_num
        self.head_num = head_num
        hidden_dim = output_dim
        graph_hidden_dim = int(hidden_dim / layer_num)
        self.gcn_dropout = nn.Dropout(0.2)
        self.graphconv = nn.ModuleList()
        for i in range(head_num):
            for j in range(layer_num):
                self.graphconv.append(GraphConv(input_dim + graph_hidden_dim * j, input_dim, graph_hidden_dim))
        self.linear_layer = nn.Linear(hidden_dim * head_num, output_dim)
        self._g = graph_hidden_dim
        if input_dim != HIDDEN or output_dim != HIDDEN or graph_hidden_dim * layer_num != HIDDEN:
            raise _lib.GcgcnError("gcgcn_b200 supports hidden size 128 with layer_num dividing 128 (G:234)")

    def forward(self, node_feat, edge_feat, adj_matrix_list):
        n = node_feat.size(0)
        bt = single_doc_batch(n, node_feat.device)
        att = getattr(adj_matrix_list, "stacked", None)
        if att is None:
            att = torch.stack([a.reshape(-1) for a in adj_matrix_list], 0)
        ebar = _take_ebar(edge_feat)
        if ebar is None:
            ebar = EdgeMeanFn.apply(_pairs2d(edge_feat), bt)
        return self.forward_batched(node_feat, ebar, att, bt)

    def forward_batched(self, x, ebar, att, batch: RaggedBatch):
        H, L = self.head_num, self.layer_num
        wn_x, w_e, winner = _pack_stack(self.graphconv, H, L, self._g)
        keep = None
        if self._dropping():
            ks = [self._keep((batch.total_nodes, self._g), self.gcn_dropout.p, x.device) for _ in range(H * L)]
            keep = None if ks[0] is None else torch.cat(ks, dim=1)
        return StackFn.apply(x, ebar, att, wn_x, w_e, winner, self.linear_layer.weight,
                             self.linear_layer.bias, batch, H, L, HIDDEN, HIDDEN, _FULL, keep)


# ------------------------------------------------------------------------------------- a5
class MultiHeadAttention(nn.Module, _KeepMixin):
    """MultiHeadAttention(head_num, att_size, dropout=0.1) -- G:122-142."""

    def __init__(self, head_num, att_size, dropout=0.1):
        super().__init__()
        assert att_size % head_num == 0
        self.hidden_size = att_size // head_num
        self.head_num = head_num
        self.linears_q = nn.ModuleList([nn.Linear(att_size, self.hidden_size) for _ in range(head_num)])
        # allocated for state_dict compatibility; the reference never uses it (G:137)
        self.linears_k = nn.ModuleList([nn.Linear(att_size, self.hidden_size) for _ in range(head_num)])
        self.dropout = nn.Dropout(p=dropout)
        if att_size != HIDDEN:
            raise _lib.GcgcnError("gcgcn_b200 supports att_size 128 (G:234)")

    def forward(self, node_feat, mask=None):
        # `mask` receives the edge tensor at the reference's call site (G:336) and is ignored (G:133)
        n = node_feat.size(0)
        bt = single_doc_batch(n, node_feat.device)
        att = self.forward_batched(node_feat, bt)
        out = AttentionList(att[h].view(n, n) for h in range(self.head_num))
        out.stacked = att
        return out

    def packed_q(self):
        """Wq [128, 128], bq [128]: the H query projections stacked (G:129), one launch each on the GPU."""
        ws, bs = [l.weight for l in self.linears_q], [l.bias for l in self.linears_q]
        if ws[0].is_cuda:
            return PackRowsFn.apply(*ws), PackRowsFn.apply(*bs)
        return torch.cat(ws, 0), torch.cat(bs, 0)

    def forward_batched(self, x, batch: RaggedBatch):
        wq, bq = self.packed_q()
        keep = None
        if self.dropout is not None and self._dropping():
            ks = [self._keep((batch.total_pairs,), self.dropout.p, x.device) for _ in range(self.head_num)]
            keep = None if ks[0] is None else torch.stack(ks, 0)
        return MhaFn.apply(x, wq, bq, batch, self.head_num, keep)


# ------------------------------------------------------------------------------------- a2
class GATAttention(nn.Module, _KeepMixin):
    """GATAttention(att_input_dim, hidden_dim, dropout=0.1) -- G:144-168.

    ``apply_mask`` (default False = reference behaviour) opts in to the evidently intended
    in-place ``masked_fill`` of G:164.
    """

    def __init__(self, att_input_dim, hidden_dim, dropout=0.1):
        super().__init__()
        self.linear_node_h = nn.Linear(att_input_dim, hidden_dim)
        self.linear_node_t = nn.Linear(att_input_dim, hidden_dim)
        self.linear_edge_r = nn.Linear(att_input_dim, hidden_dim)
        self.wt = nn.Linear(hidden_dim * 3, 1)
        self.dropout = nn.Dropout(p=dropout)
        self.apply_mask = False
        self._hid = hidden_dim
        if att_input_dim != HIDDEN:
            raise _lib.GcgcnError("gcgcn_b200 supports att_input_dim 128 (G:234)")

    def collapse(self):
        """energy_ij = wt.[Wh x_j + bh ; Wt x_j + bt ; Wr e_ij + br] + b  ==  u.x_j + v.e_ij + c."""
        if self.wt.weight.is_cuda:          # one launch forward, one backward (no library gemv / dot calls)
            return GatCollapseFn.apply(self.linear_node_h.weight, self.linear_node_h.bias, self.linear_node_t.weight,
                                       self.linear_node_t.bias, self.linear_edge_r.weight, self.linear_edge_r.bias,
                                       self.wt.weight, self.wt.bias)
        h = self._hid                        # (host-side restatement, used by the CPU state_dict / shape tests only)
        w = self.wt.weight[0]
        w1, w2, w3 = w[:h], w[h:2 * h], w[2 * h:]
        u = self.linear_node_h.weight.t().mv(w1) + self.linear_node_t.weight.t().mv(w2)
        v = self.linear_edge_r.weight.t().mv(w3)
        c = (w1.dot(self.linear_node_h.bias) + w2.dot(self.linear_node_t.bias)
             + w3.dot(self.linear_edge_r.bias) + self.wt.bias[0])
        return u, v, c

    def forward(self, node_feat, edge_feat, mask=None):
        n = node_feat.size(0)
        bt = single_doc_batch(n, node_feat.device)
        att, ebar = self.forward_batched(node_feat, _pairs2d(edge_feat), bt, mask)
        _stash_ebar(edge_feat, ebar)
        return att.view(n, n)

    def forward_batched(self, x, edge, batch: RaggedBatch, mask=None):
        u, v, c = self.collapse()
        keep = None
        if self.dropout is not None and self._dropping():
            keep = self._keep((batch.total_pairs,), self.dropout.p, x.device)
        mask_u8 = None
        if self.apply_mask and mask is not None:
            mask_u8 = mask.reshape(-1).to(torch.uint8)
        return GatFn.apply(x, edge, u, v, c, batch, mask_u8, self.apply_mask, keep)


# ------------------------------------------------------------------------------------- a7 + a1 + a8
class GraphBlocks(nn.Module, _KeepMixin):
    """The graph hot path of one model, batched over documents: pooling -> CAGGC -> MAGGC ->
    classifier-side pair gathers, with the reference model's own attribute names so that the
    hot-path slice of a reference ``state_dict`` loads unchanged (G:254-262):
    ``get_weighted_adj_matrix``, ``get_adj_matrix.0``, ``graphcnn.0``, ``graphcnn.1``.

    layer_num/head_num: 2/8 for GCGCN_glove (G:250-251), 4/4 for the BERT variant (B:247-248).
    """

    def __init__(self, layer_num=2, head_num=8, alpha=1.0, hidden_size=HIDDEN, graph_hop=2, overlap=True):
        super().__init__()
        # overlap: stream the hop-1 edge tensor (mean forward, broadcast-write backward: pure HBM work)
        # on a side stream while the CAGGC block (tensor-core / latency-bound work) runs on the main one
        self.overlap = overlap
        # fused: route dropout-free passes through the block-level C-ABI entry points (gcgcn_caggc_*,
        # gcgcn_mha_stack_*) whose kernels keep the attention maps and their gradients on chip
        self.fused = True
        self.last_drop = {"caggc": None, "maggc": None}    # (seed, p_att, p_gcn) of the last train-mode block calls
        self._side = {}
        if graph_hop != 2:
            raise _lib.GcgcnError("graph_hop = 2 (config/Config.py:71) is the only supported depth")
        self.layerNum, self.headNum, self.alpha = layer_num, head_num, alpha
        self.get_weighted_adj_matrix = GATAttention(hidden_size, hidden_size)
        self.get_adj_matrix = nn.ModuleList([MultiHeadAttention(head_num, hidden_size)])
        self.graphcnn = nn.ModuleList([GraphConvolution(layer_num, hidden_size, hidden_size),
                                       MultiGraphConvolution(layer_num, head_num, hidden_size, hidden_size)])
        self.dropout = nn.Dropout(0.2)   # G:232

    def _blend(self, new, old):
        out = new if self.alpha == 1.0 else self.alpha * new + (1 - self.alpha) * old   # G:339
        keep = self._keep(out.shape, self.dropout.p, out.device) if self._dropping() else None
        return out if keep is None else out * keep                                     # G:341

    def _draw(self, att_p, gcn_p):
        if not self.training or (att_p <= 0.0 and gcn_p <= 0.0):
            return None
        return (int(torch.randint(0, 2 ** 62, (1,)).item()), att_p, gcn_p)

    def _fusable(self, x):
        # Block-level entry points: no mask tensors.  In train mode their kernels regenerate the dropout keep
        # factors from a per-call seed (drawn from torch's CPU generator, so torch.manual_seed reproduces a run);
        # masks injected by a test (inject_keep) take the one-entry-point-per-module route instead.
        gat, mha = self.get_weighted_adj_matrix, self.get_adj_matrix[0]
        cag, mag = self.graphcnn
        injected = any(getattr(m, "_injected", None) for m in (self, gat, mha, cag, mag))
        return self.fused and not injected and not gat.apply_mask and x.is_cuda

    def hop0(self, x0, e0, batch: RaggedBatch, adj=None):
        """CAGGC hop (G:330-341 at i = 0): returns (y1, a0)."""
        mask = None if adj is None else torch.eq(adj, 0)                                # G:330
        gat, cag = self.get_weighted_adj_matrix, self.graphcnn[0]
        cag_ok = self._fusable(x0) and (not self.training or block_supported(batch, 1, cag.layer_num, False))
        if cag_ok:
            u, v, c = gat.collapse()
            wn_x, w_e, winner = _pack_stack(cag.graphconv, 1, cag.layer_num, cag._g)
            drop = self._draw(gat.dropout.p if gat.training else 0.0, cag.gcn_dropout.p if cag.training else 0.0)
            self.last_drop["caggc"] = drop
            new, a0 = CaggcFn.apply(x0, e0, u, v, c, wn_x, w_e, winner, cag.linear_layer.weight,
                                    cag.linear_layer.bias, batch, cag.layer_num, drop)          # G:332-333
        else:
            a0, ebar0 = gat.forward_batched(x0, e0, batch, mask)                                # G:332
            new = cag.forward_batched(x0, ebar0, a0.view(1, -1), batch)                         # G:333
        return self._blend(new, x0), a0

    def hop1(self, y1, e1, batch: RaggedBatch, ebar1=None):
        """MAGGC hop (G:334-341 at i = 1): returns (y2, a1).  ``ebar1`` = an edge mean of e1 computed ahead."""
        mha, mag = self.get_adj_matrix[0], self.graphcnn[1]
        mag_ok = self._fusable(y1) and block_supported(batch, mag.head_num, mag.layer_num, True)
        a1 = None if mag_ok else mha.forward_batched(y1, batch)                                 # G:336
        if ebar1 is None:
            ebar1 = EdgeMeanFn.apply(e1, batch)
        if mag_ok:
            wq, bq = mha.packed_q()
            wn_x, w_e, winner = _pack_stack(mag.graphconv, mag.head_num, mag.layer_num, mag._g)
            drop = self._draw(mha.dropout.p if mha.training else 0.0, mag.gcn_dropout.p if mag.training else 0.0)
            self.last_drop["maggc"] = drop
            new, a1 = MhaStackFn.apply(y1, ebar1, wq, bq, wn_x, w_e, winner, mag.linear_layer.weight,
                                       mag.linear_layer.bias, batch, mag.head_num, mag.layer_num, drop)  # G:336-337
        else:
            new = mag.forward_batched(y1, ebar1, a1, batch)                                     # G:337
        return self._blend(new, y1), a1

    def forward(self, x0, e0, e1, batch: RaggedBatch, adj=None, with_node_feats: bool = True):
        """x0 [total_nodes,128]; e0, e1 [total_pairs,128] (fp32 or bf16); adj [total_pairs] or None.
        Returns y1, y2 and node_feats = cat[x0, x0, y1] (append-before-update, G:338; the classifier's input, skipped
        with ``with_node_feats=False``)."""
        ebar1 = None
        if self.overlap and e1.is_cuda:
            main = torch.cuda.current_stream(e1.device)
            side = self._side.get(e1.device)
            if side is None:
                side = self._side[e1.device] = torch.cuda.Stream(e1.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ebar1 = EdgeMeanFn.apply(e1, batch)
            ebar1.record_stream(main)
        y1, a0 = self.hop0(x0, e0, batch, adj)
        if ebar1 is not None:
            torch.cuda.current_stream(e1.device).wait_stream(self._side[e1.device])
        y2, a1 = self.hop1(y1, e1, batch, ebar1)
        out = {"y1": y1, "y2": y2, "a0": a0, "a1": a1}
        if with_node_feats:
            out["node_feats"] = torch.cat([x0, x0, y1], 1)
        return out


def pool_nodes(context: torch.Tensor, table: PoolTable) -> torch.Tensor:
    """Mention->entity pooling (G:297-298) over the concatenated context of a batch."""
    return PoolFn.apply(context, table)


def pair_gather(node_feats_with_type: torch.Tensor, dis_embed: Optional[torch.Tensor], tables: PairTables,
                batch: RaggedBatch):
    """Classifier-side h/t pair tensors (G:351-352); ``dis_embed=None`` gives the in-loop form (G:321-322)."""
    return PairGatherFn.apply(node_feats_with_type, dis_embed, tables, batch)


def pair_dense(node_feats_with_type: torch.Tensor, dense_layer: nn.Linear, dis_embed: torch.Tensor,
               tables: PairTables, batch: RaggedBatch):
    """entity_feature_h, entity_feature_t of G:351-355 -- ``tanh(dense_layer(cat(F[j], dis[10 + rp])))`` and
    ``tanh(dense_layer(cat(F[i], dis[10 - rp])))`` -- without forming the two gathered ``[n, n, 424]`` tensors:
    ``dense_layer`` is split by input columns into a node part and a distance part, both applied at node / table
    level (``[rows, 404] x [404, 128]`` and ``[21, 20] x [20, 128]`` on gcgcn_gemm), and one kernel adds the two
    128-vectors per pair and applies tanh.  Same values as ``tanh(dense_layer(pair_gather(...)))``."""
    fw = node_feats_with_type.shape[1]
    w = dense_layer.weight
    U = LinearFn.apply(node_feats_with_type, w[:, :fw], None)
    Vd = LinearFn.apply(dis_embed, w[:, fw:], dense_layer.bias)
    return PairDenseFn.apply(U, Vd, tables, batch)
