"""CPU oracle for the edge-feature producer and the relation classifier of GCGCN.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import it, and only as the checker.  Nothing under ``gcgcn_b200/`` imports it.

It restates, in stock fp32 PyTorch on the CPU and "as written" (dense ``[n, n, S, L, 128]`` tensors, no use of
the algebraic slack the CUDA path exploits), what the reference computes between the document encoder and the
logits: G = /root/reference/models/GCGCN_glove.py.

    word_attention       G:178-192   WordAttention.forward
    sentence_attention   G:201-214   SentenceAttention.forward
    edge_features        G:299-327   one hop of the in-loop producer of ``context_sent_att`` (the hot path's edge_feat)
    classifier           G:343-358   node_feats -> logits (dense_layer, Bilinear, Linear)
    graph_head           G:293-358   everything after ``context_output`` for one document
    loss_as_written      C:355-364   the trainer's per-pair BCE loop on sigmoid(logits)

Parity pin: ``tests/golden/make_golden_edge.py`` (build container only) builds the UNMODIFIED reference models --
GCGCN_glove and the BERT variant, loaded by path -- loads the deterministic head weights of
``tests/helpers.head_state`` into them, runs their own ``forward`` and the trainer's literal loss loop, and requires
``graph_head`` + ``loss_as_written`` below to reproduce logits, both hops' ``context_sent_att``, the loss, d loss / d
``context_output`` and every head parameter's gradient (worst relative difference 2.9e-7); the same run writes
``tests/golden/edge_head.npz`` / ``edge_head_bert.npz``, which the CPU and GPU tests compare against.

Parameters use the reference's ``state_dict`` key names relative to each module (``attention_sent.weight`` ...).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import gcgcn_oracle as O

Tensor = torch.Tensor
Params = Dict[str, Tensor]


def _lin(x: Tensor, p: Params, name: str) -> Tensor:
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def word_attention(att_padding: Tensor, ctx_before: Tensor, dis_embedding: Tensor, p: Params) -> Tensor:
    """WordAttention.forward, G:178-192.  att_padding [n,n,S,L,1] bool (True = padded), ctx_before [n,n,S,L,128]
    (an expanded view of the same [L,128] matrix), dis_embedding [n,n,S,L,20] -> [n,n,S,128]."""
    sent_feat = _lin(ctx_before, p, "attention_sent")                       # G:179
    dis_feat = _lin(dis_embedding, p, "attention_pos")                      # G:180
    all_feat = _lin(torch.tanh(sent_feat + dis_feat), p, "attention_all")   # G:183
    att_padding = att_padding.expand_as(all_feat)                           # G:185
    all_feat = all_feat.masked_fill(att_padding, -100000.0)                 # G:186
    att = F.softmax(all_feat, dim=3).expand_as(ctx_before)                  # G:187
    return torch.sum(att * ctx_before, dim=3)                               # G:188


def sentence_attention(att_padding: Tensor, ctx_word_att: Tensor, node_embedding: Tensor, p: Params) -> Tensor:
    """SentenceAttention.forward, G:201-214.  att_padding [n,n,S,1] bool, ctx_word_att / node_embedding [n,n,S,128]
    -> [n,n,128].  Quirks kept: ``sent_num`` counts the PADDED slots (G:206), the weights are relu, not softmax
    (G:212), and the division is by (sent_num + 1e-10) (G:213)."""
    sent_feat = _lin(ctx_word_att, p, "attention_sent")                     # G:202
    dis_feat = _lin(node_embedding, p, "attention_pos")                     # G:203
    all_feat = _lin(torch.tanh(sent_feat + dis_feat), p, "attention_all")   # G:204
    sent_num = att_padding.sum(dim=2)                                       # G:206
    all_feat = all_feat.masked_fill(att_padding, -100000.0)                 # G:209
    att = torch.relu(all_feat).expand_as(ctx_word_att)                      # G:212
    return torch.sum(att * ctx_word_att, dim=2) / (sent_num + 1e-10)        # G:213


def edge_features(ctx: Tensor, node_feat: Tensor, sen: Tensor, pos_h: Tensor, pos_t: Tensor, dis_table: Tensor,
                  p_word: Params, p_sent: Params, p_lw: Params, p_ls: Params) -> Tensor:
    """One hop of G:299-327: ``context_sent_att`` [n,n,128] from ctx [L,128] (``context_output[0]``), the current
    node features [n,128], the dense pair-context tensors sen (bool) / pos_h / pos_t [n,n,S,L] and the distance
    embedding table [21,20].  p_word / p_sent / p_lw / p_ls = word_attention[i], sentence_attention[i],
    linear_word_att[i], linear_sentence_att[i].  Note the token-0 quirk of G:302: a slot counts as padding unless
    its sentence contains token 0."""
    n, S, L = sen.size(0), sen.size(2), sen.size(3)
    word_pad = ~sen.unsqueeze(4)                                            # G:299
    ctx_before = ctx.unsqueeze(0).unsqueeze(0).unsqueeze(0).expand(n, n, S, -1, -1)     # G:300
    sent_pad = ~sen[:, :, :, 0:1]                                           # G:302
    dis_h = F.embedding(pos_h, dis_table)                                   # G:304
    dis_t = F.embedding(pos_t, dis_table)                                   # G:305
    cwa_h = word_attention(word_pad, ctx_before, dis_h, p_word)             # G:313
    cwa_t = word_attention(word_pad, ctx_before, dis_t, p_word)             # G:314
    cwa = torch.cat([cwa_h, cwa_t], 3)                                      # G:316
    cwa = F.linear(cwa, p_lw["weight"], p_lw["bias"])                       # G:317
    emb_h = node_feat.unsqueeze(0).unsqueeze(2).expand_as(cwa)              # G:320
    emb_t = node_feat.unsqueeze(1).unsqueeze(2).expand_as(cwa)              # G:321
    csa_h = sentence_attention(sent_pad, cwa, emb_h, p_sent)                # G:323
    csa_t = sentence_attention(sent_pad, cwa, emb_t, p_sent)                # G:324
    csa = torch.cat([csa_h, csa_t], 2)                                      # G:325
    return F.linear(csa, p_ls["weight"], p_ls["bias"])                      # G:326


def classifier(node_feats: Tensor, node_type: Tensor, rel_pos: Tensor, params: Params, dis_plus: int = 10):
    """G:343-358: node_feats [n,384] = cat[x0,x0,y1] -> logits [n,n,R].  params: ner_emb.weight, dis_embed.weight,
    dense_layer.*, bili_layer_01.*, classification_layer_01.* (top-level key names)."""
    n = node_feats.size(0)
    type_feats = F.embedding(node_type, params["ner_emb.weight"])           # G:345 (padding_idx only affects grads)
    with_type = torch.cat([node_feats, type_feats], 1)                      # G:346
    rp_h = F.embedding(dis_plus + rel_pos, params["dis_embed.weight"])      # G:306
    rp_t = F.embedding(dis_plus - rel_pos, params["dis_embed.weight"])      # G:307
    pos_h = torch.cat([with_type.unsqueeze(0).expand(n, -1, -1), rp_h], -1)  # G:351
    pos_t = torch.cat([with_type.unsqueeze(1).expand(-1, n, -1), rp_t], -1)  # G:352
    fh = torch.tanh(_lin(pos_h, params, "dense_layer"))                     # G:354
    ft = torch.tanh(_lin(pos_t, params, "dense_layer"))                     # G:355
    feat = torch.cat([fh, ft], -1)                                          # G:356
    bil = F.bilinear(fh.contiguous(), ft.contiguous(), params["bili_layer_01.weight"], params["bili_layer_01.bias"])
    return bil + _lin(feat, params, "classification_layer_01"), fh, ft     # G:358


def loss_as_written(logits: Tensor, label: Tensor) -> Tensor:
    """The trainer's loss, C:355-364: ``predict_re = sigmoid(logits)`` then, pair by pair, BCELoss (mean over the R
    relation slots) summed over ALL ordered pairs i != j and divided by n(n-1) (the label-mask branch at C:359 is
    commented out in the reference, so every off-diagonal pair counts).  Restated without the Python double loop."""
    n = logits.size(0)
    pred = torch.sigmoid(logits)                                            # C:355
    per_pair = F.binary_cross_entropy(pred, label, reduction="none").mean(dim=-1)   # BCELoss() per pair, C:361
    off = ~torch.eye(n, dtype=torch.bool)
    return per_pair[off].sum() / max(int(off.sum()), 1)


def graph_head(ctx: Tensor, node_pos: Tensor, sen: Tensor, pos_h: Tensor, pos_t: Tensor, adj: Tensor,
               node_type: Tensor, rel_pos: Tensor, params: Params, layers: int, heads: int, alpha: float = 1.0,
               cls_feat: Optional[Tensor] = None) -> dict:
    """G:293-358 for one document in eval mode: pooling, two hops of (edge-feature producer -> graph block),
    classifier.  ``params`` uses the top-level model's key names.  ``cls_feat`` [768] (BERT variant, B:277) adds
    ``linear_cls(cls_feat)`` to every pair's logits (B:346-347)."""
    x0 = O.pool_nodes(node_pos, ctx.unsqueeze(0))                           # G:297-298
    node_feat, feats, edges = x0, [x0], []
    for i in range(2):                                                      # G:310 (graph_hop = 2)
        e = edge_features(ctx, node_feat, sen, pos_h, pos_t, params["dis_embed.weight"],
                          O._sub(params, f"word_attention.{i}"), O._sub(params, f"sentence_attention.{i}"),
                          O._sub(params, f"linear_word_att.{i}"), O._sub(params, f"linear_sentence_att.{i}"))
        edges.append(e)
        if i < 1:
            mask = torch.eq(adj, 0)                                         # G:330
            a = O.gat_attention(node_feat, e, O._sub(params, "get_weighted_adj_matrix"), mask)   # G:332
            new = O.caggc_conv(node_feat, e, a, O._sub(params, "graphcnn.0"), layers)             # G:333
        else:
            atts = O.mha_attention(node_feat, O._sub(params, "get_adj_matrix.0"), heads)          # G:336
            new = O.maggc_conv(node_feat, e, atts, O._sub(params, "graphcnn.1"), layers, heads)    # G:337
        feats.append(node_feat)                                             # G:338 (append BEFORE the update)
        node_feat = alpha * new + (1 - alpha) * node_feat                   # G:339
    node_feats = torch.cat(feats, 1)                                        # G:343
    logits, fh, ft = classifier(node_feats, node_type, rel_pos, params)
    if cls_feat is not None:
        n = logits.size(0)
        logits = logits + _lin(cls_feat, params, "linear_cls").unsqueeze(0).unsqueeze(0).expand(n, n, -1)   # B:346-347
    return {"x0": x0, "e0": edges[0], "e1": edges[1], "y1": feats[2], "y2": node_feat, "logits": logits,
            "entity_feature_h": fh, "entity_feature_t": ft}
