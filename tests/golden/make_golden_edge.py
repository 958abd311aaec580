"""Golden vectors for the edge-feature producer and the classifier, generated from the UNMODIFIED reference model
(build container only), and the pin of oracle/edge_oracle.py against it.

Run:  python tests/golden/make_golden_edge.py
Builds the reference's GCGCN_glove (loaded by path, oracle/reference_loader.py), loads the deterministic head weights
of tests/helpers.head_state into it, runs its own ``forward`` on synthetic pickle-style records
(gcgcn_b200.synthetic.make_record -> the dense tensors of Config.from_list_to_tensor via oracle/featurize_oracle.py,
which is pinned to the reference's own function) and the trainer's loss (config/Config.py:355-364, the literal double
loop), and stores per document:
    ctx            context_output[0]  [L, 128]  (the encoder is out of scope: its output is the head's input)
    e0, e1         context_sent_att of hop 0 / 1 (forward hooks on linear_sentence_att[i])
    logits, loss   the model's output and the trainer's loss
    dctx           d loss / d context_output
    g_<param>      gradients of the small head parameters; strided samples of the large ones
It then checks that oracle/edge_oracle.graph_head reproduces logits, e0, e1, loss and dctx (the pin).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import reference_loader as R            # noqa: E402
from oracle import edge_oracle as EO                # noqa: E402
from oracle import featurize_oracle as FO           # noqa: E402
from gcgcn_b200 import synthetic as S               # noqa: E402
from helpers import head_labels, head_shapes, head_state   # noqa: E402

# (record seed, n, L, S): small graphs (the dense reference path is O(n^2 S L 128)); every one has pairs whose
# sentence contains token 0 (active slots) and pairs that share only later sentences
DOCS = [(8, 7, 150, None), (10, 9, 230, None), (2, 6, 90, 3), (10, 10, 600, None)]
BIG_STRIDE = 997


BERT_DOCS = [(8, 7, 150, None), (2, 6, 90, 3)]


def build_reference_bert(state):
    m = R.bert_module()
    cfg = types.SimpleNamespace(entity_type_size=20, coref_size=20, max_length=512, keep_prob=1.0, graph_hop=2, dis_size=20,
                                dis_num=21, dis_plus=10, relation_num=97, alpha=1.0)
    torch.manual_seed(0)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = m.GraphCNN_multihead_bert_gate_cls(cfg)
    sd = model.state_dict()
    for k, shape in head_shapes(4, 4, cls_dim=768).items():
        assert tuple(sd[k].shape) == tuple(shape), (k, sd[k].shape, shape)
    missing = model.load_state_dict(state, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    return model.eval()


def build_reference(state):
    m = R.glove_module()
    rng = np.random.RandomState(0)
    cfg = types.SimpleNamespace(data_word_vec=rng.randn(1000, 100).astype(np.float32), entity_type_size=20,
                                coref_size=20, max_length=512, keep_prob=1.0, graph_hop=2, dis_size=20, dis_num=21,
                                dis_plus=10, relation_num=97, alpha=1.0)
    torch.manual_seed(0)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = m.GCGCN_glove(cfg)
    sd = model.state_dict()
    for k, shape in head_shapes().items():
        assert tuple(sd[k].shape) == tuple(shape), (k, sd[k].shape, shape)
    missing = model.load_state_dict(state, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    return model.eval()


def reference_pass(model, item, labels, shapes=None):
    shapes = head_shapes() if shapes is None else shapes
    t = FO.from_list_to_tensor(item)
    L = min(len(item["document"]), 512)
    tt = lambda k, dt: torch.from_numpy(np.asarray(t[k])).to(dt)
    captured = {}

    def grab_pre(_, __, out):
        out.retain_grad()
        captured["pre"] = out
    hooks = [model.linear_re.register_forward_hook(grab_pre)]
    if hasattr(model, "bert"):                # BERT variant: the first-token feature of the (stubbed) encoder, B:277
        hooks.append(model.bert.register_forward_hook(
            lambda _, __, out: captured.__setitem__("cls_feat", out[0][0, 0, :].detach().clone())))
    for i in range(2):
        hooks.append(model.linear_sentence_att[i].register_forward_hook(
            lambda _, __, out, i=i: captured.__setitem__(f"e{i}", out.detach().clone())))
    model.zero_grad()
    logits = model(torch.LongTensor(item["document"][:512]), torch.LongTensor(item["document_ner"][:512]),
                   torch.LongTensor(item["document_pos"][:512]), tt("adj_matrix", torch.float32),
                   tt("sen_matrix", torch.bool), tt("pos_matrix_h", torch.int64), tt("pos_matrix_t", torch.int64),
                   tt("node_pos", torch.float32), tt("node_type", torch.int64), tt("node_relative_pos", torch.int64))
    for h in hooks:
        h.remove()
    # the trainer's loss, written as the trainer writes it (C:355-364)
    bce = torch.nn.BCELoss()
    pred = torch.sigmoid(logits)
    n = labels.size(0)
    temp = torch.zeros(1, requires_grad=True)
    for hi in range(n):
        for tj in range(n):
            if hi == tj:
                continue
            temp = temp + bce(pred[hi][tj], labels[hi][tj])
    loss = temp / (n * n - n)
    loss.backward()
    pre = captured["pre"]
    ctx = torch.tanh(pre.detach())[0]
    dctx = (pre.grad[0].double() / (1.0 - ctx.double() ** 2)).float()      # through tanh: d/dctx = d/dpre / (1 - ctx^2)
    # ner_emb is shared with the (out-of-scope) encoder input (G:286), so its gradient is not a head-only quantity
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()
             if k in shapes and k != "ner_emb.weight" and p.grad is not None}
    return {"cls_feat": captured.get("cls_feat"), "ctx": ctx, "e0": captured["e0"], "e1": captured["e1"], "logits": logits.detach(), "loss": loss.detach(),
            "dctx": dctx, "grads": grads, "tensors": t, "L": L}


def oracle_pass(state, ref, labels, layers=2, heads=8):
    ps = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    t = ref["tensors"]
    ctx = ref["ctx"].clone().requires_grad_(True)
    tt = lambda k, dt: torch.from_numpy(np.asarray(t[k])).to(dt)
    r = EO.graph_head(ctx, tt("node_pos", torch.float32), tt("sen_matrix", torch.bool), tt("pos_matrix_h", torch.int64),
                      tt("pos_matrix_t", torch.int64), tt("adj_matrix", torch.float32), tt("node_type", torch.int64),
                      tt("node_relative_pos", torch.int64), ps, layers, heads, cls_feat=ref.get("cls_feat"))
    loss = EO.loss_as_written(r["logits"], labels)
    loss.backward()
    return r, loss.detach(), ctx.grad, {k: v.grad for k, v in ps.items() if v.grad is not None}


def main():
    state = head_state(0)
    model = build_reference(state)
    out = {}
    worst = 0.0
    for idx, (seed, n, L, Sx) in enumerate(DOCS):
        item = S.make_record(seed, n=n, L=L, S=Sx)
        labels = head_labels(seed, n)
        ref = reference_pass(model, item, labels)
        sen = np.asarray(ref["tensors"]["sen_matrix"])
        active = int(sen[:, :, :, 0].sum())
        sent_num = (~sen[:, :, :, 0]).sum(-1)
        assert active > 0, f"record {seed}: no active slot"
        assert sent_num.min() >= 1, f"record {seed}: a pair with every slot active (division by 1e-10)"
        r, loss, dctx, grads = oracle_pass(state, ref, labels)
        diffs = {"logits": (r["logits"] - ref["logits"]).abs().max().item(),
                 "e0": (r["e0"] - ref["e0"]).abs().max().item(), "e1": (r["e1"] - ref["e1"]).abs().max().item(),
                 "loss": (loss - ref["loss"]).abs().max().item(),
                 "dctx": (dctx - ref["dctx"]).abs().max().item() / max(ref["dctx"].abs().max().item(), 1e-30)}
        for k, g in ref["grads"].items():
            # (softmax-shift parameters such as the GAT biases have a true gradient of 0: pure rounding noise, so the
            # scale of the comparison is floored at 1e-3)
            d = (grads[k] - g).abs().max().item() / max(g.abs().max().item(), 1e-3)
            diffs["g_" + k] = d
        bad = {k: v for k, v in diffs.items() if v > 2e-5}
        print(f"record {seed}: n={n} L={ref['L']} S={sen.shape[2]} active slots={active} "
              f"max oracle-vs-reference diff {max(diffs.values()):.2e}")
        assert not bad, bad
        worst = max(worst, max(diffs.values()))
        p = f"d{idx}_"
        out[p + "meta"] = np.asarray([seed, n, L, -1 if Sx is None else Sx, active], dtype=np.int64)
        for k in ("ctx", "e0", "e1", "logits", "loss", "dctx"):
            out[p + k] = ref[k].numpy()
        for k, g in ref["grads"].items():
            a = g.numpy().reshape(-1)
            out[p + "g_" + k] = a if a.size <= 40000 else a[::BIG_STRIDE].copy()
    np.savez_compressed(os.path.join(HERE, "edge_head.npz"), **out)
    print(f"wrote edge_head.npz ({os.path.getsize(os.path.join(HERE, 'edge_head.npz')) / 1e6:.2f} MB); "
          f"oracle pinned: worst relative diff {worst:.2e}")

    # ---- the BERT variant (models/GraphCNN_multihead_bert_gate_cls.py): L_s = 4, H = 4, + linear_cls on the encoder's
    # first-token feature.  BertModel is the stub of oracle/reference_loader.py (the encoder is out of scope: its output
    # is captured and becomes the head's input).
    state = head_state(0, 4, 4, cls_dim=768)
    shapes = head_shapes(4, 4, cls_dim=768)
    model = build_reference_bert(state)
    out, worst = {}, 0.0
    for idx, (seed, n, L, Sx) in enumerate(BERT_DOCS):
        item = S.make_record(seed, n=n, L=L, S=Sx)
        labels = head_labels(seed, n)
        ref = reference_pass(model, item, labels, shapes)
        r, loss, dctx, grads = oracle_pass(state, ref, labels, 4, 4)
        diffs = {"logits": (r["logits"] - ref["logits"]).abs().max().item(), "loss": (loss - ref["loss"]).abs().max().item(),
                 "dctx": (dctx - ref["dctx"]).abs().max().item() / max(ref["dctx"].abs().max().item(), 1e-30)}
        for k, g in ref["grads"].items():
            diffs["g_" + k] = (grads[k] - g).abs().max().item() / max(g.abs().max().item(), 1e-3)
        bad = {k: v for k, v in diffs.items() if v > 2e-5}
        print(f"bert record {seed}: n={n} max oracle-vs-reference diff {max(diffs.values()):.2e}")
        assert not bad, bad
        worst = max(worst, max(diffs.values()))
        p = f"d{idx}_"
        out[p + "meta"] = np.asarray([seed, n, L, -1 if Sx is None else Sx, 0], dtype=np.int64)
        for k in ("ctx", "cls_feat", "logits", "loss", "dctx"):
            out[p + k] = ref[k].numpy()
        for k, g in ref["grads"].items():
            a = g.numpy().reshape(-1)
            out[p + "g_" + k] = a if a.size <= 40000 else a[::BIG_STRIDE].copy()
    np.savez_compressed(os.path.join(HERE, "edge_head_bert.npz"), **out)
    print(f"wrote edge_head_bert.npz ({os.path.getsize(os.path.join(HERE, 'edge_head_bert.npz')) / 1e6:.2f} MB); "
          f"oracle pinned: worst relative diff {worst:.2e}")


if __name__ == "__main__":
    main()
