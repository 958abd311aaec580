"""Pin the oracle restatement against the reference's own modules (build container only).

Run:  python oracle/pin_against_reference.py
Compares, on the 12-document synthetic batch and both model
configurations (GloVe: L_s=2,H=8; BERT variant: L_s=4,H=4):
  * every oracle function's forward output with the matching reference module
    (G:18-168) in eval mode and in train mode with injected keep masks,
  * gradients w.r.t. node features, edge features and every parameter,
  * pooling (G:297-298), pair gathers (G:351-352, 321-322) and the host index
    builders (C:169-176, 207-217) against literal executions of the reference lines.
Prints the max abs difference per check and exits non-zero on any mismatch above
``TOL`` (0.0: bit-exact is expected because the same ATen ops are issued).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gcgcn_oracle as O            # noqa: E402
from oracle import reference_loader as R        # noqa: E402
from gcgcn_b200 import synthetic as S           # noqa: E402

TOL = 0.0


class _MaskDrop(torch.nn.Module):
    """Stands in for an nn.Dropout instance on a reference module: multiplies by the
    next queued keep-scale mask (SURVEY.md section 8c 'Dropout parity')."""

    def __init__(self, masks):
        super().__init__()
        self.masks = list(masks)
        self.i = 0

    def forward(self, x):
        m = self.masks[self.i % len(self.masks)]
        self.i += 1
        return x * m


def _sd(mod):
    return {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}


def _maxdiff(a, b):
    return float((a - b).abs().max()) if a.numel() else 0.0


def run(variant: str, layers: int, heads: int, report):
    gat, mha, cag, mag = R.build_graph_modules(layers, heads, seed=0, variant=variant)
    docs = S.make_batch()
    worst = 0.0
    for train in (False, True):
        for d in docs:
            keep = S.make_keep_masks(d.doc_id, d.n, layers, heads) if train else None
            # ---- reference, module by module, wired like G:329-341
            if train:
                gat.dropout = _MaskDrop([keep["gat"]])
                cag.gcn_dropout = _MaskDrop(keep["cag"])
                mha.dropout = _MaskDrop(keep["mha"])
                mag.gcn_dropout = _MaskDrop([m for hm in keep["mag"] for m in hm])
            else:
                gat.dropout = torch.nn.Dropout(0.1).eval()
                cag.gcn_dropout = torch.nn.Dropout(0.2).eval()
                mha.dropout = torch.nn.Dropout(0.1).eval()
                mag.gcn_dropout = torch.nn.Dropout(0.2).eval()
            for m in (gat, mha, cag, mag):
                m.zero_grad()
            x0 = d.x0.clone().requires_grad_(True)
            e0 = d.e0.clone().requires_grad_(True)
            e1 = d.e1.clone().requires_grad_(True)
            mask = torch.eq(d.adj, 0)
            a0 = gat(x0, e0, mask)
            y1 = cag(x0, e0, a0)
            if train:
                y1 = y1 * keep["out0"]
            a1 = mha(y1, e1)
            y2 = mag(y1, e1, a1)
            if train:
                y2 = y2 * keep["out1"]
            gen = torch.Generator().manual_seed(99 + d.doc_id)
            dy1 = torch.randn(y1.shape, generator=gen)
            dy2 = torch.randn(y2.shape, generator=gen)
            ((y1 * dy1).sum() + (y2 * dy2).sum()).backward()

            # ---- oracle
            ps = [_sd(m) for m in (gat, cag, mha, mag)]
            ox0 = d.x0.clone().requires_grad_(True)
            oe0 = d.e0.clone().requires_grad_(True)
            oe1 = d.e1.clone().requires_grad_(True)
            r = O.graph_blocks(ox0, oe0, oe1, d.adj, ps[0], ps[1], ps[2], ps[3], layers, heads,
                               alpha=1.0, keep=keep)
            ((r["y1"] * dy1).sum() + (r["y2"] * dy2).sum()).backward()

            checks = {
                "a0": _maxdiff(a0, r["a0"]), "y1": _maxdiff(y1, r["y1"]), "y2": _maxdiff(y2, r["y2"]),
                "a1": max(_maxdiff(p, q) for p, q in zip(a1, r["a1"])),
                "dx0": _maxdiff(x0.grad, ox0.grad), "de0": _maxdiff(e0.grad, oe0.grad),
                "de1": _maxdiff(e1.grad, oe1.grad),
            }
            for mod, p in zip((gat, cag, mha, mag), ps):
                for k, v in mod.named_parameters():
                    if v.grad is None:
                        assert p[k].grad is None, f"{k}: reference grad None, oracle not"
                        continue
                    checks["d" + k] = max(checks.get("d" + k, 0.0), _maxdiff(v.grad, p[k].grad))
            w = max(checks.values())
            worst = max(worst, w)
            report(f"{variant} L{layers}H{heads} train={int(train)} doc{d.doc_id:02d} n={d.n:2d} "
                   f"max|diff|={w:.3e}")
    return worst


def run_pool_and_gathers(report):
    """Literal executions of G:297-298, G:306-307, G:351-352, G:321-322, C:169-176, C:207-217."""
    worst = 0.0
    g = torch.Generator().manual_seed(5)
    dis = torch.randn(21, 20, generator=g)
    ner = torch.randn(7, 20, generator=g)
    ner[0] = 0
    for d in S.make_batch():
        # C:169-176, 223
        node_pos = np.zeros((d.n, d.L))
        for node in range(d.n):
            for position in d.spans[node]:
                node_pos[node, position[0]:position[1]] = 1.0 / (position[1] - position[0])
            node_pos[node, :] *= 1.0 / len(d.spans[node])
        node_pos = torch.FloatTensor(node_pos[:, :512])
        worst = max(worst, _maxdiff(node_pos, O.build_node_pos(d.spans, d.L)))
        # G:297-298
        context_output = d.ctx.unsqueeze(0)
        node_feat = node_pos.unsqueeze(2).expand(-1, -1, 128) * context_output.expand(d.n, -1, -1)
        node_feat = node_feat.sum(dim=1)
        worst = max(worst, _maxdiff(node_feat, O.pool_nodes(node_pos, context_output)))
        # C:106-116, C:207-217
        dis2idx = np.zeros((1024), dtype="int64")
        dis2idx[1] = 1
        dis2idx[2:] = 2
        dis2idx[4:] = 3
        dis2idx[8:] = 4
        dis2idx[16:] = 5
        dis2idx[32:] = 6
        dis2idx[64:] = 7
        dis2idx[128:] = 8
        dis2idx[256:] = 9
        dis2idx[512:] = 10
        nrp = np.zeros((d.n, d.n))
        for a in range(d.n):
            for b in range(d.n):
                if a == b:
                    continue
                rel = d.first_pos[a] - d.first_pos[b]
                nrp[a, b] = -dis2idx[-rel] if rel < 0 else dis2idx[rel]
        nrp = torch.LongTensor(nrp)
        assert torch.equal(nrp, O.build_node_relative_pos(d.first_pos))
        # G:306-307, 344-352
        feats = torch.cat([d.x0, d.x0, torch.tanh(d.x0)], 1)
        fwt = torch.cat([feats, torch.nn.functional.embedding(d.node_type, ner)], 1)
        rel_h = torch.nn.functional.embedding(10 + nrp, dis)
        rel_t = torch.nn.functional.embedding(10 - nrp, dis)
        ph = torch.cat([fwt.unsqueeze(0).expand(d.n, -1, -1), rel_h], -1)
        pt = torch.cat([fwt.unsqueeze(1).expand(-1, d.n, -1), rel_t], -1)
        oh, ot = O.pair_gather_classifier(O.node_feats_with_type(feats, d.node_type, ner), nrp, dis)
        worst = max(worst, _maxdiff(ph, oh), _maxdiff(pt, ot))
        # G:321-322
        cw = torch.zeros(d.n, d.n, d.S, 128)
        eh = d.x0.unsqueeze(0).unsqueeze(2).expand_as(cw)
        et = d.x0.unsqueeze(1).unsqueeze(2).expand_as(cw)
        qh, qt = O.pair_gather_inloop(d.x0, d.S)
        worst = max(worst, _maxdiff(eh, qh), _maxdiff(et, qt))
    report(f"pool / node_pos / rel_pos / pair gathers: max|diff|={worst:.3e}")
    return worst


def main():
    if not R.available():
        print("reference tree not present; nothing to pin against")
        return 2
    torch.set_num_threads(1)     # one thread -> run-to-run identical reduction order
    lines = []
    worst = max(run("glove", 2, 8, lines.append), run("bert", 4, 4, lines.append),
                run_pool_and_gathers(lines.append))
    print("\n".join(lines))
    print(f"WORST max|diff| = {worst:.3e} (tolerance {TOL})")
    return 0 if worst <= TOL else 1


if __name__ == "__main__":
    sys.exit(main())
