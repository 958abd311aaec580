"""Generate the committed golden vectors from the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden.py
Imports the reference's own modules from /root/reference (by path, see
oracle/reference_loader.py), runs them on the synthetic DocRED-shaped batch of SURVEY.md
section 8d and writes small .npz fixtures next to this script.  Inputs are NOT stored: they are
regenerated from seeds by gcgcn_b200.synthetic (CPU torch generators are deterministic for a
given torch build; an input checksum is stored to detect drift).

Fixtures
  graph_blocks_<variant>.npz   variant glove (L_s=2,H=8, G:250-251) / bert (L_s=4,H=4, B:247-248)
      eval mode, 12 docs: y1, y2, dx0 (node rows concatenated), de0/de1 checksums,
      parameter gradients summed over the batch (checksums + strided samples),
      train mode (keep masks injected), docs 0/5/11: y1, y2, dx0
      weights_sha256: hash of the reference modules' parameters under torch.manual_seed(0)
  pool_pairs.npz               pooled x0 per doc, node_relative_pos, pair-gather samples
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_loader as R          # noqa: E402
from oracle.pin_against_reference import _MaskDrop  # noqa: E402
from gcgcn_b200 import synthetic as S             # noqa: E402

SAMPLE_STRIDE = 37
TRAIN_DOCS = (0, 5, 11)


def weights_sha256(mods) -> str:
    h = hashlib.sha256()
    for m in mods:
        for k, v in m.state_dict().items():
            h.update(k.encode())
            h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def upstream(doc_id, shape1, shape2):
    gen = torch.Generator().manual_seed(99 + doc_id)
    return torch.randn(shape1, generator=gen), torch.randn(shape2, generator=gen)


def run_reference(mods, d, keep):
    gat, mha, cag, mag = mods
    if keep is not None:
        gat.dropout = _MaskDrop([keep["gat"]])
        cag.gcn_dropout = _MaskDrop(keep["cag"])
        mha.dropout = _MaskDrop(keep["mha"])
        mag.gcn_dropout = _MaskDrop([m for hm in keep["mag"] for m in hm])
    else:
        gat.dropout = torch.nn.Dropout(0.1).eval()
        cag.gcn_dropout = torch.nn.Dropout(0.2).eval()
        mha.dropout = torch.nn.Dropout(0.1).eval()
        mag.gcn_dropout = torch.nn.Dropout(0.2).eval()
    x0 = d.x0.clone().requires_grad_(True)
    e0 = d.e0.clone().requires_grad_(True)
    e1 = d.e1.clone().requires_grad_(True)
    a0 = gat(x0, e0, torch.eq(d.adj, 0))          # G:330-332
    y1 = cag(x0, e0, a0)                          # G:333, alpha = 1 (C:74)
    if keep is not None:
        y1 = y1 * keep["out0"]                    # G:341
    a1 = mha(y1, e1)                              # G:336
    y2 = mag(y1, e1, a1)                          # G:337
    if keep is not None:
        y2 = y2 * keep["out1"]
    dy1, dy2 = upstream(d.doc_id, y1.shape, y2.shape)
    ((y1 * dy1).sum() + (y2 * dy2).sum()).backward()
    return y1.detach(), y2.detach(), x0.grad, e0.grad, e1.grad


def graph_blocks(variant, layers, heads):
    mods = R.build_graph_modules(layers, heads, seed=0, variant=variant)
    out = {"weights_sha256": np.frombuffer(weights_sha256(mods).encode(), dtype=np.uint8),
           "layers": np.int64(layers), "heads": np.int64(heads)}
    docs = S.make_batch()
    for m in mods:
        m.zero_grad()
    y1s, y2s, dxs, de_sums, in_sums = [], [], [], [], []
    for d in docs:
        y1, y2, dx, de0, de1 = run_reference(mods, d, None)
        y1s.append(y1), y2s.append(y2), dxs.append(dx)
        de_sums.append([float(de0.double().sum()), float(de0.double().abs().sum()),
                        float(de1.double().sum()), float(de1.double().abs().sum())])
        in_sums.append([float(d.x0.double().sum()), float(d.e0.double().sum()), float(d.e1.double().sum())])
    out["y1"] = torch.cat(y1s).numpy()
    out["y2"] = torch.cat(y2s).numpy()
    out["dx0"] = torch.cat(dxs).numpy()
    out["de_checksums"] = np.asarray(de_sums)
    out["input_checksums"] = np.asarray(in_sums)
    names, stats, samples = [], [], []
    for prefix, m in zip(("get_weighted_adj_matrix", "get_adj_matrix.0", "graphcnn.0", "graphcnn.1"), mods):
        for k, p in m.named_parameters():
            if p.grad is None:
                continue     # linears_k: never used, never receives a gradient (quirk 3)
            g = p.grad.double().reshape(-1)
            names.append(f"{prefix}.{k}")
            stats.append([float(g.sum()), float(g.abs().sum()), float(g.norm())])
            samples.append(p.grad.reshape(-1)[::SAMPLE_STRIDE].numpy())
    out["grad_names"] = np.asarray(names)
    out["grad_stats"] = np.asarray(stats)
    out["grad_samples"] = np.concatenate(samples)
    out["grad_sample_sizes"] = np.asarray([s.size for s in samples])
    out["no_grad_params"] = np.asarray(
        [f"get_adj_matrix.0.{k}" for k, p in mods[1].named_parameters() if p.grad is None])
    for i in TRAIN_DOCS:
        d = docs[i]
        keep = S.make_keep_masks(d.doc_id, d.n, layers, heads)
        y1, y2, dx, _, _ = run_reference(mods, d, keep)
        out[f"train{i}_y1"], out[f"train{i}_y2"], out[f"train{i}_dx0"] = y1.numpy(), y2.numpy(), dx.numpy()
    np.savez_compressed(os.path.join(HERE, f"graph_blocks_{variant}.npz"), **out)
    print(variant, "->", {k: getattr(v, "shape", None) for k, v in out.items() if k in ("y1", "grad_samples")})


def pool_and_pairs():
    out = {}
    g = torch.Generator().manual_seed(5)
    dis = torch.randn(21, 20, generator=g)
    ner = torch.randn(7, 20, generator=g)
    ner[0] = 0
    out["dis_table"], out["ner_table"] = dis.numpy(), ner.numpy()
    x0s, rps, rows_h, rows_t, sums = [], [], [], [], []
    for d in S.make_batch():
        # literal C:169-176 + C:223
        node_pos = np.zeros((d.n, d.L))
        for node in range(d.n):
            for position in d.spans[node]:
                node_pos[node, position[0]:position[1]] = 1.0 / (position[1] - position[0])
            node_pos[node, :] *= 1.0 / len(d.spans[node])
        node_pos = torch.FloatTensor(node_pos[:, :512])
        # literal G:297-298
        context_output = d.ctx.unsqueeze(0)
        node_feat = node_pos.unsqueeze(2).expand(-1, -1, 128) * context_output.expand(d.n, -1, -1)
        node_feat = node_feat.sum(dim=1)
        x0s.append(node_feat)
        # literal C:106-116, 207-217
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
        rps.append(nrp.reshape(-1))
        # literal G:344-352 with node_feats = cat[x0, x0, tanh(x0)]
        feats = torch.cat([node_feat, node_feat, torch.tanh(node_feat)], 1)
        fwt = torch.cat([feats, torch.nn.functional.embedding(d.node_type, ner)], 1)
        ph = torch.cat([fwt.unsqueeze(0).expand(d.n, -1, -1), torch.nn.functional.embedding(10 + nrp, dis)], -1)
        pt = torch.cat([fwt.unsqueeze(1).expand(-1, d.n, -1), torch.nn.functional.embedding(10 - nrp, dis)], -1)
        ph, pt = ph.reshape(d.n * d.n, -1), pt.reshape(d.n * d.n, -1)
        pick = torch.arange(0, d.n * d.n, 7)
        rows_h.append(ph[pick]), rows_t.append(pt[pick])
        sums.append([float(ph.double().sum()), float(pt.double().sum())])
    out["x0"] = torch.cat(x0s).numpy()
    out["rel_pos"] = torch.cat(rps).numpy()
    out["pair_h_rows"] = torch.cat(rows_h).numpy()
    out["pair_t_rows"] = torch.cat(rows_t).numpy()
    out["pair_checksums"] = np.asarray(sums)
    np.savez_compressed(os.path.join(HERE, "pool_pairs.npz"), **out)
    print("pool_pairs ->", out["x0"].shape, out["pair_h_rows"].shape)


def main():
    if not R.available():
        raise SystemExit("reference tree not present")
    torch.set_num_threads(1)
    graph_blocks("glove", 2, 8)
    graph_blocks("bert", 4, 4)
    pool_and_pairs()


if __name__ == "__main__":
    main()
