"""Per-kernel device times of one graph-block step (fwd+bwd) at the entity-sweep points of BASELINE.json configs[3].
python scripts/profile_sweep.py [n] [heads]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcgcn_b200 import _lib  # noqa: E402
from gcgcn_b200.batch import RaggedBatch  # noqa: E402
from gcgcn_b200.modules import GraphBlocks  # noqa: E402


def point(n, h, edge_gb=4.0, overlap=True):
    dev = torch.device("cuda:0")
    ndoc = max(1, int(edge_gb * 1e9 // (n * n * 512)))
    torch.manual_seed(0)
    gb = GraphBlocks(2, h, overlap=overlap).to(dev).eval()
    bt = RaggedBatch(np.full(ndoc, n, dtype=np.int64), dev)
    x0 = torch.tanh(torch.randn(bt.total_nodes, 128, device=dev)).requires_grad_(True)
    e0 = torch.randn(bt.total_pairs, 128, device=dev).requires_grad_(True)
    e1 = torch.randn(bt.total_pairs, 128, device=dev).requires_grad_(True)
    dy1 = torch.randn(bt.total_nodes, 128, device=dev)
    dy2 = torch.randn(bt.total_nodes, 128, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        if rep == 2:
            _lib.timing_begin(st)
        x0.grad = e0.grad = e1.grad = None
        out = gb(x0, e0, e1, bt, with_node_feats=False)
        torch.autograd.backward([out["y1"], out["y2"]], [dy1, dy2])
    t = _lib.timing_end(st)
    tot = sum(v[1] for v in t.values())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for rep in range(5):
        x0.grad = e0.grad = e1.grad = None
        out = gb(x0, e0, e1, bt, with_node_feats=False)
        torch.autograd.backward([out["y1"], out["y2"]], [dy1, dy2])
    ev1.record()
    torch.cuda.synchronize()
    print(f"overlap={overlap} step {ev0.elapsed_time(ev1) / 5:.3f} ms")
    print(f"n={n} heads={h} docs={ndoc} rows={bt.total_nodes} pairs={bt.total_pairs}: gcgcn kernels {tot:.3f} ms")
    for k, v in sorted(t.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {k:34s} x{v[0]:3d} {v[1]:8.3f} ms  {100 * v[1] / tot:5.1f}%")
    del x0, e0, e1, gb
    torch.cuda.empty_cache()


if __name__ == "__main__":
    if len(sys.argv) >= 3:
        point(int(sys.argv[1]), int(sys.argv[2]), overlap=True)
        point(int(sys.argv[1]), int(sys.argv[2]), overlap=False)
    else:
        for n in (42, 128, 256):
            for h in (4, 8):
                point(n, h)
