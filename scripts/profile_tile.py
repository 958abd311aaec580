"""Bench-size forward(+backward) passes of the graph blocks with the packed-tile MAGGC kernels on -- the target of the
ncu captures of csrc/gcn_tile.cu.  python scripts/profile_tile.py [docs] [--bwd]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcgcn_b200 import _lib, synthetic as S  # noqa: E402
from gcgcn_b200.batch import RaggedBatch  # noqa: E402
from gcgcn_b200.modules import GraphBlocks  # noqa: E402


def main():
    docs = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 6144
    bwd = "--bwd" in sys.argv
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    bt = RaggedBatch(S.shard_doc_sizes(docs), dev)
    gb = GraphBlocks(2, 8).to(dev).eval()
    x0 = torch.randn(bt.total_nodes, 128, device=dev, requires_grad=bwd)
    e0 = torch.randn(bt.total_pairs, 128, device=dev, requires_grad=bwd)
    e1 = torch.randn(bt.total_pairs, 128, device=dev, requires_grad=bwd)
    _lib.set_tile_blocks(True)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        if rep == 2:
            _lib.timing_begin(st)
        with torch.set_grad_enabled(bwd):
            out = gb(x0, e0, e1, bt)
        if bwd:
            (out["y1"].sum() + out["y2"].sum()).backward()
    t = _lib.timing_end(st)
    torch.cuda.synchronize()
    print({k: round(v[1], 3) for k, v in t.items() if "block" in k or "tile" in k}, f"tiles {bt.num_tiles}")
    if "--all" in sys.argv:
        tot = sum(v[1] for v in t.values())
        print(f"all gcgcn kernels of the step: {tot:.3f} ms")
        for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
            print(f"   {k:34s} x{v[0]:3d} {v[1]:8.3f} ms  {100 * v[1] / tot:5.1f}%")


if __name__ == "__main__":
    main()
