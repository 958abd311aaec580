"""Per-kernel device time of the on-device pipeline pooling -> producer -> CAGGC -> producer -> MAGGC (fwd+bwd) that
bench.py's e2e leg runs, with the library's own per-launch events (gcgcn_timing_begin/end).
    python scripts/profile_pipeline.py [tiles] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcgcn_b200 import _lib, synthetic                       # noqa: E402
from gcgcn_b200.batch import PoolTable, RaggedBatch         # noqa: E402
from gcgcn_b200.edgefeat import EdgeFeatures, EdgeTables    # noqa: E402
from gcgcn_b200.modules import GraphBlocks, pool_nodes      # noqa: E402

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
docs = synthetic.make_batch()
wires = [synthetic.make_wire(d) for d in docs] * tiles
bt = RaggedBatch([d.n for d in docs] * tiles, dev)
pool = PoolTable.from_spans([d.spans for d in docs] * tiles, [d.L for d in docs] * tiles, device=dev)
edge = EdgeTables(wires, bt, dev)
rows_tab = PoolTable.concat(pool, edge.gather, dev)
torch.manual_seed(0)
gb = GraphBlocks(2, 8).to(dev).eval()
producer = EdgeFeatures().to(dev)
dis = torch.randn(21, 20, device=dev, requires_grad=True)
ctx0 = torch.tanh(torch.randn(pool.total_tokens, 128, device=dev))
dy1, dy2 = torch.randn(bt.total_nodes, 128, device=dev), torch.randn(bt.total_nodes, 128, device=dev)
params = list(gb.parameters()) + list(producer.parameters()) + [dis]


def step():
    for p in params:
        p.grad = None
    ctx = ctx0.detach().requires_grad_(True)
    rows = pool_nodes(ctx, rows_tab)
    x0, act = rows[:bt.total_nodes], rows[bt.total_nodes:]
    e0 = producer(0, None, x0, dis, edge, ctx_act=act)
    y1, _ = gb.hop0(x0, e0, bt)
    e1 = producer(1, None, y1, dis, edge, ctx_act=act)
    y2, _ = gb.hop1(y1, e1, bt)
    torch.autograd.backward([y1, y2], [dy1, dy2])


for _ in range(3):
    step()
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    step()
ev1.record()
torch.cuda.synchronize()
print(f"{bt.num_docs} documents, {edge.num_slots} active slots: {ev0.elapsed_time(ev1) / steps:.3f} ms per step")
_lib.timing_begin(st)
for _ in range(steps):
    step()
kern = _lib.timing_end(st)
tot = sum(v[1] for v in kern.values())
print(f"{'kernel':44s} {'launches/step':>13s} {'ms/step':>9s} {'share':>7s}")
for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} {v[0] / steps:13.1f} {v[1] / steps:9.4f} {v[1] / tot:7.3f}")
