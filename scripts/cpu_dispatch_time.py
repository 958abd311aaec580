import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from gcgcn_b200 import synthetic
from gcgcn_b200.batch import RaggedBatch
from gcgcn_b200.modules import GraphBlocks
dev = torch.device("cuda:0")
torch.manual_seed(0)
gb = GraphBlocks(2, 8).to(dev).eval()
for tile in (512, 256):
    bt = RaggedBatch(synthetic.shard_doc_sizes(12 * tile), dev)
    g = torch.Generator(device=dev).manual_seed(1)
    x0 = torch.randn(bt.total_nodes, 128, device=dev, generator=g).requires_grad_(True)
    e0 = torch.randn(bt.total_pairs, 128, device=dev, generator=g).requires_grad_(True)
    e1 = torch.randn(bt.total_pairs, 128, device=dev, generator=g).requires_grad_(True)
    dy = torch.randn(bt.total_nodes, 128, device=dev, generator=g)
    params = [p for n, p in gb.named_parameters() if "linears_k" not in n]
    def step():
        x0.grad = e0.grad = e1.grad = None
        for p in params: p.grad = None
        out = gb(x0, e0, e1, bt)
        torch.autograd.backward([out["y1"], out["y2"]], [dy, dy])
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"tile {tile}: cpu dispatch {1e3*(t1-t0)/10:.2f} ms/step, wall {1e3*(t2-t0)/10:.2f} ms/step")
