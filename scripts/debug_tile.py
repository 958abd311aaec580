"""Stage-by-stage comparison of the saved tensors of gcgcn_mha_stack_fwd with the packed-tile path on and off."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcgcn_b200 import _lib
from gcgcn_b200.batch import RaggedBatch
from gcgcn_b200 import functional as Fn


class Ctx:
    def save_for_backward(self, *t): self.saved = t
    def mark_non_differentiable(self, *a): pass


def run(sizes, on):
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    bt = RaggedBatch(sizes, dev)
    M, H, L = bt.total_nodes, 8, 2
    x = torch.randn(M, 128, device=dev)
    ebar = torch.randn(M, 128, device=dev)
    Wq, bq = torch.randn(128, 128, device=dev) * 0.1, torch.randn(128, device=dev) * 0.1
    WnX, We = torch.randn(128, H * 128, device=dev) * 0.1, torch.randn(128, H * 128, device=dev) * 0.1
    Winner = torch.randn(H, L, 128, 64, device=dev) * 0.1
    Wout, bout = torch.randn(128, H * 128, device=dev) * 0.05, torch.randn(128, device=dev) * 0.1
    _lib.set_tile_blocks(on)
    ctx = Ctx()
    y, P = Fn.MhaStackFn.forward(ctx, x, ebar, Wq, bq, WnX, We, Winner, Wout, bout, bt, H, L)
    torch.cuda.synchronize()
    _lib.set_tile_blocks(False)
    names = "x ebar Wq WnX We Winner Wout q P Z G F".split()
    return dict(zip(names, ctx.saved)), y


for sizes in ([5], [19] * 12, [64, 31, 1]):
    a, ya = run(sizes, False)
    b, yb = run(sizes, True)
    M = a["Z"].shape[0]
    def d(u, v): return float((u - v).abs().max())
    Za, Zb = a["Z"].view(M, 8, 2, 64), b["Z"].view(M, 8, 2, 64)
    Ga, Gb = a["G"].view(M, 8, 2, 64), b["G"].view(M, 8, 2, 64)
    Fa, Fb = a["F"].view(M, 8, 2, 64), b["F"].view(M, 8, 2, 64)
    print(sizes[:4], "P", d(a["P"], b["P"]), "Z0", d(Za[:, :, 0], Zb[:, :, 0]), "g0", d(Ga[:, :, 0], Gb[:, :, 0]),
          "F0", d(Fa[:, :, 0], Fb[:, :, 0]), "Z1", d(Za[:, :, 1], Zb[:, :, 1]), "g1", d(Ga[:, :, 1], Gb[:, :, 1]),
          "F1", d(Fa[:, :, 1], Fb[:, :, 1]), "y", d(ya, yb))
    if os.environ.get("GCGCN_TILE_DEBUG"):
        # F of the tile run holds N_l = P Z_l; rebuild it from the saved P and Z of the reference run
        bt = RaggedBatch(sizes, torch.device("cuda:0"))
        n0 = sizes[0]
        for l in range(2):
            Pm = a["P"][0, : n0 * n0].view(n0, n0)
            Nref = Pm @ Za[:n0, 0, l]
            Nt = Fb[:n0, 0, l]
            print(f" N_{l} head0 doc0: max|ref| {float(Nref.abs().max()):.3f} max|diff| {d(Nref, Nt):.3e}")
            print("   ref row0:", [round(float(v), 3) for v in Nref[0, :10]])
            print("   tile row0:", [round(float(v), 3) for v in Nt[0, :10]])
            print("   tile row1:", [round(float(v), 3) for v in Nt[1, :10]])
            print("   Z row0:", [round(float(v), 3) for v in Za[0, 0, l, :10]])
    if os.environ.get("TILE_DEBUG"):
        print(" g0 ref row0 head0:", Ga[0, 0, 0, :8].tolist())
        print(" g0 tile row0 head0:", Gb[0, 0, 0, :8].tolist())
        diff = (Ga[:, :, 0] - Gb[:, :, 0]).abs()
        print(" g0 diff per row (head 0):", [round(float(v), 4) for v in diff[:, 0].max(dim=1).values[:12]])
        print(" g0 diff per col (head 0, row 0):", [round(float(v), 3) for v in diff[0, 0, :64]])
