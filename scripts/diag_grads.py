"""GPU diagnostic: per-document parameter-gradient error of the batched path vs the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import oracle_blocks, maxdiff
from gpu_common import device_blocks, run_blocks
from gcgcn_b200 import synthetic as S

gb, state = device_blocks(2, 8)
tot_o, tot_g = {}, {}
for i in range(24):
    d = S.make_doc(i)
    res = run_blocks(gb, [d])
    r = oracle_blocks(d, state, 2, 8)
    worst = ("", 0.0, 0.0)
    for k, v in res["dparams"].items():
        if v is None:
            continue
        o = r["dparams"][k]
        tot_o[k] = tot_o.get(k, 0) + o
        tot_g[k] = tot_g.get(k, 0) + v.cpu()
        e = maxdiff(v, o)
        if e > worst[1]:
            worst = (k, e, float(o.abs().max()))
    print(f"doc{i:02d} n={d.n:2d} worst {worst[0]:55s} err={worst[1]:.2e} |ref|max={worst[2]:.2e} "
          f"a0max={float(res['a0'].max()):.4f}")
print("summed:")
for k in tot_o:
    print(f"  {k:60s} err={maxdiff(tot_g[k], tot_o[k]):.2e} |ref|max={float(tot_o[k].abs().max()):.2e}")
