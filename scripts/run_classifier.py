"""Classifier forward (+ backward with --bwd) on P pairs, for profiling the bilinear kernels.
    python scripts/run_classifier.py [pairs] [iters] [--bwd]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcgcn_b200 import _lib                                   # noqa: E402
from gcgcn_b200.classifier import relation_logits            # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
P = int(args[0]) if args else 99312
iters = int(args[1]) if len(args) > 1 else 5
bwd = "--bwd" in sys.argv
dev = torch.device("cuda:0")
torch.manual_seed(0)
bili, cls = torch.nn.Bilinear(128, 128, 97).to(dev), torch.nn.Linear(256, 97).to(dev)
h = torch.tanh(torch.randn(P, 128, device=dev)).requires_grad_(bwd)
t = torch.tanh(torch.randn(P, 128, device=dev)).requires_grad_(bwd)
up = torch.randn(P, 97, device=dev)


def step():
    z = relation_logits(h, t, bili, cls)
    if bwd:
        torch.autograd.grad(z, [h, t] + list(bili.parameters()) + list(cls.parameters()), up)


for _ in range(2):
    step()
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(iters):
    step()
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / iters
flop = 2.0 * P * (128 * 128 * 97 + 256 * 97) * (3 if bwd else 1)
print(f"{P} pairs, {'fwd+bwd' if bwd else 'fwd'}: {ms:.3f} ms, {flop / ms / 1e9:.1f} fp32-equivalent TFLOP/s "
      f"({3 * flop / ms / 1e9:.1f} executed TF32 TFLOP/s)")
_lib.timing_begin(st)
for _ in range(iters):
    step()
kern = _lib.timing_end(st)
for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:40s} {v[0] / iters:6.1f} launches  {v[1] / iters:9.4f} ms" + (f"  {3 * v[2] / v[1] / 1e9:8.1f} TF32 TFLOP/s" if v[2] else ""))
