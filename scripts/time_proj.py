"""Device time of the projection GEMMs over all node rows of a step: [rows, K] x [K, N] with a weight-like second
operand (y = F Wout^T, dx = dZ WnX^T: K = 1024, N = 128; Z = x WnX: K = 128, N = 1024)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcgcn_b200 import _lib, functional as F  # noqa: E402

dev = torch.device("cuda:0")
M = 119808
st = torch.cuda.current_stream().cuda_stream
for K, N, tb in ((1024, 128, True), (1024, 128, False), (128, 1024, False), (128, 128, False), (256, 128, True)):
    a = torch.randn(M, K, device=dev)
    b = torch.randn((N, K) if tb else (K, N), device=dev) * 0.1
    for rep in range(6):
        if rep == 3:
            _lib.timing_begin(st)
        out = F.gemm(a, b, False, tb)
    t = _lib.timing_end(st)
    tot = sum(v[1] for v in t.values()) / 3
    mb = (a.numel() + out.numel()) * 4 / 1e6
    fl = 2.0 * M * N * K * 3
    print(f"M={M} K={K} N={N} tb={tb}: {tot * 1e3:.1f} us  {mb / tot / 1e3:.2f} TB/s of A+C bytes, {fl / tot / 1e9:.0f} executed TF32 TFLOP/s  "
          + " ".join(f"{k}:{v[1] / 3 * 1e3:.1f}us" for k, v in t.items()))
