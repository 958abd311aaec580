"""GPU micro-benchmark + accuracy check of gcgcn_gemm on the hot path's shapes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gcgcn_b200 import functional as F, _lib

dev = "cuda:0"
M = 124416
shapes = [  # (name, ta, tb, M, N, K)
    ("Zx  nn M x1024x128", 0, 0, M, 1024, 128),
    ("out nt M x128x1024", 0, 1, M, 128, 1024),
    ("dF  nn M x1024x128", 0, 0, M, 1024, 128),
    ("dx  nt M x128x1024", 0, 1, M, 128, 1024),
    ("dW  tn 128x1024xM ", 1, 0, 128, 1024, M),
    ("q   nt M x128x128 ", 0, 1, M, 128, 128),
    ("dWi tn 64x64xM    ", 1, 0, 64, 64, M),
]
g = torch.Generator(device=dev).manual_seed(0)
pick = sys.argv[1:] 
for name, ta, tb, m, n, k in shapes:
    if pick and not any(p in name for p in pick):
        continue
    a = torch.randn((k, m) if ta else (m, k), device=dev, generator=g)
    b = torch.randn((n, k) if tb else (k, n), device=dev, generator=g)
    out = F.gemm(a, b, ta, tb)
    torch.cuda.synchronize()
    # accuracy on a sample of rows vs fp64
    rows = torch.randint(0, m, (64,), device=dev)
    aa = (a.t() if ta else a)[rows].double()
    bb = (b.t() if tb else b).double()
    ref = aa @ bb
    err = float((out[rows].double() - ref).abs().max())
    rel = err / float(ref.abs().max())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        F.gemm(a, b, ta, tb, out=out)
    ev0.record()
    iters = 10
    for _ in range(iters):
        F.gemm(a, b, ta, tb, out=out)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    print(f"{name}  {ms*1e3:9.1f} us  {2.0*m*n*k/ms/1e9:8.1f} TFLOP/s(fp32-equiv)  max|err|={err:.2e} rel={rel:.1e}", flush=True)
