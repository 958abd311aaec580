"""CUDA-graph replay of one static-shape pass of the hot path.

The path is ~70 kernels per (forward + backward) pass, each a few tens of microseconds at DocRED
sizes, so the host (Python autograd + ctypes) can become the pacing item.  When the ragged batch
layout is fixed (same ``RaggedBatch``, same tensors) the whole pass -- every kernel of this
library, the side-stream overlap and torch's small glue ops -- is captured once and replayed with
one ``cudaGraphLaunch``.  Inputs are *static tensors*: refresh them with ``copy_`` before
``replay()``; outputs and gradients are read from the same static tensors afterwards.

Nothing here changes numerics: a replay issues exactly the kernels the eager call issued.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedPass:
    """Capture ``fn()`` (which must only touch static CUDA tensors) and replay it.

    fn is run ``warmup`` times on a side stream first (lazy one-time work such as
    cudaFuncSetAttribute, workspace growth and pointer-table uploads must not happen under
    capture), then once more under ``torch.cuda.graph``.  The value returned by the captured call
    is kept in ``self.out``.
    """

    def __init__(self, fn: Callable[[], object], device: torch.device, warmup: int = 3):
        self.device = device
        cur = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()
        torch.cuda.synchronize(device)

    def replay(self):
        self.graph.replay()
        return self.out
