"""Load the reference's model files from /root/reference *by path* (test infrastructure).

Only usable in the build container: /root/reference does not exist on the GPU box,
so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  It is
used by ``oracle/pin_against_reference.py`` and ``tests/golden/make_golden.py`` to pin
the oracle and to generate the committed golden vectors, and by the CPU tests that
are skipped when the reference tree is absent.

Two non-invasive shims:
  1. the model files import ``pytorch_pretrained_bert`` at import time (G:13, B:13),
     which is not installed -> a stub module is put in ``sys.modules`` first;
  2. ``import models`` fails on a syntax error (models/__init__.py:7), so the two
     model files are loaded with ``importlib`` from their paths.
Nothing under /root/reference is modified or copied.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GCGCN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "GCGCN_glove.py"))


def _install_bert_stub():
    if "pytorch_pretrained_bert" in sys.modules:
        return
    import torch

    stub = types.ModuleType("pytorch_pretrained_bert")

    class BertModel(torch.nn.Module):  # stands in for B:228; the encoder is out of scope
        def __init__(self, width=768):
            super().__init__()
            self.width = width

        @classmethod
        def from_pretrained(cls, path):
            return cls()

        def forward(self, ids, output_all_encoded_layers=False):
            g = torch.Generator().manual_seed(int(ids.sum()) % (1 << 31))
            return torch.randn(ids.size(0), ids.size(1), self.width, generator=g), None

    stub.BertModel = BertModel
    sys.modules["pytorch_pretrained_bert"] = stub


def _load(fname: str, modname: str):
    if modname in sys.modules:
        return sys.modules[modname]
    _install_bert_stub()
    path = os.path.join(REFERENCE_ROOT, "models", fname)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    sys.modules[modname] = mod
    return mod


def glove_module():
    """models/GCGCN_glove.py as a module object (graph classes G:18-168)."""
    return _load("GCGCN_glove.py", "_gcgcn_ref_glove")


def bert_module():
    """models/GraphCNN_multihead_bert_gate_cls.py as a module object (B:18-172)."""
    return _load("GraphCNN_multihead_bert_gate_cls.py", "_gcgcn_ref_bert")


def build_graph_modules(layers: int, heads: int, seed: int = 0, hidden: int = 128, variant="glove"):
    """Instantiate the four hot-path modules in the order the reference model does
    (G:254-262), under ``torch.manual_seed(seed)``."""
    import torch

    m = glove_module() if variant == "glove" else bert_module()
    torch.manual_seed(seed)
    gat = m.GATAttention(hidden, hidden)
    mha = m.MultiHeadAttention(heads, hidden)
    cag = m.GraphConvolution(layers, hidden, hidden)
    mag = m.MultiGraphConvolution(layers, heads, hidden, hidden)
    for mod in (gat, mha, cag, mag):
        mod.eval()
    return gat, mha, cag, mag
