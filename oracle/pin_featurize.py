"""Pin oracle/featurize_oracle.py against the reference's own `Config.from_list_to_tensor` (config/Config.py:162-233)
and (with --write) regenerate tests/golden/featurize.npz from the reference.

config/Config.py cannot be imported in this container (matplotlib / torch_geometric / pytorch_pretrained_bert are
absent), so the function's source is cut out of the file with `ast` and compiled as-is against a stand-in `self`
that carries only the attributes the function reads (max_length, max_num, dis2idx, dis_plus; C:67, 78, 106-118).
Run from the repo root:  python oracle/pin_featurize.py [--write]
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("GCGCN_REFERENCE", "/root/reference")


def reference_function():
    src = open(os.path.join(REF, "config", "Config.py"), encoding="utf-8").read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "from_list_to_tensor":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np, "torch": torch}
            exec(compile(mod, "Config.py:from_list_to_tensor", "exec"), ns)
            return ns["from_list_to_tensor"]
    raise RuntimeError("from_list_to_tensor not found in the reference")


def reference_self(max_length=512, max_num=5):
    from oracle.featurize_oracle import make_dis2idx
    return types.SimpleNamespace(max_length=max_length, max_num=max_num, dis2idx=make_dis2idx(), dis_plus=10)


def compare(item, max_length, max_num, fn):
    from oracle.featurize_oracle import from_list_to_tensor
    ref = fn(reference_self(max_length, max_num), item)
    got = from_list_to_tensor(item, max_length, max_num)
    for k, v in got.items():
        r = ref[k].numpy()
        assert r.dtype == v.dtype and r.shape == v.shape, (k, r.dtype, v.dtype, r.shape, v.shape)
        assert np.array_equal(r, v), k
    return ref


def main():
    from gcgcn_b200 import synthetic
    fn = reference_function()
    cases = [(0, 512, 5), (1, 512, 5), (2, 40, 2), (3, 512, 5), (4, 64, 3), (5, 512, 5)]
    for seed, max_length, max_num in cases:
        compare(synthetic.make_record(seed), max_length, max_num, fn)
    print(f"oracle/featurize_oracle.py == reference from_list_to_tensor on {len(cases)} records (all 7 outputs, exact)")
    if "--write" in sys.argv:
        item = synthetic.make_record(7, n=6, L=60, S=3)
        ref = fn(reference_self(48, 2), item)
        out = {k: v.numpy() for k, v in ref.items() if torch.is_tensor(v) and k in
               ("adj_matrix", "sen_matrix", "pos_matrix_h", "pos_matrix_t", "node_pos", "node_type", "node_relative_pos")}
        path = os.path.join(ROOT, "tests", "golden", "featurize.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
