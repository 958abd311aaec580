"""-m gpu: the packed-tile tcgen05 forward kernel of the MAGGC block (csrc/gcn_tile.cu) against the per-document
mma.sync kernels on the same inputs, and against the oracle."""
import pytest
import torch

from helpers import FP32_TOL, assert_close, oracle_blocks
from gpu_common import device_blocks, run_blocks
from gcgcn_b200 import _lib, synthetic as S
from gcgcn_b200.batch import TILE_ROWS, RaggedBatch

pytestmark = pytest.mark.gpu

CASES = {
    "one-small": [5],
    "two-large": [64, 64],
    "exact-tile": [32, 32, 32, 48, 48],
    "ragged": [1, 2, 3, 42, 1, 7, 33, 2, 64, 5],
    "many-tiles": [19] * 40,
    "straddling": [3, 17, 64, 63, 2, 61, 6],
}


@pytest.fixture
def tile_switch():
    before = _lib.set_tile_blocks(True)
    yield
    _lib.set_tile_blocks(before)


@pytest.mark.parametrize("case", list(CASES))
def test_tile_kernel_matches_per_document_kernels(case, tile_switch):
    sizes = CASES[case]
    gb, _ = device_blocks(2, 8)
    docs = [S.make_doc(900 + i, n=n, L=32) for i, n in enumerate(sizes)]
    _lib.set_tile_blocks(False)
    ref = run_blocks(gb, docs)
    _lib.set_tile_blocks(True)
    st = torch.cuda.current_stream().cuda_stream
    _lib.timing_begin(st)
    out = run_blocks(gb, docs)
    names = _lib.timing_end(st)
    assert "tile_fwd" in names, f"the packed-tile kernel must be what ran: {sorted(names)}"
    bt = ref["bt"]
    assert bt.num_tiles >= 1 and int(bt.tile_doc_host[-1]) == len(sizes)
    # forward: continuous in the inputs, so the two paths agree to fp32 rounding everywhere
    assert_close(out["y1"], ref["y1"], 0.0, "y1 (the CAGGC hop does not change)")
    assert_close(out["y2"], ref["y2"], 2e-5, "y2")
    assert_close(out["a1"], ref["a1"], 1e-6, "attention maps")
    # backward goes through relu'(g): a pre-activation within rounding of zero may fall on the other side of the kink
    # in one of the two paths (about one element in 1e6), which changes that document's gradients by O(1e-3).  Require
    # all documents but at most one to agree.
    off = 0
    for k in ("dx0", "de0", "de1"):
        parts_o = bt.split_nodes(out[k]) if k == "dx0" else bt.split_pairs(out[k])
        parts_r = bt.split_nodes(ref[k]) if k == "dx0" else bt.split_pairs(ref[k])
        bad = [b for b, (o, r) in enumerate(zip(parts_o, parts_r)) if float((o - r).abs().max()) > 2e-5]
        assert len(bad) <= 1, f"{k}: documents {bad} differ between the two forward kernels"
        off = max(off, len(bad))
    if off == 0:
        for k, v in ref["dparams"].items():
            if v is not None:
                scale = max(float(v.abs().max()), 1.0)
                assert float((out["dparams"][k] - v).abs().max()) <= 2e-5 * scale, "d" + k


@pytest.mark.parametrize("heads", [8, 4], ids=["8x16", "4x32"])
def test_tile_kernel_against_the_oracle(heads, tile_switch):
    gb, state = device_blocks(2, heads)
    sizes = [7, 11, 42, 3, 29, 5, 64, 1]
    docs = [S.make_doc(300 + i, n=n, L=32) for i, n in enumerate(sizes)]
    st = torch.cuda.current_stream().cuda_stream
    _lib.timing_begin(st)
    res = run_blocks(gb, docs)
    assert "tile_fwd" in _lib.timing_end(st), "the packed-tile kernel must be what ran"
    bt = res["bt"]
    for b, d in enumerate(docs):
        r = oracle_blocks(d, state, 2, heads)
        assert_close(bt.split_nodes(res["y2"])[b], r["y2"], FP32_TOL, f"doc{b} y2")
        for h in range(heads):
            assert_close(bt.split_pairs(res["a1"][h])[b], r["a1"][h], 1e-5, f"doc{b} a1[{h}]")
        assert_close(bt.split_nodes(res["dx0"])[b], r["dx0"], FP32_TOL, f"doc{b} dx0")
        assert_close(bt.split_pairs(res["de1"])[b], r["de1"], FP32_TOL, f"doc{b} de1")


def test_documents_too_large_for_a_tile_take_the_per_document_path(tile_switch):
    """A document of more than TILE_ROWS entities switches the packing hint off for the whole batch."""
    bt = RaggedBatch([TILE_ROWS + 1, 4], "cuda:0")
    assert bt.num_tiles == 0
    gb, state = device_blocks(2, 8)
    docs = [S.make_doc(500, n=TILE_ROWS + 1, L=32), S.make_doc(501, n=4, L=32)]
    res = run_blocks(gb, docs)
    r = oracle_blocks(docs[1], state, 2, 8)
    assert_close(res["bt"].split_nodes(res["y2"])[1], r["y2"], FP32_TOL, "y2 of the small document")
