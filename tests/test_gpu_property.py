"""-m gpu: property test of the graph blocks over ragged batches (n in [1, 256], mixed sizes in one
batch, both variants).  hypothesis draws the entity counts and the document seeds; outputs must match the oracle at
1e-4 for EVERY drawn batch, gradients for every drawn batch whose documents are well conditioned -- relu' jumps at 0, so
a document with a pre-activation inside float rounding of 0 (helpers.relu_margin, an oracle-side probe) may
legitimately flip one unit; such documents keep their forward check and are left out of the gradient check, and the
test counts how often that happens (it must stay rare)."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from helpers import FP32_TOL, assert_close, oracle_blocks, relu_margin
from gpu_common import device_blocks, run_blocks
from gcgcn_b200 import synthetic as S

pytestmark = pytest.mark.gpu

SIZES = st.one_of(st.integers(1, 64), st.sampled_from([1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65, 66, 96, 97, 127,
                                                       128, 129, 200, 255, 256]))
STATS = {"docs": 0, "ill": 0}


@settings(max_examples=14, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(sizes=st.lists(SIZES, min_size=1, max_size=5), seed=st.integers(0, 10_000), variant=st.sampled_from([(2, 8), (4, 4)]))
def test_ragged_batches_match_the_oracle(sizes, seed, variant):
    layers, heads = variant
    if sum(n * n for n in sizes) > 140_000:                # keep the CPU oracle (as-written n^2 linears) in seconds
        sizes = sorted(sizes)[:2]
    gb, state = device_blocks(layers, heads)
    docs = [S.make_doc(seed * 7 + i, n=n, L=16) for i, n in enumerate(sizes)]
    res = run_blocks(gb, docs)
    bt = res["bt"]
    total, all_well = {}, True
    for b, d in enumerate(docs):
        r = oracle_blocks(d, state, layers, heads)
        for k in ("y1", "y2"):
            assert_close(bt.split_nodes(res[k])[b], r[k], FP32_TOL, f"n={d.n} {k}")
        well = relu_margin(d, state, layers, heads) > 2e-6
        STATS["docs"] += 1
        STATS["ill"] += 0 if well else 1
        all_well &= well
        if well:
            assert_close(bt.split_nodes(res["dx0"])[b], r["dx0"], FP32_TOL, f"n={d.n} dx0")
            for k in ("de0", "de1"):
                assert_close(bt.split_pairs(res[k])[b], r[k], FP32_TOL, f"n={d.n} {k}")
        for k, v in r["dparams"].items():
            if v is not None:
                total[k] = total.get(k, 0) + v
    if all_well:
        for k, v in res["dparams"].items():
            if v is not None:
                assert_close(v, total[k], max(5, len(docs)) * FP32_TOL, f"sizes={sizes} d{k}")
        assert res["dparams"]["get_adj_matrix.0.linears_k.0.weight"] is None      # quirk 3 survives batching


def test_most_documents_are_well_conditioned():
    """A document of n entities has about 1150 n relu pre-activations (GloVe variant); with a density of ~0.5 per unit
    around 0, one of them lies within 2e-6 of the kink with probability ~1 - exp(-0.0023 n): a few percent at DocRED
    sizes, about half of the 256-entity documents.  The gradient check above therefore covers most, not all, draws."""
    assert STATS["docs"] >= 14
    assert STATS["ill"] <= STATS["docs"] // 2, STATS
