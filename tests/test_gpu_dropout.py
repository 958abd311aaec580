"""-m gpu: train-mode dropout of the block kernels.  The kernels regenerate keep factors from (seed, stream,
index); gcgcn_dropout_mask materialises the same streams, which are then injected into the keep-mask-in route
(the route test_gpu_blocks pins to the oracle with injected masks): both routes must agree element for element."""
import pytest
import torch

from helpers import FP32_TOL, VARIANTS, assert_close, upstream
from gpu_common import DEV, cat_inputs, device_blocks
from gcgcn_b200 import synthetic as S
from gcgcn_b200.batch import RaggedBatch
from gcgcn_b200.functional import dropout_mask

pytestmark = pytest.mark.gpu


def test_dropout_streams_are_deterministic_unbiased_and_distinct():
    n = 1 << 20
    for p in (0.1, 0.2, 0.5):
        m = dropout_mask(1234, 3, p, n, DEV)
        assert torch.equal(m, dropout_mask(1234, 3, p, n, DEV))
        vals = torch.unique(m)
        assert vals.numel() == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - 1 / (1 - p)) < 1e-6
        drop_rate = float((m == 0).float().mean())
        assert abs(drop_rate - p) < 4 * (p * (1 - p) / n) ** 0.5 + 1e-4          # 4 sigma
        assert abs(float(m.mean()) - 1.0) < 5e-3
    a, b, c = dropout_mask(1, 1, 0.2, n, DEV), dropout_mask(1, 2, 0.2, n, DEV), dropout_mask(2, 1, 0.2, n, DEV)
    for u, v in ((a, b), (a, c)):                                                 # independent streams / seeds
        both = float(((u == 0) & (v == 0)).float().mean())
        assert abs(both - 0.04) < 2e-3
    assert torch.equal(dropout_mask(7, 4, 0.0, 1000, DEV), torch.ones(1000, device=DEV))


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_in_kernel_dropout_equals_the_keep_mask_route(variant):
    layers, heads = VARIANTS[variant]
    g = 128 // layers
    docs = S.make_batch()
    bt = RaggedBatch([d.n for d in docs], DEV)
    ups = [upstream(d.doc_id, (d.n, 128), (d.n, 128)) for d in docs]
    dy1 = torch.cat([u[0] for u in ups]).to(DEV)
    dy2 = torch.cat([u[1] for u in ups]).to(DEV)
    M, P = bt.total_nodes, bt.total_pairs

    def run(gb, inject):
        x0, e0, e1, _ = cat_inputs(docs)
        gb.zero_grad()
        if inject is not None:
            inject(gb)
        out = gb(x0, e0, e1, bt)
        torch.autograd.backward([out["y1"], out["y2"]], [dy1, dy2])
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().clone() for k, p in gb.named_parameters() if p.grad is not None}
        return {"y1": out["y1"].detach(), "y2": out["y2"].detach(), "dx0": x0.grad, "de0": e0.grad, "de1": e1.grad}, grads

    gb, _ = device_blocks(layers, heads)
    gb.train()
    gb.dropout.p = 0.0                       # the block-output dropout (G:341) stays a torch op on both routes
    torch.manual_seed(99)
    fused, fgrads = run(gb, None)
    dc, dm = gb.last_drop["caggc"], gb.last_drop["maggc"]
    assert dc is not None and dm is not None and dc[0] != dm[0]
    # the eval result must differ (dropout really happened) ...
    gb.eval()
    plain, _ = run(gb, None)
    assert float((plain["y2"] - fused["y2"]).abs().max()) > 1e-2
    # ... and the same masks through the keep-mask-in route must reproduce the fused result
    gb.train()
    gb.fused = False

    def inject(m):
        gat, mha = m.get_weighted_adj_matrix, m.get_adj_matrix[0]
        cag, mag = m.graphcnn
        gat.inject_keep([dropout_mask(dc[0], 1, dc[1], P, DEV)])
        kc = dropout_mask(dc[0], 2, dc[2], M * 128, DEV).view(M, 128)
        cag.inject_keep([kc[:, l * g:(l + 1) * g].contiguous() for l in range(layers)])
        ka = dropout_mask(dm[0], 3, dm[1], heads * P, DEV).view(heads, P)
        mha.inject_keep([ka[h].contiguous() for h in range(heads)])
        km = dropout_mask(dm[0], 4, dm[2], M * heads * 128, DEV).view(M, heads * 128)
        mag.inject_keep([km[:, k * g:(k + 1) * g].contiguous() for k in range(heads * layers)])

    masked, mgrads = run(gb, inject)
    for k in ("y1", "y2"):
        assert_close(fused[k], masked[k], 2e-5, k)
    for k in ("dx0", "de0", "de1"):
        assert_close(fused[k], masked[k], FP32_TOL, k)
    for k, v in mgrads.items():
        assert_close(fgrads[k], v, 5 * FP32_TOL, "d" + k)


def test_train_mode_is_reproducible_under_manual_seed():
    gb, _ = device_blocks(2, 8)
    gb.train()
    docs = S.make_batch((9, 21, 40))
    bt = RaggedBatch([d.n for d in docs], DEV)
    outs = []
    for _ in range(2):
        torch.manual_seed(5)
        x0, e0, e1, _ = cat_inputs(docs, requires_grad=False)
        with torch.no_grad():
            outs.append(gb(x0, e0, e1, bt)["y2"].clone())
    assert torch.equal(outs[0], outs[1])
    torch.manual_seed(6)
    with torch.no_grad():
        other = gb(x0, e0, e1, bt)["y2"]
    assert not torch.equal(outs[0], other)
