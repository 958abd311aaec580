"""-m gpu: the block-level C-ABI entry points (gcgcn_caggc_fwd/bwd, gcgcn_maggc_fwd/bwd -- SURVEY.md 8b
minimum set) called the way a C caller would, with raw pointers and caller-owned arenas, against the
per-module route (which test_gpu_blocks pins to the oracle and the reference's golden vectors)."""
import pytest
import torch

from helpers import FP32_TOL, VARIANTS, assert_close, upstream
from gpu_common import DEV, cat_inputs, device_blocks
from gcgcn_b200 import _lib, synthetic as S
from gcgcn_b200.batch import RaggedBatch
from gcgcn_b200.functional import _p, _stream, workspace
from gcgcn_b200.modules import _pack_stack

pytestmark = pytest.mark.gpu


def _arena(nbytes):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=DEV)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_block_composites_match_the_module_route(variant):
    layers, heads = VARIANTS[variant]
    gb, _ = device_blocks(layers, heads)
    gb.fused = False
    docs = S.make_batch()
    bt = RaggedBatch([d.n for d in docs], DEV)
    x0, e0, e1, _ = cat_inputs(docs)
    ups = [upstream(d.doc_id, (d.n, 128), (d.n, 128)) for d in docs]
    dy1 = torch.cat([u[0] for u in ups]).to(DEV)
    dy2 = torch.cat([u[1] for u in ups]).to(DEV)
    out = gb(x0, e0, e1, bt)
    torch.autograd.backward([out["y1"], out["y2"]], [dy1, dy2])
    ref = {"y1": out["y1"].detach(), "y2": out["y2"].detach(), "dx0": x0.grad, "de0": e0.grad, "de1": e1.grad}

    lib = _lib.load()
    M, P = bt.total_nodes, bt.total_pairs
    st = _stream(DEV)
    ws = workspace(DEV, lib.gcgcn_workspace_bytes(M, P, heads))
    gat, mha = gb.get_weighted_adj_matrix, gb.get_adj_matrix[0]
    cag, mag = gb.graphcnn
    with torch.no_grad():
        u, v, c = (t.contiguous() for t in gat.collapse())
        c = c.reshape(1)
        cw = [t.contiguous() if t is not None else None for t in _pack_stack(cag.graphconv, 1, layers, cag._g)]
        mw = [t.contiguous() if t is not None else None for t in _pack_stack(mag.graphconv, heads, layers, mag._g)]
        wq = torch.cat([l.weight for l in mha.linears_q], 0).contiguous()
        bq = torch.cat([l.bias for l in mha.linears_q], 0).contiguous()
        cwo, cbo = cag.linear_layer.weight.contiguous(), cag.linear_layer.bias.contiguous()
        mwo, mbo = mag.linear_layer.weight.contiguous(), mag.linear_layer.bias.contiguous()
        xd, e0d, e1d = x0.detach(), e0.detach(), e1.detach()

        saved_c = _arena(lib.gcgcn_block_saved_bytes(M, P, 1))
        saved_m = _arena(lib.gcgcn_block_saved_bytes(M, P, heads))
        y1, y2 = torch.empty(M, 128, device=DEV), torch.empty(M, 128, device=DEV)
        _lib.call("gcgcn_caggc_fwd", bt.ref, layers, _p(xd), _p(e0d), _lib.F32, _p(u), _p(v), _p(c), _p(cw[0]),
                  _p(cw[1]), _p(cw[2]), _p(cwo), _p(cbo), _p(y1), _p(saved_c), None, ws.data_ptr(), ws.numel(), st)
        _lib.call("gcgcn_maggc_fwd", bt.ref, layers, heads, _p(y1), _p(e1d), _lib.F32, _p(wq), _p(bq), _p(mw[0]),
                  _p(mw[1]), _p(mw[2]), _p(mwo), _p(mbo), _p(y2), _p(saved_m), None, ws.data_ptr(), ws.numel(), st)
        assert_close(y1, ref["y1"], 1e-5, "caggc_fwd y1")
        assert_close(y2, ref["y2"], 1e-5, "maggc_fwd y2")

        def like(t):
            return None if t is None else torch.empty_like(t)

        dx1, de1 = torch.empty_like(xd), torch.empty_like(e1d)
        g_m = [like(wq), torch.empty(128, device=DEV), like(mw[0]), like(mw[1]), like(mw[2]), like(mwo),
               torch.empty(128, device=DEV)]
        _lib.call("gcgcn_maggc_bwd", bt.ref, layers, heads, _p(y1), _lib.F32, _p(wq), _p(mw[0]), _p(mw[1]), _p(mw[2]),
                  _p(mwo), _p(dy2), _p(saved_m), _p(dx1), _p(de1), *[_p(t) for t in g_m], None, ws.data_ptr(),
                  ws.numel(), st)
        dy1_total = dy1 + dx1                       # y1 feeds the MAGGC block and the output (alpha = 1, G:339)
        dx0, de0 = torch.empty_like(xd), torch.empty_like(e0d)
        du, dv, dc = torch.empty(128, device=DEV), torch.empty(128, device=DEV), torch.empty(1, device=DEV)
        g_c = [like(cw[0]), like(cw[1]), like(cw[2]), like(cwo), torch.empty(128, device=DEV)]
        _lib.call("gcgcn_caggc_bwd", bt.ref, layers, _p(xd), _p(e0d), _lib.F32, _p(u), _p(v), _p(cw[0]), _p(cw[1]),
                  _p(cw[2]), _p(cwo), _p(dy1_total), _p(saved_c), _p(dx0), _p(de0), _p(du), _p(dv), _p(dc),
                  *[_p(t) for t in g_c], None, ws.data_ptr(), ws.numel(), st)
        torch.cuda.synchronize()
        assert_close(de1, ref["de1"], FP32_TOL, "maggc_bwd de1")
        assert_close(dx0, ref["dx0"], FP32_TOL, "caggc_bwd dx0")
        assert_close(de0, ref["de0"], FP32_TOL, "caggc_bwd de0")
        # parameter gradients of the output linears (layouts identical on both routes)
        assert_close(g_m[5], mag.linear_layer.weight.grad, 5 * FP32_TOL, "maggc dWout")
        assert_close(g_c[3], cag.linear_layer.weight.grad, 5 * FP32_TOL, "caggc dWout")
        assert_close(g_m[0], torch.cat([l.weight.grad for l in mha.linears_q], 0), 5 * FP32_TOL, "dWq")


def test_block_supported_reports_the_kernel_envelope():
    lib = _lib.load()
    small = RaggedBatch([42, 7, 19], DEV)
    big = RaggedBatch([128, 7], DEV)
    assert lib.gcgcn_block_supported(small.ref, 8, 2, 1) == 1
    assert lib.gcgcn_block_supported(small.ref, 4, 4, 1) == 1
    assert lib.gcgcn_block_supported(small.ref, 2, 2, 1) == 0        # head width 64 is not instantiated
    assert lib.gcgcn_block_supported(small.ref, 1, 2, 0) == 1
    assert lib.gcgcn_block_supported(big.ref, 8, 2, 1) == 0          # > 64 nodes: per-op route
    x = torch.zeros(small.total_nodes, 128, device=DEV)
    with pytest.raises(_lib.GcgcnError):                              # and the fused entry point refuses, loudly
        _lib.call("gcgcn_mha_stack_fwd", big.ref, 8, 2, *([_p(x)] * 15), None, None, 0, _stream(DEV))
