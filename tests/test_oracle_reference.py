"""Pins the oracle against the reference's own modules when /root/reference is present
(build container).  Skipped on the GPU box, where only the committed golden vectors exist."""
import pytest
import torch

from oracle import reference_loader as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not mounted")


def test_oracle_is_bit_exact_against_reference():
    from oracle import pin_against_reference as pin
    torch.set_num_threads(1)
    lines = []
    worst = max(pin.run("glove", 2, 8, lines.append), pin.run_pool_and_gathers(lines.append))
    assert worst == 0.0, "\n".join(lines)


def test_dropin_state_dict_round_trips_with_reference():
    from gcgcn_b200.modules import GraphBlocks
    for variant, (layers, heads) in {"glove": (2, 8), "bert": (4, 4)}.items():
        ref = R.build_graph_modules(layers, heads, seed=3, variant=variant)
        gb = GraphBlocks(layers, heads)
        mine = (gb.get_weighted_adj_matrix, gb.get_adj_matrix[0], gb.graphcnn[0], gb.graphcnn[1])
        for r, m in zip(ref, mine):
            m.load_state_dict(r.state_dict(), strict=True)       # reference -> drop-in
            r.load_state_dict(m.state_dict(), strict=True)       # drop-in -> reference
            assert [n for n, _ in r.named_parameters()] == [n for n, _ in m.named_parameters()]
