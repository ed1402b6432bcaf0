"""Host-side logic of the round-2 additions that needs no GPU: the gates of the fused kernels, the serial fall-back of
ops.fork_join / ops.GraphReplay, the tiles-per-replay rule of the frame pipeline."""
import torch
import torch.nn as nn


def test_fork_join_runs_both_branches_serially_without_a_capture():
    from realcamnet_b200 import ops

    order = []
    ra, rb = ops.fork_join(lambda: order.append("a") or 1, lambda: order.append("b") or 2, torch.device("cpu"))
    assert (ra, rb) == (1, 2) and sorted(order) == ["a", "b"]


def test_graph_replay_is_a_plain_call_on_cpu_tensors():
    from realcamnet_b200 import ops

    class M(ops.GraphReplay, nn.Module):
        def __init__(self):
            super().__init__()
            self.w = nn.Parameter(torch.ones(3))

        def forward(self, x):
            return self._graph_call(lambda xs: xs[0] * self.w, [x])

    m = M().enable_cuda_graphs(True)
    x = torch.arange(3.0)
    assert torch.equal(m(x), x)
    with torch.no_grad():
        m.w.mul_(2)
    assert torch.equal(m(x), 2 * x)


def test_fused_kernel_gates_follow_shapes_and_engine():
    from realcamnet_b200 import LiteISP, ops
    from realcamnet_b200.layers import Linear, conv3x3

    old = ops.get_engine()
    try:
        ops.set_engine("bf16x3")
        lsc = LiteISP.Lens_Shading_Correction(2, 128, 128)
        cf = conv3x3(4, 128)
        ok = torch.zeros(1, 2, 64, 128)
        assert ops.fused_ingest_ok(lsc.layers(), ok, cf)
        assert not ops.fused_ingest_ok(lsc.layers(), torch.zeros(1, 2, 64, 96), cf)        # W % 64 != 0
        assert not ops.fused_ingest_ok(lsc.layers(), torch.zeros(1, 2, 63, 128), cf)       # odd H
        assert not ops.fused_ingest_ok(lsc.layers(), ok, conv3x3(4, 64))                   # other conv_first width
        assert not ops.fused_ingest_ok(LiteISP.Lens_Shading_Correction(2, 48, 48).layers(), ok)   # LiteISP's 48-wide MLP
        fc1, fc2 = Linear(64, 256), Linear(256, 64)
        assert not ops.mlp_fused_ok(fc1, fc2, None)                                        # no operand planes, no LayerNorm rows
        assert not ops.mlp_fused_ok(Linear(128, 512), Linear(512, 128), None)
        ops.set_engine("fp32")                                                             # exact-arithmetic engine: per-layer path
        assert not ops.fused_ingest_ok(lsc.layers(), ok, cf)
    finally:
        ops.set_engine(old)


def test_tiles_per_replay_rule():
    from realcamnet_b200 import frame

    assert frame.batch_size_for(40, 8) == 8 and frame.batch_size_for(40, 20) == 20 and frame.batch_size_for(10, 8) == 5
    assert frame.batch_size_for(5, 8) == 5 and frame.batch_size_for(7, 4) == 1 and frame.batch_size_for(1, 8) == 1
    assert frame.batch_size_for(40) == frame.batch_size_for(40, frame.MAX_BATCH)
