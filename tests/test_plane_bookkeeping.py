"""Host-side bookkeeping of the tcgen05 operand planes (no kernels run): plane widths, channel-slice views, pixel strides."""
import pytest
import torch

from realcamnet_b200 import ops


def test_plane_channels_are_k_chunk_multiples():
    assert [ops.plane_channels(c) for c in (1, 2, 4, 16, 17, 32, 33, 64, 80, 128, 200, 320)] == \
        [16, 16, 16, 16, 32, 32, 64, 64, 128, 128, 256, 320]


def test_split_operand_views_and_strides():
    hi = torch.zeros(2, 6, 10, 128, dtype=torch.bfloat16)
    sp = ops.SplitOperand(hi, torch.zeros_like(hi), (None, 2, 6, 10, 128, None, 128, 1, False))
    assert sp.ld == 128
    left, right = sp.channels(0, 64), sp.channels(64, 128)
    assert left.ld == 128 and right.ld == 128 and tuple(left.hi.shape) == (2, 6, 10, 64)
    assert right.hi.data_ptr() - hi.data_ptr() == 64 * 2 and right.key[4] == 64 and right.key[6] == 64
    right.hi.fill_(1)                                      # a view: writes land in the parent planes
    assert float(hi[..., 64:].float().sum()) == 2 * 6 * 10 * 64 and float(hi[..., :64].float().sum()) == 0
    with pytest.raises(ValueError):
        sp.channels(0, 48)                                 # 48 is not a valid plane width
    s2 = ops.SplitOperand(torch.zeros(8, 3, 5, 64, dtype=torch.bfloat16), None, (None, 2, 6, 10, 64, None, 64, 2, False))
    with pytest.raises(ValueError):
        s2.channels(0, 32)                                 # polyphase planes cannot be sliced
    assert ops.plane_ld(torch.zeros(1, 1, 1, 32, dtype=torch.bfloat16)) == 32
    with pytest.raises(ValueError):
        ops.plane_ld(torch.zeros(1, 4, 4, 32, dtype=torch.bfloat16).permute(0, 2, 1, 3)[:, :, ::2])
