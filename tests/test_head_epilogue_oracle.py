"""CPU: the head-epilogue oracle.  torch's own nn.UpsamplingBilinear2d is the reference
operator (smp SegmentationHead, lib/pose_regressor.py:633-666); oracle/head_epilogue_ref.c restates its arithmetic and
must agree with it bit for bit; the xyz split restates lib/pose_regressor.py:729-732."""
import numpy as np
import pytest
import torch

from helpers import port, syn
from oracle import native


@pytest.mark.parametrize("shape,scale", [((2, 7, 24, 32), 4), ((1, 67, 30, 40), 4), ((1, 3, 120, 160), 4), ((2, 2, 50, 66), 2),
                                         ((1, 2, 40, 30), 3), ((1, 1, 100, 7), 4)])
def test_c_restatement_equals_torch_upsampling(shape, scale):
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape)))
    want = torch.nn.UpsamplingBilinear2d(scale_factor=scale)(x)
    got = native.upsample_bilinear(x, scale)
    assert got.shape == want.shape and torch.equal(got, want)


def test_upsampling_endpoints_and_constants():
    x = torch.randn(1, 2, 9, 40, generator=torch.Generator().manual_seed(1))
    up = native.upsample_bilinear(x, 4)
    # align_corners: the four corners are copied; a constant plane stays constant up to one rounding of w0 + w1
    assert torch.equal(up[..., 0, 0], x[..., 0, 0]) and torch.equal(up[..., -1, -1], x[..., -1, -1])
    c = native.upsample_bilinear(torch.full((1, 1, 10, 40), 3.25), 4)
    assert float((c - 3.25).abs().max()) <= 4e-7


def test_split_xyz_is_the_reference_indexing():
    xyz = torch.randn(2, 18, 5, 6)
    # lib/pose_regressor.py:729-732 verbatim semantics
    xy_index = np.array([i for i in range(xyz.shape[1]) if i % 3 != 0]) - 1
    z_index = np.array([i for i in range(xyz.shape[1]) if i % 3 == 0]) + 2
    xy, z = port.split_xyz(xyz)
    assert torch.equal(xy, xyz[:, xy_index]) and torch.equal(z, xyz[:, z_index])


def test_lowres_scene_lands_where_the_full_resolution_one_does():
    frames = [[(30, 30, 14, 1), (90, 40, 18, 3)], [(60, 60, 20, 2)]]
    low = syn.render_lowres_heads(frames, 96, 128, 4, seed=3)
    assert low["mask"].shape == (2, 7, 24, 32) and low["xy"].shape == (2, 12, 24, 32)
    inv_k = torch.inverse(syn.camera_intrinsics())
    cat, agg = port.pose_recover_lowres(low, inv_k, 32, idx_source=port.seeded_idx_source(7))
    assert agg["class_ids"].tolist() == [1, 3, 2] and agg["sample_ids"].tolist() == [0, 0, 1]
    centres = agg["xy"]
    want = torch.tensor([[30.0, 30.0], [90.0, 40.0], [60.0, 60.0]])
    assert float((centres - want).abs().max()) < 3.0
