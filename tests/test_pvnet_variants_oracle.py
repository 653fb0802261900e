"""CPU, build container only: the oracle restatements of the other PVNet drivers against
the reference's own ransac_voting_gpu.py imported unmodified, on identical fixed pixel pairs."""
import pytest
import torch

import helpers
from helpers import port, syn
from oracle import ref_import

pytestmark = [pytest.mark.skipif(not ref_import.available(), reason="reference sources not on this machine"),
              pytest.mark.filterwarnings("ignore")]


def class_scene(vn=1, seed=2):
    frames = [[(30, 30, 14, 1), (90, 40, 18, 3), (60, 75, 12, 2)], [(40, 50, 20, 1), (100, 30, 9, 2)], [(64, 48, 22, 3)]]
    logits = syn.render_heads(frames, 96, 128, seed=seed)
    cat = port.class_compression(logits, 7)
    vertex = cat["xy"].permute(0, 2, 3, 1).unsqueeze(3)
    if vn > 1:
        vertex = torch.cat([vertex, vertex.flip(-1) * torch.tensor([1.0, -1.0])], dim=3)
    return cat["mask"], vertex.contiguous()


@pytest.mark.parametrize("vn", [1, 2])
def test_v2_equals_reference(vn):
    ref = ref_import.load()
    mask, vertex = class_scene(vn)
    hn = 40
    a = port.ransac_voting_layer_v2(mask, vertex, 4, hn, idx_source=port.seeded_idx_source(9))
    with ref_import.fixed_idxs(port.seeded_idx_source(9), hn, vn):
        b = ref.rvg.ransac_voting_layer_v2(mask, vertex, 4, hn)
    assert a.shape == b.shape == (3, 3, vn, 2) and torch.equal(a, b)


def test_hypothesis_dump_equals_reference():
    ref = ref_import.load()
    mask, vertex = class_scene()
    hn = 48
    a_h, a_c = port.ransac_voting_hypothesis(mask, vertex, hn, idx_source=port.seeded_idx_source(4))
    with ref_import.fixed_idxs(port.seeded_idx_source(4), hn):
        b_h, b_c = ref.rvg.ransac_voting_hypothesis(mask, vertex, hn)
    assert torch.equal(a_h, b_h) and a_c.dtype == b_c.dtype and torch.equal(a_c, b_c)
    assert torch.equal(a_c[2], torch.ones_like(a_c[2])) and int(a_h[2].abs().sum()) == 0     # image 2 has no class-1 pixels


def test_distribution_estimators_equal_reference():
    ref = ref_import.load()
    mask, vertex = class_scene()
    mask, vertex = mask[:2], vertex[:2]                       # both images have class-1 pixels (see the skip-shape quirk)
    kw = dict(round_hyp_num=32, min_hyp_num=96, topk=24)
    a_mean, a_cov = port.estimate_voting_distribution(mask, vertex, idx_source=port.seeded_idx_source(6), **kw)
    with ref_import.fixed_idxs(port.seeded_idx_source(6), 32):
        b_mean, b_cov = ref.rvg.estimate_voting_distribution(mask, vertex, **kw)
    assert torch.equal(a_mean, b_mean) and torch.equal(a_cov, b_cov)
    _, a_cov2 = port.estimate_voting_distribution_with_mean(mask, vertex, a_mean, idx_source=port.seeded_idx_source(7), **kw)
    with ref_import.fixed_idxs(port.seeded_idx_source(7), 32):
        _, b_cov2 = ref.rvg.estimate_voting_distribution_with_mean(mask, vertex, b_mean, **kw)
    assert torch.equal(a_cov2, b_cov2)


def instance_scene(vn=1):
    frames = [[(30, 30, 14, 1)], [(64, 48, 22, 3)], [(100, 60, 1.0, 2)], []]
    logits = syn.render_heads(frames, 96, 128, seed=7)
    cat = port.class_compression(logits, 7)
    vertex = cat["xy"].permute(0, 2, 3, 1).unsqueeze(3)
    if vn > 1:
        vertex = torch.cat([vertex, vertex.flip(-1) * torch.tensor([1.0, -1.0])], dim=3)
    return (cat["mask"] != 0).float(), vertex.contiguous()


@pytest.mark.parametrize("vn", [1, 2])
def test_v4_v5_equal_reference(vn):
    ref = ref_import.load()
    mask, vertex = instance_scene(vn)
    hn = 40
    a_pts, a_var = port.ransac_voting_layer_v4(mask, vertex, hn, idx_source=port.seeded_idx_source(3))
    with ref_import.fixed_idxs(port.seeded_idx_source(3), hn, vn), ref_import.legacy_uint8_masks():
        b_pts, b_var = ref.rvg.ransac_voting_layer_v4(mask, vertex, hn)
    assert torch.equal(a_pts, b_pts) and torch.equal(a_var, b_var)
    assert torch.equal(a_var[3], torch.ones(vn)) and float(a_var[0, 0]) < 1.0
    a_pts, a_conf = port.ransac_voting_layer_v5(mask, vertex, hn, max_num=30000, idx_source=port.seeded_idx_source(5))
    with ref_import.fixed_idxs(port.seeded_idx_source(5), hn, vn), ref_import.legacy_uint8_masks():
        b_pts, b_conf = ref.rvg.ransac_voting_layer_v5(mask, vertex, hn, max_num=30000)
    assert torch.equal(a_pts, b_pts) and torch.equal(a_conf, b_conf)
    assert torch.equal(a_conf[3], torch.zeros(vn)) and float(a_conf[:2, 0].min()) > 0.3


def dense_scene(vn=1):
    """every image has at least a handful of mask pixels (generate_hypothesis / v6 have no working skip branch per image)"""
    frames = [[(30, 30, 14, 1)], [(64, 48, 22, 3)], [(100, 60, 6, 2), (40, 40, 9, 1)]]
    logits = syn.render_heads(frames, 96, 128, seed=11)
    cat = port.class_compression(logits, 7)
    vertex = cat["xy"].permute(0, 2, 3, 1).unsqueeze(3)
    if vn > 1:
        vertex = torch.cat([vertex, vertex.flip(-1) * torch.tensor([1.0, -1.0])], dim=3)
    return (cat["mask"] != 0).float(), vertex.contiguous()


@pytest.mark.parametrize("vn", [1, 2])
def test_v6_equals_reference(vn, capsys):
    ref = ref_import.load()
    mask, vertex = dense_scene(vn)
    hn = 40
    # (a) no sub-sampling (max_num above the batch's foreground count)
    a_pts, a_conf = port.ransac_voting_layer_v6(mask, vertex, hn, max_num=30000, idx_source=port.seeded_idx_source(8))
    with ref_import.fixed_idxs(port.seeded_idx_source(8), hn, vn), ref_import.legacy_uint8_masks():
        b_pts, b_conf = ref.rvg.ransac_voting_layer_v6(mask, vertex, hn, max_num=30000)
    assert a_pts.shape == b_pts.shape == (3, vn, 2) and torch.equal(a_pts, b_pts) and torch.equal(a_conf, b_conf)
    assert float(a_conf[:2, 0].min()) > 0.3
    # (b) the whole batch's count decides: sub-sampled by max_num / sum(mask) in every image (same RNG stream on both sides)
    torch.manual_seed(77)
    a_pts, a_conf = port.ransac_voting_layer_v6(mask, vertex, hn, max_num=900, idx_source=port.seeded_idx_source(8))
    torch.manual_seed(77)
    with ref_import.fixed_idxs(port.seeded_idx_source(8), hn, vn), ref_import.legacy_uint8_masks():
        b_pts, b_conf = ref.rvg.ransac_voting_layer_v6(mask, vertex, hn, max_num=900)
    assert torch.equal(a_pts, b_pts) and torch.equal(a_conf, b_conf)
    # (c) whole batch below min_num: zeros for every image
    a_pts, a_conf = port.ransac_voting_layer_v6(mask, vertex, hn, min_num=10 ** 6, idx_source=port.seeded_idx_source(8))
    with ref_import.legacy_uint8_masks():
        b_pts, b_conf = ref.rvg.ransac_voting_layer_v6(mask, vertex, hn, min_num=10 ** 6)
    assert torch.equal(a_pts, b_pts) and torch.equal(a_conf, b_conf) and int(a_pts.abs().sum()) == 0
    capsys.readouterr()            # the reference prints the mask's device (:880)


def test_center_motion_and_hypothesis_driver_equal_reference():
    ref = ref_import.load()
    mask, vertex = instance_scene()
    # ransac_voting_center: only the images below min_num contribute (a zero mask each); an image at or above min_num makes
    # the reference raise (its masked_select at :636 broadcasts a [h,w,1,1] mask against the [h,w,2] field)
    a = port.ransac_voting_center(mask[2:], vertex[2:, :, :, 0], 16, min_num=100)
    with ref_import.legacy_uint8_masks():
        b = ref.rvg.ransac_voting_center(mask[2:], vertex[2:, :, :, 0], 16, min_num=100)
        with pytest.raises(RuntimeError):
            ref.rvg.ransac_voting_center(mask, vertex[:, :, :, 0], 16, min_num=100)
    assert len(a) == len(b) == 2 and all(torch.equal(x, y) and x.shape == (96, 128) for x, y in zip(a, b))
    # ransac_motion_voting
    mask2, vertex2 = instance_scene(2)
    with ref_import.legacy_uint8_masks():
        want = ref.rvg.ransac_motion_voting(mask2, vertex2)
    got = port.ransac_motion_voting(mask2, vertex2)
    assert got.shape == want.shape == (4, 2, 2) and torch.equal(got, want) and int(got[3].abs().sum()) == 0
    # generate_hypothesis (the driver at :991, not the native kernel)
    mask3, vertex3 = dense_scene(2)
    hn = 24
    a_h, a_c = port.generate_hypothesis(mask3, vertex3, hn, idx_source=port.seeded_idx_source(12))
    with ref_import.fixed_idxs(port.seeded_idx_source(12), hn, 2), ref_import.legacy_uint8_masks():
        b_h, b_c = ref.rvg.generate_hypothesis(mask3, vertex3, hn)
    assert a_h.shape == (3, hn, 2, 2) and torch.equal(a_h, b_h) and a_c.dtype == b_c.dtype and torch.equal(a_c, b_c)
    with pytest.raises(NameError), ref_import.legacy_uint8_masks():
        ref.rvg.generate_hypothesis(mask, vertex, hn)                 # image 3 of instance_scene is empty
    with pytest.raises(NameError):
        port.generate_hypothesis(mask, vertex, hn)


def test_motion_mean_host_logic_ignores_values_outside_the_mask():
    """The drop-in's batched masked sums (device-agnostic torch) against the oracle's per-image loop, with NaN / inf planted
    outside the mask: the reference indexes the masked pixels, so they must not matter."""
    from fastposecnn_b200.ransac_voting_gpu_layer.ransac_voting_gpu import _motion_mean
    mask, vertex = instance_scene(2)
    vertex = vertex.clone()
    outside = mask == 0
    vertex[outside] = float("nan")
    vertex[0, 0, 0] = float("inf")
    want = port.ransac_motion_voting(mask, vertex)
    got = _motion_mean(mask, vertex)
    assert torch.isfinite(got).all() and got.shape == want.shape
    assert float(((got - want).norm(dim=-1) / want.norm(dim=-1).clamp_min(1e-6)).max()) <= 1e-5
    assert int(got[3].abs().sum()) == 0
