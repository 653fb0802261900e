"""CPU: known-answer tests derived for the path (SURVEY.md section 8c i-viii; the reference ships none)."""
import math

import torch

import helpers
from helpers import port, syn
from oracle import native


def _radial_field(h, w, cx, cy):
    ys = torch.arange(h, dtype=torch.float32).view(h, 1).expand(h, w)
    xs = torch.arange(w, dtype=torch.float32).view(1, w).expand(h, w)
    d = torch.stack([cx - xs, cy - ys], 0)
    return d / d.norm(dim=0, keepdim=True).clamp_min(1e-12)


def test_kat_i_toy_5x5_votes_for_centre():
    """lib/hough_voting.py:598-618: 5x5 mask with a 3x3 interior, unit vectors toward (2,2)."""
    mask = torch.zeros(1, 5, 5)
    mask[0, 1:4, 1:4] = 1
    field = _radial_field(5, 5, 2.0, 2.0) * mask
    vertex = field.unsqueeze(0).permute(0, 2, 3, 1).unsqueeze(3)
    out = port.ransac_voting_layer_v3(mask, vertex, 64, idx_source=port.seeded_idx_source(0))
    assert torch.allclose(out[0, 0], torch.tensor([2.0, 2.0]), atol=1e-4)


def test_kat_ii_disc_exact_field_all_vote():
    h, w, cx, cy, r = 64, 80, 40.0, 30.0, 12.0
    ys = torch.arange(h).view(h, 1)
    xs = torch.arange(w).view(1, w)
    mask = (((xs - cx) ** 2 + (ys - cy) ** 2) <= r * r).float().unsqueeze(0)
    field = _radial_field(h, w, cx, cy) * mask
    vertex = field.unsqueeze(0).permute(0, 2, 3, 1).unsqueeze(3)
    det = []
    out = port.ransac_voting_layer_v3(mask, vertex, 32, idx_source=port.seeded_idx_source(1), details=det)
    assert torch.allclose(out[0, 0], torch.tensor([cx, cy]), atol=1e-3)
    tn = det[0]["tn"]
    hyp, counts = det[0]["hyp"][:, 0], det[0]["counts"][:, 0]
    good = (hyp - torch.tensor([cx, cy])).norm(dim=1) < 1e-3
    assert good.sum() > 20
    assert (counts[good] >= tn - 1).all()          # every pixel but the centre one (zero direction) votes


def test_kat_iii_iv_v_instance_semantics():
    mask = torch.zeros(2, 12, 16, dtype=torch.int64)
    mask[0, 2:5, 2:5] = 5
    mask[0, 2:5, 5:8] = 2                      # touches the class-5 block -> one instance, class 2
    mask[0, 8:10, 2:4] = 3
    mask[0, 10:12, 4:6] = 4                    # diagonal contact only -> separate instances
    mask[1, 2:5, 2:5] = 1                      # same place, next frame -> separate instance, sample id 1
    cat = {"mask": mask, "quaternion": torch.randn(2, 4, 12, 16), "scales": torch.rand(2, 3, 12, 16),
           "xy": torch.randn(2, 2, 12, 16), "z": torch.randn(2, 12, 16)}
    agg = port.aggregate(cat)
    assert agg["class_ids"].tolist() == [2, 3, 4, 1]
    assert agg["sample_ids"].tolist() == [0, 0, 0, 1]
    assert helpers.oracle_tns(agg) == [18, 4, 4, 9]


def test_kat_vi_parallel_directions_degenerate():
    direct = torch.tensor([[[1.0, 0.0]], [[1.0, 0.0]], [[0.0, 1.0]]])
    coords = torch.tensor([[0.0, 0.0], [5.0, 3.0], [2.0, 2.0]])
    idxs = torch.tensor([[[0, 1]], [[0, 2]], [[1, 1]]], dtype=torch.int32)
    hyp = native.ransac_voting.generate_hypothesis(direct, coords, idxs)
    assert hyp[0, 0].tolist() == [0.0, 0.0]            # parallel lines
    assert hyp[1, 0].tolist() == [2.0, 0.0]            # y = 0 meets x = 2
    assert hyp[2, 0].tolist() == [0.0, 0.0]            # same pixel twice


def test_kat_vii_small_instance_skipped_without_draw():
    mask = torch.zeros(2, 8, 8)
    mask[0, 1, 1:5] = 1                                # 4 px < min_num
    mask[1, 2:6, 2:6] = 1
    vertex = torch.randn(2, 8, 8, 1, 2)
    draws = []

    def src(i, hn, vn, tn):
        draws.append((i, tn))
        return torch.zeros((hn, vn, 2), dtype=torch.int32)
    out = port.ransac_voting_layer_v3(mask, vertex, 8, idx_source=src)
    assert out[0].tolist() == [[0.0, 0.0]] and draws == [(1, 16)]


def test_kat_viii_unit_pose():
    q = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    xy = torch.tensor([[319.5, 239.5]])
    z = torch.tensor([[1000.0]])
    inv_k = torch.inverse(syn.camera_intrinsics())
    R, T, RT = port.batchwise_get_RT(q, xy, z, inv_k)
    assert torch.allclose(T, torch.tensor([[0.0, 0.0, 1.0]]), atol=1e-6)
    want = torch.eye(4)
    want[:3, :3] = R[0]
    want[:3, 3] = -(R[0] @ T[0])
    assert torch.allclose(RT[0], want, atol=1e-6)
    # quaternion (cos t/2, 0, 0, sin t/2): the reference's matrix (with its final transpose) is orthonormal
    t = 0.7
    q2 = torch.tensor([[math.cos(t / 2), 0.0, 0.0, math.sin(t / 2)]])
    R2 = port.quats_2_rotation_matrix(q2)[0]
    assert torch.allclose(R2 @ R2.T, torch.eye(3), atol=1e-6)


def test_c_kernels_fma_variant_differs_only_in_last_bits():
    g = torch.Generator().manual_seed(0)
    tn, hn = 4000, 64
    coords = torch.stack([torch.randint(0, 640, (tn,), generator=g), torch.randint(0, 480, (tn,), generator=g)], 1).float()
    d = torch.tensor([320.0, 240.0]) - coords + torch.randn(tn, 2, generator=g) * 4
    direct = (d / d.norm(dim=1, keepdim=True).clamp_min(1e-9)).unsqueeze(1).contiguous()
    idxs = torch.randint(0, tn, (hn, 1, 2), generator=g, dtype=torch.int32)
    a = native.ransac_voting.generate_hypothesis(direct, coords, idxs)
    b = native.ransac_voting_fma.generate_hypothesis(direct, coords, idxs)
    assert (a - b).abs().max() < 1e-2 and helpers.rel_err(a.reshape(hn, 2), b.reshape(hn, 2)) < 1e-4
    ia = torch.zeros((hn, 1, tn), dtype=torch.uint8)
    ib = torch.zeros((hn, 1, tn), dtype=torch.uint8)
    native.ransac_voting.voting_for_hypothesis(direct, coords, a, ia, 0.999)
    native.ransac_voting_fma.voting_for_hypothesis(direct, coords, a, ib, 0.999)
    assert (ia != ib).float().mean() < 1e-4 and int(ia.sum()) > 0
