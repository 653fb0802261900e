"""CPU: property tests (hypothesis) of the oracle pieces the next rows rest on -- they must hold for ANY input, not only the
seeded scenes: IoU against the per-pair definition, the pairing rule, bilinear up-sampling of affine planes, label order."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from helpers import port
from oracle import native

masks_st = st.integers(0, 2 ** 31 - 1).map(lambda s: np.random.default_rng(s))


@settings(max_examples=40, deadline=None, derandomize=True)
@given(rng=masks_st, n1=st.integers(0, 5), n2=st.integers(0, 5), h=st.integers(1, 9), w=st.integers(1, 40))
def test_iou_matrix_is_the_pairwise_definition(rng, n1, n2, h, w):
    a = torch.from_numpy((rng.random((n1, h, w)) > 0.5).astype(np.float32))
    b = torch.from_numpy((rng.random((n2, h, w)) > 0.6).astype(np.float32))
    got = port.batchwise_get_2d_iou(a, b)
    assert got.shape == (n1, n2)
    for i in range(n1):
        for j in range(n2):
            want = port.torch_get_2d_iou(a[i], b[j])
            assert torch.equal(got[i, j], want) or (torch.isnan(got[i, j]) and torch.isnan(want))


@settings(max_examples=40, deadline=None, derandomize=True)
@given(rng=masks_st, ng=st.integers(1, 8), np_=st.integers(0, 8))
def test_pairing_rule(rng, ng, np_):
    gc = torch.from_numpy(rng.integers(1, 4, ng))
    pc = torch.from_numpy(rng.integers(1, 4, np_))
    iou = torch.from_numpy(rng.choice([0.0, 0.25, 0.5, 0.5, 0.75], size=(ng, np_)).astype(np.float32))
    best, order = port.match_pairs(gc, pc, iou)
    for i in range(ng):
        same = [j for j in range(np_) if int(pc[j]) == int(gc[i])]
        top = max((float(iou[i, j]) for j in same), default=0.0)
        if top > 0:
            j = int(best[i])
            assert int(pc[j]) == int(gc[i]) and float(iou[i, j]) == top
            assert all(float(iou[i, k]) < top for k in same if k < j)            # first maximum
        else:
            assert int(best[i]) == -1
    kept = [i for i in range(ng) if int(best[i]) >= 0]
    assert order.tolist() == sorted(kept, key=lambda i: (int(gc[i]), i))


@settings(max_examples=25, deadline=None, derandomize=True)
@given(hl=st.integers(2, 12), wl=st.integers(2, 12), scale=st.integers(1, 5), a=st.floats(-2, 2), b=st.floats(-2, 2), c=st.floats(-5, 5))
def test_upsampling_reproduces_affine_planes_and_corners(hl, wl, scale, a, b, c):
    ys, xs = torch.meshgrid(torch.arange(hl, dtype=torch.float32), torch.arange(wl, dtype=torch.float32), indexing="ij")
    plane = (a * xs + b * ys + c).reshape(1, 1, hl, wl).contiguous()
    up = native.upsample_bilinear(plane, scale)
    h, w = hl * scale, wl * scale
    sy, sx = (hl - 1) / max(h - 1, 1), (wl - 1) / max(w - 1, 1)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float64) * sy, torch.arange(w, dtype=torch.float64) * sx, indexing="ij")
    want = a * xx + b * yy + c                                                   # align_corners: an affine plane stays affine
    assert float((up[0, 0].double() - want).abs().max()) <= 2e-5 * (1 + abs(a) * wl + abs(b) * hl + abs(c))
    # the first corner is copied exactly; the last one only up to the rounding of scale * (out - 1) (ATeN's rule, kept)
    assert float(up[0, 0, 0, 0]) == float(plane[0, 0, 0, 0])
    assert abs(float(up[0, 0, -1, -1]) - float(plane[0, 0, -1, -1])) <= 1e-5 * (1 + abs(float(plane[0, 0, -1, -1])))
    assert float(up.min()) >= float(plane.min()) - 1e-5 and float(up.max()) <= float(plane.max()) + 1e-5


@settings(max_examples=30, deadline=None, derandomize=True)
@given(rng=masks_st, b=st.integers(1, 3), h=st.integers(1, 12), w=st.integers(1, 16))
def test_labels_are_raster_ordered_and_never_cross_images(rng, b, h, w):
    fg = torch.from_numpy(rng.random((b, h, w)) > 0.55)
    lab, total = port.label_instances(fg)
    assert torch.equal(lab != 0, fg)
    firsts = []
    for k in range(1, total + 1):
        idx = torch.nonzero(lab == k)
        assert len(set(idx[:, 0].tolist())) == 1                                  # one image per component
        firsts.append(tuple(idx[0].tolist()))                                     # nonzero() is raster ordered
    assert firsts == sorted(firsts)
    # 4-connectivity: horizontally / vertically adjacent foreground pixels share a label
    assert torch.equal(lab[:, :, 1:][fg[:, :, 1:] & fg[:, :, :-1]], lab[:, :, :-1][fg[:, :, 1:] & fg[:, :, :-1]])
    assert torch.equal(lab[:, 1:][fg[:, 1:] & fg[:, :-1]], lab[:, :-1][fg[:, 1:] & fg[:, :-1]])
