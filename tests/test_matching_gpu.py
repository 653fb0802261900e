"""GPU: matching drop-ins (fastposecnn_b200.matching, gpu_tensor_funcs.batchwise_get_2d_iou) against the oracle
restatement and the committed reference outputs (tests/golden/matching_*.npz) -- IoU values, pairings and output
order bit-exact."""
import pytest
import torch

import helpers
from helpers import port, syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def to_dev(d):
    return {k: v.to(DEV) for k, v in d.items()}


def same_dict(got, want):
    if want is None or got is None:
        assert got is None and want is None
        return
    assert set(got.keys()) == set(want.keys())
    for k in want:
        g = got[k].cpu()
        assert g.dtype == want[k].dtype and g.shape == want[k].shape and torch.equal(g, want[k]), k


@pytest.mark.parametrize("name", helpers.MATCHING_SCENES)
def test_golden_iou_and_matches(name):
    import fastposecnn_b200 as fp
    preds, gts, iou, matches = helpers.load_matching_golden(name)
    got_iou = fp.batchwise_get_2d_iou(gts["instance_masks"].to(DEV), preds["instance_masks"].to(DEV))
    assert got_iou.dtype == torch.float32 and torch.equal(got_iou.cpu(), iou)
    same_dict(fp.batchwise_find_matches(to_dev(preds), to_dev(gts)), matches)


@pytest.mark.parametrize("name", helpers.MATCHING_SCENES)
def test_fill_missing_variant(name):
    from fastposecnn_b200 import matching
    preds, gts, _, _ = helpers.load_matching_golden(name)
    if hasattr(matching.get_standard_preds, "standard_preds"):
        del matching.get_standard_preds.standard_preds
    same_dict(matching.batchwise_find_matches2(to_dev(preds), to_dev(gts)), port.batchwise_find_matches2(preds, gts))


def test_mask_dtypes_and_single_pair():
    import fastposecnn_b200 as fp
    g = torch.Generator().manual_seed(3)
    a = (torch.rand((5, 37, 75), generator=g) > 0.6)
    b = (torch.rand((4, 37, 75), generator=g) > 0.5)
    a[4] = False
    b[0] = False                                   # empty vs empty -> NaN, empty vs non-empty -> 0
    want = port.batchwise_get_2d_iou(a, b)
    for cast in (lambda t: t, lambda t: t.float() * 2.5, lambda t: t.to(torch.uint8), lambda t: t.long() * -3):
        got = fp.batchwise_get_2d_iou(cast(a).to(DEV), cast(b).to(DEV)).cpu()
        assert torch.equal(torch.isnan(got), torch.isnan(want))
        assert torch.equal(torch.nan_to_num(got, nan=-1.0), torch.nan_to_num(want, nan=-1.0))
    one = fp.torch_get_2d_iou(a[1].float().to(DEV), b[2].float().to(DEV))
    assert one.dim() == 0 and torch.equal(one.cpu(), port.torch_get_2d_iou(a[1], b[2]))
    nan_mask = torch.zeros((1, 37, 75))
    nan_mask[0, 3, 3] = float("nan")               # NaN is "non-zero" for logical_and, so the pixel counts
    assert float(fp.batchwise_get_2d_iou(nan_mask.to(DEV), nan_mask.to(DEV))[0, 0]) == 1.0


def test_label_volume_predictions_match_dense_ones():
    """Predictions straight from pose_recover (label volume, no dense masks) pair exactly like their dense twins."""
    import fastposecnn_b200 as fp
    frames = [[(30, 30, 14, 1), (90, 40, 18, 3), (60, 75, 12, 6)], [(40, 50, 20, 2), (100, 30, 12, 1)], [], [(64, 48, 22, 3)]]
    logits = syn.render_heads(frames, 96, 128, seed=8)
    inv_k = torch.inverse(syn.camera_intrinsics())
    sparse = fp.pose_recover({k: v.to(DEV) for k, v in logits.items()}, inv_k.to(DEV), 32, seed=1234)
    dense = fp.pose_recover({k: v.to(DEV) for k, v in logits.items()}, inv_k.to(DEV), 32, materialize_dense=True, seed=1234)
    assert "instance_masks" not in sparse
    _, gts = helpers.matching_scene("shifted")
    gts = to_dev(gts)
    a = fp.batchwise_find_matches(sparse, gts)
    b = fp.batchwise_find_matches({k: v for k, v in dense.items() if k != "labels"}, gts)
    assert a is not None and set(a.keys()) == set(b.keys())
    for k in b:
        if k in ("xy", "T", "RT"):     # two runs sample different hypothesis pairs on the device
            continue
        assert torch.equal(a[k], b[k]), k
    # and the oracle agrees on who pairs with whom
    want = port.batchwise_find_matches({k: v.cpu() for k, v in dense.items() if k != "labels"}, {k: v.cpu() for k, v in gts.items()})
    for k in ("sample_ids", "class_ids", "symmetric_ids", "instance_masks", "quaternion", "scales"):
        assert torch.equal(a[k].cpu(), want[k]), k


def test_many_instances_random_blobs():
    """A few hundred random rectangles/ellipses per side across classes and frames: IoU matrix and pairing vs oracle."""
    import fastposecnn_b200 as fp
    g = torch.Generator().manual_seed(11)
    h, w, n = 120, 200, 150
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")

    def blobs(count):
        cy, cx = torch.randint(0, h, (count,), generator=g), torch.randint(0, w, (count,), generator=g)
        ry, rx = torch.randint(1, 25, (count,), generator=g), torch.randint(1, 40, (count,), generator=g)
        m = (((yy[None] - cy[:, None, None]).float() / ry[:, None, None]) ** 2 +
             ((xx[None] - cx[:, None, None]).float() / rx[:, None, None]) ** 2) <= 1.0
        return m.float(), torch.randint(1, 5, (count,), generator=g)

    pm, pc = blobs(n)
    gm, gc = blobs(n + 17)
    want_iou = port.batchwise_get_2d_iou(gm, pm)
    got_iou = fp.batchwise_get_2d_iou(gm.to(DEV), pm.to(DEV)).cpu()
    assert torch.equal(got_iou, want_iou)
    best, order = port.match_pairs(gc, pc, want_iou)
    from fastposecnn_b200 import matching
    bp, bi, pairs, nm = matching.match_instances(matching.pack_masks(gm.to(DEV)), gc.to(DEV), matching.pack_masks(pm.to(DEV)), pc.to(DEV))
    assert torch.equal(bp.cpu().long(), best)
    m = int(nm.item())
    assert m == order.numel()
    assert torch.equal(pairs[:m, 0].cpu().long(), order) and torch.equal(pairs[:m, 1].cpu().long(), best[order])
    rows = torch.arange(gm.shape[0])[best >= 0]
    assert torch.equal(bi.cpu()[rows], want_iou[rows, best[rows]])


def test_odd_image_size_label_volume_path():
    """70x101 frames: h*w is not a multiple of 4 (scalar paint kernel), w is not a multiple of 32 (ragged last bit-plane word)."""
    import fastposecnn_b200 as fp
    frames, h, w = helpers.scenes()["odd_width"]
    logits = syn.render_heads(frames, h, w, seed=4)
    inv_k = torch.inverse(syn.camera_intrinsics())
    sparse = fp.pose_recover({k: v.to(DEV) for k, v in logits.items()}, inv_k.to(DEV), 32, seed=1234)
    dense = {k: v.clone() for k, v in fp.pose_recover({k: v.to(DEV) for k, v in logits.items()}, inv_k.to(DEV), 32, seed=1234,
                                                      materialize_dense=True).items() if k not in ("labels", "xy_mask")}
    gts = {k: v.clone() for k, v in dense.items() if k != "cat_mask"}
    gts["instance_masks"] = torch.roll(gts["instance_masks"], shifts=(1, 2), dims=(1, 2)).contiguous()
    gts["symmetric_ids"] = gts["class_ids"] % 2
    a = fp.batchwise_find_matches({k: v for k, v in sparse.items()}, gts)
    b = fp.batchwise_find_matches(dense, gts)
    want = port.batchwise_find_matches({k: v.cpu() for k, v in dense.items()}, {k: v.cpu() for k, v in gts.items()})
    assert a is not None and want is not None
    for k in ("sample_ids", "class_ids", "symmetric_ids", "instance_masks", "quaternion", "scales", "z"):
        assert torch.equal(a[k], b[k]) and torch.equal(b[k].cpu(), want[k]), k
    want_iou = port.batchwise_get_2d_iou(gts["instance_masks"].cpu(), dense["instance_masks"].cpu())
    assert torch.equal(fp.batchwise_get_2d_iou(gts["instance_masks"], dense["instance_masks"]).cpu(), want_iou)
