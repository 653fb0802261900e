"""CPU: the matching restatement in oracle/port.py against the reference's own
lib/matching.py + gpu_tensor_funcs.batchwise_get_2d_iou imported unmodified, and against the committed fixture."""
import pytest
import torch

import helpers
from helpers import port
from oracle import ref_import

needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference sources not on this machine")


def same_dict(a, b):
    if a is None or b is None:
        assert a is None and b is None
        return
    assert set(a.keys()) == set(b.keys())
    for k in b:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


@needs_ref
@pytest.mark.parametrize("name", helpers.MATCHING_SCENES)
@pytest.mark.filterwarnings("ignore")
def test_iou_and_matches_equal_reference(name):
    ref = ref_import.load()
    preds, gts = helpers.matching_scene(name)
    a = port.batchwise_get_2d_iou(gts["instance_masks"], preds["instance_masks"])
    b = ref.gtf.batchwise_get_2d_iou(gts["instance_masks"], preds["instance_masks"])
    assert a.dtype == b.dtype and torch.equal(a, b)
    assert torch.equal(port.torch_get_2d_iou(gts["instance_masks"][0], preds["instance_masks"][0]),
                       ref.gtf.torch_get_2d_iou(gts["instance_masks"][0], preds["instance_masks"][0]))
    same_dict(port.batchwise_find_matches(preds, gts), ref.mg.batchwise_find_matches(preds, gts))


@needs_ref
@pytest.mark.parametrize("name", helpers.MATCHING_SCENES)
@pytest.mark.filterwarnings("ignore")
def test_fill_missing_variant_equals_reference(name):
    ref = ref_import.load()
    preds, gts = helpers.matching_scene(name)
    if hasattr(ref.mg.get_standard_preds, "standard_preds"):
        del ref.mg.get_standard_preds.standard_preds      # cached on the function (lib/matching.py:187)
    same_dict(port.batchwise_find_matches2(preds, gts), ref.mg.batchwise_find_matches2(preds, gts))


def test_empty_inputs_return_none():
    preds, gts = helpers.matching_scene("shifted")
    assert port.batchwise_find_matches(None, gts) is None
    assert port.batchwise_find_matches(preds, {}) is None
    empty = {k: v[:0] for k, v in preds.items()}
    assert port.batchwise_find_matches(empty, gts) is None
    p2, g2 = helpers.matching_scene("no_overlap")
    assert port.batchwise_find_matches(p2, g2) is None


def test_iou_of_empty_masks_is_nan():
    z = torch.zeros((1, 4, 4))
    assert torch.isnan(port.batchwise_get_2d_iou(z, z)).all()


@pytest.mark.parametrize("name", helpers.MATCHING_SCENES)
def test_golden_fixture(name):
    preds, gts, iou, matches = helpers.load_matching_golden(name)
    assert torch.equal(port.batchwise_get_2d_iou(gts["instance_masks"], preds["instance_masks"]), iou)
    same_dict(port.batchwise_find_matches(preds, gts), matches)
