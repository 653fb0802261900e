"""GPU: the fused path against the committed golden fixtures (outputs of the reference's own Python,
tests/golden/*.npz) -- independent of the oracle port; runs on the GPU box, where the reference is absent."""
import os

import numpy as np
import pytest
import torch

import helpers
from helpers import syn
from test_oracle_golden import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_fused_matches_reference_golden(path):
    from fastposecnn_b200.pose_recovery import pose_recover
    g, frames, h, w, hn, rseed, iseed = load_golden(path)
    logits = syn.render_heads(frames, h, w, seed=rseed)
    dev = torch.device("cuda:0")
    idxs = syn.presampled_idxs([int(v) for v in g["agg_mask_sizes"]], hn, seed=iseed).reshape(-1, hn, 2)
    out = pose_recover({k: v.to(dev) for k, v in logits.items()}, torch.inverse(syn.camera_intrinsics()).to(dev), hn,
                       idxs=idxs.to(dev), materialize_dense=True)
    assert np.array_equal(out["cat_mask"].cpu().numpy(), g["cat_mask"])
    assert np.array_equal(out["labels"].cpu().numpy(), g["labels"])
    assert np.array_equal(out["class_ids"].cpu().numpy(), g["agg_class_ids"].astype(np.int64))
    assert np.array_equal(out["sample_ids"].cpu().numpy(), g["agg_sample_ids"])
    assert np.array_equal(out["mask_sizes"].cpu().numpy(), g["agg_mask_sizes"])
    for k in ("quaternion", "scales", "z", "xy", "hypothesis", "R", "T", "RT"):
        ref = torch.from_numpy(g["agg_" + k])
        assert out[k].shape == ref.shape, k
        assert helpers.rel_err(out[k], ref) <= helpers.REL_TOL, k
    # dense reference-layout outputs on request
    n = len(g["agg_class_ids"])
    assert out["instance_masks"].shape == (n, h, w) and out["xy_mask"].shape == (n, 2, h, w)
    lab = torch.from_numpy(g["labels"])
    for i in range(n):
        assert torch.equal(out["instance_masks"][i].cpu() != 0, lab[int(g["agg_sample_ids"][i])] == i + 1)
