"""CPU: the oracle port (oracle/port.py) against the committed golden fixtures, which are outputs of the
reference's own Python (oracle/make_golden.py).  Integers bit-exact; floats to 1e-5 relative (torch-CPU
reductions may vectorise differently on another host CPU)."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

import helpers
from helpers import port, syn

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "g_*.npz")))


def load_golden(path):
    g = np.load(path, allow_pickle=False)
    frames = ast.literal_eval(str(g["frames_json"]))
    return g, frames, int(g["h"]), int(g["w"]), int(g["hn"]), int(g["render_seed"]), int(g["idx_seed"])


def test_fixtures_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_port_matches_reference_golden(path):
    g, frames, h, w, hn, rseed, iseed = load_golden(path)
    logits = syn.render_heads(frames, h, w, seed=rseed)
    chk = np.array([float(v.double().sum()) for v in logits.values()])
    assert np.allclose(chk, g["input_checksum"], rtol=1e-9), "synthetic generator changed: regenerate the fixtures"
    inv_k = torch.inverse(syn.camera_intrinsics())
    cat, agg = port.pose_recover(logits, inv_k, hn, idx_source=port.seeded_idx_source(iseed))
    assert np.array_equal(cat["mask"].numpy().astype(np.uint8), g["cat_mask"])
    assert np.allclose(cat["xy"].numpy(), g["cat_xy"], rtol=0, atol=2e-7)
    assert np.array_equal(agg["class_ids"].numpy().astype(np.int64), g["agg_class_ids"].astype(np.int64))
    assert np.array_equal(agg["sample_ids"].numpy(), g["agg_sample_ids"])
    assert np.array_equal(agg["instance_masks"].sum(dim=(-2, -1)).numpy().astype(np.int64), g["agg_mask_sizes"])
    lab, total = port.label_instances(cat["mask"] != 0)
    assert np.array_equal(lab.numpy(), g["labels"])
    for k in ("quaternion", "scales", "z", "xy", "hypothesis", "R", "T", "RT"):
        ref = torch.from_numpy(g["agg_" + k])
        assert agg[k].shape == ref.shape
        assert helpers.rel_err(agg[k], ref) <= 1e-5, k
