"""CPU, build container only: oracle/port.py == the reference's own Python imported unmodified
(oracle/ref_import.py), bit for bit, on seeded scenes.  Skipped where /root/reference is absent."""
import pytest
import torch

import helpers
from helpers import port, syn
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference sources not on this machine")


@pytest.mark.parametrize("name", list(helpers.scenes().keys()))
@pytest.mark.filterwarnings("ignore")
def test_port_equals_reference(name):
    frames, h, w = helpers.scenes()[name]
    logits = syn.render_heads(frames, h, w, seed=11)
    inv_k = torch.inverse(syn.camera_intrinsics())
    hn = 48
    cat_a, agg_a = port.pose_recover(logits, inv_k, hn, idx_source=port.seeded_idx_source(5))
    cat_b, agg_b = ref_import.reference_pose_recover(logits, inv_k, hn, port.seeded_idx_source(5))
    for k in cat_b:
        assert torch.equal(cat_a[k], cat_b[k]), f"cat[{k}]"
    for k in agg_b:
        assert agg_a[k].dtype == agg_b[k].dtype and torch.equal(agg_a[k], agg_b[k]), f"agg[{k}]"


@pytest.mark.filterwarnings("ignore")
def test_v1_equals_reference():
    ref = ref_import.load()
    frames, h, w = helpers.scenes()["three_frames_one_empty"]
    logits = syn.render_heads(frames, h, w, seed=2)
    cat = port.class_compression(logits, 7)
    vertex = cat["xy"].permute(0, 2, 3, 1).unsqueeze(3).contiguous()
    hn = 40
    a = port.ransac_voting_layer(cat["mask"], vertex, 7, hn, idx_source=port.seeded_idx_source(9))
    with ref_import.fixed_idxs(port.seeded_idx_source(9), hn):
        b = ref.rvg.ransac_voting_layer(cat["mask"], vertex, 7, hn)
    assert torch.equal(a, b)
