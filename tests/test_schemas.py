"""CPU: the dict contracts (fastposecnn_b200/type_hinting.py) against what the oracle -- i.e. the reference -- produces."""
import pytest
import torch

import helpers
from helpers import port, syn
from fastposecnn_b200 import type_hinting as th


def test_oracle_results_satisfy_the_schemas():
    frames, h, w = helpers.scenes()["three_frames_one_empty"]
    logits = syn.render_heads(frames, h, w, seed=1)
    dims = th.validate(logits, "LogitData")
    assert dims == {"b": 3, "C": 7, "h": h, "w": w, "K": 6}
    cat, agg, _ = helpers.run_oracle(logits, 16)
    th.validate(cat, "CategoricalData", b=3, h=h, w=w)
    agg = dict(agg, class_ids=agg["class_ids"].long())
    assert th.validate(agg, "AggData", h=h, w=w)["N"] == 4
    low = syn.render_lowres_heads(frames, h, w, 4)
    th.validate(low, "LowResLogitData", h=h, w=w, S=4, b=3)
    preds, gts, _, matches = helpers.load_matching_golden("shifted")
    assert th.validate(matches, "MatchedData", h=96, w=128)["M"] == 4
    assert set(th.MatchedData.__annotations__) >= {"symmetric_ids", "RT"} and set(th.AggData.__annotations__) >= {"R", "T", "RT"}


def test_violations_are_reported():
    logits = syn.render_heads([[(10, 10, 4, 1)]], 32, 32)
    bad = dict(logits, xy=logits["xy"][:, :-1])
    with pytest.raises(ValueError, match="2K"):
        th.validate(bad, "LogitData")
    with pytest.raises(ValueError, match="int64"):
        th.validate({"mask": torch.zeros(1, 4, 4)}, "CategoricalData")
    with pytest.raises(ValueError, match="dims"):
        th.validate({"quaternion": torch.zeros(3, 4, 1)}, "AggData")
    with pytest.raises(ValueError, match="dim h"):
        th.validate({"mask": torch.zeros(1, 7, 8, 8), "z": torch.zeros(1, 6, 9, 8)}, "LogitData")
