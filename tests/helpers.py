"""Shared test helpers: scenes, oracle runs, comparison utilities."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fastposecnn_b200 import synthetic as syn  # noqa: E402
from oracle import port  # noqa: E402

REL_TOL = 1e-4   # BASELINE.json north_star: poses within 1e-4 relative
ABS_FLOOR = 1e-6


def scenes():
    """name -> (frames, h, w).  Covers SURVEY.md section 8c KATs iii-v, vii plus ragged/empty frames."""
    s = {}
    s["three_frames_one_empty"] = ([[(30, 30, 14, 1), (90, 40, 18, 3), (60, 75, 12, 6)], [(40, 50, 20, 2)], []], 96, 128)
    # two touching discs of classes 2 and 5 -> ONE instance whose class id is 2 (min non-zero class)
    s["touching_classes_2_5"] = ([[(30, 40, 12, 5), (52, 40, 12, 2), (100, 60, 10, 4)]], 96, 128)
    # tiny instance (< min_num pixels) next to a regular one; radius 1 disc = 5 px, radius 0.9 = 1 px
    s["tiny_instances"] = ([[(20, 20, 0.9, 1), (40, 20, 1.0, 2), (80, 50, 15, 3)], [(64, 48, 25, 6)]], 96, 128)
    # width not a multiple of 4 (scalar arg-max path) and not a multiple of 32
    s["odd_width"] = ([[(25, 25, 11, 1), (70, 30, 13, 2)], [(45, 40, 16, 4)]], 70, 101)
    # a wide instance crossing several 128-pixel spans and many rows
    s["wide"] = ([[(200, 60, 55, 3), (420, 64, 50, 1)]], 128, 512)
    return s


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max over rows of ||a-b|| / max(||b||, floor)  (vectors compared by norm)."""
    a = a.detach().double().cpu().reshape(a.shape[0], -1) if a.dim() > 0 and a.shape[0] else a.detach().double().cpu().reshape(0, 1)
    b = b.detach().double().cpu().reshape(a.shape[0], -1)
    if a.shape[0] == 0:
        return 0.0
    num = (a - b).norm(dim=1)
    den = b.norm(dim=1).clamp_min(ABS_FLOOR)
    return float((num / den).max())


def run_oracle(logits, hn, seed=1234, num_classes=None, **kw):
    """Oracle (oracle/port.py) on CPU tensors with seeded fixed idxs.  Returns (cat, agg, details)."""
    inv_k = torch.inverse(syn.camera_intrinsics())
    details = []
    cat, agg = port.pose_recover(logits, inv_k, hn, num_of_classes=num_classes,
                                 idx_source=port.seeded_idx_source(seed), details=details, **kw)
    return cat, agg, details


def oracle_tns(agg):
    return [int(v) for v in agg["instance_masks"].sum(dim=(-2, -1)).tolist()]


def matching_scene(name="shifted", h=96, w=128, hn=16):
    """(preds, gts) AggData pairs on CPU for the matching tests: both sides come from the oracle path on two renders of
    related scenes (gt discs shifted / missing / re-classed / duplicated across frames), so every stacked key exists."""
    if name == "shifted":
        pred_frames = [[(30, 30, 14, 1), (90, 40, 18, 3), (60, 75, 12, 6)], [(40, 50, 20, 2), (100, 30, 12, 1)], [(64, 48, 22, 3)]]
        gt_frames = [[(33, 31, 14, 1), (88, 44, 17, 3), (20, 80, 8, 6)], [(42, 50, 19, 2), (100, 70, 12, 1)], [(60, 50, 22, 3), (15, 15, 9, 5)]]
    elif name == "cross_frame_ties":
        # the same class at the same place in every frame: every gt ties with several predictions -> first one wins
        pred_frames = [[(40, 40, 15, 2), (90, 60, 12, 4)] for _ in range(3)]
        gt_frames = [[(40, 40, 15, 2), (90, 60, 12, 4)] for _ in range(2)]
    elif name == "no_overlap":
        pred_frames = [[(30, 30, 10, 1)], [(90, 60, 10, 2)]]
        gt_frames = [[(90, 60, 10, 1)], [(30, 30, 10, 2)]]
    elif name == "class_without_preds":
        pred_frames = [[(30, 30, 12, 1), (80, 30, 12, 1)], [(60, 60, 15, 1)]]
        gt_frames = [[(31, 30, 12, 1), (80, 31, 12, 4)], [(100, 20, 9, 1), (60, 61, 15, 1)]]
    else:
        raise KeyError(name)
    out = []
    for seed, frames in ((21, pred_frames), (22, gt_frames)):
        logits = syn.render_heads(frames, h, w, seed=seed)
        _, agg, _ = run_oracle(logits, hn, seed=seed)
        agg = {k: v for k, v in agg.items()}
        agg["symmetric_ids"] = agg["class_ids"] % 2
        out.append(agg)
    return out[0], out[1]


MATCHING_SCENES = ("shifted", "cross_frame_ties", "no_overlap", "class_without_preds")


def load_matching_golden(name):
    """tests/golden/matching_<name>.npz (oracle/make_golden.py:matching_main) -> (preds, gts, iou, matches or None)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", f"matching_{name}.npz"))
    sides = {"preds": {}, "gts": {}, "matches": {}}
    for key in z.files:
        if "__" in key:
            side, k = key.split("__", 1)
            t = torch.from_numpy(z[key])
            sides[side][k] = t.float() if k == "instance_masks" else t
    return sides["preds"], sides["gts"], torch.from_numpy(z["iou"]), (sides["matches"] if bool(z["has_matches"]) else None)
