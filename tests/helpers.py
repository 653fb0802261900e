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
