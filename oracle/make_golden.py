"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Regenerates tests/golden/*.npz: outputs of the reference's OWN Python (imported unmodified from
/root/reference by oracle/ref_import.py) on small seeded scenes, with the fixed pre-sampled pixel pairs
of oracle.port.seeded_idx_source.  Runs only in the build container (the GPU box has no reference);
the fixtures travel.  Usage:  python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fastposecnn_b200 import synthetic as syn  # noqa: E402
from oracle import port, ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
HN = 32
IDX_SEED = 1234
SCENES = {
    # name: (frames, h, w, render seed)
    "g_three_frames": ([[(30, 30, 14, 1), (90, 40, 18, 3), (60, 75, 12, 6)], [(40, 50, 20, 2)], []], 96, 128, 3),
    "g_touching": ([[(30, 40, 12, 5), (52, 40, 12, 2), (100, 60, 10, 4)]], 96, 128, 4),
    "g_tiny_odd": ([[(20, 20, 0.9, 1), (40, 20, 1.0, 2), (75, 45, 15, 3)], [(50, 35, 22, 6)]], 70, 101, 5),
}


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    inv_k = torch.inverse(syn.camera_intrinsics())
    for name, (frames, h, w, seed) in SCENES.items():
        logits = syn.render_heads(frames, h, w, seed=seed)
        cat, agg = ref_import.reference_pose_recover(logits, inv_k, HN, port.seeded_idx_source(IDX_SEED))
        out = {"h": h, "w": w, "hn": HN, "render_seed": seed, "idx_seed": IDX_SEED,
               "frames_json": np.array(repr(frames))}
        # inputs are regenerated from the seed in the tests; store a checksum so a changed generator is caught
        out["input_checksum"] = np.array([float(v.double().sum()) for v in logits.values()])
        out["cat_mask"] = cat["mask"].numpy().astype(np.uint8)
        out["cat_xy"] = cat["xy"].numpy()
        out["cat_quaternion_sum"] = cat["quaternion"].double().sum(dim=(0, 2, 3)).numpy()
        for k in ("class_ids", "sample_ids", "quaternion", "scales", "z", "xy", "hypothesis", "R", "T", "RT"):
            out["agg_" + k] = agg[k].numpy()
        out["agg_mask_sizes"] = agg["instance_masks"].sum(dim=(-2, -1)).numpy().astype(np.int64)
        # instance masks as a label image (bit-exact reconstruction, small)
        lab = torch.zeros((len(frames), h, w), dtype=torch.int32)
        for i in range(agg["instance_masks"].shape[0]):
            lab[int(agg["sample_ids"][i])][agg["instance_masks"][i] != 0] = i + 1
        out["labels"] = lab.numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
        print(name, "N =", agg["class_ids"].shape[0], "centres", agg["xy"].tolist())


def matching_main():
    """tests/golden/matching_*.npz: (preds, gts) -> the reference's own batchwise_get_2d_iou / batchwise_find_matches
    (lib/matching.py:226-325) outputs.  Masks are stored as uint8 (they are 0/1), everything else as produced."""
    import helpers
    ref = ref_import.load()
    for name in helpers.MATCHING_SCENES:
        preds, gts = helpers.matching_scene(name)
        out = {}
        for side, d in (("preds", preds), ("gts", gts)):
            for k, v in d.items():
                out[f"{side}__{k}"] = v.numpy().astype(np.uint8) if k == "instance_masks" else v.numpy()
        out["iou"] = ref.gtf.batchwise_get_2d_iou(gts["instance_masks"], preds["instance_masks"]).numpy()
        m = ref.mg.batchwise_find_matches(preds, gts)
        out["has_matches"] = np.array(m is not None)
        if m is not None:
            for k, v in m.items():
                out[f"matches__{k}"] = v.numpy().astype(np.uint8) if k == "instance_masks" else v.numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"matching_{name}.npz"), **out)
        print("matching", name, "M =", 0 if m is None else m["class_ids"].shape[0])


if __name__ == "__main__":
    if not ref_import.available():
        raise SystemExit("reference sources not found; golden fixtures can only be regenerated in the build container")
    main()
    matching_main()
