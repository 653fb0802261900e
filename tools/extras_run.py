"""One pass over the kernels of the 'next' rows (matching, head-epilogue fusion) at cfg2 scale; target of ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastposecnn_b200 as fp  # noqa: E402
from fastposecnn_b200 import matching, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
wl = syn.WORKLOADS["cfg2"]
b = int(os.environ.get("B", wl.batch))
inv_k = torch.inverse(syn.camera_intrinsics()).to(dev)
low = syn.render_lowres_heads([wl.discs()] * b, wl.h, wl.w, 4, wl.num_classes, seed=1000, device=dev)
from fastposecnn_b200.aggregation_layer import materialize_instance_masks  # noqa: E402
from fastposecnn_b200.pose_recovery import get_engine  # noqa: E402
for _ in range(2):
    preds = fp.pose_recover(low, inv_k, wl.hyps, upsample=4)                                # k_argmax_runs_up4, k_gather<3>
    up_q = fp.upsample_bilinear(low["quaternion"], 4)                                      # k_upsample_bilinear (24 of the 67 channels)
eng = get_engine(b, wl.h, wl.w, wl.num_classes, wl.hyps, dev, want_labels=True, upsample=4)
preds["instance_masks"] = materialize_instance_masks(eng.labels, eng.pose_table, int(preds["class_ids"].shape[0]))
n = int(preds["class_ids"].shape[0])
gts = {k: v.clone() for k, v in preds.items() if k not in ("labels", "cat_mask", "xy_mask")}
gts["instance_masks"] = torch.roll(gts["instance_masks"], shifts=(3, -2), dims=(1, 2)).contiguous()
gts["symmetric_ids"] = gts["class_ids"] % 2
preds = {k: v for k, v in preds.items() if k != "xy_mask"}
for _ in range(2):
    gs = matching.pack_masks(gts["instance_masks"])                                        # k_pack_masks_v4
    ps = matching.pack_labels(preds["labels"], n)                                          # k_pack_labels
    iou = matching.mask_iou(gs, ps)                                                        # k_mask_iou
    m = fp.batchwise_find_matches({k: v for k, v in preds.items() if k != "instance_masks"}, gts)   # k_match_best/order, k_paint_instances
print("instances", n, "matches", int(m["class_ids"].shape[0]), "iou diag mean", float(iou.diagonal().mean()))
