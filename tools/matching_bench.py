"""Times the matching row at cfg2 scale on one GPU: pack (dense and label-volume),
IoU matrix, pairing; next to the reference's algorithm (expand to [n1,n2,h,w]) written in torch on the same GPU."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastposecnn_b200 as fp  # noqa: E402
from fastposecnn_b200 import matching, synthetic as syn  # noqa: E402


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def reference_style(gts, preds):
    """lib/matching.py:253-296 as written (per class, dense expand) -- for timing only."""
    n = 0
    for c in torch.unique(gts["class_ids"]):
        gi = torch.where(gts["class_ids"] == c)[0]
        pi = torch.where(preds["class_ids"] == c)[0]
        if gi.shape[0] == 0 or pi.shape[0] == 0:
            continue
        m1, m2 = gts["instance_masks"][gi], preds["instance_masks"][pi]
        n1, h, w = m1.shape
        n2 = m2.shape[0]
        e1 = m1.unsqueeze(1).expand((n1, n2, h, w))
        e2 = m2.expand((n1, n2, h, w))
        iou = torch.logical_and(e1, e2).sum(dim=(2, 3)) / torch.logical_or(e1, e2).sum(dim=(2, 3))
        v, j = torch.max(iou, dim=1)
        n += int((v > 0).sum())
    return n


def main():
    dev = torch.device("cuda:0")
    b = int(os.environ.get("B", 32))
    wl = syn.WORKLOADS["cfg2"]
    logits = syn.render_workload(wl, b, seed=0, device=dev)
    inv_k = torch.inverse(syn.camera_intrinsics()).to(dev)
    preds = fp.pose_recover(logits, inv_k, wl.hyps, materialize_dense=True)
    n = int(preds["class_ids"].shape[0])
    gts = {k: v.clone() for k, v in preds.items() if k not in ("labels", "cat_mask")}
    gts["instance_masks"] = torch.roll(gts["instance_masks"], shifts=(3, -2), dims=(1, 2)).contiguous()
    gts["symmetric_ids"] = gts["class_ids"] % 2
    h, w = gts["instance_masks"].shape[1:]
    out = {"n_gt": n, "n_pred": n, "h": h, "w": w, "frames": b}
    out["pack_dense_ms"] = timed(lambda: matching.pack_masks(gts["instance_masks"]))
    out["pack_dense_GBps"] = n * h * w * 4 / out["pack_dense_ms"] / 1e6
    out["pack_labels_ms"] = timed(lambda: matching.pack_labels(preds["labels"], n))
    gs, ps = matching.pack_masks(gts["instance_masks"]), matching.pack_labels(preds["labels"], n)
    assert torch.equal(ps.bits, matching.pack_masks(preds["instance_masks"]).bits)
    out["iou_matrix_ms"] = timed(lambda: matching.mask_iou(gs, ps))
    out["pairing_ms"] = timed(lambda: matching.match_instances(gs, gts["class_ids"], ps, preds["class_ids"]))
    sparse_preds = {k: v for k, v in preds.items() if k != "instance_masks"}
    small = lambda d: {k: v for k, v in d.items() if k != "xy_mask"}
    res = {}
    out["find_matches_dense_ms"] = timed(lambda: res.__setitem__("d", fp.batchwise_find_matches(small(preds), gts)), iters=5, warm=2)
    out["find_matches_labels_ms"] = timed(lambda: res.__setitem__("s", fp.batchwise_find_matches(small(sparse_preds), gts)), iters=5, warm=2)
    out["matches"] = int(res["d"]["class_ids"].shape[0])
    for k in res["d"]:
        assert torch.equal(res["d"][k], res["s"][k]), k
    if os.environ.get("REF", "1") == "1":
        out["reference_style_torch_gpu_ms"] = timed(lambda: res.__setitem__("r", reference_style(gts, preds)), iters=2, warm=1)
        assert res["r"] == out["matches"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
