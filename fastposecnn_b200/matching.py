"""Drop-in for the reference's ``lib/matching.py``: pairing of ground-truth and predicted instances.

Same names and semantics as the reference (``batchwise_find_matches`` lib/matching.py:226-325,
``batchwise_find_matches2`` :64-182, ``get_standard_preds`` :184-223, ``stack_and_store_data`` :40-58), including
two things a reader might not expect and the parity tests pin:

* instances are paired by CLASS only -- the reference never looks at ``sample_ids``, so a ground truth can pair with
  a prediction of another frame that overlaps it in image coordinates, and ties go to the first prediction;
* ``batchwise_find_matches2`` indexes ``gts`` with class-local positions for its unmatched rows (:166-172).

What changes is the cost.  The reference builds ``[n_gt, n_pred, h, w]`` logical_and / logical_or volumes per class
(with a host sync per class for ``torch.unique`` / ``torch.where``); here each mask is read once and bit-packed
(``fpc_pack_masks``; predictions coming from ``pose_recover`` are packed straight from its label volume by
``fpc_pack_labels`` and never exist as dense masks), one kernel pairs all classes (``fpc_match_instances``) and a
single 4-byte read returns the number of matches.  IoU values are the same correctly-rounded fp32 quotients, so
pairings are bit-identical.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib

KEYS_TO_STACK = [
    "instance_masks",           # class
    "quaternion", "R",          # rotation
    "scales",                   # size
    "xy", "z", "T",             # translation
    "RT",                       # transformation
]


class MaskSet:
    """Bit planes ``bits [n,h,ceil(w/32)] int32`` + ``meta [n,8] int32`` (include/fpc_b200.h, matching section)."""

    def __init__(self, n: int, h: int, w: int, device):
        self.n, self.h, self.w = n, h, w
        self.device = torch.device(device)
        self.bits = torch.empty((n, h, (w + 31) // 32), dtype=torch.int32, device=self.device)
        self.meta = torch.empty((n, _lib.MASK_META), dtype=torch.int32, device=self.device)

    @property
    def counts(self) -> torch.Tensor:
        return self.meta[:, 0]


def pack_masks(masks: torch.Tensor) -> MaskSet:
    """Dense ``[n,h,w]`` masks (non-zero = set) -> MaskSet.  float32, uint8 and bool are read in place."""
    if not isinstance(masks, torch.Tensor) or masks.dim() != 3:
        raise RuntimeError("masks must be a [n,h,w] tensor")
    if not masks.is_cuda:
        raise RuntimeError("masks must be a CUDA tensor (fastposecnn_b200 has no CPU path)")
    if masks.dtype == torch.bool:
        masks = masks.contiguous().view(torch.uint8)
    elif masks.dtype not in (torch.float32, torch.uint8):
        masks = (masks != 0).view(torch.uint8)
    masks = masks.contiguous()
    n, h, w = masks.shape
    out = MaskSet(n, h, w, masks.device)
    elem = _lib.MASK_F32 if masks.dtype == torch.float32 else _lib.MASK_U8
    with torch.cuda.device(masks.device):
        _lib.check(_lib.lib().fpc_pack_masks(masks.data_ptr(), elem, n, h, w, out.bits.data_ptr(), out.meta.data_ptr(),
                                             _lib.current_stream(masks.device)))
    return out


def pack_labels(labels: torch.Tensor, n: int) -> MaskSet:
    """Label volume ``[b,h,w] int32`` (0 = background, k = instance k-1; ``pose_recover(...)['labels']``) -> MaskSet of
    instances ``0..n-1``."""
    labels = _lib.require_cuda(labels, "labels", torch.int32)
    b, h, w = labels.shape
    out = MaskSet(n, h, w, labels.device)
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().fpc_pack_labels(labels.data_ptr(), b, h, w, n, out.bits.data_ptr(), out.meta.data_ptr(),
                                              _lib.current_stream(labels.device)))
    return out


def mask_iou(a: MaskSet, b: MaskSet) -> torch.Tensor:
    """``[a.n, b.n]`` float32 IoU matrix, bit-identical to gpu_tensor_funcs.batchwise_get_2d_iou of the reference."""
    if (a.h, a.w) != (b.h, b.w):
        raise RuntimeError(f"mask sizes differ: {(a.h, a.w)} vs {(b.h, b.w)}")
    iou = torch.empty((a.n, b.n), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().fpc_mask_iou(a.bits.data_ptr(), a.meta.data_ptr(), a.n, b.bits.data_ptr(), b.meta.data_ptr(), b.n,
                                           a.h, a.w, iou.data_ptr(), _lib.current_stream(a.device)))
    return iou


def _as_mask_set(agg: Dict[str, torch.Tensor]) -> MaskSet:
    if "mask_set" in agg:
        return agg["mask_set"]
    if "instance_masks" in agg:
        return pack_masks(agg["instance_masks"])
    if "labels" in agg:
        return pack_labels(agg["labels"], int(agg["class_ids"].shape[0]))
    raise KeyError("need 'instance_masks' (dense) or 'labels' (label volume) to match instances")


def match_instances(gt_set: MaskSet, gt_class: torch.Tensor, pred_set: MaskSet, pred_class: torch.Tensor):
    """One launch pair for all classes.  Returns device tensors ``(best_pred [n_gt] int32 (-1 = none), best_iou [n_gt],
    pairs [n_gt,2] int32, n_matches [1] int32)``; the first ``n_matches`` rows of ``pairs`` are (gt, pred) indices in
    the reference's output order.  Nothing is read back here."""
    if (gt_set.h, gt_set.w) != (pred_set.h, pred_set.w):
        raise RuntimeError(f"mask sizes differ: {(gt_set.h, gt_set.w)} vs {(pred_set.h, pred_set.w)}")
    dev = gt_set.device
    gt_class = _lib.require_cuda(gt_class.to(torch.int64).contiguous(), "gts['class_ids']")
    pred_class = _lib.require_cuda(pred_class.to(torch.int64).contiguous(), "preds['class_ids']")
    ng, np_ = gt_set.n, pred_set.n
    best_pred = torch.empty((ng,), dtype=torch.int32, device=dev)
    best_iou = torch.empty((ng,), dtype=torch.float32, device=dev)
    pairs = torch.empty((ng, 2), dtype=torch.int32, device=dev)
    n_matches = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fpc_match_instances(
            gt_set.bits.data_ptr(), gt_set.meta.data_ptr(), gt_class.data_ptr(), ng,
            pred_set.bits.data_ptr(), pred_set.meta.data_ptr(), pred_class.data_ptr(), np_, gt_set.h, gt_set.w,
            best_pred.data_ptr(), best_iou.data_ptr(), pairs.data_ptr(), n_matches.data_ptr(), _lib.current_stream(dev)))
    return best_pred, best_iou, pairs, n_matches


def _stack_rows(gts, preds, key: str, gi: torch.Tensor, pj: torch.Tensor) -> torch.Tensor:
    """``stack((gts[key][gi], preds[key][pj]))`` written once, straight into the ``[2,M,...]`` result; the dense masks of
    label-volume predictions are painted for the matched rows only (``fpc_paint_instances``)."""
    g = gts[key]
    out = torch.empty((2, gi.shape[0]) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
    torch.index_select(g, 0, gi, out=out[0])
    if key == "instance_masks" and key not in preds:
        labels = _lib.require_cuda(preds["labels"], "preds['labels']", torch.int32)
        if out.dtype != torch.float32 or tuple(labels.shape[1:]) != tuple(g.shape[1:]):
            raise RuntimeError("gts['instance_masks'] must be float32 with the label volume's [h,w]")
        frames = preds["sample_ids"].to(torch.int64)[pj].contiguous()
        b, h, w = labels.shape
        with torch.cuda.device(labels.device):
            _lib.check(_lib.lib().fpc_paint_instances(labels.data_ptr(), b, h, w, frames.data_ptr(), pj.data_ptr(), int(pj.shape[0]),
                                                      out[1].data_ptr(), _lib.current_stream(labels.device)))
    else:
        p = preds[key]
        if p.dtype != out.dtype or p.device != out.device or tuple(p.shape[1:]) != tuple(g.shape[1:]):
            return torch.stack((out[0], p[pj]))        # let torch raise / promote exactly as the reference's stack would
        torch.index_select(p, 0, pj, out=out[1])
    return out


def stack_and_store_data(pred_gt_matches, gts, preds, gts_instances, preds_instances):
    """lib/matching.py:40-58: for every key of ``gts`` in KEYS_TO_STACK append ``stack((gt rows, pred rows))``."""
    for data_key in gts.keys():
        if data_key in KEYS_TO_STACK:
            pred_gt_matches.setdefault(data_key, []).append(_stack_rows(gts, preds, data_key, gts_instances, preds_instances))


def batchwise_find_matches(preds, gts):
    """lib/matching.py:226-325 -> MatchedData dict (``sample_ids / class_ids / symmetric_ids [M]`` from the ground truth,
    stacked ``[2,M,...]`` tensors for KEYS_TO_STACK) or ``None`` (no predictions, nothing matched, empty inputs)."""
    if not preds or not gts:
        return None
    if preds["class_ids"].shape[0] == 0:
        return None
    gt_set, pred_set = _as_mask_set(gts), _as_mask_set(preds)
    _, _, pairs, n_matches = match_instances(gt_set, gts["class_ids"], pred_set, preds["class_ids"])
    m = int(n_matches.item())                      # the one device->host read
    if m == 0:
        return None
    gi, pj = pairs[:m, 0].long().contiguous(), pairs[:m, 1].long().contiguous()
    out = {"sample_ids": gts["sample_ids"][gi], "class_ids": gts["class_ids"][gi], "symmetric_ids": gts["symmetric_ids"][gi]}
    for key in gts.keys():
        if key in KEYS_TO_STACK:
            out[key] = _stack_rows(gts, preds, key, gi, pj)
    return out


def get_standard_preds(gts, n_of_data):
    """lib/matching.py:184-223: ``n_of_data`` copies of the stand-in prediction (zeros; quaternion (1,0,0,0), RT = I,
    z = 1000), cached on the function like the reference does."""
    if not hasattr(get_standard_preds, "standard_preds"):
        base = {k: torch.zeros_like(gts[k][0]) for k in gts.keys() if k in KEYS_TO_STACK}
        base["quaternion"][0] = 1
        base["RT"] = torch.eye(4, device=gts["RT"].device)
        base["z"][0] = 1000
        get_standard_preds.standard_preds = base
    dev = gts["instance_masks"].device
    return {k: get_standard_preds.standard_preds[k].unsqueeze(0).repeat_interleave(n_of_data, dim=0).to(dev)
            for k in gts.keys() if k in KEYS_TO_STACK}


def batchwise_find_matches2(preds, gts):
    """lib/matching.py:64-182: like batchwise_find_matches, but every ground truth appears -- unmatched ones are paired
    with the standard prediction.  Per class (ascending): matched rows, then unmatched rows."""
    n_gt, n_pred = int(gts["class_ids"].shape[0]), int(preds["class_ids"].shape[0])
    dev = gts["class_ids"].device
    if n_pred:
        best_pred, _, _, _ = match_instances(_as_mask_set(gts), gts["class_ids"], _as_mask_set(preds), preds["class_ids"])
        best = best_pred.cpu().tolist()
    else:
        best = [-1] * n_gt
    gt_cls, pred_cls = gts["class_ids"].cpu().tolist(), set(preds["class_ids"].cpu().tolist())
    out = {"sample_ids": [], "class_ids": [], "symmetric_ids": []}

    def shared(idx, c):
        out["sample_ids"].append(gts["sample_ids"][idx])
        out["symmetric_ids"].append(gts["symmetric_ids"][idx])
        out["class_ids"].append(torch.full((idx.shape[0],), c, dtype=gts["class_ids"].dtype, device=dev))

    for c in sorted(set(gt_cls)):
        members = [i for i in range(n_gt) if gt_cls[i] == c]
        if c not in pred_cls:
            idx = torch.tensor(members, dtype=torch.int64, device=dev)
            shared(idx, c)
            stack_and_store_data(out, gts, get_standard_preds(gts, len(members)), idx, torch.arange(len(members), device=dev))
            continue
        hit = [(i, best[i]) for i in members if best[i] >= 0]
        miss_local = [k for k, i in enumerate(members) if best[i] < 0]
        gi = torch.tensor([i for i, _ in hit], dtype=torch.int64, device=dev)
        shared(gi, c)
        stack_and_store_data(out, gts, preds, gi, torch.tensor([j for _, j in hit], dtype=torch.int64, device=dev))
        if not miss_local:
            continue
        shared(torch.tensor([members[k] for k in miss_local], dtype=torch.int64, device=dev), c)
        local = torch.tensor(miss_local, dtype=torch.int64, device=dev)       # the reference's class-local indexing
        stack_and_store_data(out, gts, get_standard_preds(gts, len(miss_local)), local, torch.arange(len(miss_local), device=dev))
    return {k: torch.cat(v, dim=0 if k in ("sample_ids", "class_ids", "symmetric_ids") else 1) for k, v in out.items()}
