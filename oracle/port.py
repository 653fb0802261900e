"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

CPU restatement (torch-CPU + scipy + the C library in this directory) of
FastPoseCNN's post-network pose-recovery path.  Each function cites the
reference lines it follows (paths relative to
``/root/reference/source_code/FastPoseCNN/``).  The op *sequence* is kept the
same as the reference wherever the float bits depend on it (torch ``norm``,
``sum``, ``matmul``, ``pinverse``, ``inverse``), so that on one machine this
port and the imported reference agree bit for bit; that agreement is what
``tests/test_oracle_vs_reference.py`` and the fixtures under ``tests/golden/``
pin (the reference ships no golden vectors of its own -- SURVEY.md section 4).

Nothing under ``fastposecnn_b200/`` imports this module.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np
import scipy.ndimage
import torch

from . import native

# lib/aggregation_layer.py:43-59 -- centre plane is the 2-D cross, the two outer
# planes are empty: 4-connectivity inside an image, nothing across images.
LABEL_STRUCTURE = np.zeros((3, 3, 3), dtype=bool)
LABEL_STRUCTURE[1] = [[0, 1, 0], [1, 1, 1], [0, 1, 0]]


# ----------------------------------------------------------------------------
# lib/gpu_tensor_funcs.py
# ----------------------------------------------------------------------------

def normalize(data: torch.Tensor, dim: int) -> torch.Tensor:
    """gpu_tensor_funcs.py:37-50 -- L2 normalise with a zero-norm -> 1 guard."""
    n = data.norm(dim=dim, keepdim=True)
    one = torch.tensor(1.0, device=n.device).float()
    return data / torch.where(n != 0, n.float(), one)


def categorical_mask(mask_logits: torch.Tensor) -> torch.Tensor:
    """lib/pose_regressor.py:449 -- argmax over classes of the log-softmax."""
    return torch.argmax(torch.nn.LogSoftmax(dim=1)(mask_logits), dim=1)


def class_compress(num_of_classes: int, cat_mask: torch.Tensor, logits: Dict[str, torch.Tensor]):
    """gpu_tensor_funcs.py:52-99 -- keep, per pixel, the channels of the predicted class."""
    b, h, w = cat_mask.shape
    onehot = torch.zeros((b, num_of_classes, h, w), device=cat_mask.device)
    onehot = onehot.scatter(1, cat_mask.unsqueeze(1), 1)[:, 1:]            # :64-65 (bg dropped)
    sel = onehot.unsqueeze(2).bool()
    out = {}
    for key, head in logits.items():
        if key == "mask":
            continue
        per_class = torch.stack(torch.chunk(head, num_of_classes - 1, dim=1), dim=1)   # :68
        picked = torch.where(sel, per_class.double(), 0.0).float()                     # :78-82
        val = picked.sum(dim=1)                                                        # :85
        if key == "z":
            val = val.squeeze(1)                                                       # :89-90
        elif key in ("quaternion", "xy"):
            val = normalize(val, dim=1)                                                # :93-94
        out[key] = val
    return out


def class_compression(logits: Dict[str, torch.Tensor], num_of_classes: int):
    """lib/pose_regressor.py:445-457 (``Model.class_compression``)."""
    cat_mask = categorical_mask(logits["mask"])
    cat = class_compress(num_of_classes, cat_mask, logits)
    cat["mask"] = cat_mask
    return cat


def quats_2_rotation_matrix(q: torch.Tensor) -> torch.Tensor:
    """gpu_tensor_funcs.py:306-326 (note the final transpose)."""
    a, b, c, d = q.unbind(dim=-1)
    aa, bb, cc, dd = torch.pow(a, 2), torch.pow(b, 2), torch.pow(c, 2), torch.pow(d, 2)
    m = torch.zeros((q.shape[0], 3, 3), device=q.device, dtype=q.dtype)
    m[:, 0, 0] = aa - bb - cc + dd
    m[:, 0, 1] = 2 * (a * b + c * d)
    m[:, 0, 2] = 2 * (a * c - b * d)
    m[:, 1, 0] = 2 * (a * b - c * d)
    m[:, 1, 1] = -aa + bb - cc + dd
    m[:, 1, 2] = 2 * (b * c + a * d)
    m[:, 2, 0] = 2 * (a * c + b * d)
    m[:, 2, 1] = 2 * (b * c - a * d)
    m[:, 2, 2] = -aa - bb + cc + dd
    return m.transpose(-2, -1)


def batchwise_get_RT(q, xys, exp_zs, inv_intrinsics):
    """gpu_tensor_funcs.py:204-235."""
    proj = xys * (exp_zs / 1000)
    homo = torch.vstack([proj.T, exp_zs.T / 1000])
    T = inv_intrinsics @ homo
    n = q.norm(dim=1)
    q = q / torch.where(n > 0, n, torch.ones_like(n)).unsqueeze(1)
    R = quats_2_rotation_matrix(q)
    inv_R = torch.inverse(R)
    bottom = torch.tensor([0, 0, 0, 1], device=q.device, dtype=q.dtype).expand((q.shape[0], 1, 4))
    inv_RT = torch.cat([torch.cat([inv_R, T.T.unsqueeze(-1)], dim=-1), bottom], dim=1)
    RT = torch.inverse(inv_RT)
    return R, T.t(), RT


def samplewise_get_RT(agg, inv_intrinsics):
    """gpu_tensor_funcs.py:237-253."""
    agg["R"], agg["T"], agg["RT"] = batchwise_get_RT(agg["quaternion"], agg["xy"], agg["z"], inv_intrinsics)
    return agg


# ----------------------------------------------------------------------------
# lib/aggregation_layer.py
# ----------------------------------------------------------------------------

def label_instances(fg: torch.Tensor):
    """aggregation_layer.py:160-183 (CPU branch :174-181): scipy.ndimage.label on
    the [b,h,w] volume with LABEL_STRUCTURE.  Returns (int32 labels, total)."""
    lab, total = scipy.ndimage.label(np.asarray(fg), structure=LABEL_STRUCTURE)
    return torch.from_numpy(lab), int(total)


def aggregate(cat: Dict[str, torch.Tensor]):
    """aggregation_layer.py:61-158 (``AggregationLayer.forward``)."""
    cat_mask = cat["mask"]
    labels, total = label_instances(cat_mask != 0)                                      # :76
    b, h, w = cat_mask.shape
    class_ids: List[torch.Tensor] = []
    planes_all: List[torch.Tensor] = []
    sample_ids: List[torch.Tensor] = []
    for bi in range(b):                                                                 # :87
        n_here = (torch.unique(labels[bi]) != 0).sum()                                  # :90
        sample_ids.append(torch.ones((n_here,), dtype=torch.int64) * bi)                # :91-98
        planes = torch.zeros((total + 1, h, w))
        planes = planes.scatter(0, labels[bi].unsqueeze(0).type(torch.int64), 1)[1:]    # :101-102
        planes = planes[planes.sum(dim=(-2, -1)) != 0]                                  # :105
        planes_all.append(planes)
        per_inst = cat_mask[bi].unsqueeze(0) * planes.bool()                            # :111
        try:
            cls = torch.stack([torch.unique(x)[1] for x in torch.unbind(per_inst)])     # :113
        except RuntimeError:
            cls = torch.empty((0,))                                                     # :115
        class_ids.append(cls)
    out = {
        "class_ids": torch.cat(class_ids, dim=0),
        "instance_masks": torch.cat(planes_all, dim=0),
        "sample_ids": torch.cat(sample_ids, dim=0),
    }
    masks = out["instance_masks"]
    for key in ("quaternion", "scales", "xy", "z"):                                     # :125
        per_inst = cat[key][out["sample_ids"]]                                          # :128
        if key == "z":
            per_inst = per_inst.unsqueeze(1)
        masked = masks.unsqueeze(1) * per_inst                                          # :135
        if key == "xy":
            out[key] = masked                                                           # :152-153
            continue
        tot = masked.sum(dim=(-2, -1))                                                  # :139
        size = masks.sum(dim=(-2, -1))
        val = torch.div(tot, size.unsqueeze(1))                                         # :141 (mask_size.T on 1-D is a no-op)
        if key == "z":
            val = torch.exp(val)                                                        # :145
        elif key == "quaternion":
            val = normalize(val, dim=1)                                                 # :149
        out[key] = val
    return out


# ----------------------------------------------------------------------------
# lib/ransac_voting_gpu_layer/ransac_voting_gpu.py
# ----------------------------------------------------------------------------

def b_inv(m: torch.Tensor) -> torch.Tensor:
    """ransac_voting_gpu.py:503-516.  ``torch.solve`` no longer exists (torch>=2),
    so the reference always lands in its ``except RuntimeError`` arm: pinverse."""
    return torch.pinverse(m)


IdxSource = Callable[[int, int, int, int], torch.Tensor]   # (instance, hn, vn, tn) -> int32 [hn,vn,2]


def seeded_idx_source(seed: int = 1234) -> IdxSource:
    """Fixed pre-sampled pixel pairs: one CPU generator, drawn in instance order,
    only for instances that reach the sampling line (tn >= min_num), exactly where
    the reference calls ``random_(0, tn)`` (ransac_voting_gpu.py:552)."""
    g = torch.Generator().manual_seed(seed)

    def draw(_i, hn, vn, tn):
        return torch.randint(0, tn, (hn, vn, 2), generator=g, dtype=torch.int32)
    return draw


def ransac_voting_layer_v3(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                           min_num=5, max_num=30000, *, idx_source: Optional[IdxSource] = None,
                           select_masks: Optional[Dict[int, torch.Tensor]] = None,
                           details: Optional[list] = None, kernels=None):
    """ransac_voting_gpu.py:518-607.

    ``idx_source`` replaces the in-place ``random_`` draw; ``select_masks[i]`` replaces the
    Bernoulli sub-sampling mask of instance ``i`` when it has more than ``max_num`` pixels.
    ``details`` (a list) receives one dict per instance with hypotheses, counts and winner.
    """
    rv = kernels or native.ransac_voting
    if idx_source is None:
        idx_source = seeded_idx_source()
    n, h, w, vn, _ = vertex.shape
    results = []
    for i in range(n):
        cur = mask[i].bool()                                                    # :532
        fg = torch.sum(cur)
        if fg < min_num:                                                        # :536-539
            results.append(torch.zeros([1, vn, 2], dtype=torch.float32))
            if details is not None:
                details.append({"tn": int(fg), "skipped": True})
            continue
        if fg > max_num:                                                        # :542-545
            if select_masks is not None and i in select_masks:
                keep = select_masks[i].bool()
            else:
                u = torch.zeros(cur.shape, dtype=torch.float32).uniform_(0, 1)
                keep = u < (max_num / fg.float())
            cur = cur * keep
        coords = torch.nonzero(cur).float()[:, [1, 0]]                          # :547-548  (x=col, y=row)
        direct = vertex[i].masked_select(cur.unsqueeze(2).unsqueeze(3))         # :549
        direct = direct.view([coords.shape[0], vn, 2])
        tn = coords.shape[0]
        idxs = idx_source(i, round_hyp_num, vn, tn).contiguous()                # :552
        best_ratio = torch.zeros([vn], dtype=torch.float32)
        best_pts = torch.zeros([vn, 2], dtype=torch.float32)
        hyp_num, it = 0, 0
        while True:                                                             # :557-581
            hyp = rv.generate_hypothesis(direct, coords, idxs)
            inl = torch.zeros([round_hyp_num, vn, tn], dtype=torch.uint8)
            rv.voting_for_hypothesis(direct, coords, hyp, inl, inlier_thresh)
            counts = torch.sum(inl, 2)
            win_counts, win_idx = torch.max(counts, 0)
            win_pts = hyp[win_idx, torch.arange(vn)]
            ratio = win_counts.float() / tn
            better = best_ratio < ratio
            best_pts[better, :] = win_pts[better, :]
            best_ratio[better] = ratio[better]
            if details is not None and it == 0:
                details.append({"tn": tn, "skipped": False, "hyp": hyp.clone(), "counts": counts.clone(),
                                "win_idx": win_idx.clone(), "win_counts": win_counts.clone(),
                                "idxs": idxs.clone()})
            hyp_num += round_hyp_num
            it += 1
            lo = torch.min(best_ratio)
            if (1 - (1 - lo ** 2) ** hyp_num) > confidence or it > max_iter:
                break
        normal = torch.zeros_like(direct)                                       # :584-586
        normal[:, :, 0] = direct[:, :, 1]
        normal[:, :, 1] = -direct[:, :, 0]
        final_inl = torch.zeros([1, vn, tn], dtype=torch.uint8)
        rv.voting_for_hypothesis(direct, coords, best_pts.unsqueeze(0).contiguous(), final_inl, inlier_thresh)
        wgt = final_inl.float().squeeze(0)                                      # [vn,tn]
        nrm = normal.permute(1, 0, 2) * wgt.unsqueeze(2)                        # :592-593
        bb = torch.sum(nrm * coords.unsqueeze(0), 2)                            # :595
        ata = torch.matmul(nrm.permute(0, 2, 1), nrm)                           # :596
        atb = torch.sum(nrm * bb.unsqueeze(2), 1)                               # :597
        refined = torch.matmul(b_inv(ata), atb.unsqueeze(2))                    # :598
        if details is not None:
            details[-1].update({"best_pts": best_pts.clone(), "refine_inliers": int(wgt.sum())})
        results.append(refined[None, :, :, 0])
    if not results:
        return torch.empty((0, vn, 2))
    return torch.cat(results)


def ransac_voting_layer(mask, vertex, class_num, round_hyp_num, inlier_thresh=0.999, confidence=0.99,
                        max_iter=20, min_num=5, max_num=30000, *, idx_source: Optional[IdxSource] = None,
                        kernels=None):
    """ransac_voting_gpu.py:11-98 (v1: per image, per class id, no refinement)."""
    rv = kernels or native.ransac_voting
    if idx_source is None:
        idx_source = seeded_idx_source()
    b, h, w, vn, _ = vertex.shape
    per_image = []
    problem = 0
    for bi in range(b):
        per_class = []
        hyp_num = 0                                   # :26 -- NOT reset per class in the reference
        for k in range(class_num - 1):
            cur = mask[bi] == k + 1
            fg = torch.sum(cur)
            if fg < min_num:
                per_class.append(torch.zeros([1, vn, 2], dtype=torch.float32))
                problem += 1
                continue
            if fg > max_num:
                u = torch.zeros(cur.shape, dtype=torch.float32).uniform_(0, 1)
                cur = cur * (u < (max_num / fg.float()))
            coords = torch.nonzero(cur).float()[:, [1, 0]]
            direct = vertex[bi].masked_select(cur.unsqueeze(2).unsqueeze(3)).view([coords.shape[0], vn, 2])
            tn = coords.shape[0]
            idxs = idx_source(problem, round_hyp_num, vn, tn).contiguous()
            problem += 1
            best_ratio = torch.zeros([vn], dtype=torch.float32)
            best_pts = torch.zeros([vn, 2], dtype=torch.float32)
            it = 0
            while True:
                hyp = rv.generate_hypothesis(direct, coords, idxs)
                inl = torch.zeros([round_hyp_num, vn, tn], dtype=torch.uint8)
                rv.voting_for_hypothesis(direct, coords, hyp, inl, inlier_thresh)
                counts = torch.sum(inl, 2)
                win_counts, win_idx = torch.max(counts, 0)
                win_pts = hyp[win_idx, torch.arange(vn)]
                ratio = win_counts.float() / tn
                better = best_ratio < ratio
                best_pts[better, :] = win_pts[better, :]
                best_ratio[better] = ratio[better]
                hyp_num += round_hyp_num
                it += 1
                lo = torch.min(best_ratio)
                if (1 - (1 - lo ** 2) ** hyp_num) > confidence or it > max_iter:
                    break
            per_class.append(best_pts.unsqueeze(0))
        per_image.append(torch.cat(per_class, 0).unsqueeze(0))
    return torch.cat(per_image, 0)


# ----------------------------------------------------------------------------
# lib/hough_voting.py and the Model mixin
# ----------------------------------------------------------------------------

def hough_voting(agg, round_hyp_num: int, **vote_kwargs):
    """hough_voting.py:41-63 (``HoughVotingLayer.forward``)."""
    uv = agg["xy"]
    vertex = uv.permute(0, 2, 3, 1).unsqueeze(3)                               # :51
    out = ransac_voting_layer_v3(mask=agg["instance_masks"], vertex=vertex,
                                 round_hyp_num=round_hyp_num, **vote_kwargs)
    agg.update({"hypothesis": out, "pruned_hypothesis": out, "xy": out.squeeze(1), "xy_mask": uv})
    return agg


def pose_recover(logits, inv_intrinsics, round_hyp_num: int, num_of_classes: Optional[int] = None,
                 **vote_kwargs):
    """lib/pose_regressor.py:445-504, 753-770: class compression -> aggregation ->
    voting -> RT, i.e. the whole hot path on CPU tensors."""
    if num_of_classes is None:
        num_of_classes = logits["mask"].shape[1]
    cat = class_compression(logits, num_of_classes)
    agg = aggregate(cat)
    agg = hough_voting(agg, round_hyp_num, **vote_kwargs)
    agg = samplewise_get_RT(agg, inv_intrinsics)
    return cat, agg


# ---------------------------------------------------------------------------------------------------------------------
# ground-truth <-> prediction matching (lib/matching.py, lib/gpu_tensor_funcs.py:380-409)
# ---------------------------------------------------------------------------------------------------------------------

# lib/matching.py:29-35 -- the keys stack_and_store_data() stacks (the longer local list at :243-250 is never used)
MATCH_STACK_KEYS = ("instance_masks", "quaternion", "R", "scales", "xy", "z", "T", "RT")


def torch_get_2d_iou(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:380-384: |A and B| / |A or B| of two masks (non-zero = set), one float32 scalar."""
    a, b = (t1 != 0), (t2 != 0)
    return torch.true_divide((a & b).sum(), (a | b).sum())


def batchwise_get_2d_iou(masks1: torch.Tensor, masks2: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:386-409: IoU of every mask of ``masks1 [n1,h,w]`` with every mask of
    ``masks2 [n2,h,w]`` -> ``[n1,n2]`` float32.  The reference expands both to ``[n1,n2,h,w]`` and sums
    ``logical_and`` / ``logical_or`` as int64; here the same integers come from one integer matrix
    product (|A and B|) and inclusion-exclusion (|A or B| = |A| + |B| - |A and B|).  int64 / int64 is torch's true
    division: both sides converted to float32, IEEE divide, 0/0 = NaN."""
    hw = masks1.shape[-2] * masks1.shape[-1]
    a = (masks1 != 0).reshape(masks1.shape[0], hw).to(torch.int64)
    b = (masks2 != 0).reshape(masks2.shape[0], hw).to(torch.int64)
    inter = a @ b.t()
    union = a.sum(dim=1, keepdim=True) + b.sum(dim=1).unsqueeze(0) - inter
    return inter / union


def match_pairs(gt_class: torch.Tensor, pred_class: torch.Tensor, iou: torch.Tensor):
    """The pairing rule of lib/matching.py:253-296 for all classes at once.  Returns ``(best_pred [n_gt] (-1 = no
    match), order [M] gt indices in the reference's output order)``: for each ground-truth instance the FIRST
    prediction of the same class (any frame -- the reference never compares sample ids) with the largest IoU, kept
    only if that IoU > 0; output order = ascending class id (torch.unique, :253), then ascending gt index."""
    n_gt = gt_class.shape[0]
    best = torch.full((n_gt,), -1, dtype=torch.int64)
    for i in range(n_gt):
        best_v = 0.0
        for j in range(pred_class.shape[0]):
            if int(pred_class[j]) != int(gt_class[i]):
                continue
            v = float(iou[i, j])
            if v > best_v:            # strict: first maximum wins; NaN (0/0) never wins and never validates
                best_v, best[i] = v, j
    keep = [i for i in range(n_gt) if best[i] >= 0]
    keep.sort(key=lambda i: (int(gt_class[i]), i))
    return best, torch.tensor(keep, dtype=torch.int64)


def batchwise_find_matches(preds, gts):
    """lib/matching.py:226-325.  ``None`` when either side is empty/None, when there are no predictions (:233-234)
    or when nothing matches (:318-319); else a dict with ``sample_ids / class_ids / symmetric_ids [M]`` taken from
    the ground truth and, for every key of ``gts`` in MATCH_STACK_KEYS, ``[2, M, ...]`` = (gt rows, matched pred
    rows)."""
    if not preds or not gts:
        return None
    if preds["class_ids"].shape[0] == 0:
        return None
    iou = batchwise_get_2d_iou(gts["instance_masks"], preds["instance_masks"])
    best, order = match_pairs(gts["class_ids"], preds["class_ids"], iou)
    if order.numel() == 0:
        return None
    out = {"sample_ids": gts["sample_ids"][order], "class_ids": gts["class_ids"][order],
           "symmetric_ids": gts["symmetric_ids"][order]}
    for key in gts.keys():
        if key in MATCH_STACK_KEYS:
            out[key] = torch.stack((gts[key][order], preds[key][best[order]]))
    return out


def standard_pred_row(gts, key: str) -> torch.Tensor:
    """lib/matching.py:184-207: the stand-in prediction for an unmatched ground truth -- zeros with the shape of
    one gt row, except quaternion = (1,0,0,0), RT = eye(4), z[0] = 1000."""
    row = torch.zeros_like(gts[key][0])
    if key == "quaternion":
        row[0] = 1
    elif key == "RT":
        row = torch.eye(4, dtype=gts[key].dtype)
    elif key == "z":
        row[0] = 1000
    return row


def batchwise_find_matches2(preds, gts):
    """lib/matching.py:64-182, the variant that fills unmatched ground truths with the standard prediction.
    Per class (ascending): matched gts first, then the unmatched ones paired with standard rows.  One quirk of the
    reference is kept on purpose: for the unmatched rows of a class that HAS predictions it indexes ``gts`` with the
    class-LOCAL positions (:166-172 pass ``invalid_gt_ids``, not ``gts_class_instances[invalid_gt_ids]``)."""
    iou = batchwise_get_2d_iou(gts["instance_masks"], preds["instance_masks"]) if preds["class_ids"].shape[0] else None
    gt_rows, pred_rows, meta_rows = [], [], []          # pred_rows: -1 = standard prediction
    for c in sorted(set(int(v) for v in gts["class_ids"])):
        members = [i for i in range(gts["class_ids"].shape[0]) if int(gts["class_ids"][i]) == c]
        cand = [j for j in range(preds["class_ids"].shape[0]) if int(preds["class_ids"][j]) == c]
        if not cand:
            for i in members:
                gt_rows.append(i); pred_rows.append(-1); meta_rows.append((i, c))
            continue
        hits, misses = [], []
        for local, i in enumerate(members):
            best_v, best_j = 0.0, -1
            for j in cand:
                v = float(iou[i, j])
                if v > best_v:
                    best_v, best_j = v, j
            (hits if best_j >= 0 else misses).append((local, i, best_j))
        for local, i, j in hits:
            gt_rows.append(i); pred_rows.append(j); meta_rows.append((i, c))
        for local, i, j in misses:
            gt_rows.append(local); pred_rows.append(-1); meta_rows.append((i, c))     # the quirk: gts[local]
    meta_idx = torch.tensor([m[0] for m in meta_rows], dtype=torch.int64)
    out = {"sample_ids": gts["sample_ids"][meta_idx], "symmetric_ids": gts["symmetric_ids"][meta_idx],
           "class_ids": torch.tensor([m[1] for m in meta_rows], dtype=gts["class_ids"].dtype)}
    g_idx = torch.tensor(gt_rows, dtype=torch.int64)
    for key in gts.keys():
        if key in MATCH_STACK_KEYS:
            std = standard_pred_row(gts, key)
            p = torch.stack([preds[key][j] if j >= 0 else std for j in pred_rows])
            out[key] = torch.stack((gts[key][g_idx], p))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the heads' epilogue (lib/pose_regressor.py:633-666, 706-741)
# ---------------------------------------------------------------------------------------------------------------------

def split_xyz(xyz_logits: torch.Tensor):
    """lib/pose_regressor.py:729-732: per class k the translation head emits (x-dir, y-dir, z) = channels 3k..3k+2."""
    c = xyz_logits.shape[1]
    xy = [i for i in range(c) if i % 3 != 0]
    z = [i for i in range(c) if i % 3 == 0]
    return xyz_logits[:, [i - 1 for i in xy]], xyz_logits[:, [i + 2 for i in z]]


def upsample_heads(lowres: Dict[str, torch.Tensor], scale: int = 4) -> Dict[str, torch.Tensor]:
    """What smp's SegmentationHead does after its 1x1 convolution (``nn.UpsamplingBilinear2d(scale_factor=4)`` then the
    identity activation; pose_regressor.py:633-666 with the default ``upsampling=4``): low-resolution LogitData
    ``[b,.,h/4,w/4]`` -> the full-resolution LogitData the path consumes.  torch's own operator is the reference here;
    ``oracle/head_epilogue_ref.c`` restates its arithmetic and is pinned to it."""
    up = torch.nn.UpsamplingBilinear2d(scale_factor=scale)
    return {k: up(v) for k, v in lowres.items()}


def pose_recover_lowres(lowres, inv_intrinsics, round_hyp_num: int, scale: int = 4, **kw):
    """Reference behaviour for low-resolution head outputs: up-sample every head map, then the path."""
    return pose_recover(upsample_heads(lowres, scale), inv_intrinsics, round_hyp_num, **kw)


# ---------------------------------------------------------------------------------------------------------------------
# the other PVNet drivers built on the same two kernels (ransac_voting_gpu.py)
# ---------------------------------------------------------------------------------------------------------------------

def _class_problem(mask_bi, vertex_bi, k, min_num, max_num):
    """The shared prologue of every driver (e.g. ransac_voting_gpu.py:117-141): pixels of class id k+1 in raster order,
    their (x = column, y = row) coordinates and directions.  ``None`` when fewer than ``min_num`` pixels."""
    cur = mask_bi == k + 1
    fg = torch.sum(cur)
    if fg < min_num:
        return None
    if fg > max_num:
        u = torch.zeros(cur.shape, dtype=torch.float32).uniform_(0, 1)
        cur = cur * (u < (max_num / fg.float()))
    coords = torch.nonzero(cur).float()[:, [1, 0]]
    vn = vertex_bi.shape[2]
    direct = vertex_bi.masked_select(cur.unsqueeze(2).unsqueeze(3)).view([coords.shape[0], vn, 2])
    return coords, direct


def _vote_round(rv, direct, coords, idxs, inlier_thresh):
    hyp = rv.generate_hypothesis(direct, coords, idxs)
    inl = torch.zeros([idxs.shape[0], direct.shape[1], direct.shape[0]], dtype=torch.uint8)
    rv.voting_for_hypothesis(direct, coords, hyp, inl, inlier_thresh)
    return hyp, inl


def ransac_voting_layer_v2(mask, vertex, class_num, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                           min_num=5, max_num=30000, refine_iter_num=1, *, idx_source: Optional[IdxSource] = None, kernels=None):
    """ransac_voting_gpu.py:100-216: v1 followed by ``refine_iter_num`` rounds of "vote for the current point, then
    least-squares intersection of the inlier rays" (``pinverse(A) @ b``, A = inlier normals, b = normal . pixel)."""
    rv = kernels or native.ransac_voting
    if idx_source is None:
        idx_source = seeded_idx_source()
    b, h, w, vn, _ = vertex.shape
    out = torch.zeros((b, class_num - 1, vn, 2), dtype=torch.float32)
    problem = 0
    for bi in range(b):
        for k in range(class_num - 1):
            prob = _class_problem(mask[bi], vertex[bi], k, min_num, max_num)
            problem += 1
            if prob is None:
                continue
            coords, direct = prob
            tn = coords.shape[0]
            idxs = idx_source(problem - 1, round_hyp_num, vn, tn).contiguous()
            hyp, inl = _vote_round(rv, direct, coords, idxs, inlier_thresh)       # further passes repeat this one (same idxs)
            win_counts, win_idx = torch.max(torch.sum(inl, 2), 0)
            pts = hyp[win_idx, torch.arange(vn)]
            pts = torch.where((win_counts.float() / tn > 0).unsqueeze(1), pts, torch.zeros_like(pts))   # :166-168, ratio must beat 0
            normal = torch.stack((direct[:, :, 1], -direct[:, :, 0]), dim=2)
            for _ in range(refine_iter_num):
                cur = torch.zeros([1, vn, tn], dtype=torch.uint8)
                rv.voting_for_hypothesis(direct, coords, pts.unsqueeze(0).contiguous(), cur, inlier_thresh)
                refined = []
                for vi in range(vn):
                    sel = cur[0, vi].bool()
                    if int(sel.sum()) == 0:
                        refined.append(torch.zeros([1, 2]))
                        continue
                    a_mat = normal[:, vi, :][sel]
                    rhs = torch.sum(a_mat * coords[sel], 1)
                    refined.append(torch.matmul(torch.pinverse(a_mat), rhs).unsqueeze(0))
                pts = torch.cat(refined, 0)
            out[bi, k] = pts
    return out


def ransac_voting_hypothesis(mask, vertex, round_hyp_num, inlier_thresh=0.999, min_num=5, max_num=30000, *,
                             idx_source: Optional[IdxSource] = None, kernels=None):
    """ransac_voting_gpu.py:218-261: every hypothesis and its vote count for class id 1 of each image ->
    ``[b,hn,vn,2]`` float32, ``[b,hn,vn]`` int64 (zeros / ones for images with fewer than ``min_num`` pixels)."""
    rv = kernels or native.ransac_voting
    if idx_source is None:
        idx_source = seeded_idx_source()
    b, h, w, vn, _ = vertex.shape
    hyps, counts = [], []
    for bi in range(b):
        prob = _class_problem(mask[bi], vertex[bi], 0, min_num, max_num)
        if prob is None:
            hyps.append(torch.zeros([1, round_hyp_num, vn, 2], dtype=torch.float32))
            counts.append(torch.ones([1, round_hyp_num, vn], dtype=torch.int64))
            continue
        coords, direct = prob
        idxs = idx_source(bi, round_hyp_num, vn, coords.shape[0]).contiguous()
        hyp, inl = _vote_round(rv, direct, coords, idxs, inlier_thresh)
        hyps.append(hyp.unsqueeze(0))
        counts.append(torch.sum(inl, 2).unsqueeze(0))
    return torch.cat(hyps, 0), torch.cat(counts, 0)


def _hypothesis_rounds(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh, min_num, max_num, skip_len, idx_source, rv):
    """Shared body of the two distribution estimators (:263-312, :333-383): ceil(min_hyp_num / round_hyp_num) rounds of
    fresh pixel pairs per image -> hypotheses ``[b,vn,H,2]`` and inlier RATIOS (count / voters) ``[b,vn,H]``."""
    b, h, w, vn, _ = vertex.shape
    rounds = int(np.ceil(min_hyp_num / round_hyp_num))
    pts, ratio = [], []
    for bi in range(b):
        prob = _class_problem(mask[bi], vertex[bi], 0, min_num, max_num)
        if prob is None:
            pts.append(torch.zeros([1, skip_len, vn, 2], dtype=torch.float32))
            ratio.append(torch.ones([1, skip_len, vn], dtype=torch.float32))
            continue
        coords, direct = prob
        tn = coords.shape[0]
        p_img, r_img = [], []
        for r in range(rounds):
            idxs = idx_source(bi * rounds + r, round_hyp_num, vn, tn).contiguous()
            hyp, inl = _vote_round(rv, direct, coords, idxs, inlier_thresh)
            p_img.append(hyp)
            r_img.append(torch.sum(inl, 2).float() / float(tn))
        pts.append(torch.cat(p_img, 0).unsqueeze(0))
        ratio.append(torch.cat(r_img, 0).unsqueeze(0))
    return torch.cat(pts, 0).permute(0, 2, 1, 3), torch.cat(ratio, 0).permute(0, 2, 1)


def estimate_voting_distribution(mask, vertex, round_hyp_num=256, min_hyp_num=4096, topk=128, inlier_thresh=0.99, min_num=5,
                                 max_num=30000, *, idx_source: Optional[IdxSource] = None, kernels=None):
    """ransac_voting_gpu.py:263-331: mean and covariance of the hypotheses, weighted by inlier ratio, over the ``topk``
    best-supported ones -> ``[b,vn,2]``, ``[b,vn,2,2]``."""
    pts, ratio = _hypothesis_rounds(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh, min_num, max_num, round_hyp_num,
                                    idx_source or seeded_idx_source(), kernels or native.ransac_voting)
    values, indexes = torch.topk(ratio, topk, dim=2, sorted=False)
    weight = torch.zeros_like(ratio).scatter_(2, indexes, values)
    total = torch.sum(weight, 2)
    mean = torch.sum(weight.unsqueeze(3) * pts, 2) / total.unsqueeze(2)
    diff = pts - mean.unsqueeze(2)
    cov = torch.matmul(diff.transpose(2, 3), diff * weight.unsqueeze(3)) / total.unsqueeze(2).unsqueeze(3)
    return mean, cov


def estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=256, min_hyp_num=4096, topk=128,
                                           inlier_thresh=0.99, min_num=5, max_num=30000, output_hyp=False, *,
                                           idx_source: Optional[IdxSource] = None, kernels=None):
    """ransac_voting_gpu.py:333-406: covariance about a GIVEN mean; hypotheses whose ratio is more than 0.1 below the best
    one get weight 0; the normaliser carries +1e-3.  Returns ``(mean, cov)``."""
    pts, ratio = _hypothesis_rounds(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh, min_num, max_num, min_hyp_num,
                                    idx_source or seeded_idx_source(), kernels or native.ransac_voting)
    ratio = ratio.clone()
    thresh = torch.max(ratio, 2)[0] - 0.1
    ratio[ratio < thresh.unsqueeze(2)] = 0.0
    diff = pts - mean.unsqueeze(2)
    cov = torch.matmul(diff.transpose(2, 3), diff * ratio.unsqueeze(3))
    cov = cov / (torch.sum(ratio, 2).unsqueeze(2).unsqueeze(3) + 1e-3)
    return mean, cov


# ---------------------------------------------------------------------------------------------------------------------
# evaluation maths on matched pairs (lib/gpu_tensor_funcs.py:411-476, 503-578, 611-652, 752-799)
# ---------------------------------------------------------------------------------------------------------------------

def get_raw_quat_distance(q0: torch.Tensor, q1: torch.Tensor) -> torch.Tensor:
    """gpu_tensor_funcs.py:434-455: min(|q0 - q1|, |q0 + q1|) converted with rad2deg (q and -q are one rotation);
    ``tensor([nan])`` for no data."""
    if q0.shape[0] == 0:
        return torch.tensor([float("nan")])
    return torch.rad2deg(torch.minimum((q0 - q1).norm(dim=-1), (q0 + q1).norm(dim=-1)))


def symmetry_rotations() -> torch.Tensor:
    """gpu_tensor_funcs.py:762-780: the 360 unit quaternions (w, 0, y, 0) of 0..359 degrees about the y axis, built in
    float32 exactly like the reference's cached ``rot_q``."""
    half = torch.deg2rad(torch.arange(0, 360).float()) / 2
    s, c = torch.sin(half), torch.cos(half)
    return torch.vstack((c, 0 * s, 1 * s, 0 * s)).T


def get_symmetric_quat_distance(q0: torch.Tensor, q1: torch.Tensor) -> torch.Tensor:
    """gpu_tensor_funcs.py:457-476 + quat_symmetric_tf :752-799: the smallest raw distance between q0 and q1 composed
    with each of the 360 rotations; the composition runs in float64 and is re-normalised by a norm rounded to float32
    (``normalize`` :37-50 casts the norm with ``.float()``), so the result is float64."""
    if q0.shape[0] == 0:
        return torch.tensor([float("nan")])
    r = symmetry_rotations().double().unsqueeze(0)                       # [1,360,4]
    a = q1.double().unsqueeze(1)                                         # [n,1,4]
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = r.unbind(-1)
    prod = torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)
    rot = normalize(prod, dim=-1)
    e_q0 = q0.unsqueeze(1).expand(q0.shape[0], 360, 4)
    return torch.min(get_raw_quat_distance(e_q0, rot), dim=-1).values


def get_quat_distance(q0, q1, symmetric_ids=None):
    """gpu_tensor_funcs.py:411-432: non-symmetric pairs first (raw distance), then symmetric pairs (symmetric distance),
    NaNs dropped -- the output is NOT in the input order."""
    if symmetric_ids is None:
        return get_raw_quat_distance(q0, q1)
    plain, sym = torch.where(symmetric_ids == 0)[0], torch.where(symmetric_ids != 0)[0]
    d = torch.cat((get_raw_quat_distance(q0[plain], q1[plain]), get_symmetric_quat_distance(q0[sym], q1[sym])), dim=0)
    return d[~torch.isnan(d)]


def get_3d_ious(rts_1, rts_2, scales_1, scales_2) -> torch.Tensor:
    """gpu_tensor_funcs.py:503-547: per pair, the 8 corners of each scaled unit cube go through inverse(RT) and are
    de-homogenised; then -- as written in the reference -- ``amax/amin(dim=0)`` reduce over the COORDINATE axis of the
    [3,8] corner matrix, so the "box" compared has 8 extents, one per corner; IoU = prod(overlap) / union of prods."""
    cube = torch.tensor([[1, 1, 1], [1, 1, -1], [-1, 1, 1], [-1, 1, -1], [1, -1, 1], [1, -1, -1], [-1, -1, 1], [-1, -1, -1]],
                        dtype=scales_1.dtype) / 2

    def corners(rt, scales):
        pts = (cube * scales.unsqueeze(0)).T                                       # [3,8]
        hom = torch.vstack([pts, torch.ones((1, 8), dtype=pts.dtype)])
        world = torch.inverse(rt) @ hom
        return world[:-1, :] / world[-1, :]
    out = []
    for i in range(rts_1.shape[0]):
        b1, b2 = corners(rts_1[i], scales_1[i]), corners(rts_2[i], scales_2[i])
        max1, min1, max2, min2 = b1.amax(dim=0), b1.amin(dim=0), b2.amax(dim=0), b2.amin(dim=0)
        ext = torch.minimum(max1, max2) - torch.maximum(min1, min2)
        inter = torch.prod(ext) if torch.amin(ext) >= 0 else torch.zeros(())
        union = torch.prod(max1 - min1) + torch.prod(max2 - min2) - inter
        out.append(inter / union)
    return torch.stack(out)


def from_Ts_get_offset_error(gt_ts, pred_ts) -> torch.Tensor:
    """gpu_tensor_funcs.py:565-567: Euclidean distance of the translations, times 10."""
    return torch.linalg.norm(gt_ts - pred_ts, dim=1) * 10


def calculate_aps(raw_data, metrics_threshold, metrics_operator):
    """gpu_tensor_funcs.py:611-652: per metric and class, the fraction of (non-NaN) values that pass each threshold
    under the metric's comparison operator, plus the mean over classes."""
    aps = {}
    for key, per_class in raw_data.items():
        thresholds, op = metrics_threshold[key], metrics_operator[key]
        aps[key] = {}
        for class_id, values in per_class.items():
            values = values[~torch.isnan(values)]
            hits = op(values.unsqueeze(0), thresholds.unsqueeze(1))
            aps[key][class_id] = torch.sum(hits, dim=1) / values.shape[0]
        aps[key]["mean"] = torch.mean(torch.stack(list(aps[key].values())).float(), dim=0)
    return aps


def _v3_with_extras(mask, vertex, round_hyp_num, inlier_thresh, min_num, max_num, idx_source, rv, batch_fg=False, select_u=None):
    """Shared body of v4 (:678-769) and v5 (:771-866): per image the binary mask's winner, its inlier-weighted normal
    equations, the residual variance about the refined point (v4) and the inlier ratio AT the refined point with the
    hard-coded threshold 0.999 (v5).  Yields (points [vn,2], var [vn], conf [vn]) or None for a skipped image."""
    b, h, w, vn, _ = vertex.shape
    for bi in range(b):
        cur = mask[bi].byte()
        # v6 (:884) counts the foreground of the WHOLE batch (``torch.sum(mask)``), v4 / v5 that of the image
        fg = torch.sum(mask) if batch_fg else torch.sum(cur)
        if fg < min_num:
            yield None
            continue
        if fg > max_num:
            u = select_u[bi] if select_u is not None else torch.zeros(cur.shape, dtype=torch.float32).uniform_(0, 1)
            cur = cur * (u < (max_num / fg.float()))
        sel = cur.bool()
        coords = torch.nonzero(sel).float()[:, [1, 0]]
        direct = vertex[bi].masked_select(sel.unsqueeze(2).unsqueeze(3)).view([coords.shape[0], vn, 2])
        tn = coords.shape[0]
        idxs = idx_source(bi, round_hyp_num, vn, tn).contiguous()
        hyp, inl = _vote_round(rv, direct, coords, idxs, inlier_thresh)
        win_counts, win_idx = torch.max(torch.sum(inl, 2), 0)
        pts = hyp[win_idx, torch.arange(vn)]
        pts = torch.where((win_counts.float() / tn > 0).unsqueeze(1), pts, torch.zeros_like(pts))
        cur_inl = torch.zeros([1, vn, tn], dtype=torch.uint8)
        rv.voting_for_hypothesis(direct, coords, pts.unsqueeze(0).contiguous(), cur_inl, inlier_thresh)
        keep = cur_inl[0].float()                                               # [vn,tn]
        normal = torch.stack((direct[:, :, 1], -direct[:, :, 0]), dim=2).permute(1, 0, 2) * keep.unsqueeze(2)   # [vn,tn,2]
        rhs = torch.sum(normal * coords.unsqueeze(0), 2)                        # [vn,tn]
        ata = torch.matmul(normal.permute(0, 2, 1), normal)
        atb = torch.sum(normal * rhs.unsqueeze(2), 1)
        refined = torch.matmul(b_inv(ata), atb.unsqueeze(2))                    # [vn,2,1]
        residual = torch.matmul(normal, refined)[:, :, 0] - rhs
        var = torch.sum(residual ** 2, 1) / torch.sum(keep, 1)
        conf_inl = torch.zeros([1, vn, tn], dtype=torch.uint8)
        rv.voting_for_hypothesis(direct, coords, refined[:, :, 0].unsqueeze(0).contiguous(), conf_inl, 0.999)
        conf = torch.sum(conf_inl.int(), 2).float()[0] / tn
        yield refined[:, :, 0], var, conf


def ransac_voting_layer_v4(mask, vertex, round_hyp_num, inlier_thresh=0.99, confidence=0.999, max_iter=20, min_num=5,
                           max_num=30000, *, idx_source: Optional[IdxSource] = None, kernels=None):
    """ransac_voting_gpu.py:678-769 -> ``[b,vn,2]`` points, ``[b,vn]`` residual variances (zeros / ones when skipped)."""
    vn = vertex.shape[3]
    pts, var = [], []
    for res in _v3_with_extras(mask, vertex, round_hyp_num, inlier_thresh, min_num, max_num, idx_source or seeded_idx_source(),
                               kernels or native.ransac_voting):
        pts.append(torch.zeros([1, vn, 2]) if res is None else res[0].unsqueeze(0))
        var.append(torch.ones([1, vn]) if res is None else res[1].unsqueeze(0))
    return torch.cat(pts), torch.cat(var)


def ransac_voting_layer_v5(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20, min_num=5,
                           max_num=100, *, idx_source: Optional[IdxSource] = None, kernels=None):
    """ransac_voting_gpu.py:771-866 -> ``[b,vn,2]`` points, ``[b,vn]`` confidences (zeros when skipped)."""
    vn = vertex.shape[3]
    pts, conf = [], []
    for res in _v3_with_extras(mask, vertex, round_hyp_num, inlier_thresh, min_num, max_num, idx_source or seeded_idx_source(),
                               kernels or native.ransac_voting):
        pts.append(torch.zeros([1, vn, 2]) if res is None else res[0].unsqueeze(0))
        conf.append(torch.zeros([1, vn]) if res is None else res[2].unsqueeze(0))
    return torch.cat(pts), torch.cat(conf)


def ransac_voting_layer_v6(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20, min_num=5,
                           max_num=100, *, idx_source: Optional[IdxSource] = None, select_u=None, kernels=None):
    """ransac_voting_gpu.py:868-966 -> ``[b,vn,2]`` points, ``[b,vn]`` confidences.  v5 with one difference: the
    foreground count that decides the skip and the sub-sampling ratio is the sum of the mask over the WHOLE batch (:884)."""
    vn = vertex.shape[3]
    pts, conf = [], []
    for res in _v3_with_extras(mask, vertex, round_hyp_num, inlier_thresh, min_num, max_num, idx_source or seeded_idx_source(),
                               kernels or native.ransac_voting, batch_fg=True, select_u=select_u):
        pts.append(torch.zeros([1, vn, 2]) if res is None else res[0].unsqueeze(0))
        conf.append(torch.zeros([1, vn]) if res is None else res[2].unsqueeze(0))
    return torch.cat(pts), torch.cat(conf)


def ransac_voting_center(mask, vertex, round_hyp_num, inlier_thresh=0.99, confidence=0.999, max_iter=20, min_num=100):
    """ransac_voting_gpu.py:609-676.  The function votes (vn = 1) and then DROPS the result: only images with fewer than
    ``min_num`` pixels append anything -- an all-zero ``[h,w]`` mask -- to the returned list (:624-628, :676)."""
    b, h, w, _ = vertex.shape
    return [torch.zeros([h, w], dtype=torch.float32) for bi in range(b) if torch.sum(mask[bi].byte()) < min_num]


def ransac_motion_voting(mask, vertex):
    """ransac_voting_gpu.py:968-989 -> ``[b,vn,2]``: mean over the mask's pixels of (pixel + its motion vector); zeros
    for an empty mask."""
    b, h, w, vn, _ = vertex.shape
    pts = []
    for bi in range(b):
        cur = mask[bi].byte().bool()
        coords = torch.nonzero(cur).float()
        if coords.shape[0] < 1:
            pts.append(torch.zeros([1, vn, 2], dtype=torch.float32))
            continue
        coords = coords[:, (1, 0)]
        pts.append(torch.mean(vertex[bi][cur] + coords.unsqueeze(1), 0).unsqueeze(0))
    return torch.cat(pts, 0)


def generate_hypothesis(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20, min_num=5,
                        max_num=30000, *, idx_source: Optional[IdxSource] = None, kernels=None):
    """ransac_voting_gpu.py:991-1043 -> all hypotheses ``[b,hn,vn,2]`` and their vote counts ``[b,hn,vn]`` (int64) of the
    binary mask of every image.  (An image with fewer than ``min_num`` pixels makes the reference raise NameError, :1010.)"""
    rv = kernels or native.ransac_voting
    idx_source = idx_source or seeded_idx_source()
    b, h, w, vn, _ = vertex.shape
    hyps, counts = [], []
    for bi in range(b):
        cur = mask[bi].byte()
        fg = torch.sum(cur)
        if fg < min_num:
            raise NameError("name 'batch_win_pts' is not defined")
        if fg > max_num:
            u = torch.zeros(cur.shape, dtype=torch.float32).uniform_(0, 1)
            cur = cur * (u < (max_num / fg.float()))
        sel = cur.bool()
        coords = torch.nonzero(sel).float()[:, [1, 0]]
        direct = vertex[bi].masked_select(sel.unsqueeze(2).unsqueeze(3)).view([coords.shape[0], vn, 2])
        idxs = idx_source(bi, round_hyp_num, vn, coords.shape[0]).contiguous()
        hyp, inl = _vote_round(rv, direct, coords, idxs, inlier_thresh)
        hyps.append(hyp)
        counts.append(torch.sum(inl, 2))
    return torch.stack(hyps), torch.stack(counts)


def from_RTs_get_T_offset_errors(gt_rts, pred_rts) -> torch.Tensor:
    """gpu_tensor_funcs.py:569-609: the camera-frame origin mapped through inverse(RT) for both sides; the reference then
    takes ONE Euclidean norm over the whole [n,3] difference (:557-559 sums every element), times 10."""
    def centres(rts):
        out = []
        for rt in rts:
            w = torch.inverse(rt) @ torch.tensor([[0.0], [0.0], [0.0], [1.0]], dtype=rt.dtype)
            out.append((w[:-1] / w[-1]).flatten())
        return torch.stack(out)
    diff = centres(gt_rts) - centres(pred_rts)
    return torch.sqrt(torch.sum(torch.pow(diff, 2))) * 10


def metric_values(matches_seq, kind: str, threshold=None) -> torch.Tensor:
    """lib/metrics.py restated as a fold over a sequence of matched-pair dicts.  ``kind``: degree_ap / iou_ap / offset_ap
    (:11-50, 91-133, 176-219: passed / seen * 100) or degree_error / iou_accuracy / offset_error (:52-89, 135-174, 221-260:
    value = (value + batch mean) / 2 from 0)."""
    correct, total, running = torch.tensor(0), torch.tensor(0), torch.tensor(0)
    for m in matches_seq:
        if m is None:
            continue
        if kind.startswith("degree") and "quaternion" in m:
            v = get_quat_distance(m["quaternion"][0], m["quaternion"][1], m["symmetric_ids"])
            passed = v < threshold if threshold is not None else None
        elif kind.startswith("iou") and "RT" in m:
            v = get_3d_ious(m["RT"][0], m["RT"][1], m["scales"][0], m["scales"][1])
            passed = v > threshold if threshold is not None else None
            v = v * 100
        elif kind == "offset_ap" and "RT" in m:
            v = from_Ts_get_offset_error(m["T"][0], m["T"][1])
            passed = v < threshold
        elif kind == "offset_error" and "RT" in m:
            v, passed = from_RTs_get_T_offset_errors(m["RT"][0], m["RT"][1]), None
        else:
            continue
        if kind.endswith("_ap"):
            correct, total = correct + torch.sum(passed.int()), total + passed.shape[0]
        else:
            running = (running + torch.mean(v)) / 2
    return (correct.float() / total.float()) * 100 if kind.endswith("_ap") else running
