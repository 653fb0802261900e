"""Drop-in for the PVNet drivers of lib/ransac_voting_gpu_layer/ransac_voting_gpu.py: ``ransac_voting_layer_v3``
(:518-607, the one HoughVotingLayer calls), ``ransac_voting_layer`` (v1, :11-98), ``b_inv`` (:503-516) and, from the
"next" rows, ``ransac_voting_layer_v2`` (:100), ``ransac_voting_hypothesis`` (:218), ``estimate_voting_distribution``
(:263), ``estimate_voting_distribution_with_mean`` (:333), ``ransac_voting_layer_v4`` (:678, + residual variance) and
``ransac_voting_layer_v5`` (:771, + confidence), ``ransac_voting_layer_v6`` (:868), ``ransac_voting_center`` (:609),
``ransac_motion_voting`` (:968) and the driver ``generate_hypothesis`` (:991).  (``ransac_voting_vanish_point_layer`` :408 references an
undefined ``class_num`` and cannot run in the reference; its two kernels are mirrored in ``ransac_voting.py``.)

Same positional/keyword signatures and result layouts.  Instead of a Python loop with ~40 small
kernels and >=4 host syncs per instance, ALL instances go through one batched launch sequence
(compaction -> hypotheses -> vote counting -> refinement) with a single host read at the end.
Extra keyword-only arguments: ``idxs`` (fixed pre-sampled pixel pairs) and ``select_mask`` /
``select_u`` (explicit sub-sampling for instances with more than ``max_num`` pixels)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from .. import _lib
from ..aggregation_layer import _pipeline_args, _read_count, grow_and_retry


def b_inv(b_mat: torch.Tensor) -> torch.Tensor:
    """Batched inverse with the reference's fallback (:503-516): ``torch.solve`` no longer exists, so the
    reference always takes its pinverse branch; kept for API completeness (tiny library call)."""
    return torch.pinverse(b_mat)


def _vote(nprob, h, w, vn, fmask, imask, nplanes_per_src, match_base, vertex, round_hyp_num, inlier_thresh, min_num,
          max_num, refine, idxs, select_u, details, arith=None):
    dev = vertex.device
    out = torch.zeros((nprob, vn, 2), dtype=torch.float32, device=dev)
    if nprob == 0:
        return out
    if nprob * h * w >= 2 ** 31:
        raise RuntimeError("ransac voting: problems*h*w must be < 2^31 per call; split the batch")
    if idxs is not None:
        idxs = _lib.require_cuda(idxs, "idxs", torch.int32, contiguous=False)
        idxs = idxs.reshape(nprob, round_hyp_num, vn, 2)
    for vi in range(vn):
        def vote_with(_cap, max_rows, max_records, vi=vi):
            with torch.cuda.device(dev):
                a, bufs = _pipeline_args(nprob, h, w, 2, round_hyp_num, nprob, dev, inlier_thresh=float(inlier_thresh),
                                         min_num=int(min_num), max_num=int(max_num), max_rows=max_rows, max_records=max_records,
                                         arith=_lib.ARITH_IEEE if arith is None else int(arith))
                hyp = torch.empty((nprob, round_hyp_num, 2), dtype=torch.float32, device=dev)
                votes = torch.empty((nprob, round_hyp_num), dtype=torch.int32, device=dev)
                a.hyp_out, a.vote_counts_out = hyp.data_ptr(), votes.data_ptr()
                extra = torch.empty((nprob, 4), dtype=torch.float32, device=dev) if details is not None else None
                if extra is not None:
                    a.extra_out = extra.data_ptr()
                keep = []
                if idxs is not None:
                    ix = idxs[:, :, vi, :].contiguous()
                    keep.append(ix)
                    a.idxs = ix.data_ptr()
                if select_u is not None:
                    a.select_u = select_u.data_ptr()
                v = vertex[..., vi, :]                       # [P,h,w,2] strided view of this keypoint
                sN, sH, sW, s2 = v.stride()
                base = v.data_ptr()
                _lib.check(_lib.lib().fpc_vote_dense(ctypes.byref(a), _lib.ptr(fmask), _lib.ptr(imask), nplanes_per_src,
                                                     match_base, base, sN, sH, sW, s2, 1 if refine else 0))
                _read_count(bufs, nprob)
            return bufs, hyp, votes, extra
        # the number of problems is known; the run / record tables (speckled masks) grow on demand
        bufs, hyp, votes, extra = grow_and_retry(vote_with, nprob, fixed=False)
        table = bufs["table_full"][1:1 + nprob]
        out[:, vi, :] = table[:, _lib.ROW_XY:_lib.ROW_XY + 2]
        if details is not None:
            ti = table.view(torch.int32)
            details.append({"hyp": hyp, "counts": votes, "win_idx": ti[:, _lib.ROW_WIN_IDX].clone(),
                            "win_counts": ti[:, _lib.ROW_WIN_COUNT].clone(), "tn": ti[:, _lib.ROW_TN].clone(),
                            "best_pts": table[:, _lib.ROW_HYP:_lib.ROW_HYP + 2].clone(),
                            "refine_inliers": ti[:, _lib.ROW_REFINE_INL].clone(),
                            "residual_var": extra[:, 0], "confidence": extra[:, 1]})
    return out


def _select_u(select_mask, select_u, shape, dev):
    if select_u is not None:
        return _lib.require_cuda(select_u, "select_u", torch.float32).reshape(shape)
    if select_mask is not None:
        # an explicit keep-mask: u = 0 where kept (always below the threshold), 2 where dropped (never)
        sm = _lib.require_cuda(select_mask, "select_mask", None, contiguous=False)
        return torch.where(sm.reshape(shape) != 0, 0.0, 2.0).to(torch.float32).contiguous()
    return None


def ransac_voting_layer_v3(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                           min_num=5, max_num=30000, *, idxs: Optional[torch.Tensor] = None,
                           select_mask: Optional[torch.Tensor] = None, select_u: Optional[torch.Tensor] = None,
                           details: Optional[list] = None, arith: Optional[int] = None):
    """
    :param mask:      [b,h,w]   (any dtype; non-zero = pixel of instance b)
    :param vertex:    [b,h,w,vn,2]  (may be a non-contiguous view)
    :param round_hyp_num: hypotheses per instance
    :return: [b,vn,2] refined centres (x = column, y = row); (0,0) for instances with < min_num pixels

    ``arith``: ``_lib.ARITH_IEEE`` (default; what a CPU build of the reference kernels computes) or
    ``_lib.ARITH_NVCC_FMA`` (what an nvcc build of them computes).
    ``confidence`` / ``max_iter`` are accepted and ignored: the reference never re-samples ``idxs`` inside
    its while loop (:552 is outside it), so every further pass recomputes the first one.
    """
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    vertex = _lib.require_cuda(vertex, "vertex", torch.float32, contiguous=False)
    b, h, w, vn, two = vertex.shape
    if two != 2 or tuple(mask.shape) != (b, h, w):
        raise RuntimeError(f"expected mask [b,h,w] and vertex [b,h,w,vn,2], got {tuple(mask.shape)}, {tuple(vertex.shape)}")
    fmask = mask if (mask.dtype == torch.float32 and mask.is_contiguous()) else mask.to(torch.float32).contiguous()
    su = _select_u(select_mask, select_u, (b, h, w), vertex.device)
    if torch.is_grad_enabled() and vertex.requires_grad and b > 0:
        # training: the refined centres stay differentiable w.r.t. the direction field (autograd.RansacV3Fn)
        from ..autograd import RansacV3Fn
        if su is not None or bool((fmask.flatten(1) != 0).sum(dim=1).max() > max_num):
            raise NotImplementedError("ransac_voting_layer_v3 backward: instances sub-sampled to max_num are not differentiable here")

        def run():
            det = details if details is not None else []
            first = len(det)
            pts = _vote(b, h, w, vn, fmask, None, 1, 0, vertex.detach(), int(round_hyp_num), inlier_thresh, min_num, max_num, True,
                        idxs, None, det, arith)
            return pts, det[first:]
        return RansacV3Fn.apply(run, fmask, vertex, inlier_thresh, _lib.ARITH_IEEE if arith is None else int(arith))
    return _vote(b, h, w, vn, fmask, None, 1, 0, vertex, int(round_hyp_num), inlier_thresh, min_num, max_num, True,
                 idxs, su, details, arith)


def ransac_voting_layer(mask, vertex, class_num, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                        min_num=5, max_num=30000, *, idxs: Optional[torch.Tensor] = None,
                        select_u: Optional[torch.Tensor] = None, details: Optional[list] = None):
    """
    :param mask:      [b,h,w] class ids
    :param vertex:    [b,h,w,vn,2]
    :return: [b,class_num-1,vn,2] winning hypothesis per (image, class) -- v1 has no refinement step

    ``idxs``: [b*(class_num-1),hn,vn,2]; ``select_u``: [b*(class_num-1),h,w].
    """
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    vertex = _lib.require_cuda(vertex, "vertex", torch.float32, contiguous=False)
    b, h, w, vn, two = vertex.shape
    if two != 2 or tuple(mask.shape) != (b, h, w):
        raise RuntimeError(f"expected mask [b,h,w] and vertex [b,h,w,vn,2], got {tuple(mask.shape)}, {tuple(vertex.shape)}")
    k = int(class_num) - 1
    imask = mask.to(torch.int32).contiguous()
    su = _select_u(None, select_u, (b * k, h, w), vertex.device)
    out = _vote(b * k, h, w, vn, None, imask, k, 1, vertex, int(round_hyp_num), inlier_thresh, min_num, max_num, False,
                idxs, su, details)
    return out.reshape(b, k, vn, 2)


# ---------------------------------------------------------------------------------------------------------------------
# The other PVNet drivers built on the same two kernels.  Same signatures and result
# layouts as the reference; each is ONE batched launch sequence per keypoint instead of a Python loop per image/class.
# ---------------------------------------------------------------------------------------------------------------------

def ransac_voting_layer_v2(mask, vertex, class_num, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                           min_num=5, max_num=30000, refine_iter_num=1, *, idxs: Optional[torch.Tensor] = None,
                           select_u: Optional[torch.Tensor] = None, details: Optional[list] = None):
    """ransac_voting_gpu.py:100-216 -- v1 plus the least-squares refinement over the winner's inliers.

    :param mask:      [b,h,w] class ids
    :param vertex:    [b,h,w,vn,2]
    :return: [b,class_num-1,vn,2]

    The reference solves ``pinverse(A) @ b`` over the inlier normals; the kernel solves the same least-squares problem
    through its 2x2 normal equations in FP64 (<= 1e-4 relative).  ``refine_iter_num`` 0 (= v1) or 1."""
    if refine_iter_num not in (0, 1):
        raise NotImplementedError("ransac_voting_layer_v2: refine_iter_num must be 0 or 1")
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    vertex = _lib.require_cuda(vertex, "vertex", torch.float32, contiguous=False)
    b, h, w, vn, two = vertex.shape
    if two != 2 or tuple(mask.shape) != (b, h, w):
        raise RuntimeError(f"expected mask [b,h,w] and vertex [b,h,w,vn,2], got {tuple(mask.shape)}, {tuple(vertex.shape)}")
    k = int(class_num) - 1
    imask = mask.to(torch.int32).contiguous()
    su = _select_u(None, select_u, (b * k, h, w), vertex.device)
    out = _vote(b * k, h, w, vn, None, imask, k, 1, vertex, int(round_hyp_num), inlier_thresh, min_num, max_num,
                bool(refine_iter_num), idxs, su, details)
    return out.reshape(b, k, vn, 2)


def _binary_mask_vote(mask, vertex, round_hyp_num, inlier_thresh, min_num, max_num, idxs, select_mask, select_u):
    """v3's batched launch sequence with the per-instance details (winner, refinement, v4 / v5 extras) kept."""
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    vertex = _lib.require_cuda(vertex, "vertex", torch.float32, contiguous=False)
    b, h, w, vn, two = vertex.shape
    if two != 2 or tuple(mask.shape) != (b, h, w):
        raise RuntimeError(f"expected mask [b,h,w] and vertex [b,h,w,vn,2], got {tuple(mask.shape)}, {tuple(vertex.shape)}")
    fmask = mask if (mask.dtype == torch.float32 and mask.is_contiguous()) else mask.to(torch.float32).contiguous()
    su = _select_u(select_mask, select_u, (b, h, w), vertex.device)
    det: list = []
    pts = _vote(b, h, w, vn, fmask, None, 1, 0, vertex, int(round_hyp_num), inlier_thresh, min_num, max_num, True, idxs, su, det)
    return pts, det


def ransac_voting_layer_v4(mask, vertex, round_hyp_num, inlier_thresh=0.99, confidence=0.999, max_iter=20, min_num=5,
                           max_num=30000, *, idxs: Optional[torch.Tensor] = None, select_mask: Optional[torch.Tensor] = None,
                           select_u: Optional[torch.Tensor] = None):
    """ransac_voting_gpu.py:678-769 -- v3 plus the variance of the inlier rays' residuals about the refined point.

    :return: ``[b,vn,2]`` refined points, ``[b,vn]`` variances (ones for instances with < ``min_num`` pixels)."""
    pts, det = _binary_mask_vote(mask, vertex, round_hyp_num, inlier_thresh, min_num, max_num, idxs, select_mask, select_u)
    return pts, torch.stack([d["residual_var"] for d in det], dim=1)


def ransac_voting_layer_v5(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20, min_num=5,
                           max_num=100, *, idxs: Optional[torch.Tensor] = None, select_mask: Optional[torch.Tensor] = None,
                           select_u: Optional[torch.Tensor] = None):
    """ransac_voting_gpu.py:771-866 -- v3 plus a confidence: the fraction of the voters that vote for the refined point at
    threshold 0.999.  Note the reference's default ``max_num=100``: larger instances are randomly sub-sampled.

    :return: ``[b,vn,2]`` refined points, ``[b,vn]`` confidences (zeros for instances with < ``min_num`` pixels)."""
    pts, det = _binary_mask_vote(mask, vertex, round_hyp_num, inlier_thresh, min_num, max_num, idxs, select_mask, select_u)
    return pts, torch.stack([d["confidence"] for d in det], dim=1)


def ransac_voting_layer_v6(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20, min_num=5,
                           max_num=100, *, idxs: Optional[torch.Tensor] = None, select_u: Optional[torch.Tensor] = None):
    """ransac_voting_gpu.py:868-966 -- v5 except that the foreground count deciding the skip (``< min_num``) and the
    sub-sampling ratio (``max_num / count``) is ``torch.sum(mask)`` over the WHOLE batch (:884), and a pixel belongs to the
    mask iff ``mask.byte()`` is non-zero (:885).  ``select_u``: the ``[b,h,w]`` uniforms of the sub-sampling (drawn on the
    device when omitted).  An image left without a single pixel gets zeros (the reference raises there).

    :return: ``[b,vn,2]`` refined points, ``[b,vn]`` confidences."""
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    vertex = _lib.require_cuda(vertex, "vertex", torch.float32, contiguous=False)
    b, h, w, vn, two = vertex.shape
    if two != 2 or tuple(mask.shape) != (b, h, w):
        raise RuntimeError(f"expected mask [b,h,w] and vertex [b,h,w,vn,2], got {tuple(mask.shape)}, {tuple(vertex.shape)}")
    total = torch.sum(mask)
    if bool(total < min_num):
        return (torch.zeros((b, vn, 2), dtype=torch.float32, device=vertex.device),
                torch.zeros((b, vn), dtype=torch.float32, device=vertex.device))
    fmask = (mask.byte() != 0).to(torch.float32)
    if bool(total > max_num):
        u = _lib.require_cuda(select_u, "select_u", torch.float32).reshape(b, h, w) if select_u is not None else \
            torch.rand((b, h, w), dtype=torch.float32, device=vertex.device)
        fmask = fmask * (u < (max_num / total.float())).to(torch.float32)
    det: list = []
    # the decisions were taken above for the whole batch: every image with a pixel left votes with all of its pixels
    pts = _vote(b, h, w, vn, fmask.contiguous(), None, 1, 0, vertex, int(round_hyp_num), inlier_thresh, 1, 2 ** 31 - 1, True,
                idxs, None, det)
    return pts, torch.stack([d["confidence"] for d in det], dim=1)


def ransac_voting_center(mask, vertex, round_hyp_num, inlier_thresh=0.99, confidence=0.999, max_iter=20, min_num=100):
    """ransac_voting_gpu.py:609-676.  The reference votes for one centre per image and then drops the result: the list it
    returns holds one all-zero ``[h,w]`` mask per image with FEWER than ``min_num`` pixels and nothing for the others
    (:624-628, :676).  Reproduced as is -- there is nothing to compute."""
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    b, h, w = mask.shape
    few = (mask.byte().flatten(1).sum(dim=1) < min_num).tolist()
    return [torch.zeros((h, w), dtype=torch.float32, device=mask.device) for bi in range(b) if few[bi]]


def _motion_mean(mask, vertex):
    """[b,vn,2] mean over the mask's pixels of (pixel (x, y) + its vector); zeros for an empty mask.  Device-agnostic torch."""
    b, h, w, vn, _ = vertex.shape
    m = (mask.byte() != 0)
    n = m.flatten(1).sum(dim=1).to(torch.float32)                                   # [b]
    ys = torch.arange(h, dtype=torch.float32, device=vertex.device).view(1, h, 1)
    xs = torch.arange(w, dtype=torch.float32, device=vertex.device).view(1, 1, w)
    zero = torch.zeros((), dtype=torch.float32, device=vertex.device)
    centre = torch.stack((torch.where(m, xs, zero).flatten(1).sum(dim=1), torch.where(m, ys, zero).flatten(1).sum(dim=1)), dim=1)
    # select, do not multiply: values outside the mask (possibly NaN / inf) must not reach the sums -- the reference indexes
    total = torch.where(m.view(b, h, w, 1, 1), vertex, zero).flatten(1, 2).sum(dim=1) + centre.unsqueeze(1)    # [b,vn,2]
    return torch.where((n > 0).view(b, 1, 1), total / n.clamp_min(1).view(b, 1, 1), torch.zeros_like(total))


def ransac_motion_voting(mask, vertex):
    """ransac_voting_gpu.py:968-989 -> ``[b,vn,2]``: mean over the mask's pixels of (pixel (x, y) + its vector); zeros for an
    empty mask.  Two batched masked sums instead of the per-image loop (sums in another order: <= 1e-5 relative)."""
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    vertex = _lib.require_cuda(vertex, "vertex", torch.float32, contiguous=False)
    return _motion_mean(mask, vertex)


def generate_hypothesis(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20, min_num=5,
                        max_num=30000, *, idxs: Optional[torch.Tensor] = None, select_mask: Optional[torch.Tensor] = None,
                        select_u: Optional[torch.Tensor] = None):
    """ransac_voting_gpu.py:991-1043 -> all hypotheses ``[b,hn,vn,2]`` and their vote counts ``[b,hn,vn]`` (int64) over
    the binary mask of every image, no winner and no refinement.  An image with fewer than ``min_num`` pixels makes the
    reference fail (NameError at :1010); here it raises RuntimeError."""
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    vertex = _lib.require_cuda(vertex, "vertex", torch.float32, contiguous=False)
    b, h, w, vn, two = vertex.shape
    if two != 2 or tuple(mask.shape) != (b, h, w):
        raise RuntimeError(f"expected mask [b,h,w] and vertex [b,h,w,vn,2], got {tuple(mask.shape)}, {tuple(vertex.shape)}")
    fmask = (mask.byte() != 0).to(torch.float32).contiguous()
    if bool((fmask.flatten(1).sum(dim=1) < min_num).any()):
        raise RuntimeError("generate_hypothesis: an image has fewer than min_num mask pixels (the reference raises NameError here)")
    su = _select_u(select_mask, select_u, (b, h, w), vertex.device)
    det: list = []
    _vote(b, h, w, vn, fmask, None, 1, 0, vertex, int(round_hyp_num), inlier_thresh, min_num, max_num, False, idxs, su, det)
    return torch.stack([d["hyp"] for d in det], dim=2), torch.stack([d["counts"] for d in det], dim=2).long()


def _class1_hypotheses(mask, vertex, hn, inlier_thresh, min_num, max_num, idxs, select_u):
    """All hypotheses and vote counts of the class-id-1 pixels of every image: ``hyp [b,hn,vn,2]``, ``counts [b,hn,vn]``
    int32, ``tn [b]`` voters per image (0 = fewer than ``min_num`` pixels: nothing was computed)."""
    mask = _lib.require_cuda(mask, "mask", None, contiguous=False)
    vertex = _lib.require_cuda(vertex, "vertex", torch.float32, contiguous=False)
    b, h, w, vn, two = vertex.shape
    if two != 2 or tuple(mask.shape) != (b, h, w):
        raise RuntimeError(f"expected mask [b,h,w] and vertex [b,h,w,vn,2], got {tuple(mask.shape)}, {tuple(vertex.shape)}")
    imask = mask.to(torch.int32).contiguous()
    su = _select_u(None, select_u, (b, h, w), vertex.device)
    det: list = []
    _vote(b, h, w, vn, None, imask, 1, 1, vertex, int(hn), inlier_thresh, min_num, max_num, False, idxs, su, det)
    hyp = torch.stack([d["hyp"] for d in det], dim=2)          # [b,hn,vn,2]
    counts = torch.stack([d["counts"] for d in det], dim=2)    # [b,hn,vn]
    return hyp, counts, det[0]["tn"]


def ransac_voting_hypothesis(mask, vertex, round_hyp_num, inlier_thresh=0.999, min_num=5, max_num=30000, *,
                             idxs: Optional[torch.Tensor] = None, select_u: Optional[torch.Tensor] = None):
    """ransac_voting_gpu.py:218-261 -> ``[b,hn,vn,2]`` hypotheses and ``[b,hn,vn]`` int64 vote counts of class id 1
    (zeros / ones for images with fewer than ``min_num`` such pixels).  ``idxs``: ``[b,hn,vn,2]``."""
    hyp, counts, tn = _class1_hypotheses(mask, vertex, round_hyp_num, inlier_thresh, min_num, max_num, idxs, select_u)
    live = (tn > 0).view(-1, 1, 1)
    return torch.where(live.unsqueeze(3), hyp, torch.zeros_like(hyp)), torch.where(live, counts.long(), torch.ones_like(counts).long())


def _hypothesis_rounds(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh, min_num, max_num, idxs, select_u):
    """``ceil(min_hyp_num / round_hyp_num)`` rounds in one launch -> points ``[b,vn,H,2]`` and inlier ratios ``[b,vn,H]``
    (count / voters); images without enough pixels get zero points and ratio 1 like the reference's skip branch."""
    rounds = -(-int(min_hyp_num) // int(round_hyp_num))
    hyp, counts, tn = _class1_hypotheses(mask, vertex, rounds * int(round_hyp_num), inlier_thresh, min_num, max_num, idxs, select_u)
    live = (tn > 0).view(-1, 1, 1)
    ratio = torch.where(live, counts.float() / tn.clamp_min(1).float().view(-1, 1, 1), torch.ones_like(counts, dtype=torch.float32))
    pts = torch.where(live.unsqueeze(3), hyp, torch.zeros_like(hyp))
    return pts.permute(0, 2, 1, 3), ratio.permute(0, 2, 1)


def estimate_voting_distribution(mask, vertex, round_hyp_num=256, min_hyp_num=4096, topk=128, inlier_thresh=0.99, min_num=5,
                                 max_num=30000, *, idxs: Optional[torch.Tensor] = None, select_u: Optional[torch.Tensor] = None):
    """ransac_voting_gpu.py:263-331 -> ``mean [b,vn,2]``, ``cov [b,vn,2,2]`` of the hypotheses weighted by inlier ratio
    over the ``topk`` best-supported ones.  ``idxs``: ``[b, rounds*round_hyp_num, vn, 2]``, rounds concatenated."""
    pts, ratio = _hypothesis_rounds(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh, min_num, max_num, idxs, select_u)
    values, indexes = torch.topk(ratio, topk, dim=2, sorted=False)
    weight = torch.zeros_like(ratio).scatter_(2, indexes, values)
    total = torch.sum(weight, 2)
    mean = torch.sum(weight.unsqueeze(3) * pts, 2) / total.unsqueeze(2)
    diff = pts - mean.unsqueeze(2)
    cov = torch.matmul(diff.transpose(2, 3), diff * weight.unsqueeze(3)) / total.unsqueeze(2).unsqueeze(3)
    return mean, cov


def estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=256, min_hyp_num=4096, topk=128,
                                           inlier_thresh=0.99, min_num=5, max_num=30000, output_hyp=False, *,
                                           idxs: Optional[torch.Tensor] = None, select_u: Optional[torch.Tensor] = None):
    """ransac_voting_gpu.py:333-406 -> ``(mean, cov)``: covariance about the given mean with ratios more than 0.1 below
    the best one zeroed and +1e-3 in the normaliser."""
    pts, ratio = _hypothesis_rounds(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh, min_num, max_num, idxs, select_u)
    thresh = torch.max(ratio, 2)[0] - 0.1
    ratio = torch.where(ratio < thresh.unsqueeze(2), torch.zeros_like(ratio), ratio)
    diff = pts - mean.unsqueeze(2)
    cov = torch.matmul(diff.transpose(2, 3), diff * ratio.unsqueeze(3))
    cov = cov / (torch.sum(ratio, 2).unsqueeze(2).unsqueeze(3) + 1e-3)
    return mean, cov
