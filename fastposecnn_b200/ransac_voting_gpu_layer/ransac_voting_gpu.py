"""Drop-in for the two PVNet drivers FastPoseCNN's path names
(lib/ransac_voting_gpu_layer/ransac_voting_gpu.py): ``ransac_voting_layer_v3`` (:518-607, the one
HoughVotingLayer calls) and ``ransac_voting_layer`` (v1, :11-98), plus ``b_inv`` (:503-516).

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
from ..aggregation_layer import _pipeline_args, _read_count


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
        with torch.cuda.device(dev):
            a, bufs = _pipeline_args(nprob, h, w, 2, round_hyp_num, nprob, dev, inlier_thresh=float(inlier_thresh),
                                     min_num=int(min_num), max_num=int(max_num),
                                     arith=_lib.ARITH_IEEE if arith is None else int(arith))
            hyp = torch.empty((nprob, round_hyp_num, 2), dtype=torch.float32, device=dev)
            votes = torch.empty((nprob, round_hyp_num), dtype=torch.int32, device=dev)
            a.hyp_out, a.vote_counts_out = hyp.data_ptr(), votes.data_ptr()
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
        table = bufs["table_full"][1:1 + nprob]
        out[:, vi, :] = table[:, _lib.ROW_XY:_lib.ROW_XY + 2]
        if details is not None:
            ti = table.view(torch.int32)
            details.append({"hyp": hyp, "counts": votes, "win_idx": ti[:, _lib.ROW_WIN_IDX].clone(),
                            "win_counts": ti[:, _lib.ROW_WIN_COUNT].clone(), "tn": ti[:, _lib.ROW_TN].clone(),
                            "best_pts": table[:, _lib.ROW_HYP:_lib.ROW_HYP + 2].clone(),
                            "refine_inliers": ti[:, _lib.ROW_REFINE_INL].clone()})
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
