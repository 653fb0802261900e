"""Same two callables as the reference's native module ``ransac_voting``
(lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:20-31, :41-55) on CUDA tensors, so the
*unmodified* reference driver ``ransac_voting_layer_v3`` can run on top of these kernels."""
from __future__ import annotations

import torch

from .. import _lib

ARITH = _lib.ARITH_IEEE   # module-level switch: _lib.ARITH_NVCC_FMA reproduces an nvcc build of the reference


def generate_hypothesis(direct: torch.Tensor, coords: torch.Tensor, idxs: torch.Tensor) -> torch.Tensor:
    """direct [tn,vn,2] f32, coords [tn,2] f32, idxs [hn,vn,2] i32 -> hypo_pts [hn,vn,2] f32."""
    direct = _lib.require_cuda(direct, "direct", torch.float32)
    coords = _lib.require_cuda(coords, "coords", torch.float32)
    idxs = _lib.require_cuda(idxs, "idxs", torch.int32)
    tn, vn = direct.shape[0], direct.shape[1]
    hn = idxs.shape[0]
    out = torch.empty((hn, vn, 2), dtype=torch.float32, device=direct.device)
    with torch.cuda.device(direct.device):
        _lib.check(_lib.lib().fpc_generate_hypothesis(direct.data_ptr(), coords.data_ptr(), idxs.data_ptr(), out.data_ptr(),
                                                      tn, vn, hn, ARITH, _lib.current_stream(direct.device)))
    return out


def voting_for_hypothesis(direct: torch.Tensor, coords: torch.Tensor, hypo_pts: torch.Tensor, inliers: torch.Tensor,
                          inlier_thresh: float) -> None:
    """Sets inliers[hi,vi,ti] = 1 (uint8, in place, caller pre-zeroes) where the cosine test passes."""
    direct = _lib.require_cuda(direct, "direct", torch.float32)
    coords = _lib.require_cuda(coords, "coords", torch.float32)
    hypo_pts = _lib.require_cuda(hypo_pts, "hypo_pts", torch.float32)
    inliers = _lib.require_cuda(inliers, "inliers", torch.uint8)
    tn, vn = direct.shape[0], direct.shape[1]
    hn = hypo_pts.shape[0]
    with torch.cuda.device(direct.device):
        _lib.check(_lib.lib().fpc_voting_for_hypothesis(direct.data_ptr(), coords.data_ptr(), hypo_pts.data_ptr(),
                                                        inliers.data_ptr(), tn, vn, hn, float(inlier_thresh), ARITH,
                                                        _lib.current_stream(direct.device)))


def generate_hypothesis_vanishing_point(direct: torch.Tensor, coords: torch.Tensor, idxs: torch.Tensor) -> torch.Tensor:
    """src/ransac_voting.cpp:62-73 -- direct [tn,vn,2], coords [tn,2], idxs [hn,vn,2] i32 -> homogeneous hypo_pts
    [hn,vn,3] (all zero where the two rays do not meet)."""
    direct = _lib.require_cuda(direct, "direct", torch.float32)
    coords = _lib.require_cuda(coords, "coords", torch.float32)
    idxs = _lib.require_cuda(idxs, "idxs", torch.int32)
    tn, vn = direct.shape[0], direct.shape[1]
    hn = idxs.shape[0]
    out = torch.empty((hn, vn, 3), dtype=torch.float32, device=direct.device)
    with torch.cuda.device(direct.device):
        _lib.check(_lib.lib().fpc_generate_hypothesis_vanishing_point(direct.data_ptr(), coords.data_ptr(), idxs.data_ptr(),
                                                                      out.data_ptr(), tn, vn, hn, ARITH,
                                                                      _lib.current_stream(direct.device)))
    return out


def voting_for_hypothesis_vanishing_point(direct: torch.Tensor, coords: torch.Tensor, hypo_pts: torch.Tensor, inliers: torch.Tensor,
                                          inlier_thresh: float) -> None:
    """src/ransac_voting.cpp:83-97 -- sets inliers[hi,vi,ti] = 1 (uint8, in place) for hypo_pts [hn,vn,3]."""
    direct = _lib.require_cuda(direct, "direct", torch.float32)
    coords = _lib.require_cuda(coords, "coords", torch.float32)
    hypo_pts = _lib.require_cuda(hypo_pts, "hypo_pts", torch.float32)
    inliers = _lib.require_cuda(inliers, "inliers", torch.uint8)
    tn, vn = direct.shape[0], direct.shape[1]
    hn = hypo_pts.shape[0]
    if hypo_pts.shape[-1] != 3:
        raise RuntimeError("hypo_pts must be [hn,vn,3]")
    with torch.cuda.device(direct.device):
        _lib.check(_lib.lib().fpc_voting_for_hypothesis_vanishing_point(direct.data_ptr(), coords.data_ptr(), hypo_pts.data_ptr(),
                                                                        inliers.data_ptr(), tn, vn, hn, float(inlier_thresh), ARITH,
                                                                        _lib.current_stream(direct.device)))
