"""Same two callables as the reference's native module ``ransac_voting``
(lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:20-31, :41-55) on CUDA tensors, so the
*unmodified* reference driver ``ransac_voting_layer_v3`` can run on top of these kernels."""
from __future__ import annotations

import torch

from .. import _lib

ARITH = _lib.ARITH_IEEE   # module-level switch: _lib.ARITH_NVCC_FMA reproduces an nvcc build of the reference


def _shapes(direct, coords, second, inliers=None, last=2, what="idxs"):
    """The kernels index direct[(ti*vn+vi)*2], coords[2*ti], second[(hi*vn+vi)*last], inliers[(hi*vn+vi)*tn+ti]: a mismatched or
    undersized buffer would be an out-of-bounds device access (the reference's shape asserts vanish under NDEBUG,
    ransac_voting_kernel.cu:61-65; here they are checked on the host and raise)."""
    if direct.dim() != 3 or direct.shape[2] != 2:
        raise RuntimeError(f"direct must be [tn,vn,2], got {tuple(direct.shape)}")
    tn, vn = int(direct.shape[0]), int(direct.shape[1])
    if tuple(coords.shape) != (tn, 2):
        raise RuntimeError(f"coords must be [{tn},2], got {tuple(coords.shape)}")
    if second.dim() != 3 or second.shape[1] != vn or second.shape[2] != last:
        raise RuntimeError(f"{what} must be [hn,{vn},{last}], got {tuple(second.shape)}")
    hn = int(second.shape[0])
    if inliers is not None and tuple(inliers.shape) != (hn, vn, tn):
        raise RuntimeError(f"inliers must be [{hn},{vn},{tn}], got {tuple(inliers.shape)}")
    return tn, vn, hn


def generate_hypothesis(direct: torch.Tensor, coords: torch.Tensor, idxs: torch.Tensor) -> torch.Tensor:
    """direct [tn,vn,2] f32, coords [tn,2] f32, idxs [hn,vn,2] i32 -> hypo_pts [hn,vn,2] f32."""
    direct = _lib.require_cuda(direct, "direct", torch.float32)
    coords = _lib.require_cuda(coords, "coords", torch.float32)
    idxs = _lib.require_cuda(idxs, "idxs", torch.int32)
    tn, vn, hn = _shapes(direct, coords, idxs)
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
    tn, vn, hn = _shapes(direct, coords, hypo_pts, inliers, what="hypo_pts")
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
    tn, vn, hn = _shapes(direct, coords, idxs)
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
    tn, vn, hn = _shapes(direct, coords, hypo_pts, inliers, last=3, what="hypo_pts")
    with torch.cuda.device(direct.device):
        _lib.check(_lib.lib().fpc_voting_for_hypothesis_vanishing_point(direct.data_ptr(), coords.data_ptr(), hypo_pts.data_ptr(),
                                                                        inliers.data_ptr(), tn, vn, hn, float(inlier_thresh), ARITH,
                                                                        _lib.current_stream(direct.device)))
