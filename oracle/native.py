"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

ctypes front-end of ``oracle/libransac_voting_ref.so`` presenting the same two
callables as the reference's pybind module ``ransac_voting``
(lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:20-31, :41-55, :102-107),
but operating on **CPU** torch tensors.  It stands in for that module both in
``oracle/port.py`` and underneath the *unmodified* reference Python driver
(``oracle/ref_import.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libransac_voting_ref.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc, no GPU needed)."""
    srcs = [os.path.join(_HERE, f) for f in ("ransac_voting_ref.c", "head_epilogue_ref.c", "vanishing_point_ref.c")]
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp, ip, up = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        for name in ("fpc_ref_generate_hypothesis", "fpc_ref_generate_hypothesis_fma"):
            f = getattr(L, name)
            f.argtypes = [fp, fp, ip, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
            f.restype = None
        for name in ("fpc_ref_voting_for_hypothesis", "fpc_ref_voting_for_hypothesis_fma"):
            f = getattr(L, name)
            f.argtypes = [fp, fp, fp, up, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]
            f.restype = None
        for name in ("fpc_ref_generate_hypothesis_vp", "fpc_ref_generate_hypothesis_vp_fma"):
            f = getattr(L, name)
            f.argtypes = [fp, fp, ip, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
            f.restype = None
        for name in ("fpc_ref_voting_for_hypothesis_vp", "fpc_ref_voting_for_hypothesis_vp_fma"):
            f = getattr(L, name)
            f.argtypes = [fp, fp, fp, up, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]
            f.restype = None
        L.fpc_ref_upsample_bilinear.argtypes = [fp, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        L.fpc_ref_upsample_bilinear.restype = None
        L.fpc_ref_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _chk(t: torch.Tensor, dtype, name: str):
    if t.is_cuda:
        raise RuntimeError(f"oracle: {name} must be a CPU tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"oracle: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        # mirrors CHECK_CONTIGUOUS (ransac_voting.cpp:8)
        raise RuntimeError(f"{name} must be contiguous")


class RansacVotingCPU:
    """Module-like object: ``generate_hypothesis`` / ``voting_for_hypothesis``."""

    def __init__(self, fma: bool = False):
        self._sfx = "_fma" if fma else ""

    def generate_hypothesis(self, direct, coords, idxs):
        _chk(direct, torch.float32, "direct")
        _chk(coords, torch.float32, "coords")
        _chk(idxs, torch.int32, "idxs")
        tn, vn = direct.shape[0], direct.shape[1]
        hn = idxs.shape[0]
        out = torch.zeros((hn, vn, 2), dtype=torch.float32)
        if hn * vn:
            getattr(lib(), "fpc_ref_generate_hypothesis" + self._sfx)(
                direct.data_ptr(), coords.data_ptr(), idxs.data_ptr(), out.data_ptr(), tn, vn, hn)
        return out

    def voting_for_hypothesis(self, direct, coords, hypo_pts, inliers, inlier_thresh):
        _chk(direct, torch.float32, "direct")
        _chk(coords, torch.float32, "coords")
        _chk(hypo_pts, torch.float32, "hypo_pts")
        _chk(inliers, torch.uint8, "inliers")
        tn, vn = direct.shape[0], direct.shape[1]
        hn = hypo_pts.shape[0]
        if hn * vn * tn:
            getattr(lib(), "fpc_ref_voting_for_hypothesis" + self._sfx)(
                direct.data_ptr(), coords.data_ptr(), hypo_pts.data_ptr(), inliers.data_ptr(),
                tn, vn, hn, float(inlier_thresh))


    def generate_hypothesis_vanishing_point(self, direct, coords, idxs):
        """K3 (src/ransac_voting.cpp:62-73): homogeneous hypotheses [hn,vn,3]."""
        _chk(direct, torch.float32, "direct")
        _chk(coords, torch.float32, "coords")
        _chk(idxs, torch.int32, "idxs")
        tn, vn = direct.shape[0], direct.shape[1]
        hn = idxs.shape[0]
        out = torch.zeros((hn, vn, 3), dtype=torch.float32)
        if hn * vn:
            getattr(lib(), "fpc_ref_generate_hypothesis_vp" + self._sfx)(
                direct.data_ptr(), coords.data_ptr(), idxs.data_ptr(), out.data_ptr(), tn, vn, hn)
        return out

    def voting_for_hypothesis_vanishing_point(self, direct, coords, hypo_pts, inliers, inlier_thresh):
        """K4 (src/ransac_voting.cpp:83-97): in-place uint8 votes [hn,vn,tn]."""
        _chk(direct, torch.float32, "direct")
        _chk(coords, torch.float32, "coords")
        _chk(hypo_pts, torch.float32, "hypo_pts")
        _chk(inliers, torch.uint8, "inliers")
        tn, vn = direct.shape[0], direct.shape[1]
        hn = hypo_pts.shape[0]
        if hn * vn * tn:
            getattr(lib(), "fpc_ref_voting_for_hypothesis_vp" + self._sfx)(
                direct.data_ptr(), coords.data_ptr(), hypo_pts.data_ptr(), inliers.data_ptr(),
                tn, vn, hn, float(inlier_thresh))


ransac_voting = RansacVotingCPU(fma=False)
ransac_voting_fma = RansacVotingCPU(fma=True)


def upsample_bilinear(x: torch.Tensor, scale: int) -> torch.Tensor:
    """C restatement of nn.UpsamplingBilinear2d(scale_factor=scale) (oracle/head_epilogue_ref.c) on a CPU tensor."""
    _chk(x, torch.float32, "x")
    hl, wl = x.shape[-2:]
    out = torch.empty(tuple(x.shape[:-2]) + (hl * scale, wl * scale), dtype=torch.float32)
    if x.numel():
        lib().fpc_ref_upsample_bilinear(x.data_ptr(), x.numel() // (hl * wl), hl, wl, int(scale), out.data_ptr())
    return out


def num_threads() -> int:
    return int(lib().fpc_ref_num_threads())
