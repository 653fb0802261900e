"""ctypes binding of ``libfpc_b200.so`` (C ABI declared in ``include/fpc_b200.h``).

There is no fallback of any kind: if the shared library is missing the first
call raises, and every entry point refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfpc_b200.so")

FPC_OK, FPC_EINVAL, FPC_ECUDA, FPC_ECAPACITY = 0, -1, -2, -3
ARITH_IEEE, ARITH_NVCC_FMA = 0, 1
POSE_ROW = 48
NUM_COUNTERS = 16
CNT_INSTANCES, CNT_ROWS, CNT_RECORDS, CNT_WORK, CNT_FLAGS = 0, 1, 2, 3, 4
FLAG_INSTANCES, FLAG_ROWS, FLAG_RECORDS = 1, 2, 4
# word offsets inside a pose-table row (include/fpc_b200.h)
ROW_CLASS, ROW_SAMPLE, ROW_COUNT, ROW_Q, ROW_SCALES, ROW_XY, ROW_Z, ROW_T, ROW_R, ROW_RT = 0, 1, 2, 3, 7, 10, 12, 13, 16, 25
ROW_HYP, ROW_WIN_IDX, ROW_WIN_COUNT, ROW_TN, ROW_REFINE_INL, ROW_BBOX = 41, 43, 44, 45, 46, 47

EXPORTS = (
    "fpc_version", "fpc_last_error", "fpc_generate_hypothesis", "fpc_voting_for_hypothesis",
    "fpc_normalize", "fpc_class_compress", "fpc_get_rt", "fpc_pose_recover_workspace_bytes",
    "fpc_pose_recover", "fpc_pose_recover_num_launches", "fpc_pose_recover_kernel_name", "fpc_bench_fp32_fma",
    "fpc_aggregate", "fpc_vote_dense", "fpc_materialize_instances",
    "fpc_pack_masks", "fpc_pack_labels", "fpc_mask_iou", "fpc_match_instances", "fpc_paint_instances", "fpc_upsample_bilinear",
    "fpc_generate_hypothesis_vanishing_point", "fpc_voting_for_hypothesis_vanishing_point",
    "fpc_pose_errors", "fpc_threshold_fraction", "fpc_label_instances", "fpc_recover_args_size", "fpc_class_compress_backward", "fpc_aggregate_backward", "fpc_vote_refine_backward", "fpc_get_rt_backward", "fpc_pose_recover_xy_backward",
)
MASK_META = 8
MASK_F32, MASK_U8 = 0, 1

_vp, _i, _ll, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


class RecoverArgs(ctypes.Structure):
    """``fpc_recover_args`` (include/fpc_b200.h)."""
    _fields_ = [
        ("b", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32), ("num_classes", ctypes.c_int32),
        ("hn", ctypes.c_int32), ("max_instances", ctypes.c_int32),
        ("max_records", ctypes.c_int64), ("max_rows", ctypes.c_int64),
        ("inlier_thresh", ctypes.c_float), ("min_num", ctypes.c_int32), ("max_num", ctypes.c_int32),
        ("arith", ctypes.c_int32), ("seed", ctypes.c_uint64),
        ("mask_logits", _vp), ("quaternion", _vp), ("scales", _vp), ("xy", _vp), ("z", _vp),
        ("inv_intrinsics", _vp), ("idxs", _vp), ("select_u", _vp),
        ("pose_table", _vp), ("counters", _vp), ("cat_mask_u8", _vp), ("labels", _vp),
        ("hyp_out", _vp), ("vote_counts_out", _vp),
        ("workspace", _vp), ("workspace_bytes", ctypes.c_size_t), ("stream", _vp),
        ("stage_events", ctypes.POINTER(_vp)), ("num_stage_events", ctypes.c_int32),
        ("upsample", ctypes.c_int32),
        ("extra_out", _vp),
        ("stage_stamps", _vp),
    ]


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"fastposecnn_b200: {LIB_PATH} is missing -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C fastposecnn_b200/csrc`. "
            "There is no CPU or PyTorch fallback for this path.")
    L = ctypes.CDLL(LIB_PATH)
    L.fpc_version.restype = _i
    L.fpc_recover_args_size.restype = ctypes.c_size_t
    if L.fpc_recover_args_size() != ctypes.sizeof(RecoverArgs):
        raise RuntimeError(f"libfpc_b200.so was built with a different fpc_recover_args ({L.fpc_recover_args_size()} bytes) than "
                           f"this binding declares ({ctypes.sizeof(RecoverArgs)}): rebuild with `make -C fastposecnn_b200/csrc`")
    L.fpc_last_error.restype = ctypes.c_char_p
    L.fpc_generate_hypothesis.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]
    L.fpc_voting_for_hypothesis.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp]
    L.fpc_generate_hypothesis_vanishing_point.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]
    L.fpc_voting_for_hypothesis_vanishing_point.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp]
    L.fpc_generate_hypothesis_vanishing_point.restype = L.fpc_voting_for_hypothesis_vanishing_point.restype = _i
    L.fpc_pose_errors.argtypes = [_vp] * 10 + [_i] + [_vp] * 6
    L.fpc_threshold_fraction.argtypes = [_vp, _i, _vp, _i, _i, _vp, _vp]
    L.fpc_pose_errors.restype = L.fpc_threshold_fraction.restype = _i
    L.fpc_class_compress_backward.argtypes = [_vp] * 11 + [_i, _i, _i, _i, _vp]
    L.fpc_aggregate_backward.argtypes = [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]
    L.fpc_vote_refine_backward.argtypes = [_vp, _vp, _ll, _ll, _ll, _ll, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _vp, _vp]
    L.fpc_get_rt_backward.argtypes = [_vp] * 7 + [_i] + [_vp] * 4
    L.fpc_get_rt_backward.restype = _i
    L.fpc_pose_recover_xy_backward.argtypes = [_vp] * 8 + [_f, _i, _i, _i, _i, _i, _i, _vp, _vp]
    L.fpc_pose_recover_xy_backward.restype = _i
    L.fpc_class_compress_backward.restype = L.fpc_aggregate_backward.restype = L.fpc_vote_refine_backward.restype = _i
    L.fpc_normalize.argtypes = [_vp, _vp, _ll, _i, _ll, _vp]
    L.fpc_class_compress.argtypes = [_vp] * 11 + [_i, _i, _i, _i, _vp]
    L.fpc_get_rt.argtypes = [_vp] * 7 + [_i, _vp]
    L.fpc_pose_recover_workspace_bytes.argtypes = [ctypes.POINTER(RecoverArgs)]
    L.fpc_pose_recover_workspace_bytes.restype = ctypes.c_size_t
    L.fpc_pose_recover.argtypes = [ctypes.POINTER(RecoverArgs)]
    L.fpc_pose_recover_num_launches.restype = _i
    L.fpc_pose_recover_kernel_name.argtypes = [_i]
    L.fpc_pose_recover_kernel_name.restype = ctypes.c_char_p
    L.fpc_bench_fp32_fma.argtypes = [_vp, _i, _i, _vp]
    L.fpc_bench_fp32_fma.restype = _i
    L.fpc_aggregate.argtypes = [ctypes.POINTER(RecoverArgs), _vp]
    L.fpc_label_instances.argtypes = [ctypes.POINTER(RecoverArgs), _vp]
    L.fpc_label_instances.restype = _i
    L.fpc_vote_dense.argtypes = [ctypes.POINTER(RecoverArgs), _vp, _vp, _i, _i, _vp, _ll, _ll, _ll, _ll, _i]
    L.fpc_materialize_instances.argtypes = [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]
    L.fpc_pack_masks.argtypes = [_vp, _i, _i, _i, _i, _vp, _vp, _vp]
    L.fpc_pack_labels.argtypes = [_vp, _i, _i, _i, _i, _vp, _vp, _vp]
    L.fpc_mask_iou.argtypes = [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp]
    L.fpc_match_instances.argtypes = [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]
    L.fpc_paint_instances.argtypes = [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp]
    L.fpc_upsample_bilinear.argtypes = [_vp, _ll, _i, _i, _i, _vp, _vp]
    L.fpc_upsample_bilinear.restype = _i
    for name in ("fpc_aggregate", "fpc_vote_dense", "fpc_materialize_instances", "fpc_pack_masks", "fpc_pack_labels",
                 "fpc_mask_iou", "fpc_match_instances", "fpc_paint_instances"):
        getattr(L, name).restype = _i
    for name in ("fpc_generate_hypothesis", "fpc_voting_for_hypothesis", "fpc_normalize", "fpc_class_compress",
                 "fpc_get_rt", "fpc_pose_recover"):
        getattr(L, name).restype = _i
    _lib = L
    return L


class CapacityError(RuntimeError):
    """A capacity-sized table overflowed (FPC_FLAG_*).  Carries the counters the kernels reported so that callers with
    default capacities can grow their buffers and retry (the reference accepts any instance count,
    lib/aggregation_layer.py:87-118)."""

    def __init__(self, message: str, flags: int, instances: int, rows: int, records: int):
        super().__init__(message)
        self.flags, self.instances, self.rows, self.records = flags, instances, rows, records

    def grown(self, max_instances: int, max_rows: int, max_records: int, P: int, h: int):
        """Capacities for the retry: at least what the counters ask for, at least double what overflowed."""
        if self.flags & FLAG_INSTANCES:
            max_instances = max(self.instances + 16, 2 * max_instances)
            max_rows = max(max_rows, min(P, max_instances * h))
            max_records = max(max_records, P + 16 * max_instances)
        if self.flags & FLAG_ROWS:
            max_rows = min(max(P, 1), max(self.rows + 16, 2 * max_rows))
        if self.flags & FLAG_RECORDS:
            max_records = max(self.records + 16, 2 * max_records)
        return int(max_instances), int(max_rows), int(max_records)


def capacity_error(c, max_instances, max_rows, max_records) -> CapacityError:
    flags = int(c[CNT_FLAGS])
    what = [n for bit, n in ((FLAG_INSTANCES, f"instances ({int(c[CNT_INSTANCES])} > max_instances={max_instances})"),
                             (FLAG_ROWS, f"rows ({int(c[CNT_ROWS])} > max_rows={max_rows})"),
                             (FLAG_RECORDS, f"records ({int(c[CNT_RECORDS])} > max_records={max_records})")) if flags & bit]
    return CapacityError("libfpc_b200 error -3 (FPC_ECAPACITY): capacity exceeded for " + ", ".join(what), flags,
                         int(c[CNT_INSTANCES]), int(c[CNT_ROWS]), int(c[CNT_RECORDS]))


_seed_counter = 0


def fresh_seed() -> int:
    """Seed of one call's on-device pixel-pair sampling when the caller fixed none: torch's seed mixed with a call counter, so
    every call (and every keypoint) draws new pairs like the reference's ``random_`` (ransac_voting_gpu.py:552)."""
    global _seed_counter
    _seed_counter += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


def check(rc: int) -> None:
    if rc != FPC_OK:
        msg = lib().fpc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libfpc_b200 error {rc}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str, dtype=None, contiguous: bool = True) -> torch.Tensor:
    """Mirrors CHECK_INPUT of the reference binding (src/ransac_voting.cpp:7-9): CUDA + contiguous."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (fastposecnn_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


def require_device_readable(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    """CUDA tensor, or a PINNED host tensor: under unified virtual addressing a page-locked host allocation is
    addressable from the device with the same pointer, so the kernels can read it in place over PCIe (zero-copy
    ingestion of host-resident head maps: only the bytes the path needs ever cross the bus)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda and not t.is_pinned():
        raise RuntimeError(f"{name} must be a CUDA tensor or a pinned host tensor (fastposecnn_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


def current_stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
