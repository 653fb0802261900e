"""Drop-in for the reference's ``lib/metrics.py`` (six ``pl.metrics.Metric`` accumulators over matched pairs) without the
pytorch-lightning dependency: same class names, constructor arguments, ``update(gt_pred_matches)`` / ``compute()``
semantics -- including the reference's running "(old + new) / 2" update of the error metrics -- on top of the evaluation
kernels (``fpc_pose_errors``: one launch per update instead of a Python loop per pair)."""
from __future__ import annotations

import torch

from . import gpu_tensor_funcs as gtf


class _Metric:
    """The sliver of ``pl.metrics.Metric`` the reference uses: named tensor states, ``update`` / ``compute`` / ``reset``."""

    def __init__(self, name: str):
        self.name = name
        self._defaults = {}

    def add_state(self, name, default, dist_reduce_fx=None):
        self._defaults[name] = (default.clone(), dist_reduce_fx)
        setattr(self, name, default.clone())

    def reset(self):
        for name, (default, _) in self._defaults.items():
            setattr(self, name, default.clone())

    def __call__(self, gt_pred_matches):
        self.update(gt_pred_matches)
        return self.compute()


def _has(matches, key) -> bool:
    return matches is not None and key in matches.keys()


class _ThresholdAP(_Metric):
    """correct / total * 100 over every pair seen (lib/metrics.py:11-50, 91-133, 176-219)."""

    def __init__(self, name, threshold):
        super().__init__(name)
        self.threshold = threshold
        self.add_state("correct", default=torch.tensor(0), dist_reduce_fx="sum")
        self.add_state("total", default=torch.tensor(0), dist_reduce_fx="sum")

    def _count(self, passed: torch.Tensor):
        self.correct = self.correct.to(passed.device) + torch.sum(passed.int())
        self.total = self.total.to(passed.device) + passed.shape[0]

    def compute(self):
        return (self.correct.float() / self.total.float()) * 100


class _RunningError(_Metric):
    """The reference's running value: new = (old + mean of this batch) / 2, starting from 0 (lib/metrics.py:52-89, 135-174,
    221-260)."""

    def __init__(self, name, state):
        super().__init__(name)
        self._state = state
        self.add_state(state, default=torch.tensor(0), dist_reduce_fx="mean")

    def _push(self, values: torch.Tensor):
        old = getattr(self, self._state)
        setattr(self, self._state, (old.to(values.device) + torch.mean(values)) / 2)

    def compute(self):
        return getattr(self, self._state)


class DegreeErrorMeanAP(_ThresholdAP):
    def __init__(self, threshold):
        super().__init__(f"degree_error_mAP_{threshold}", threshold)

    def update(self, gt_pred_matches):
        if _has(gt_pred_matches, "quaternion"):
            q = gt_pred_matches["quaternion"]
            self._count(gtf.get_quat_distance(q[0], q[1], gt_pred_matches["symmetric_ids"]) < self.threshold)


class DegreeError(_RunningError):
    def __init__(self):
        super().__init__("degree_error", "error")

    def update(self, gt_pred_matches):
        if _has(gt_pred_matches, "quaternion"):
            q = gt_pred_matches["quaternion"]
            self._push(gtf.get_quat_distance(q[0], q[1], gt_pred_matches["symmetric_ids"]))


class Iou3dAP(_ThresholdAP):
    def __init__(self, threshold):
        super().__init__(f"3D_iou_mAP_{threshold}", threshold)

    def update(self, gt_pred_matches):
        if _has(gt_pred_matches, "RT"):
            rt, sc = gt_pred_matches["RT"], gt_pred_matches["scales"]
            self._count(gtf.get_3d_ious(rt[0], rt[1], sc[0], sc[1]) > self.threshold)


class Iou3dAccuracy(_RunningError):
    def __init__(self):
        super().__init__("3D_iou_accuracy", "accuracy")

    def update(self, gt_pred_matches):
        if _has(gt_pred_matches, "RT"):
            rt, sc = gt_pred_matches["RT"], gt_pred_matches["scales"]
            self._push(gtf.get_3d_ious(rt[0], rt[1], sc[0], sc[1]) * 100)


class OffsetAP(_ThresholdAP):
    def __init__(self, threshold):
        super().__init__(f"offset_error_mAP_{threshold}cm", threshold)

    def update(self, gt_pred_matches):
        if _has(gt_pred_matches, "RT"):
            t = gt_pred_matches["T"]
            self._count(gtf.from_Ts_get_offset_error(t[0], t[1]) < self.threshold)


class OffsetError(_RunningError):
    def __init__(self):
        super().__init__("offset_error", "error")

    def update(self, gt_pred_matches):
        if _has(gt_pred_matches, "RT"):
            rt = gt_pred_matches["RT"]
            self._push(gtf.from_RTs_get_T_offset_errors(rt[0], rt[1]))
