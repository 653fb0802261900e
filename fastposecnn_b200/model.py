"""The five boundary methods of the reference's ``Model`` mixin (lib/pose_regressor.py:443-504) plus the fused
fast entry.  ``PoseRecovery(HPARAM, classes, intrinsics)`` can stand where the reference's
``PoseRegressor`` inherits ``Model``: same method names, arguments and dict contracts."""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import gpu_tensor_funcs as gtf
from . import type_hinting as th
from .aggregation_layer import AggregationLayer
from .hough_voting import HoughVotingLayer
from .pose_recovery import pose_recover


class Model(object):
    """Mixin: expects ``self.classes``, ``self.HPARAM``, ``self.inv_intrinsics``, ``self.aggregation_layer``,
    ``self.hough_voting_layer`` (as the reference's PoseRegressor sets up, lib/pose_regressor.py:582-700)."""

    def class_compression(self, logits: th.LogitData) -> th.CategoricalData:
        return gtf.class_compression(logits, self.classes)                      # :445-457

    def aggregate(self, data: th.CategoricalData) -> th.AggData:
        return self.aggregation_layer.forward(data)                             # :459-465

    def hough_voting(self, agg_data: th.AggData) -> th.AggData:
        return self.hough_voting_layer(agg_data)                                # :467-473

    def perform_RT_calculation(self, agg_data: th.AggData) -> th.AggData:
        return gtf.samplewise_get_RT(agg_data, self.inv_intrinsics)             # :475-481

    def agg_hough_and_generate_RT(self, categorical_data: th.CategoricalData) -> Union[None, th.AggData]:
        if not self.HPARAM.PERFORM_AGGREGATION:                                 # :484-504
            return None
        agg_data = self.aggregate(categorical_data)
        if self.HPARAM.PERFORM_HOUGH_VOTING:
            agg_data = self.hough_voting(agg_data)
            if self.HPARAM.PERFORM_RT_CALCULATION:
                agg_data = self.perform_RT_calculation(agg_data)
        return agg_data


class PoseRecovery(torch.nn.Module, Model):
    """Stand-alone owner of the path: ``forward(logits)`` runs the reference's staged sequence
    (drop-in mode) and ``recover(logits)`` the fused one (no dense intermediates)."""

    def __init__(self, HPARAM, classes: int, intrinsics: torch.Tensor):
        super().__init__()
        self.HPARAM = HPARAM
        self.classes = classes
        self.intrinsics = intrinsics
        self.inv_intrinsics = torch.inverse(intrinsics)
        self.aggregation_layer = AggregationLayer(HPARAM, classes)
        self.hough_voting_layer = HoughVotingLayer(HPARAM)

    def _sync_device(self, device):
        if self.intrinsics.device != device:                                    # lib/pose_regressor.py:749-751
            self.intrinsics = self.intrinsics.to(device)
            self.inv_intrinsics = torch.inverse(self.intrinsics)

    def forward(self, logits: th.LogitData):
        self._sync_device(logits["mask"].device)
        categorical_data = self.class_compression(logits)
        agg_pred = self.agg_hough_and_generate_RT(categorical_data)
        return {"logits": logits, "categorical": categorical_data, "aggregated": agg_pred}

    def recover(self, logits: th.LogitData, idxs: Optional[torch.Tensor] = None, **kw):
        self._sync_device(logits["mask"].device)
        return pose_recover(logits, self.inv_intrinsics, self.HPARAM.HV_NUM_OF_HYPOTHESES, idxs=idxs, **kw)
