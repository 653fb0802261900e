"""Drop-in for the reference's lib/aggregation_layer.py: ``AggregationLayer(HPARAM, classes)(cat_data)``."""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from . import type_hinting as th
from .pose_recovery import table_to_agg


def _pipeline_args(b, h, w, num_classes, hn, max_instances, dev, **kw):
    """RecoverArgs + the buffers it points to for one call of a pipeline entry point."""
    a = _lib.RecoverArgs()
    P = b * h * w
    a.b, a.h, a.w, a.num_classes, a.hn = b, h, w, num_classes, hn
    a.max_instances = max_instances
    a.max_records = kw.get("max_records") or max(P, 1) + 16 * max_instances
    a.max_rows = kw.get("max_rows") or max(min(P, max_instances * h), 1)
    a.inlier_thresh = kw.get("inlier_thresh", 0.999)
    a.min_num, a.max_num = kw.get("min_num", 5), kw.get("max_num", 30000)
    seed = kw.get("seed")
    a.arith, a.seed = kw.get("arith", _lib.ARITH_IEEE), (_lib.fresh_seed() if seed is None else int(seed))
    L = _lib.lib()
    nbytes = L.fpc_pose_recover_workspace_bytes(ctypes.byref(a))
    if nbytes == 0:
        raise RuntimeError("libfpc_b200: " + L.fpc_last_error().decode())
    bufs = {
        "workspace": torch.empty(nbytes, dtype=torch.uint8, device=dev),
        "table_full": torch.zeros((max_instances + 1, _lib.POSE_ROW), dtype=torch.float32, device=dev),
        "labels": torch.empty((b, h, w), dtype=torch.int32, device=dev),
    }
    a.workspace, a.workspace_bytes = bufs["workspace"].data_ptr(), nbytes
    a.pose_table = bufs["table_full"][1:].data_ptr()
    a.counters = bufs["table_full"].data_ptr()
    a.labels = bufs["labels"].data_ptr()
    a.stream = _lib.current_stream(dev)
    bufs["caps"] = (int(a.max_instances), int(a.max_rows), int(a.max_records), P, h)
    return a, bufs


def grow_and_retry(call, max_instances, fixed: bool):
    """``call(max_instances, max_rows, max_records)`` with default table capacities; on overflow (CapacityError) the tables
    grow to what the kernels' counters ask for and the call is repeated -- the reference accepts any instance count
    (lib/aggregation_layer.py:87-118).  ``fixed``: the caller chose the capacity, overflow is an error."""
    caps = (max_instances, None, None)
    for attempt in range(5):
        try:
            return call(*caps)
        except _lib.CapacityError as e:
            if fixed or attempt == 4:
                raise
            caps = e.grown_caps


def _read_count(bufs, max_instances) -> int:
    c = bufs["table_full"][0, :_lib.NUM_COUNTERS].view(torch.int32).cpu()
    if int(c[_lib.CNT_FLAGS]):
        mi, mr, mrec, P, h = bufs["caps"]
        err = _lib.capacity_error(c, mi, mr, mrec)
        err.grown_caps = err.grown(mi, mr, mrec, P, h)
        raise err
    return int(c[_lib.CNT_INSTANCES])


def materialize_instance_masks(labels: torch.Tensor, table: torch.Tensor, n: int) -> torch.Tensor:
    """instance_masks [N,h,w] float32 0/1 (lib/aggregation_layer.py:101-105) from the label volume."""
    b, h, w = labels.shape
    out = torch.empty((n, h, w), dtype=torch.float32, device=labels.device)
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().fpc_materialize_instances(labels.data_ptr(), table.data_ptr(), None, out.data_ptr(), None,
                                                        n, h, w, _lib.current_stream(labels.device)))
    return out


def materialize_xy_mask(labels: torch.Tensor, table: torch.Tensor, xy_cat: torch.Tensor, n: int) -> torch.Tensor:
    """masked direction field [N,2,h,w] (lib/aggregation_layer.py:152-153)."""
    b, h, w = labels.shape
    xy_cat = _lib.require_cuda(xy_cat, "xy", torch.float32)
    out = torch.empty((n, 2, h, w), dtype=torch.float32, device=labels.device)
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().fpc_materialize_instances(labels.data_ptr(), table.data_ptr(), xy_cat.data_ptr(), None,
                                                        out.data_ptr(), n, h, w, _lib.current_stream(labels.device)))
    return out


class AggregationLayer(nn.Module):
    """Same constructor and call signature as the reference (lib/aggregation_layer.py:34-61).

    ``forward(cat_data)`` -> AggData with ``class_ids``/``sample_ids`` (int64), ``instance_masks``
    [N,h,w] float32, ``quaternion`` [N,4], ``scales`` [N,3], ``z`` [N,1] (= exp of the mean) and the
    masked dense ``xy`` [N,2,h,w] that HoughVotingLayer consumes.  Deviation: ``class_ids`` is always
    int64 (the reference silently yields float32 when a frame has no instance, :115)."""

    def __init__(self, HPARAM, classes, max_instances: int = None):
        super().__init__()
        self.HPARAM = HPARAM
        self.classes = classes  # including background
        self.max_instances = max_instances
        from .hough_voting import HoughVotingLayer
        self.hough_voting_layer = HoughVotingLayer(self.HPARAM)

    def forward(self, cat_data: th.CategoricalData) -> th.AggData:
        f32 = torch.float32
        cat_mask = _lib.require_cuda(cat_data["mask"], "cat_data['mask']", None, contiguous=False)
        if cat_mask.dtype != torch.int64:
            cat_mask = cat_mask.to(torch.int64)
        cat_mask = cat_mask.contiguous()
        q = _lib.require_cuda(cat_data["quaternion"], "cat_data['quaternion']", f32)
        s = _lib.require_cuda(cat_data["scales"], "cat_data['scales']", f32)
        xy = _lib.require_cuda(cat_data["xy"], "cat_data['xy']", f32)
        z = _lib.require_cuda(cat_data["z"], "cat_data['z']", f32)
        b, h, w = cat_mask.shape
        if tuple(q.shape) != (b, 4, h, w) or tuple(s.shape) != (b, 3, h, w) or tuple(xy.shape) != (b, 2, h, w) \
                or tuple(z.shape) != (b, h, w):
            raise RuntimeError("AggregationLayer: cat_data tensors have inconsistent shapes")
        dev = cat_mask.device
        def run():
            return grow_and_retry(run_with, self.max_instances or max(1024, 128 * b), fixed=self.max_instances is not None)

        def run_with(cap, max_rows, max_records):
            with torch.cuda.device(dev):
                a, bufs = _pipeline_args(b, h, w, int(self.classes), 1, cap, dev, max_rows=max_rows, max_records=max_records)
                a.quaternion, a.scales, a.xy, a.z = q.data_ptr(), s.data_ptr(), xy.data_ptr(), z.data_ptr()
                extra = torch.zeros((cap, 4), dtype=f32, device=dev)
                a.extra_out = extra.data_ptr()
                _lib.check(_lib.lib().fpc_aggregate(ctypes.byref(a), cat_mask.data_ptr()))
                n = _read_count(bufs, cap)
            table = bufs["table_full"][1:]
            full = table_to_agg(table, n)
            agg = {k: full[k] for k in ("class_ids", "sample_ids", "quaternion", "scales", "z")}
            agg["instance_masks"] = materialize_instance_masks(bufs["labels"], table, n)
            agg["xy"] = materialize_xy_mask(bufs["labels"], table, xy, n)
            return agg, bufs["labels"], full["mask_sizes"].to(torch.int32), extra[:n, 2].clone()

        if torch.is_grad_enabled() and any(t.requires_grad for t in (q, s, xy, z)):
            # training: the masked means (and the masked xy) stay differentiable w.r.t. the class-compressed fields
            from .autograd import AggregateFn
            holder = {}

            def run_and_keep():
                res = run()
                holder["agg"] = res[0]
                return res
            q_o, s_o, z_o, xy_o = AggregateFn.apply(run_and_keep, q, s, xy, z)
            agg = dict(holder["agg"])
            agg.update({"quaternion": q_o, "scales": s_o, "z": z_o, "xy": xy_o})
            return agg
        return run()[0]

    def batchwise_break_segmentation_mask(self, class_mask: torch.Tensor):
        """lib/aggregation_layer.py:160-183: 4-connected labelling of a [b,h,w] foreground volume, labels in
        raster order of each component's first pixel, no links across images -> (int32 labels, count)."""
        class_mask = _lib.require_cuda(class_mask, "class_mask", None, contiguous=False)
        b, h, w = class_mask.shape
        dev = class_mask.device
        cat = (class_mask != 0).to(torch.int64).contiguous()
        def label_with(cap, max_rows, max_records):
            with torch.cuda.device(dev):
                a, bufs = _pipeline_args(b, h, w, 2, 1, cap, dev, max_rows=max_rows, max_records=max_records)
                _lib.check(_lib.lib().fpc_label_instances(ctypes.byref(a), cat.data_ptr()))
                return bufs["labels"], _read_count(bufs, cap)
        return grow_and_retry(label_with, self.max_instances or max(1024, 128 * b), fixed=self.max_instances is not None)
