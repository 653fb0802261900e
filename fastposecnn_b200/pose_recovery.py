"""Fused entry of the path: per-pixel head outputs -> per-instance pose table.

Replaces, in one stream of 15 kernel launches with no host synchronisation,
``Model.class_compression`` -> ``aggregate`` -> ``hough_voting`` ->
``perform_RT_calculation`` (lib/pose_regressor.py:445-504 of the reference).
The dense intermediates of the reference (``instance_masks [N,h,w]``,
``xy_mask [N,2,h,w]``, the ``hn x tn`` inlier matrix) are never materialised
unless asked for.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import RecoverArgs


class PoseRecoveryEngine:
    """Owns every device buffer one shape of the path needs (workspace, pose table, counters)
    so that repeated calls allocate nothing and the launch sequence can be CUDA-graphed."""

    def __init__(self, b: int, h: int, w: int, num_classes: int, hn: int, device, *, max_instances: Optional[int] = None,
                 max_records: Optional[int] = None, max_rows: Optional[int] = None, inlier_thresh: float = 0.999,
                 min_num: int = 5, max_num: int = 30000, arith: int = _lib.ARITH_IEEE, seed: Optional[int] = None,
                 want_labels: bool = False, upsample: int = 1):
        self.device = torch.device(device)
        # head-epilogue fusion: inputs are the heads' low-resolution outputs [b,.,h/upsample,w/upsample]; (h, w) stay the
        # full output resolution every table, mask and coordinate refers to
        self.upsample = int(upsample)
        if self.upsample < 1 or h % self.upsample or w % self.upsample:
            raise RuntimeError(f"h, w ({h}, {w}) must be multiples of upsample ({upsample})")
        if self.device.type != "cuda":
            raise RuntimeError("PoseRecoveryEngine needs a CUDA device (no CPU path)")
        self.b, self.h, self.w, self.num_classes, self.hn = b, h, w, num_classes, hn
        P = b * h * w
        self.max_instances = int(max_instances if max_instances is not None else max(1024, 128 * b))
        self.max_records = int(max_records if max_records is not None else P + 16 * self.max_instances)  # ranges padded to 16
        self.max_rows = int(max_rows if max_rows is not None else min(P, self.max_instances * h))
        self.inlier_thresh, self.min_num, self.max_num = float(inlier_thresh), int(min_num), int(max_num)
        # seed=None: every launch() draws fresh pixel pairs (a captured graph replays the pairs of its capture)
        self.arith, self.seed = int(arith), (None if seed is None else int(seed))
        L = _lib.lib()
        a = self._base_args()
        nbytes = L.fpc_pose_recover_workspace_bytes(ctypes.byref(a))
        if nbytes == 0:
            raise RuntimeError("libfpc_b200: " + L.fpc_last_error().decode())
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        # one buffer = header row (the counters) + pose rows: it is also the all-gather send buffer (sharding.py)
        self.table_full = torch.zeros((self.max_instances + 1, _lib.POSE_ROW), dtype=torch.float32, device=self.device)
        self.pose_table = self.table_full[1:]
        self.counters = self.table_full[0, :_lib.NUM_COUNTERS].view(torch.int32)
        self.counters_host = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int32).pin_memory()
        self.cat_mask_u8 = torch.empty((b, h, w), dtype=torch.uint8, device=self.device)
        # the scipy-style label volume (instance id + 1 per pixel) is an optional output: one extra kernel
        self.labels = torch.empty((b, h, w), dtype=torch.int32, device=self.device) if want_labels else None
        self.hyp = torch.empty((self.max_instances, hn, 2), dtype=torch.float32, device=self.device)
        self.votes = torch.empty((self.max_instances, hn), dtype=torch.int32, device=self.device)
        self.num_launches = int(L.fpc_pose_recover_num_launches())
        self._fetch_event = None
        self._graph = None
        self._idxs_keepalive = None
        self.extra_out = None          # optional [max_instances,4] f32 (fpc_recover_args.extra_out), set by callers that need it

    def _base_args(self) -> RecoverArgs:
        a = RecoverArgs()
        a.b, a.h, a.w, a.num_classes, a.hn = self.b, self.h, self.w, self.num_classes, self.hn
        a.max_instances, a.max_records, a.max_rows = self.max_instances, self.max_records, self.max_rows
        a.inlier_thresh, a.min_num, a.max_num = self.inlier_thresh, self.min_num, self.max_num
        a.arith, a.seed = self.arith, (_lib.fresh_seed() if self.seed is None else self.seed)
        a.upsample = self.upsample
        return a

    def launch(self, logits: Dict[str, torch.Tensor], inv_intrinsics: torch.Tensor, idxs: Optional[torch.Tensor] = None,
               select_u: Optional[torch.Tensor] = None, stage_events=None, stage_stamps: Optional[torch.Tensor] = None) -> None:
        """Enqueues the kernels on the current stream.  No synchronisation.  ``stage_stamps``: optional int64 device tensor
        ``[num_launches + 1]`` that receives %globaltimer before / after every kernel (graph-capturable timeline)."""
        b, C, K = self.b, self.num_classes, self.num_classes - 1
        h, w = self.h // self.upsample, self.w // self.upsample      # resolution of the head maps handed in
        f32 = torch.float32
        mask = _lib.require_device_readable(logits["mask"], "logits['mask']", f32)
        quat = _lib.require_device_readable(logits["quaternion"], "logits['quaternion']", f32)
        scales = _lib.require_device_readable(logits["scales"], "logits['scales']", f32)
        xy = _lib.require_device_readable(logits["xy"], "logits['xy']", f32)
        z = _lib.require_device_readable(logits["z"], "logits['z']", f32)
        # torch.inverse returns a column-major tensor; a 3x3 copy is free
        inv_k = _lib.require_cuda(inv_intrinsics, "inv_intrinsics", f32, contiguous=False).contiguous()
        self._invk_keepalive = inv_k
        for t, c, name in ((mask, C, "mask"), (quat, 4 * K, "quaternion"), (scales, 3 * K, "scales"),
                           (xy, 2 * K, "xy"), (z, K, "z")):
            if tuple(t.shape) != (b, c, h, w):
                raise RuntimeError(f"logits['{name}'] must be [{b},{c},{h},{w}], got {tuple(t.shape)}")
        if tuple(inv_k.shape) != (3, 3):
            raise RuntimeError("inv_intrinsics must be [3,3]")
        a = self._base_args()
        a.mask_logits, a.quaternion, a.scales, a.xy, a.z = mask.data_ptr(), quat.data_ptr(), scales.data_ptr(), xy.data_ptr(), z.data_ptr()
        a.inv_intrinsics = inv_k.data_ptr()
        if idxs is not None:
            idxs = _lib.require_cuda(idxs, "idxs", torch.int32)
            if idxs.numel() != 0 and idxs.numel() != idxs.shape[0] * self.hn * 2:
                raise RuntimeError(f"idxs must be [N,{self.hn},(1,)2] int32, got {tuple(idxs.shape)}")
            if idxs.shape[0] > self.max_instances:
                idxs = idxs[: self.max_instances].contiguous()   # the capacity flag reports the overflow
            if idxs.shape[0] < self.max_instances:
                # the kernel indexes idxs[instance] for every live instance; pad to capacity once
                pad = torch.zeros((self.max_instances, self.hn, 2), dtype=torch.int32, device=self.device)
                pad[: idxs.shape[0]] = idxs.reshape(idxs.shape[0], self.hn, 2)
                idxs = pad
            self._idxs_keepalive = idxs
            a.idxs = idxs.data_ptr()
        if select_u is not None:
            select_u = _lib.require_cuda(select_u, "select_u", f32)
            if tuple(select_u.shape) != (b, self.h, self.w):
                raise RuntimeError("select_u must be [b,h,w]")
            a.select_u = select_u.data_ptr()
        a.pose_table, a.counters = self.pose_table.data_ptr(), self.counters.data_ptr()
        a.cat_mask_u8 = self.cat_mask_u8.data_ptr()
        if self.labels is not None:
            a.labels = self.labels.data_ptr()
        a.hyp_out, a.vote_counts_out = self.hyp.data_ptr(), self.votes.data_ptr()
        if self.extra_out is not None:
            a.extra_out = self.extra_out.data_ptr()
        a.workspace, a.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel()
        a.stream = _lib.current_stream(self.device)
        if stage_stamps is not None:
            if stage_stamps.dtype != torch.int64 or stage_stamps.numel() < self.num_launches + 1 or not stage_stamps.is_cuda:
                raise RuntimeError("stage_stamps must be a CUDA int64 tensor of num_launches + 1 elements")
            a.stage_stamps = stage_stamps.data_ptr()
        if stage_events is not None:
            # torch creates the cudaEvent_t lazily on the first record(); callers pass recorded events
            arr = (ctypes.c_void_p * len(stage_events))(*[int(e.cuda_event) for e in stage_events])
            self._events_keepalive = arr
            a.stage_events = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
            a.num_stage_events = len(stage_events)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fpc_pose_recover(ctypes.byref(a)))

    def capture(self, logits, inv_intrinsics, idxs=None, select_u=None, stage_stamps=None) -> None:
        """Captures the launch sequence for these (fixed-address) inputs into a CUDA graph; ``replay()`` then
        re-issues all kernels with one driver call.  The inputs' storage must stay alive and in place.
        ``stage_stamps`` (int64 device tensor, one more element than kernels): %globaltimer stamps between the kernels, re-written
        by every replay (tools/timeline.py; CUDA events recorded inside a graph cannot be used for timing)."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.launch(logits, inv_intrinsics, idxs=idxs, select_u=select_u)      # warm-up: pads idxs, fixes addresses
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self.launch(logits, inv_intrinsics, idxs=self._idxs_keepalive if idxs is not None else None, select_u=select_u,
                            stage_stamps=stage_stamps)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph = graph
        self._graph_inputs = (logits, inv_intrinsics, idxs, select_u)

    def replay(self) -> None:
        if self._graph is None:
            raise RuntimeError("PoseRecoveryEngine.replay(): call capture() first")
        self._graph.replay()

    def enqueue_fetch(self) -> None:
        """Enqueues the path's single device->host read (the 16 counters: N and the capacity flags) behind the
        kernels and records an event; ``wait_count`` later blocks on that event only, so the next batch can
        already be in the stream (see ``PoseRecoveryPipeline``)."""
        self.counters_host.copy_(self.counters, non_blocking=True)
        if self._fetch_event is None:
            self._fetch_event = torch.cuda.Event()
        self._fetch_event.record(torch.cuda.current_stream(self.device))

    def fetch_count(self) -> int:
        """launch()'s companion: enqueue the read of N and wait for it."""
        self.enqueue_fetch()
        return self.wait_count()

    def wait_count(self) -> int:
        self._fetch_event.synchronize()
        c = self.counters_host
        if int(c[_lib.CNT_FLAGS]):
            raise _lib.capacity_error(c, self.max_instances, self.max_rows, self.max_records)
        return int(c[_lib.CNT_INSTANCES])

    def table_to_agg(self, n: int, table: Optional[torch.Tensor] = None, sample_offset: int = 0) -> Dict[str, torch.Tensor]:
        """Slices the first ``n`` pose-table rows into the reference's AggData keys (lib/type_hinting.py:19-32)."""
        return table_to_agg(self.pose_table if table is None else table, n, sample_offset)


def table_to_agg(table: torch.Tensor, n: int, sample_offset: int = 0) -> Dict[str, torch.Tensor]:
    t = table[:n]
    ti = t.view(torch.int32)
    L = _lib
    xy = t[:, L.ROW_XY:L.ROW_XY + 2].contiguous()
    agg = {
        "class_ids": ti[:, L.ROW_CLASS].to(torch.int64),
        "sample_ids": ti[:, L.ROW_SAMPLE].to(torch.int64) + sample_offset,
        "quaternion": t[:, L.ROW_Q:L.ROW_Q + 4].contiguous(),
        "scales": t[:, L.ROW_SCALES:L.ROW_SCALES + 3].contiguous(),
        "xy": xy,
        "z": t[:, L.ROW_Z:L.ROW_Z + 1].contiguous(),
        "R": t[:, L.ROW_R:L.ROW_R + 9].reshape(n, 3, 3),
        "T": t[:, L.ROW_T:L.ROW_T + 3].contiguous(),
        "RT": t[:, L.ROW_RT:L.ROW_RT + 16].reshape(n, 4, 4),
        "hypothesis": xy.unsqueeze(1),
        "pruned_hypothesis": xy.unsqueeze(1),
        # extras (not in the reference's AggData)
        "mask_sizes": ti[:, L.ROW_COUNT].to(torch.int64),
        "win_hypothesis": t[:, L.ROW_HYP:L.ROW_HYP + 2].contiguous(),
        "win_idx": ti[:, L.ROW_WIN_IDX].to(torch.int64),
        "win_counts": ti[:, L.ROW_WIN_COUNT].to(torch.int64),
        "tn": ti[:, L.ROW_TN].to(torch.int64),
        "refine_inliers": ti[:, L.ROW_REFINE_INL].to(torch.int64),
    }
    return agg


class PoseRecoveryPipeline:
    """``depth`` engines used round-robin, each on its OWN stream: batch k+1 is enqueued before the host waits for
    batch k's instance count, and kernels of different batches overlap on the device -- the HBM-bound arg-max / gather
    of one batch run next to the FP32-bound vote kernel of another.  Every batch still performs its own device->host
    read; results of the last ``depth`` batches stay valid (each engine owns its tables)."""

    def __init__(self, depth: int, *engine_args, multi_stream: bool = True, **engine_kw):
        self.engines = [PoseRecoveryEngine(*engine_args, **engine_kw) for _ in range(depth)]
        dev = self.engines[0].device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(depth)] if multi_stream else [None] * depth
        self._k = 0
        self._pending = []

    def capture(self, logits, inv_intrinsics, idxs=None, select_u=None) -> None:
        for e in self.engines:
            e.capture(logits, inv_intrinsics, idxs=idxs, select_u=select_u)

    def submit(self, logits=None, inv_intrinsics=None, idxs=None, select_u=None, stage_events=None, after_launch=None,
               replay: bool = False, before_launch=None):
        """Enqueue one batch (``replay=True``: re-issue the captured graph of this slot's engine).  Returns
        (engine, N) of the OLDEST in-flight batch once ``depth`` are in flight, else None."""
        slot = self._k % len(self.engines)
        eng, stream = self.engines[slot], self.streams[slot]
        self._k += 1
        if stream is not None:
            stream.wait_stream(torch.cuda.current_stream(eng.device))      # inputs produced on the caller's stream
        ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
        with ctx:
            if before_launch is not None:
                before_launch(eng)
            if replay:
                eng.replay()
            else:
                eng.launch(logits, inv_intrinsics, idxs=idxs, select_u=select_u, stage_events=stage_events)
            if after_launch is not None:
                after_launch(eng)
            eng.enqueue_fetch()
        self._pending.append(eng)
        if len(self._pending) >= len(self.engines):
            old = self._pending.pop(0)
            return old, old.wait_count()
        return None

    def drain(self):
        out = [(e, e.wait_count()) for e in self._pending]
        self._pending = []
        return out

    def join(self) -> None:
        """Makes the caller's current stream wait for everything enqueued on the slot streams."""
        for s in self.streams:
            if s is not None:
                torch.cuda.current_stream(self.engines[0].device).wait_stream(s)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _recover_training(eng, logits, heads, inv_intrinsics, idxs, select_u, C, device):
    """quaternion / scales / z / xy of the result stay differentiable w.r.t. their head maps (autograd.PoseRecoverFn); each
    call gets its own tables so that the saved tensors survive later calls."""
    from .autograd import PoseRecoverFn
    holder = {}

    def run():
        extra = torch.zeros((eng.max_instances, 4), dtype=torch.float32, device=device)
        eng.extra_out = extra
        try:
            eng.launch({k: v.detach() for k, v in logits.items()}, inv_intrinsics, idxs=idxs, select_u=select_u)
        finally:
            eng.extra_out = None
        n_ = eng.fetch_count()
        agg_ = eng.table_to_agg(n_, table=eng.pose_table[:n_].clone())
        if n_ and int(agg_["mask_sizes"].max()) > eng.max_num:
            raise NotImplementedError("pose_recover backward: an instance was sub-sampled to max_num voters, which is not "
                                      "differentiable here; raise max_num")
        holder["agg"], holder["n"] = agg_, n_
        return agg_, eng.labels.clone(), eng.cat_mask_u8.clone(), agg_["mask_sizes"].to(torch.int32), extra[:n_, 2].clone()
    q_o, s_o, z_o, xy_o = PoseRecoverFn.apply(run, C, eng.inlier_thresh, eng.arith, *heads)
    agg, n = holder["agg"], holder["n"]
    agg.update({"quaternion": q_o, "scales": s_o, "z": z_o, "xy": xy_o})
    # R / T / RT through the differentiable batchwise_get_RT: rotation / translation losses reach q, z and the xy head
    from .gpu_tensor_funcs import batchwise_get_RT
    agg["R"], agg["T"], agg["RT"] = batchwise_get_RT(q_o, xy_o, z_o, inv_intrinsics)
    return agg, n


ENGINE_CACHE_SIZE = 4                                   # engines kept by get_engine (least recently used is dropped)
_engines: "OrderedDict[tuple, PoseRecoveryEngine]" = OrderedDict()
_grown_caps: Dict[tuple, dict] = {}                     # shape key -> capacities a default-capacity engine had to grow to
_CAP_KEYS = ("max_instances", "max_rows", "max_records")


def get_engine(b, h, w, num_classes, hn, device, **kw) -> PoseRecoveryEngine:
    """Engines are cached by shape (they own multi-GB workspaces); at most ``ENGINE_CACHE_SIZE`` are kept, so a
    last-batch-of-epoch shape does not pin a second workspace forever."""
    if kw.get("upsample", 1) == 1:
        kw.pop("upsample", None)          # the default: same cache entry whether or not it was spelled out
    key = (b, h, w, num_classes, hn, str(torch.device(device)), tuple(sorted(kw.items())))
    eng = _engines.get(key)
    if eng is None:
        eng = PoseRecoveryEngine(b, h, w, num_classes, hn, device, **kw)
        _engines[key] = eng
        while len(_engines) > max(1, ENGINE_CACHE_SIZE):
            _engines.popitem(last=False)
    else:
        _engines.move_to_end(key)
    return eng


def pose_recover(logits: Dict[str, torch.Tensor], inv_intrinsics: torch.Tensor, hn: int, idxs: Optional[torch.Tensor] = None,
                 *, select_u: Optional[torch.Tensor] = None, materialize_dense: bool = False, upsample: int = 1,
                 **engine_kw) -> Dict[str, torch.Tensor]:
    """logits (LogitData, lib/type_hinting.py:5-10) -> AggData with the reference's keys.

    ``idxs``: fixed pre-sampled hypothesis pixel pairs ``[N,hn,2]`` (or ``[N,hn,1,2]``) int32, instance
    order; ``None`` samples on the device.  ``materialize_dense=True`` additionally returns the
    reference's dense ``instance_masks [N,h,w]`` and ``xy_mask [N,2,h,w]``.

    ``upsample=S > 1`` (head-epilogue fusion): ``logits`` are the heads' LOW-RESOLUTION outputs ``[b,.,h/S,w/S]`` (after
    the 1x1 convolutions, before ``nn.UpsamplingBilinear2d(scale_factor=S)``); the up-sampling is evaluated inside the
    kernels and the full-resolution head maps never exist.  Results refer to the full ``[h,w]`` resolution.

    Every returned tensor is OWNED by the caller (a copy of the cached engine's tables), like the reference's fresh
    tensors: a later call with the same shape does not change earlier results.  Like the reference
    (lib/aggregation_layer.py:87-118) any number of instances is accepted: with default capacities the tables grow and
    the call is repeated; capacities passed explicitly (``max_instances`` / ``max_rows`` / ``max_records``) are hard
    limits and overflow raises ``CapacityError`` (FPC_ECAPACITY)."""
    mask = logits["mask"]
    b, C, h, w = mask.shape
    h, w = h * upsample, w * upsample
    # head maps may also be PINNED HOST tensors (read in place over PCIe); the device then comes from inv_intrinsics
    device = mask.device if mask.is_cuda else inv_intrinsics.device
    auto_grow = not any(k in engine_kw for k in _CAP_KEYS)
    shape_key = (b, h, w, C, hn, str(torch.device(device)), upsample, tuple(sorted(engine_kw.items())))
    heads = [logits[k] for k in ("quaternion", "scales", "z", "xy")]
    training = torch.is_grad_enabled() and upsample == 1 and any(t.requires_grad for t in heads)
    for attempt in range(5):
        caps = _grown_caps.get(shape_key, {}) if auto_grow else {}
        eng = get_engine(b, h, w, C, hn, device, want_labels=True, upsample=upsample, **engine_kw, **caps)
        try:
            if training:
                agg, n = _recover_training(eng, logits, heads, inv_intrinsics, idxs, select_u, C, device)
            else:
                eng.launch(logits, inv_intrinsics, idxs=idxs, select_u=select_u)
                n = eng.fetch_count()
                agg = eng.table_to_agg(n, table=eng.pose_table[:n].clone())
            break
        except _lib.CapacityError as e:
            if not auto_grow or attempt == 4:
                raise
            mi, mr, mrec = e.grown(eng.max_instances, eng.max_rows, eng.max_records, b * h * w, h)
            _grown_caps[shape_key] = {"max_instances": mi, "max_rows": mr, "max_records": mrec}
            for k in [k for k, v in _engines.items() if v is eng]:
                del _engines[k]                          # the outgrown engine's workspace is released
            del eng
    agg["cat_mask"] = eng.cat_mask_u8.clone()
    agg["labels"] = eng.labels.clone()
    if materialize_dense:
        # optional reference-layout dense outputs (one extra class-compression pass for the xy field)
        from .aggregation_layer import materialize_instance_masks, materialize_xy_mask
        from .gpu_tensor_funcs import class_compression
        agg["instance_masks"] = materialize_instance_masks(eng.labels, eng.pose_table, n)
        if upsample > 1:
            from .head_epilogue import upsample_bilinear
            logits = {k: upsample_bilinear(v, upsample) for k, v in logits.items()}
        agg["xy_mask"] = materialize_xy_mask(eng.labels, eng.pose_table, class_compression(logits, C)["xy"], n)
    return agg
