"""Drop-in for the hot subset of the reference's lib/gpu_tensor_funcs.py: same names, argument
meaning and return layouts, backed by sm_100a kernels through libfpc_b200.so (no torch fallback)."""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib
from . import type_hinting as th


def normalize(data: torch.Tensor, dim: int) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:37-50 -- L2-normalise along ``dim``; zero vectors stay zero."""
    data = _lib.require_cuda(data, "data", torch.float32, contiguous=False).contiguous()
    dim = dim % data.dim()
    outer = 1
    for s in data.shape[:dim]:
        outer *= s
    inner = 1
    for s in data.shape[dim + 1:]:
        inner *= s
    out = torch.empty_like(data)
    with torch.cuda.device(data.device):
        _lib.check(_lib.lib().fpc_normalize(data.data_ptr(), out.data_ptr(), outer, data.shape[dim], inner,
                                            _lib.current_stream(data.device)))
    return out


def _compress(num_of_classes: int, mask_logits, cat_mask, logits) -> th.CategoricalData:
    f32 = torch.float32
    quat = _lib.require_cuda(logits["quaternion"], "logits['quaternion']", f32)
    scales = _lib.require_cuda(logits["scales"], "logits['scales']", f32)
    xy = _lib.require_cuda(logits["xy"], "logits['xy']", f32)
    z = _lib.require_cuda(logits["z"], "logits['z']", f32)
    b, _, h, w = quat.shape
    K = num_of_classes - 1
    for t, c, name in ((quat, 4 * K, "quaternion"), (scales, 3 * K, "scales"), (xy, 2 * K, "xy"), (z, K, "z")):
        if tuple(t.shape) != (b, c, h, w):
            raise RuntimeError(f"logits['{name}'] must be [{b},{c},{h},{w}], got {tuple(t.shape)}")
    dev = quat.device
    cat_out = None
    if cat_mask is None:
        mask_logits = _lib.require_cuda(mask_logits, "logits['mask']", f32)
        if tuple(mask_logits.shape) != (b, num_of_classes, h, w):
            raise RuntimeError(f"logits['mask'] must be [{b},{num_of_classes},{h},{w}]")
        cat_out = torch.empty((b, h, w), dtype=torch.int64, device=dev)
    else:
        cat_mask = _lib.require_cuda(cat_mask, "cat_mask", torch.int64)
        if tuple(cat_mask.shape) != (b, h, w):
            raise RuntimeError(f"cat_mask must be [{b},{h},{w}]")
    def run():
        q_out = torch.empty((b, 4, h, w), dtype=f32, device=dev)
        s_out = torch.empty((b, 3, h, w), dtype=f32, device=dev)
        xy_out = torch.empty((b, 2, h, w), dtype=f32, device=dev)
        z_out = torch.empty((b, h, w), dtype=f32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fpc_class_compress(
                _lib.ptr(mask_logits) if cat_mask is None else None, _lib.ptr(cat_mask),
                quat.data_ptr(), scales.data_ptr(), xy.data_ptr(), z.data_ptr(),
                _lib.ptr(cat_out), q_out.data_ptr(), s_out.data_ptr(), xy_out.data_ptr(), z_out.data_ptr(),
                b, num_of_classes, h, w, _lib.current_stream(dev)))
        res = {"quaternion": q_out, "scales": s_out, "xy": xy_out, "z": z_out}
        if cat_out is not None:
            res["mask"] = cat_out
        else:
            res["_cat_mask"] = cat_mask
        return res

    if torch.is_grad_enabled() and any(t.requires_grad for t in (quat, scales, xy, z)):
        # training: keep the graph (backward = fpc_class_compress_backward)
        from .autograd import ClassCompressFn
        q_o, s_o, xy_o, z_o, cat_o = ClassCompressFn.apply(run, num_of_classes, quat, scales, xy, z)
        out = {"quaternion": q_o, "scales": s_o, "xy": xy_o, "z": z_o}
        if cat_out is not None:
            out["mask"] = cat_o
        return out
    out = run()
    out.pop("_cat_mask", None)
    return out


def class_compress(num_of_classes: int, cat_mask: torch.Tensor, logits: th.LogitData) -> th.CategoricalData:
    """lib/gpu_tensor_funcs.py:52-99 -- per pixel keep the channels of the predicted class; quaternion
    and xy are L2-normalised per pixel.  Returns the same keys as the reference (no 'mask')."""
    return _compress(num_of_classes, None, cat_mask, logits)


def class_compression(logits: th.LogitData, num_of_classes: int) -> th.CategoricalData:
    """Model.class_compression (lib/pose_regressor.py:445-457): arg-max of the mask logits fused with
    class_compress; result includes 'mask' (int64 [b,h,w])."""
    return _compress(num_of_classes, logits["mask"], None, logits)


def batchwise_get_RT(q: torch.Tensor, xys: torch.Tensor, exp_zs: torch.Tensor, inv_intrinsics: torch.Tensor):
    """lib/gpu_tensor_funcs.py:204-235 -> (R [n,3,3], T [n,3], RT [n,4,4])."""
    f32 = torch.float32
    q = _lib.require_cuda(q, "q", f32, contiguous=False).contiguous()
    xys = _lib.require_cuda(xys, "xys", f32, contiguous=False).contiguous()
    exp_zs = _lib.require_cuda(exp_zs, "exp_zs", f32, contiguous=False).contiguous()
    inv_k = _lib.require_cuda(inv_intrinsics, "inv_intrinsics", f32, contiguous=False).contiguous()
    n = q.shape[0]
    if tuple(q.shape) != (n, 4) or tuple(xys.shape) != (n, 2) or exp_zs.numel() != n or tuple(inv_k.shape) != (3, 3):
        raise RuntimeError("batchwise_get_RT: expected q [n,4], xys [n,2], exp_zs [n,1], inv_intrinsics [3,3]")
    def run():
        R = torch.empty((n, 3, 3), dtype=f32, device=q.device)
        T = torch.empty((n, 3), dtype=f32, device=q.device)
        RT = torch.empty((n, 4, 4), dtype=f32, device=q.device)
        with torch.cuda.device(q.device):
            _lib.check(_lib.lib().fpc_get_rt(q.data_ptr(), xys.data_ptr(), exp_zs.data_ptr(), inv_k.data_ptr(),
                                             R.data_ptr(), T.data_ptr(), RT.data_ptr(), n, _lib.current_stream(q.device)))
        return R, T, RT

    if torch.is_grad_enabled() and any(t.requires_grad for t in (q, xys, exp_zs)):
        from .autograd import GetRTFn
        return GetRTFn.apply(run, q, xys, exp_zs, inv_k)
    return run()


def samplewise_get_RT(agg_data: Dict[str, torch.Tensor], inv_intrinsics: torch.Tensor):
    """lib/gpu_tensor_funcs.py:237-253 -- adds 'R', 'T', 'RT' to agg_data."""
    R, T, RT = batchwise_get_RT(agg_data["quaternion"], agg_data["xy"], agg_data["z"], inv_intrinsics)
    agg_data["R"], agg_data["T"], agg_data["RT"] = R, T, RT
    return agg_data


def quats_2_rotation_matrix(q: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:306-326, quaternions [n,4] -> [n,3,3].  The reference's formula is quadratic in q and does
    not normalise: R_ref(q) = |q|^2 R(q / |q|).  The kernel behind ``batchwise_get_RT`` normalises (as :215-218 does before
    calling this), so the factor is applied here -- identical for unit quaternions, faithful for the others."""
    n = q.shape[0]
    zeros2 = torch.zeros((n, 2), dtype=torch.float32, device=q.device)
    ones = torch.ones((n, 1), dtype=torch.float32, device=q.device)
    eye = torch.eye(3, dtype=torch.float32, device=q.device)
    R, _, _ = batchwise_get_RT(q, zeros2, ones, eye)
    return R * (q.detach().float() ** 2).sum(dim=-1).reshape(n, 1, 1) if n else R


def batchwise_get_2d_iou(batch_masks1: torch.Tensor, batch_masks2: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:386-409 -- IoU of every ``[n1,h,w]`` mask with every ``[n2,h,w]`` mask -> ``[n1,n2]``
    float32 (0/0 = NaN).  Each mask is read once and bit-packed; pairs cost popc(a & b) over the overlap of their
    bounding boxes instead of the reference's ``[n1,n2,h,w]`` logical volumes; values are bit-identical."""
    from .matching import mask_iou, pack_masks
    return mask_iou(pack_masks(batch_masks1), pack_masks(batch_masks2))


def torch_get_2d_iou(tensor1: torch.Tensor, tensor2: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:380-384 -- IoU of two ``[h,w]`` masks, a 0-d float32 tensor."""
    return batchwise_get_2d_iou(tensor1.reshape(1, *tensor1.shape[-2:]), tensor2.reshape(1, *tensor2.shape[-2:]))[0, 0]


# ---------------------------------------------------------------------------------------------------------------------
# Evaluation maths on matched pairs: one launch of fpc_pose_errors instead of a Python loop
# ---------------------------------------------------------------------------------------------------------------------
_SYM_ROTATIONS: Dict[str, torch.Tensor] = {}


def _symmetry_rotations(device) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:762-780: the 360 quaternions (w, 0, y, 0) of 0..359 degrees about y, built in float32 on
    the host exactly like the reference's cached ``rot_q``, then widened -- [360,4] float64 on ``device``."""
    key = str(device)
    if key not in _SYM_ROTATIONS:
        half = torch.deg2rad(torch.arange(0, 360).float()) / 2
        s, c = torch.sin(half), torch.cos(half)
        _SYM_ROTATIONS[key] = torch.vstack((c, 0 * s, 1 * s, 0 * s)).T.double().contiguous().to(device)
    return _SYM_ROTATIONS[key]


def _pose_errors(q0=None, q1=None, symmetric_ids=None, rts=None, scales=None, ts=None, want_raw=False, want_sym=False,
                 want_centers=False):
    f32 = torch.float32
    ref = q0 if q0 is not None else (rts[0] if rts is not None else ts[0])
    dev, m = ref.device, int(ref.shape[0])

    def chk(t, name, shape):
        t = _lib.require_cuda(t.contiguous(), name, f32)
        if tuple(t.shape) != (m,) + shape:
            raise RuntimeError(f"{name} must be [{m},{','.join(map(str, shape))}], got {tuple(t.shape)}")
        return t
    out = {}
    args = [None] * 10
    if q0 is not None:
        args[0], args[1] = chk(q0, "q0", (4,)), chk(q1, "q1", (4,))
        if symmetric_ids is not None:
            args[2] = _lib.require_cuda(symmetric_ids.to(torch.int64).contiguous(), "symmetric_ids")
        if want_sym:
            args[3] = _symmetry_rotations(dev)
            out["sym"] = torch.empty((m,), dtype=torch.float64, device=dev)
        if want_raw:
            out["raw"] = torch.empty((m,), dtype=f32, device=dev)
    if rts is not None:
        args[4], args[5] = chk(rts[0], "RTs_1", (4, 4)), chk(rts[1], "RTs_2", (4, 4))
        if scales is not None:
            args[6], args[7] = chk(scales[0], "scales_1", (3,)), chk(scales[1], "scales_2", (3,))
            out["iou"] = torch.empty((m,), dtype=f32, device=dev)
        if want_centers:
            out["centers"] = torch.empty((m, 6), dtype=f32, device=dev)
    if ts is not None:
        args[8], args[9] = chk(ts[0], "gt_Ts", (3,)), chk(ts[1], "pred_Ts", (3,))
        out["offset"] = torch.empty((m,), dtype=f32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fpc_pose_errors(*[_lib.ptr(a) for a in args], m, _lib.ptr(out.get("raw")), _lib.ptr(out.get("sym")),
                                              _lib.ptr(out.get("iou")), _lib.ptr(out.get("offset")), _lib.ptr(out.get("centers")),
                                              _lib.current_stream(dev)))
    return out


def get_raw_quat_distance(q0: torch.Tensor, q1: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:434-455 -- ``[n]`` float32; ``tensor([nan])`` for no data."""
    if q0.shape[0] == 0:
        return torch.tensor([float("nan")], device=q0.device)
    return _pose_errors(q0, q1, want_raw=True)["raw"]


def get_symmetric_quat_distance(q0: torch.Tensor, q1: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:457-476 -- ``[n]`` float64: the smallest distance over 360 rotations of q1 about y."""
    if q0.shape[0] == 0:
        return torch.tensor([float("nan")], device=q0.device)
    return _pose_errors(q0, q1, want_sym=True)["sym"]


def get_quat_distance(q0, q1, symmetric_ids=None):
    """lib/gpu_tensor_funcs.py:411-432 -- non-symmetric pairs first, then symmetric ones, NaNs dropped (the reference's
    output order, not the input order)."""
    if symmetric_ids is None:
        return get_raw_quat_distance(q0, q1)
    if q0.shape[0] == 0:
        return torch.zeros((0,), dtype=torch.float32, device=q0.device)
    res = _pose_errors(q0, q1, symmetric_ids, want_raw=True, want_sym=True)
    is_sym = symmetric_ids != 0
    nan = torch.tensor([float("nan")], device=q0.device)
    plain = res["raw"][~is_sym] if bool((~is_sym).any()) else nan
    sym = res["sym"][is_sym] if bool(is_sym.any()) else nan
    d = torch.cat((plain, sym), dim=0)
    return d[~torch.isnan(d)]


def get_3d_ious(RTs_1, RTs_2, scales_1, scales_2) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:540-549 (+ :503-530) -- ``[n]`` float32, one launch instead of a Python loop with two 4x4
    inversions per pair."""
    if RTs_1.shape[0] == 0:
        return torch.stack([])          # what the reference's torch.stack([]) does: raises
    return _pose_errors(rts=(RTs_1, RTs_2), scales=(scales_1, scales_2))["iou"]


def get_3d_iou(RT_1, RT_2, scales_1, scales_2) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:532-538 (symmetry flag off, as in the reference) -- a 0-d tensor."""
    return get_3d_ious(RT_1.unsqueeze(0), RT_2.unsqueeze(0), scales_1.unsqueeze(0), scales_2.unsqueeze(0))[0]


def from_Ts_get_offset_error(gt_Ts, pred_Ts) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:565-567 -- ``|gt - pred| * 10`` per pair."""
    if gt_Ts.shape[0] == 0:
        return torch.zeros((0,), dtype=torch.float32, device=gt_Ts.device)
    return _pose_errors(ts=(gt_Ts, pred_Ts))["offset"]


def from_RTs_get_T_offset_errors(gt_RTs, pred_RTs) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:569-609 -- the camera-frame origin mapped to world coordinates by inverse(RT), ground truth
    against prediction.  As in the reference the distance is taken over ALL pairs at once (``get_offset_error_from_centroid``
    sums the squared differences of the whole [n,3] arrays), so the result is ONE number (times 10), not one per pair."""
    if gt_RTs.shape[0] == 0:
        return torch.stack([])                      # the reference's torch.stack([]) raises here too
    c = _pose_errors(rts=(gt_RTs, pred_RTs), want_centers=True)["centers"]
    return torch.sqrt(torch.sum(torch.pow(c[:, :3] - c[:, 3:], 2))) * 10


def calculate_aps(raw_data, metrics_threshold, metrics_operator):
    """lib/gpu_tensor_funcs.py:611-652 -- per metric and class the fraction of non-NaN values passing each threshold
    (``torch.less`` / ``torch.lt`` or ``torch.greater`` / ``torch.gt``), plus the mean over classes."""
    ops = {torch.less: 0, torch.lt: 0, torch.greater: 1, torch.gt: 1}
    aps = {}
    for key, per_class in raw_data.items():
        if metrics_operator[key] not in ops:
            raise NotImplementedError("calculate_aps: operator must be torch.less or torch.greater")
        aps[key] = {}
        for class_id, values in per_class.items():
            v = _lib.require_cuda(values.double().contiguous(), f"raw_data[{key}][{class_id}]")
            t = metrics_threshold[key].to(v.device).double().contiguous()
            out = torch.empty((t.shape[0],), dtype=torch.float32, device=v.device)
            with torch.cuda.device(v.device):
                _lib.check(_lib.lib().fpc_threshold_fraction(v.data_ptr(), int(v.shape[0]), t.data_ptr(), int(t.shape[0]),
                                                             ops[metrics_operator[key]], out.data_ptr(), _lib.current_stream(v.device)))
            aps[key][class_id] = out
        aps[key]["mean"] = torch.mean(torch.stack(list(aps[key].values())).float(), dim=0)
    return aps
