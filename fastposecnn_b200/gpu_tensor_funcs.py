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
    out = {"quaternion": q_out, "scales": s_out, "xy": xy_out, "z": z_out}
    if cat_out is not None:
        out["mask"] = cat_out
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
    R = torch.empty((n, 3, 3), dtype=f32, device=q.device)
    T = torch.empty((n, 3), dtype=f32, device=q.device)
    RT = torch.empty((n, 4, 4), dtype=f32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.lib().fpc_get_rt(q.data_ptr(), xys.data_ptr(), exp_zs.data_ptr(), inv_k.data_ptr(),
                                         R.data_ptr(), T.data_ptr(), RT.data_ptr(), n, _lib.current_stream(q.device)))
    return R, T, RT


def samplewise_get_RT(agg_data: Dict[str, torch.Tensor], inv_intrinsics: torch.Tensor):
    """lib/gpu_tensor_funcs.py:237-253 -- adds 'R', 'T', 'RT' to agg_data."""
    R, T, RT = batchwise_get_RT(agg_data["quaternion"], agg_data["xy"], agg_data["z"], inv_intrinsics)
    agg_data["R"], agg_data["T"], agg_data["RT"] = R, T, RT
    return agg_data


def quats_2_rotation_matrix(q: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:306-326 for already normalised quaternions [n,4] -> [n,3,3]."""
    n = q.shape[0]
    zeros2 = torch.zeros((n, 2), dtype=torch.float32, device=q.device)
    ones = torch.ones((n, 1), dtype=torch.float32, device=q.device)
    eye = torch.eye(3, dtype=torch.float32, device=q.device)
    R, _, _ = batchwise_get_RT(q, zeros2, ones, eye)
    return R


def batchwise_get_2d_iou(batch_masks1: torch.Tensor, batch_masks2: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:386-409 -- IoU of every ``[n1,h,w]`` mask with every ``[n2,h,w]`` mask -> ``[n1,n2]``
    float32 (0/0 = NaN).  Each mask is read once and bit-packed; pairs cost popc(a & b) over the overlap of their
    bounding boxes instead of the reference's ``[n1,n2,h,w]`` logical volumes; values are bit-identical."""
    from .matching import mask_iou, pack_masks
    return mask_iou(pack_masks(batch_masks1), pack_masks(batch_masks2))


def torch_get_2d_iou(tensor1: torch.Tensor, tensor2: torch.Tensor) -> torch.Tensor:
    """lib/gpu_tensor_funcs.py:380-384 -- IoU of two ``[h,w]`` masks, a 0-d float32 tensor."""
    return batchwise_get_2d_iou(tensor1.reshape(1, *tensor1.shape[-2:]), tensor2.reshape(1, *tensor2.shape[-2:]))[0, 0]
