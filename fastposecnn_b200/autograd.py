"""Training support: ``torch.autograd.Function`` shells around the forward kernels, with hand-written backward kernels
(``csrc/fpc_backward.cu``).  The reference trains its aggregated losses (lib/loss.py:155-545) through plain torch ops; with
these, ``class_compress`` / ``class_compression`` and ``AggregationLayer.forward`` of the drop-in stay differentiable with
respect to the regression head maps (quaternion, scales, xy, z).  The mask head is trained by its own pixel-wise losses
(lib/loss.py:26-101); the arg-max that selects classes has no gradient in the reference either."""
from __future__ import annotations

import torch

from . import _lib


class ClassCompressFn(torch.autograd.Function):
    """(quaternion, scales, xy, z raw head maps) -> class-compressed fields; ``run`` does the forward launch."""

    @staticmethod
    def forward(ctx, run, num_of_classes, quat, scales, xy, z):
        out = run()
        cat = out["mask"] if "mask" in out else out["_cat_mask"]
        ctx.save_for_backward(cat, quat, xy)
        ctx.num_of_classes = num_of_classes
        ctx.shapes = (quat.shape, scales.shape, xy.shape, z.shape)
        ctx.mark_non_differentiable(cat)
        return out["quaternion"], out["scales"], out["xy"], out["z"], cat

    @staticmethod
    def backward(ctx, g_q, g_s, g_xy, g_z, _g_cat):
        cat, quat, xy = ctx.saved_tensors
        b, h, w = cat.shape
        dev = cat.device
        f32 = torch.float32

        def prep(g):
            return None if g is None else g.to(f32).contiguous()
        g_q, g_s, g_xy, g_z = prep(g_q), prep(g_s), prep(g_xy), prep(g_z)
        d = [torch.empty(s, dtype=f32, device=dev) for s in ctx.shapes]
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fpc_class_compress_backward(
                cat.data_ptr(), quat.data_ptr(), xy.data_ptr(), _lib.ptr(g_q), _lib.ptr(g_s), _lib.ptr(g_xy), _lib.ptr(g_z),
                d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), b, ctx.num_of_classes, h, w,
                _lib.current_stream(dev)))
        return None, None, d[0], d[1], d[2], d[3]


class AggregateFn(torch.autograd.Function):
    """Class-compressed fields -> per-instance (quaternion, scales, z) and the masked dense xy; ``run`` does the forward
    launches and returns (agg dict, labels [b,h,w] i32, pixel counts [n] i32, |mean quaternion| [n])."""

    @staticmethod
    def forward(ctx, run, q, s, xy, z):
        agg, labels, counts, qnorm = run()
        ctx.save_for_backward(labels, counts, qnorm, agg["quaternion"], agg["z"])
        ctx.shape = tuple(labels.shape)
        ctx.extra = {k: v for k, v in agg.items() if k not in ("quaternion", "scales", "z", "xy")}
        return agg["quaternion"], agg["scales"], agg["z"], agg["xy"]

    @staticmethod
    def backward(ctx, g_q, g_s, g_z, g_xy):
        labels, counts, qnorm, q_hat, z_val = ctx.saved_tensors
        b, h, w = ctx.shape
        dev, f32 = labels.device, torch.float32
        n = int(counts.shape[0])
        inv_c = (1.0 / counts.to(f32).clamp_min(1)).unsqueeze(1)
        G = torch.zeros((n, 8), dtype=f32, device=dev)
        if g_q is not None:
            g = g_q.to(f32)
            proj = (g - q_hat * (q_hat * g).sum(dim=1, keepdim=True)) / qnorm.unsqueeze(1).clamp_min(1e-30)
            G[:, 0:4] = torch.where(qnorm.unsqueeze(1) != 0, proj, g) * inv_c       # zero-norm guard of normalize(): divide by 1
        if g_s is not None:
            G[:, 4:7] = g_s.to(f32) * inv_c
        if g_z is not None:
            G[:, 7:8] = g_z.to(f32).reshape(n, 1) * z_val.reshape(n, 1) * inv_c    # z = exp(mean)
        g_xy = None if g_xy is None else g_xy.to(f32).contiguous()
        d_q = torch.empty((b, 4, h, w), dtype=f32, device=dev)
        d_s = torch.empty((b, 3, h, w), dtype=f32, device=dev)
        d_xy = torch.empty((b, 2, h, w), dtype=f32, device=dev)
        d_z = torch.empty((b, h, w), dtype=f32, device=dev)
        G = G.contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fpc_aggregate_backward(labels.data_ptr(), G.data_ptr(), _lib.ptr(g_xy), n, d_q.data_ptr(),
                                                         d_s.data_ptr(), d_xy.data_ptr(), d_z.data_ptr(), b, h, w,
                                                         _lib.current_stream(dev)))
        return None, d_q, d_s, d_xy, d_z


class PoseRecoverFn(torch.autograd.Function):
    """The fused path as one differentiable node: raw head maps -> per-instance (quaternion, scales, z, voted centre xy).
    Backward = the kernels above chained: instance gradient -> class-compressed fields -> predicted class's channels of the
    raw maps for q / scales / z; the refinement-solve backward on the label volume for xy (inlier set held constant)."""

    @staticmethod
    def forward(ctx, run, num_of_classes, inlier_thresh, arith, quat, scales, z, xy_head):
        agg, labels, cat_u8, counts, qnorm = run()
        ctx.save_for_backward(labels, cat_u8, counts, qnorm, agg["quaternion"], agg["z"], quat, xy_head, agg["xy"], agg["win_hypothesis"],
                              agg["tn"].to(torch.int32), agg["sample_ids"].to(torch.int32))
        ctx.num_of_classes, ctx.inlier_thresh, ctx.arith = num_of_classes, float(inlier_thresh), int(arith)
        ctx.shapes = (quat.shape, scales.shape, z.shape)
        return agg["quaternion"], agg["scales"], agg["z"], agg["xy"]

    @staticmethod
    def backward(ctx, g_q, g_s, g_z, g_xy):
        labels, cat_u8, counts, qnorm, q_hat, z_val, quat, xy_head, refined, win, tn, frame_of = ctx.saved_tensors
        b, h, w = labels.shape
        dev, f32 = labels.device, torch.float32
        n = int(counts.shape[0])
        L, st = _lib.lib(), _lib.current_stream(dev)
        inv_c = (1.0 / counts.to(f32).clamp_min(1)).unsqueeze(1)
        G = torch.zeros((n, 8), dtype=f32, device=dev)
        if g_q is not None:
            g = g_q.to(f32)
            proj = (g - q_hat * (q_hat * g).sum(dim=1, keepdim=True)) / qnorm.unsqueeze(1).clamp_min(1e-30)
            G[:, 0:4] = torch.where(qnorm.unsqueeze(1) != 0, proj, g) * inv_c
        if g_s is not None:
            G[:, 4:7] = g_s.to(f32) * inv_c
        if g_z is not None:
            G[:, 7:8] = g_z.to(f32).reshape(n, 1) * z_val.reshape(n, 1) * inv_c
        c_q = torch.empty((b, 4, h, w), dtype=f32, device=dev)
        c_s = torch.empty((b, 3, h, w), dtype=f32, device=dev)
        c_xy = torch.empty((b, 2, h, w), dtype=f32, device=dev)
        c_z = torch.empty((b, h, w), dtype=f32, device=dev)
        d_q, d_s, d_z = (torch.empty(sh, dtype=f32, device=dev) for sh in ctx.shapes)
        K = ctx.num_of_classes - 1
        d_xy_unused = torch.empty((b, 2 * K, h, w), dtype=f32, device=dev)
        d_xy = torch.empty((b, 2 * K, h, w), dtype=f32, device=dev)
        cat = cat_u8.to(torch.int64)
        gx = torch.zeros((n, 2), dtype=f32, device=dev) if g_xy is None else g_xy.to(f32).contiguous()
        live = (tn > 0).to(torch.int32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(L.fpc_aggregate_backward(labels.data_ptr(), G.contiguous().data_ptr(), None, n, c_q.data_ptr(), c_s.data_ptr(),
                                                c_xy.data_ptr(), c_z.data_ptr(), b, h, w, st))
            # the fused forward normalises each pixel's quaternion before averaging: same Jacobian as class_compress
            _lib.check(L.fpc_class_compress_backward(cat.data_ptr(), quat.data_ptr(), None, c_q.data_ptr(), c_s.data_ptr(), None,
                                                     c_z.data_ptr(), d_q.data_ptr(), d_s.data_ptr(), d_xy_unused.data_ptr(),
                                                     d_z.data_ptr(), b, ctx.num_of_classes, h, w, st))
            _lib.check(L.fpc_pose_recover_xy_backward(labels.data_ptr(), cat_u8.data_ptr(), xy_head.data_ptr(), frame_of.data_ptr(),
                                                      win.contiguous().data_ptr(), refined.contiguous().data_ptr(), gx.data_ptr(),
                                                      live.data_ptr(), ctx.inlier_thresh, n, b, ctx.num_of_classes, h, w, ctx.arith,
                                                      d_xy.data_ptr(), st))
        return None, None, None, None, d_q, d_s, d_z, d_xy


class RansacV3Fn(torch.autograd.Function):
    """``ransac_voting_layer_v3`` differentiable w.r.t. ``vertex``: the refinement solve over the winner's inliers
    (ransac_voting_gpu.py:584-598) is differentiated with the inlier set held constant, as autograd does in the reference
    (the set enters there as a 0/1 weight).  ``run`` does the forward and returns (points [n,vn,2], per-keypoint details)."""

    @staticmethod
    def forward(ctx, run, fmask, vertex, inlier_thresh, arith):
        pts, det = run()
        ctx.save_for_backward(fmask, vertex, pts.detach(), torch.stack([d["best_pts"] for d in det], dim=1),
                              torch.stack([d["tn"] for d in det], dim=1).to(torch.int32))
        ctx.inlier_thresh, ctx.arith = float(inlier_thresh), int(arith)
        return pts

    @staticmethod
    def backward(ctx, g_pts):
        fmask, vertex, pts, win, tn = ctx.saved_tensors
        n, h, w, vn, _ = vertex.shape
        dev, f32 = vertex.device, torch.float32
        grad = torch.empty((n, h, w, vn, 2), dtype=f32, device=dev)
        tmp = torch.empty((n, h, w, 2), dtype=f32, device=dev)
        g_pts = g_pts.to(f32)
        for vi in range(vn):
            v = vertex[..., vi, :]
            sN, sH, sW, s2 = v.stride()
            wp, rp, gx = win[:, vi].contiguous(), pts[:, vi].contiguous(), g_pts[:, vi].contiguous()
            live = (tn[:, vi] > 0).to(torch.int32).contiguous()
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().fpc_vote_refine_backward(fmask.data_ptr(), v.data_ptr(), sN, sH, sW, s2, wp.data_ptr(),
                                                               rp.data_ptr(), gx.data_ptr(), live.data_ptr(), ctx.inlier_thresh, n, h, w,
                                                               ctx.arith, tmp.data_ptr(), _lib.current_stream(dev)))
            grad[..., vi, :] = tmp
        return None, None, grad, None, None


class GetRTFn(torch.autograd.Function):
    """``batchwise_get_RT`` differentiable w.r.t. (q, xys, exp_zs); ``run`` does the forward launch."""

    @staticmethod
    def forward(ctx, run, q, xys, exp_zs, inv_k):
        R, T, RT = run()
        ctx.save_for_backward(q, xys, exp_zs, inv_k)
        return R, T, RT

    @staticmethod
    def backward(ctx, g_R, g_T, g_RT):
        q, xys, exp_zs, inv_k = ctx.saved_tensors
        n, dev, f32 = q.shape[0], q.device, torch.float32

        def prep(g):
            return None if g is None else g.to(f32).contiguous()
        g_R, g_T, g_RT = prep(g_R), prep(g_T), prep(g_RT)
        d_q = torch.empty((n, 4), dtype=f32, device=dev)
        d_xy = torch.empty((n, 2), dtype=f32, device=dev)
        d_z = torch.empty(exp_zs.shape, dtype=f32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fpc_get_rt_backward(q.data_ptr(), xys.data_ptr(), exp_zs.data_ptr(), inv_k.data_ptr(), _lib.ptr(g_R),
                                                      _lib.ptr(g_T), _lib.ptr(g_RT), n, d_q.data_ptr(), d_xy.data_ptr(), d_z.data_ptr(),
                                                      _lib.current_stream(dev)))
        return None, d_q, d_xy, d_z, None
