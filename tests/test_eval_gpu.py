"""GPU: evaluation maths drop-ins against the oracle restatements (pinned bit-for-bit to the
imported reference in tests/test_eval_oracle.py): <= 1e-4 relative, AP fractions exact."""
import pytest
import torch

from helpers import port
from test_eval_oracle import pairs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(a, b, tol=1e-4):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape
    return float(((a - b).abs() / b.abs().clamp_min(1e-6)).max()) <= tol


def test_quaternion_distances():
    from fastposecnn_b200 import gpu_tensor_funcs as gtf
    q0, q1, sym, *_ = pairs(m=100, seed=1)
    g0, g1, gs = q0.to(DEV), q1.to(DEV), sym.to(DEV)
    raw = gtf.get_raw_quat_distance(g0, g1)
    assert raw.dtype == torch.float32 and close(raw, port.get_raw_quat_distance(q0, q1))
    s = gtf.get_symmetric_quat_distance(g0, g1)
    assert s.dtype == torch.float64 and close(s, port.get_symmetric_quat_distance(q0, q1), 1e-9)
    d, want = gtf.get_quat_distance(g0, g1, gs), port.get_quat_distance(q0, q1, sym)
    assert d.dtype == want.dtype and close(d, want)
    zero = torch.zeros_like(gs)
    d0 = gtf.get_quat_distance(g0, g1, zero)
    assert d0.dtype == torch.float32 and close(d0, port.get_quat_distance(q0, q1, torch.zeros_like(sym)))
    assert close(gtf.get_quat_distance(g0, g1), port.get_quat_distance(q0, q1))
    assert torch.isnan(gtf.get_raw_quat_distance(g0[:0], g1[:0])).all()
    # opposite quaternions are the same rotation: distance 0
    assert float(gtf.get_raw_quat_distance(g0, -g0).abs().max()) == 0.0


def test_3d_iou_offsets_and_aps():
    from fastposecnn_b200 import gpu_tensor_funcs as gtf
    q0, q1, sym, rt0, rt1, s0, s1, t0, t1 = pairs(m=64, seed=3)
    want = port.get_3d_ious(rt0, rt1, s0, s1)
    got = gtf.get_3d_ious(rt0.to(DEV), rt1.to(DEV), s0.to(DEV), s1.to(DEV))
    assert got.dtype == torch.float32 and got.shape == (64,)
    assert float((got.cpu() - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-7
    one = gtf.get_3d_iou(rt0[5].to(DEV), rt1[5].to(DEV), s0[5].to(DEV), s1[5].to(DEV))
    assert one.dim() == 0 and abs(float(one) - float(want[5])) <= 1e-4 * float(want.abs().max()) + 1e-7
    same = gtf.get_3d_ious(rt0.to(DEV), rt0.to(DEV), s0.to(DEV), s0.to(DEV))
    assert float((same - 1).abs().max()) < 1e-5                                   # a box against itself
    assert close(gtf.from_Ts_get_offset_error(t0.to(DEV), t1.to(DEV)), port.from_Ts_get_offset_error(t0, t1))
    deg = port.get_raw_quat_distance(q0, q1)
    raw = {"degree_error": {1: deg[:30], 2: deg[30:]}, "iou_3d": {1: want[:30], 2: torch.cat((want[30:], torch.tensor([float("nan")])))}}
    thr = {"degree_error": torch.linspace(0, 60, 7), "iou_3d": torch.linspace(0, 1, 5)}
    ops = {"degree_error": torch.less, "iou_3d": torch.greater}
    ref = port.calculate_aps(raw, thr, ops)
    got = gtf.calculate_aps({k: {c: v.to(DEV) for c, v in d.items()} for k, d in raw.items()}, thr, ops)
    for k in ref:
        for c in ref[k]:
            assert got[k][c].dtype == ref[k][c].dtype and torch.equal(got[k][c].cpu(), ref[k][c]), (k, c)
