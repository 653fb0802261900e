"""CPU, build container only: the oracle restatements of the evaluation maths against the
reference's own gpu_tensor_funcs.py imported unmodified."""
import pytest
import torch

from helpers import port
from oracle import ref_import

pytestmark = [pytest.mark.skipif(not ref_import.available(), reason="reference sources not on this machine"),
              pytest.mark.filterwarnings("ignore")]


def pairs(m=40, seed=0):
    g = torch.Generator().manual_seed(seed)
    q0 = torch.nn.functional.normalize(torch.randn(m, 4, generator=g), dim=1)
    q1 = torch.nn.functional.normalize(q0 + 0.3 * torch.randn(m, 4, generator=g), dim=1)
    sym = (torch.rand(m, generator=g) > 0.5).long()
    inv_k = torch.inverse(torch.tensor([[577.5, 0, 319.5], [0, 577.5, 239.5], [0, 0, 1.0]]))
    xy0 = torch.rand(m, 2, generator=g) * torch.tensor([640.0, 480.0])
    z0 = 800 + 400 * torch.rand(m, 1, generator=g)
    _, t0, rt0 = port.batchwise_get_RT(q0, xy0, z0, inv_k)
    _, t1, rt1 = port.batchwise_get_RT(q1, xy0 + 3 * torch.randn(m, 2, generator=g), z0 + 20 * torch.randn(m, 1, generator=g), inv_k)
    s0 = 0.2 + torch.rand(m, 3, generator=g)
    s1 = s0 * (1 + 0.1 * torch.randn(m, 3, generator=g))
    return q0, q1, sym, rt0, rt1, s0, s1, t0, t1


def test_quaternion_distances_equal_reference():
    ref = ref_import.load()
    q0, q1, sym, *_ = pairs()
    assert torch.equal(port.get_raw_quat_distance(q0, q1), ref.gtf.get_raw_quat_distance(q0, q1))
    if hasattr(ref.gtf.quat_symmetric_tf, "rot_q"):
        del ref.gtf.quat_symmetric_tf.rot_q
    a, b = port.get_symmetric_quat_distance(q0, q1), ref.gtf.get_symmetric_quat_distance(q0, q1)
    assert a.dtype == b.dtype == torch.float64 and torch.equal(a, b)
    a, b = port.get_quat_distance(q0, q1, sym), ref.gtf.get_quat_distance(q0, q1, sym)
    assert a.dtype == b.dtype and torch.equal(a, b)
    none_sym = torch.zeros_like(sym)
    a, b = port.get_quat_distance(q0, q1, none_sym), ref.gtf.get_quat_distance(q0, q1, none_sym)
    assert a.dtype == b.dtype == torch.float32 and torch.equal(a, b)
    assert torch.isnan(port.get_raw_quat_distance(q0[:0], q1[:0])).all()


def test_3d_iou_offsets_and_aps_equal_reference():
    ref = ref_import.load()
    q0, q1, sym, rt0, rt1, s0, s1, t0, t1 = pairs(m=24, seed=3)
    a, b = port.get_3d_ious(rt0, rt1, s0, s1), ref.gtf.get_3d_ious(rt0, rt1, s0, s1)
    assert a.shape == b.shape == (24,) and torch.equal(a, b)
    assert float(a.max()) > 0.05
    assert torch.equal(port.from_Ts_get_offset_error(t0, t1), ref.gtf.from_Ts_get_offset_error(t0, t1))
    raw = {"degree_error": {1: port.get_raw_quat_distance(q0[:10], q1[:10]), 2: port.get_raw_quat_distance(q0[10:], q1[10:])},
           "iou_3d": {1: a[:10], 2: torch.cat((a[10:], torch.tensor([float("nan")])))}}
    thr = {"degree_error": torch.linspace(0, 60, 7), "iou_3d": torch.linspace(0, 1, 5)}
    ops = {"degree_error": torch.less, "iou_3d": torch.greater}
    x, y = port.calculate_aps(raw, thr, ops), ref.gtf.calculate_aps(raw, thr, ops)
    for k in y:
        for c in y[k]:
            assert torch.equal(x[k][c], y[k][c]), (k, c)
