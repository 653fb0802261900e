"""lib/metrics.py row: the oracle fold (oracle/port.py:metric_values) pinned to the reference's own metric classes (imported
with a stub base class in place of pl.metrics.Metric); the drop-in classes against the oracle on the GPU."""
import pytest
import torch

from helpers import port
from oracle import ref_import
from test_eval_oracle import pairs


def batches():
    out = []
    for seed, m in ((0, 24), (1, 7), (2, 40)):
        q0, q1, sym, rt0, rt1, s0, s1, t0, t1 = pairs(m=m, seed=seed)
        out.append({"quaternion": torch.stack((q0, q1)), "symmetric_ids": sym, "RT": torch.stack((rt0, rt1)),
                    "scales": torch.stack((s0, s1)), "T": torch.stack((t0, t1))})
    out.insert(1, None)                                        # a step without matches
    return out


CASES = [("DegreeErrorMeanAP", "degree_ap", 10), ("DegreeError", "degree_error", None), ("Iou3dAP", "iou_ap", 0.25),
         ("Iou3dAccuracy", "iou_accuracy", None), ("OffsetAP", "offset_ap", 5), ("OffsetError", "offset_error", None)]


@pytest.mark.skipif(not ref_import.available(), reason="reference sources not on this machine")
@pytest.mark.filterwarnings("ignore")
@pytest.mark.parametrize("cls,kind,thr", CASES)
def test_oracle_fold_equals_reference_metric_classes(cls, kind, thr):
    ref = ref_import.load()
    metric = getattr(ref.metrics, cls)(thr) if thr is not None else getattr(ref.metrics, cls)()
    for b in batches():
        metric.update(b)
    want = metric.compute()
    got = port.metric_values(batches(), kind, thr)
    assert got.dtype == want.dtype and torch.equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("cls,kind,thr", CASES)
def test_drop_in_metric_classes(cls, kind, thr):
    from fastposecnn_b200 import metrics
    metric = getattr(metrics, cls)(thr) if thr is not None else getattr(metrics, cls)()
    for b in batches():
        metric.update(None if b is None else {k: v.to("cuda:0") for k, v in b.items()})
    got = metric.compute().cpu().double()
    want = port.metric_values(batches(), kind, thr).double()
    assert abs(float(got) - float(want)) <= 1e-4 * max(1.0, abs(float(want)))
    metric.reset()
    assert float(getattr(metric, "total", getattr(metric, "error", getattr(metric, "accuracy", torch.tensor(0))))) == 0.0
