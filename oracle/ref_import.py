"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Imports the reference's OWN Python modules, unmodified, from where they lie
(``/root/reference/source_code/FastPoseCNN/lib``) so that ``oracle/port.py`` can
be pinned against them and golden fixtures can be generated
(``oracle/make_golden.py``).  Only possible in the build container: the GPU box
has no ``/root/reference``, so nothing that runs there may import this module.

What is stubbed, and why it does not change the path's behaviour on CPU tensors:

* ``cupy``, ``cupyx(.scipy.ndimage)`` -- only used on the CUDA branch of
  ``batchwise_break_segmentation_mask`` (aggregation_layer.py:163-172); CPU
  tensors take the scipy branch (:174-181).
* ``skimage``, ``matplotlib`` -- visualisation imports at module top.
* ``ransac_voting_gpu_layer.ransac_voting`` -- the pybind/CUDA extension
  (src/ransac_voting.cpp) cannot build against torch 2.11 (``THCState``) and
  there is no GPU here; ``oracle/native.py`` supplies the CPU restatement of its
  two kernels behind the same two function names.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from contextlib import contextmanager

import torch

from . import native

_SRC_LIB = "/root/reference/source_code/FastPoseCNN/lib"
# On the GPU box there is no /root/reference: `__graft_entry__.build()` installs the reference's own, unmodified Python
# modules of this path under the git-ignored baseline/_ref/ (which travels with the gpurun snapshot), so that the reference
# arm of bench.py runs the reference's code there too.
_INSTALLED_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "FastPoseCNN", "lib")
REFERENCE_LIB = _SRC_LIB if os.path.isdir(_SRC_LIB) else _INSTALLED_LIB
MODULES = ("gpu_tensor_funcs.py", "aggregation_layer.py", "hough_voting.py", "matching.py", "metrics.py", "type_hinting.py",
           os.path.join("ransac_voting_gpu_layer", "ransac_voting_gpu.py"))


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_LIB, "aggregation_layer.py"))


def install(dst_lib: str = _INSTALLED_LIB) -> bool:
    """Copies the reference modules the path needs, byte for byte, into baseline/_ref (build container only)."""
    import shutil
    if not os.path.isdir(_SRC_LIB):
        return False
    for rel in MODULES:
        dst = os.path.join(dst_lib, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(_SRC_LIB, rel), dst)
    return True


_loaded = None


def load():
    """Returns a namespace with the reference modules: gtf, agg, hv, rvg, mg (matching), metrics."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference sources are not present on this machine")

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    stub("cupy")
    cpx = stub("cupyx")
    cpx_scipy = stub("cupyx.scipy")
    cpx_nd = stub("cupyx.scipy.ndimage")
    cpx.scipy = cpx_scipy
    cpx_scipy.ndimage = cpx_nd
    sk = stub("skimage")
    sk.io = stub("skimage.io")
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")

    pkg = stub("ransac_voting_gpu_layer")
    pkg.__path__ = [os.path.join(REFERENCE_LIB, "ransac_voting_gpu_layer")]
    native_mod = types.ModuleType("ransac_voting_gpu_layer.ransac_voting")
    native_mod.generate_hypothesis = native.ransac_voting.generate_hypothesis
    native_mod.voting_for_hypothesis = native.ransac_voting.voting_for_hypothesis
    native_mod.generate_hypothesis_vanishing_point = native.ransac_voting.generate_hypothesis_vanishing_point
    native_mod.voting_for_hypothesis_vanishing_point = native.ransac_voting.voting_for_hypothesis_vanishing_point
    sys.modules["ransac_voting_gpu_layer.ransac_voting"] = native_mod
    pkg.ransac_voting = native_mod

    # the reference does sys.path.append(os.getenv("TOOLS_DIR")); give it a harmless dir
    os.environ.setdefault("TOOLS_DIR", "/nonexistent-tools-dir")
    if REFERENCE_LIB not in sys.path:
        sys.path.insert(0, REFERENCE_LIB)

    ns = types.SimpleNamespace()
    ns.gtf = importlib.import_module("gpu_tensor_funcs")
    ns.rvg = importlib.import_module("ransac_voting_gpu_layer.ransac_voting_gpu")
    ns.hv = importlib.import_module("hough_voting")
    ns.agg = importlib.import_module("aggregation_layer")
    ns.mg = importlib.import_module("matching")
    # lib/metrics.py subclasses pl.metrics.Metric (pytorch-lightning is not installed): a base with the two things it uses
    pl = stub("pytorch_lightning")

    class _Metric:
        def __init__(self, *a, **k):
            pass

        def add_state(self, name, default, dist_reduce_fx=None):
            setattr(self, name, default)
    pl.metrics = types.SimpleNamespace(Metric=_Metric)
    ns.metrics = importlib.import_module("metrics")
    _loaded = ns
    return ns


@contextmanager
def fixed_idxs(idx_source, hn: int, vn: int = 1):
    """While active, the reference's in-place draw
    ``torch.zeros([hn,vn,2], int32).random_(0, tn)`` (ransac_voting_gpu.py:552)
    is answered from ``idx_source(instance_counter, hn, vn, tn)`` instead of the RNG."""
    original = torch.Tensor.random_
    counter = {"i": 0}

    def patched(self, *args, **kwargs):
        if self.dtype == torch.int32 and self.dim() == 3 and tuple(self.shape) == (hn, vn, 2) and len(args) == 2:
            lo, hi = args
            assert lo == 0
            self.copy_(idx_source(counter["i"], hn, vn, int(hi)))
            counter["i"] += 1
            return self
        return original(self, *args, **kwargs)

    torch.Tensor.random_ = patched
    try:
        yield
    finally:
        torch.Tensor.random_ = original


@contextmanager
def legacy_uint8_masks():
    """torch <= 1.x let ``masked_select`` take a uint8 mask (the reference's v4/v5/v6 build theirs with ``.byte()``,
    ransac_voting_gpu.py:691,783); torch 2 insists on bool.  While active, uint8 masks are converted -- same selection."""
    original = torch.Tensor.masked_select

    def patched(self, mask):
        return original(self, mask.bool() if mask.dtype == torch.uint8 else mask)

    torch.Tensor.masked_select = patched
    try:
        yield
    finally:
        torch.Tensor.masked_select = original


class _HP:
    """Anything with the attributes the path reads (config.py:80-82,93)."""

    def __init__(self, hn):
        self.HV_NUM_OF_HYPOTHESES = hn
        self.PERFORM_AGGREGATION = True
        self.PERFORM_HOUGH_VOTING = True
        self.PERFORM_RT_CALCULATION = True


def reference_pose_recover(logits, inv_intrinsics, hn: int, idx_source, num_of_classes=None):
    """The reference's own code for lib/pose_regressor.py:445-504 on CPU tensors."""
    ref = load()
    if num_of_classes is None:
        num_of_classes = logits["mask"].shape[1]
    cat_mask = torch.argmax(torch.nn.LogSoftmax(dim=1)(logits["mask"]), dim=1)      # pose_regressor.py:449
    cat = ref.gtf.class_compress(num_of_classes, cat_mask, logits)
    cat.update({"mask": cat_mask})
    layer = ref.agg.AggregationLayer(_HP(hn), num_of_classes)
    agg = layer.forward(cat)
    voter = ref.hv.HoughVotingLayer(_HP(hn))
    with fixed_idxs(idx_source, hn):
        agg = voter(agg)
    agg = ref.gtf.samplewise_get_RT(agg, inv_intrinsics)
    return cat, agg
