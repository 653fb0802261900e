"""CPU: the C-ABI library loads and exports every symbol include/fpc_b200.h declares; host-only argument
checks return error codes (never exit); the Python layer refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fpc_b200.h")).read()
    return sorted(set(re.findall(r"FPC_API\s+[\w\s\*]+?\b(fpc_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    from fastposecnn_b200 import _lib
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"libfpc_b200.so does not export {s}"
    assert sorted(_lib.EXPORTS) == syms
    assert L.fpc_version() == 100


def test_struct_layout_and_workspace_query():
    from fastposecnn_b200 import _lib
    L = _lib.lib()
    a = _lib.RecoverArgs()
    a.b, a.h, a.w, a.num_classes, a.hn = 2, 48, 64, 7, 32
    a.max_instances, a.max_records, a.max_rows = 64, 2 * 48 * 64, 2 * 48 * 64
    n = L.fpc_pose_recover_workspace_bytes(ctypes.byref(a))
    assert n > 2 * 48 * 64 * 9 and n % 256 == 0
    a.hn = 0
    assert L.fpc_pose_recover_workspace_bytes(ctypes.byref(a)) == 0
    assert b"hn must be positive" in L.fpc_last_error()


def test_error_codes_not_exit():
    from fastposecnn_b200 import _lib
    L = _lib.lib()
    assert L.fpc_normalize(None, None, -1, 2, 3, None) == _lib.FPC_EINVAL
    assert L.fpc_generate_hypothesis(None, None, None, None, 5, 1, 4, 0, None) == _lib.FPC_EINVAL
    assert L.fpc_generate_hypothesis(None, None, None, None, 5, 1, 0, 0, None) == _lib.FPC_OK        # empty: nothing to do
    assert L.fpc_get_rt(None, None, None, None, None, None, None, 0, None) == _lib.FPC_OK
    assert L.fpc_pose_recover(None) == _lib.FPC_EINVAL
    with pytest.raises(RuntimeError, match="libfpc_b200 error -1"):
        _lib.check(L.fpc_pose_recover(None))
    assert L.fpc_pack_masks(None, 0, -1, 4, 4, None, None, None) == _lib.FPC_EINVAL
    assert L.fpc_pack_masks(None, 0, 0, 4, 4, None, None, None) == _lib.FPC_OK                       # no masks: nothing to do
    assert L.fpc_pack_masks(None, 0, 3, 4, 4, None, None, None) == _lib.FPC_EINVAL
    assert L.fpc_mask_iou(None, None, 0, None, None, 5, 4, 4, None, None) == _lib.FPC_OK
    assert L.fpc_match_instances(None, None, None, 2, None, None, None, 2, 4, 4, None, None, None, None, None) == _lib.FPC_EINVAL
    assert L.fpc_pose_recover_num_launches() == 16
    assert L.fpc_pose_recover_kernel_name(13) == b"k_vote"


def test_no_cpu_fallback():
    import fastposecnn_b200 as fp
    x = torch.randn(2, 4, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        fp.normalize(x, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        fp.ransac_voting_layer_v3(torch.ones(1, 8, 8), torch.randn(1, 8, 8, 1, 2), 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        fp.PoseRecoveryEngine(1, 8, 8, 7, 16, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        fp.batchwise_get_2d_iou(torch.ones(2, 8, 8), torch.ones(3, 8, 8))
    agg = {"class_ids": torch.ones(1, dtype=torch.int64), "instance_masks": torch.ones(1, 8, 8)}
    with pytest.raises(RuntimeError, match="CUDA"):
        fp.batchwise_find_matches(agg, agg)
    logits = fp.synthetic.render_heads([[(4, 4, 2, 1)]], 8, 8)
    with pytest.raises(RuntimeError):
        fp.class_compression(logits, 7)


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "fastposecnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert not re.search(r"^\s*(from|import)\s+scipy\b", text, re.M), f
