"""GPU: BASELINE.json configs[4] in miniature -- a random-init torch encoder/decoder (examples/network_feed.py, not
part of the product) feeds the fused path; its irregular, noisy head maps are compared against the CPU oracle."""
import os
import sys

import pytest
import torch

import helpers
from helpers import port, syn

sys.path.insert(0, os.path.join(helpers.ROOT, "examples"))
pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_random_init_network_feeds_the_path():
    pytest.importorskip("torchvision")
    from network_feed import TorchFeeder
    from fastposecnn_b200.pose_recovery import PoseRecoveryEngine
    torch.manual_seed(0)
    net = TorchFeeder().to(DEV).eval()
    with torch.no_grad():
        logits = net(torch.randn(2, 3, 96, 128, device=DEV))
    # spread the mask logits so several classes win somewhere, and quantise them (argmax(log_softmax) == argmax)
    logits["mask"] = torch.round((logits["mask"] - logits["mask"].mean(dim=(2, 3), keepdim=True)) * 8 * 1024) / 1024
    cpu = {k: v.cpu() for k, v in logits.items()}
    hn = 32
    cat = port.class_compression(cpu, 7)
    agg = port.aggregate(cat)
    n = agg["class_ids"].shape[0]
    assert n >= 1
    eng = PoseRecoveryEngine(2, 96, 128, 7, hn, DEV, max_instances=n + 8, want_labels=True)
    eng.launch(logits, torch.inverse(syn.camera_intrinsics()).to(DEV))
    assert eng.fetch_count() == n
    out = eng.table_to_agg(n)
    lab_ref, _ = port.label_instances(cat["mask"] != 0)
    assert torch.equal(eng.cat_mask_u8.cpu().long(), cat["mask"])
    assert torch.equal(eng.labels.cpu(), lab_ref.to(torch.int32))
    assert torch.equal(out["class_ids"].cpu(), agg["class_ids"].long())
    assert torch.equal(out["sample_ids"].cpu(), agg["sample_ids"])
    assert out["mask_sizes"].cpu().tolist() == helpers.oracle_tns(agg)
    for k in ("quaternion", "scales", "z"):
        assert helpers.rel_err(out[k], agg[k]) <= helpers.REL_TOL, k
    assert torch.isfinite(out["RT"]).all() and torch.isfinite(out["xy"]).all()


def test_fused_head_epilogue_equals_upsampled_flow_on_network_outputs():
    """The same random-init network, two flows on the GPU: heads up-sampled by torch then the full-resolution path, and
    head convolutions only + the low-resolution path.  Class map, labels and every instance table must be identical."""
    pytest.importorskip("torchvision")
    from network_feed import TorchFeeder
    import fastposecnn_b200 as fp
    torch.manual_seed(1)
    net = TorchFeeder().to(DEV).eval()
    imgs = torch.randn(2, 3, 96, 128, device=DEV)
    with torch.no_grad():
        low = net.lowres(imgs)
        # spread the mask logits so that several classes win somewhere (at LOW resolution: both flows see the same values)
        low["mask"] = (low["mask"] - low["mask"].mean(dim=(2, 3), keepdim=True)) * 8
        full = {k: torch.nn.UpsamplingBilinear2d(scale_factor=4)(v) for k, v in low.items()}
    inv_k = torch.inverse(syn.camera_intrinsics()).to(DEV)
    a = {k: v.clone() for k, v in fp.pose_recover(full, inv_k, 32, seed=1234).items()}
    b = fp.pose_recover(low, inv_k, 32, upsample=4, seed=1234)
    assert a["class_ids"].shape[0] >= 1
    for k in ("cat_mask", "labels", "class_ids", "sample_ids", "mask_sizes", "quaternion", "scales", "z"):
        assert torch.equal(a[k], b[k]), k
