"""GPU: head-epilogue fusion.  The fused path fed with the heads' LOW-RESOLUTION outputs
must give what the reference flow gives: up-sample every head map (torch's nn.UpsamplingBilinear2d, the operator smp's
SegmentationHead uses), then the path.  Integer results bit-exact, poses <= 1e-4."""
import pytest
import torch

import helpers
from helpers import port, syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape,scale", [((2, 7, 24, 32), 4), ((1, 67, 30, 40), 4), ((1, 3, 120, 160), 4), ((2, 2, 50, 66), 2),
                                         ((1, 2, 40, 30), 3), ((3, 1, 17, 33), 2)])
def test_upsample_operator_is_bit_identical_to_torch(shape, scale):
    import fastposecnn_b200 as fp
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape)))
    got = fp.upsample_bilinear(x.to(DEV), scale)
    want_cuda = torch.nn.UpsamplingBilinear2d(scale_factor=scale)(x.to(DEV))          # ATen's CUDA kernel
    assert got.shape == want_cuda.shape and torch.equal(got, want_cuda)
    from oracle import native
    assert torch.equal(got.cpu(), native.upsample_bilinear(x, scale))                  # the oracle's C restatement
    if shape[-2] * shape[-1] * scale * scale >= 3200:                                  # ATen's vectorised CPU kernel
        assert torch.equal(got.cpu(), torch.nn.UpsamplingBilinear2d(scale_factor=scale)(x))


def _run_lowres(frames, h, w, scale, hn=48, seed=5, idx_seed=1234):
    import fastposecnn_b200 as fp
    low = syn.render_lowres_heads(frames, h, w, scale, seed=seed)
    inv_k = torch.inverse(syn.camera_intrinsics())
    details = []
    cat, agg = port.pose_recover_lowres(low, inv_k, hn, scale, idx_source=port.seeded_idx_source(idx_seed), details=details)
    idxs = syn.presampled_idxs(helpers.oracle_tns(agg), hn, seed=idx_seed).reshape(-1, hn, 2)
    out = fp.pose_recover({k: v.to(DEV) for k, v in low.items()}, inv_k.to(DEV), hn, idxs=idxs.to(DEV), upsample=scale)
    return cat, agg, details, out


SCENES = {
    "three_frames": ([[(30, 30, 14, 1), (90, 40, 18, 3), (60, 75, 12, 6)], [(40, 50, 20, 2)], []], 96, 128, 4),
    "touching": ([[(30, 40, 12, 5), (52, 40, 12, 2), (100, 60, 10, 4)]], 96, 128, 4),
    "wide_512": ([[(200, 60, 55, 3), (420, 64, 50, 1)]], 128, 512, 4),
    "scale2_scalar_width": ([[(25, 25, 11, 1), (70, 30, 13, 2)], [(45, 40, 16, 4)]], 70, 102, 2),   # w % 4 != 0
    "scale3": ([[(30, 30, 12, 1), (80, 50, 15, 5)]], 96, 120, 3),
    "scale8": ([[(60, 60, 25, 2), (160, 90, 30, 6)]], 160, 256, 8),
}


@pytest.mark.parametrize("name", list(SCENES.keys()))
def test_fused_lowres_path_vs_reference_flow(name):
    frames, h, w, scale = SCENES[name]
    cat, agg, details, out = _run_lowres(frames, h, w, scale)
    n = agg["class_ids"].shape[0]
    assert n > 0
    assert torch.equal(out["cat_mask"].cpu().long(), cat["mask"]), "class map differs"
    lab_ref, total = port.label_instances(cat["mask"] != 0)
    assert total == n and torch.equal(out["labels"].cpu(), lab_ref.to(torch.int32))
    assert torch.equal(out["class_ids"].cpu(), agg["class_ids"].long())
    assert torch.equal(out["sample_ids"].cpu(), agg["sample_ids"])
    assert out["mask_sizes"].cpu().tolist() == helpers.oracle_tns(agg)
    for key in ("quaternion", "scales", "z", "xy", "R", "T", "RT"):
        e = helpers.rel_err(out[key], agg[key])
        assert e <= helpers.REL_TOL, f"{key}: rel err {e:.3e}"


def test_lowres_equals_full_resolution_path_on_upsampled_maps():
    """Same device, same kernels downstream: interpolating inside the kernels == running the full-resolution path on
    maps up-sampled by torch's CUDA operator -- every table word identical (same fixed idxs)."""
    import fastposecnn_b200 as fp
    frames, h, w, scale = SCENES["three_frames"]
    low = {k: v.to(DEV) for k, v in syn.render_lowres_heads(frames, h, w, scale, seed=9).items()}
    inv_k = torch.inverse(syn.camera_intrinsics()).to(DEV)
    full = {k: torch.nn.UpsamplingBilinear2d(scale_factor=scale)(v) for k, v in low.items()}
    a = fp.pose_recover(full, inv_k, 64, seed=1234)
    n = a["class_ids"].shape[0]
    idxs = syn.presampled_idxs(a["mask_sizes"].cpu().tolist(), 64).reshape(-1, 64, 2).to(DEV)
    a = {k: v.clone() for k, v in fp.pose_recover(full, inv_k, 64, idxs=idxs).items()}
    b = fp.pose_recover(low, inv_k, 64, idxs=idxs, upsample=scale)
    assert b["class_ids"].shape[0] == n
    for k in ("cat_mask", "labels", "class_ids", "sample_ids", "mask_sizes", "quaternion", "scales", "z", "xy", "R", "T", "RT"):
        assert torch.equal(a[k], b[k]), k


def test_reference_head_modules_feed_the_fused_path(monkeypatch):
    """lowres_logits() runs only the 1x1 convolutions of SegmentationHead-shaped modules; the fused path on that equals the
    reference flow conv -> up-sample -> xyz split -> path."""
    import fastposecnn_b200 as fp
    torch.manual_seed(0)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)        # fp32 convolutions: the CPU flow is the yardstick
    b, cin, hl, wl, scale, C = 2, 16, 24, 32, 4, 7
    K = C - 1

    def head(cout):
        return torch.nn.Sequential(torch.nn.Conv2d(cin, cout, kernel_size=1), torch.nn.UpsamplingBilinear2d(scale_factor=scale),
                                   torch.nn.Identity())
    heads = {"mask": head(C), "rotation": head(4 * K), "translation": head(3 * K), "scales": head(3 * K)}
    # decoder outputs that make a few blobs win the arg-max: strong class-specific channels
    dec = {k: torch.randn(b, cin, hl, wl) for k in heads}
    with torch.no_grad():
        heads["mask"][0].weight.zero_()
        heads["mask"][0].bias.zero_()
        for c in range(C):
            heads["mask"][0].weight[c, c, 0, 0] = 1.0
        dec["mask"].mul_(0.05)
        dec["mask"][:, 0] += 1.0
        dec["mask"][0, 2, 5:12, 6:15] += 4.0
        dec["mask"][0, 5, 14:20, 20:29] += 4.0
        dec["mask"][1, 1, 8:18, 10:22] += 4.0
        # reference flow on the CPU
        full = {k: heads[k](dec[k]) for k in heads}
        xy, z = port.split_xyz(full["translation"])
        ref_logits = {"mask": full["mask"], "quaternion": full["rotation"], "scales": full["scales"], "xy": xy.contiguous(), "z": z.contiguous()}
        inv_k = torch.inverse(syn.camera_intrinsics())
        cat, agg = port.pose_recover(ref_logits, inv_k, 32, idx_source=port.seeded_idx_source(3))
        # fused flow on the GPU: convolutions by torch (cuDNN), everything after them by the library
        g_heads = {k: v.to(DEV) for k, v in heads.items()}
        low = fp.lowres_logits(g_heads, {k: v.to(DEV) for k, v in dec.items()})
    assert low["mask"].shape == (b, C, hl, wl) and low["xy"].shape == (b, 2 * K, hl, wl) and low["z"].shape == (b, K, hl, wl)
    idxs = syn.presampled_idxs(helpers.oracle_tns(agg), 32, seed=3).reshape(-1, 32, 2).to(DEV)
    out = fp.pose_recover(low, inv_k.to(DEV), 32, idxs=idxs, upsample=scale)
    # the conv runs on different hardware (cuDNN vs CPU): logits agree to rounding, blobs are 4 units above background
    assert torch.equal(out["cat_mask"].cpu().long(), cat["mask"])
    assert torch.equal(out["class_ids"].cpu(), agg["class_ids"].long()) and agg["class_ids"].tolist() == [2, 5, 1]
    for key in ("quaternion", "scales", "z"):
        assert helpers.rel_err(out[key], agg[key]) <= 1e-3, key


def test_bad_shapes_raise():
    import fastposecnn_b200 as fp
    low = {k: v.to(DEV) for k, v in syn.render_lowres_heads([[(30, 30, 10, 1)]], 96, 128, 4).items()}
    inv_k = torch.inverse(syn.camera_intrinsics()).to(DEV)
    bad = dict(low)
    bad["xy"] = low["xy"][..., :-1].contiguous()
    with pytest.raises(RuntimeError, match="xy"):
        fp.pose_recover(bad, inv_k, 16, upsample=4)
    with pytest.raises(RuntimeError, match="multiples"):
        fp.PoseRecoveryEngine(1, 98, 128, 7, 16, DEV, upsample=4)
