"""GPU: BASELINE.json's full sizes, checked through size-independent properties (the oracle would take minutes):
known disc centres, instance order, determinism (bit-identical reruns), vote-count invariants, and the
1280x960 / 1024-hypothesis stress layout."""
import pytest
import torch

from helpers import syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(wl, batch, seed=0, **kw):
    from fastposecnn_b200.pose_recovery import PoseRecoveryEngine
    logits = syn.render_workload(wl, batch=batch, seed=seed, device=DEV)
    inv_k = torch.inverse(syn.camera_intrinsics()).to(DEV)
    eng = PoseRecoveryEngine(batch, wl.h, wl.w, wl.num_classes, wl.hyps, DEV, **kw)
    tns = [syn.disc_pixel_count(cx, cy, r, wl.h, wl.w) for (cx, cy, r, _c) in wl.discs()] * batch
    idxs = syn.presampled_idxs(tns, wl.hyps).reshape(-1, wl.hyps, 2).to(DEV)
    eng.launch(logits, inv_k, idxs=idxs)
    n = eng.fetch_count()
    return eng, eng.table_to_agg(n), tns, (logits, inv_k, idxs)


def _check_layout(wl, batch, agg, tns):
    discs = wl.discs()
    n = batch * len(discs)
    assert agg["class_ids"].shape[0] == n
    assert agg["class_ids"].cpu().tolist() == [d[3] for d in discs] * batch
    assert agg["sample_ids"].cpu().tolist() == [b for b in range(batch) for _ in discs]
    assert agg["mask_sizes"].cpu().tolist() == tns
    cent = torch.tensor([[d[0], d[1]] for d in discs]).repeat(batch, 1)
    assert (agg["xy"].cpu() - cent).abs().max() < 0.15
    assert (agg["z"].cpu() - 992.27).abs().max() < 5.0            # exp(6.9) mm
    q = agg["quaternion"].cpu()
    assert torch.allclose(q.norm(dim=1), torch.ones(n), atol=1e-5)
    R = agg["R"].cpu()
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(n, 3, 3), atol=1e-4)
    RT = agg["RT"].cpu()
    assert torch.allclose(RT[:, :3, 3], -(R @ agg["T"].cpu().unsqueeze(2)).squeeze(2), atol=1e-4)
    assert (agg["win_counts"] <= agg["tn"]).all() and (agg["win_counts"] > 0.5 * agg["tn"]).all()
    assert (agg["refine_inliers"] >= agg["win_counts"] * 0.9).all()


def test_cfg2_full_batch_properties_and_determinism():
    wl = syn.WORKLOADS["cfg2"]
    eng, agg, tns, (logits, inv_k, idxs) = _run(wl, wl.batch)
    _check_layout(wl, wl.batch, agg, tns)
    n = agg["class_ids"].shape[0]
    t0 = eng.pose_table[:n].clone()
    v0 = eng.votes[:n].clone()
    assert int(v0.max()) <= max(tns) and int(v0.min()) >= 0
    for _ in range(3):                                             # bit-identical reruns: no order-dependent float sums
        eng.launch(logits, inv_k, idxs=idxs)
        assert eng.fetch_count() == n
        assert torch.equal(eng.pose_table[:n].view(torch.int32), t0.view(torch.int32))
        assert torch.equal(eng.votes[:n], v0)
    # the scalar and the packed-f32x2 vote loops, and device-side sampling, agree on the centres
    eng.launch(logits, inv_k, idxs=None)
    assert eng.fetch_count() == n
    assert (eng.table_to_agg(n)["xy"] - agg["xy"]).abs().max() < 0.2


def test_cfg4_stress_1280x960_hn1024():
    wl = syn.WORKLOADS["cfg4"]
    eng, agg, tns, _ = _run(wl, 2)
    assert max(tns) < 30000                                         # below max_num: no sub-sampling involved
    _check_layout(wl, 2, agg, tns)


def test_many_small_components_noise_mask():
    """A noisy mask (thousands of specks) must neither crash nor mislabel: compare with scipy on the GPU box."""
    from fastposecnn_b200.pose_recovery import PoseRecoveryEngine
    from helpers import port
    g = torch.Generator().manual_seed(0)
    b, h, w = 2, 120, 160
    logits = syn.render_heads([[], []], h, w, seed=1)
    logits["mask"][:, 1:] += (torch.rand(b, 6, h, w, generator=g) < 0.08) * 3.0
    cat = port.class_compression(logits, 7)
    lab_ref, total = port.label_instances(cat["mask"] != 0)
    eng = PoseRecoveryEngine(b, h, w, 7, 16, DEV, max_instances=total + 8, want_labels=True)
    eng.launch({k: v.to(DEV) for k, v in logits.items()}, torch.inverse(syn.camera_intrinsics()).to(DEV))
    assert eng.fetch_count() == total and total > 1000
    assert torch.equal(eng.labels.cpu(), lab_ref.to(torch.int32))
    agg = port.aggregate(cat)
    out = eng.table_to_agg(total)
    assert torch.equal(out["class_ids"].cpu(), agg["class_ids"].long())
    assert out["mask_sizes"].cpu().tolist() == [int(v) for v in agg["instance_masks"].sum(dim=(-2, -1)).tolist()]


@pytest.mark.parametrize("seed,h,w", [(0, 120, 160), (1, 97, 131), (2, 240, 320)])
def test_irregular_blobs_with_holes_and_concavities(seed, h, w):
    """Thresholded smooth random fields: big concave components, holes, U-shapes that merge late, several runs per
    row per instance, touching classes.  Everything integer must match scipy / the oracle bit for bit."""
    import torch.nn.functional as F
    from fastposecnn_b200.pose_recovery import PoseRecoveryEngine
    from helpers import port
    g = torch.Generator().manual_seed(seed)
    b = 3
    logits = syn.render_heads([[]] * b, h, w, seed=seed)
    field = torch.randn(b, 6, h // 8 + 2, w // 8 + 2, generator=g)
    field = F.interpolate(field, size=(h, w), mode="bicubic", align_corners=False)
    logits["mask"][:, 1:] += torch.round(field * 2.0 * 1024) / 1024       # blobs of every class, quantised like the generator
    cat = port.class_compression(logits, 7)
    lab_ref, total = port.label_instances(cat["mask"] != 0)
    assert total >= 3
    eng = PoseRecoveryEngine(b, h, w, 7, 32, DEV, max_instances=total + 8, want_labels=True)
    eng.launch({k: v.to(DEV) for k, v in logits.items()}, torch.inverse(syn.camera_intrinsics()).to(DEV))
    assert eng.fetch_count() == total
    assert torch.equal(eng.cat_mask_u8.cpu().long(), cat["mask"])
    assert torch.equal(eng.labels.cpu(), lab_ref.to(torch.int32))
    agg = port.aggregate(cat)
    out = eng.table_to_agg(total)
    assert torch.equal(out["class_ids"].cpu(), agg["class_ids"].long())
    assert torch.equal(out["sample_ids"].cpu(), agg["sample_ids"])
    assert out["mask_sizes"].cpu().tolist() == [int(v) for v in agg["instance_masks"].sum(dim=(-2, -1)).tolist()]
    for k in ("quaternion", "scales", "z"):
        import helpers
        assert helpers.rel_err(out[k], agg[k]) <= helpers.REL_TOL, k
    # drop-in aggregation on the same categorical data: dense instance masks bit-exact
    import fastposecnn_b200 as fp

    class HP:
        HV_NUM_OF_HYPOTHESES = 32
    got = fp.AggregationLayer(HP, 7, max_instances=total + 8)({k: v.to(DEV) for k, v in cat.items()})
    assert torch.equal(got["instance_masks"].cpu(), agg["instance_masks"])
    assert torch.equal(got["xy"].cpu(), agg["xy"])
