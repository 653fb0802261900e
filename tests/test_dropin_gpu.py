"""GPU parity at every drop-in boundary: each reference operator's replacement is fed
ORACLE-produced inputs and the same fixed pre-sampled pixel pairs, and compared with the oracle's output.
Bit-exact: cat_mask, labels, class/sample ids, instance masks, hypotheses, per-hypothesis vote counts,
winner index, refinement inlier count.  <= 1e-4 relative: every float result."""
import pytest
import torch

import helpers
from helpers import port, syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HN = 64


def _gpu(d):
    return {k: v.to(DEV) for k, v in d.items()}


def _pixel_rel_err(a, b, dim=1):
    """max over pixels of ||a-b|| / max(||b||, floor) for [b,k,h,w] fields."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    num = (a - b).norm(dim=dim)
    den = b.norm(dim=dim).clamp_min(1e-6)
    return float((num / den).max())


@pytest.mark.parametrize("name", list(helpers.scenes().keys()))
def test_class_compression_and_normalize(name):
    import fastposecnn_b200 as fp
    frames, h, w = helpers.scenes()[name]
    logits = syn.render_heads(frames, h, w, seed=5)
    cat = port.class_compression(logits, 7)
    out = fp.class_compression(_gpu(logits), 7)
    assert out["mask"].dtype == torch.int64 and torch.equal(out["mask"].cpu(), cat["mask"])
    for k in ("quaternion", "scales", "xy"):
        assert out[k].shape == cat[k].shape
        assert _pixel_rel_err(out[k], cat[k]) <= helpers.REL_TOL, k
    assert out["z"].shape == cat["z"].shape and torch.equal(out["z"].cpu(), cat["z"])     # pure selection: exact
    # class_compress with a given categorical mask (lib/gpu_tensor_funcs.py:52) returns no 'mask' key
    out2 = fp.class_compress(7, cat["mask"].to(DEV), _gpu(logits))
    assert set(out2.keys()) == {"quaternion", "scales", "xy", "z"}
    assert torch.equal(out2["scales"].cpu(), cat["scales"])
    # normalize
    x = torch.randn(3, 4, 17, 5)
    x[0, :, 3, 2] = 0
    assert _pixel_rel_err(fp.normalize(x.to(DEV), 1), port.normalize(x, 1)) <= 1e-6
    assert torch.equal(fp.normalize(x.to(DEV), 1)[0, :, 3, 2].cpu(), torch.zeros(4))


@pytest.mark.parametrize("name", list(helpers.scenes().keys()))
def test_aggregation_layer(name):
    import fastposecnn_b200 as fp
    frames, h, w = helpers.scenes()[name]
    logits = syn.render_heads(frames, h, w, seed=5)
    cat = port.class_compression(logits, 7)
    ref = port.aggregate(cat)

    class HP:
        HV_NUM_OF_HYPOTHESES = HN
    layer = fp.AggregationLayer(HP, 7)
    out = layer(_gpu(cat))
    assert out["class_ids"].dtype == torch.int64 and out["sample_ids"].dtype == torch.int64
    assert torch.equal(out["class_ids"].cpu(), ref["class_ids"].long())
    assert torch.equal(out["sample_ids"].cpu(), ref["sample_ids"])
    assert torch.equal(out["instance_masks"].cpu(), ref["instance_masks"])
    assert torch.equal(out["xy"].cpu(), ref["xy"])                       # mask * value: exact
    assert out["z"].shape == ref["z"].shape and out["quaternion"].shape == ref["quaternion"].shape
    for k in ("quaternion", "scales", "z"):
        assert helpers.rel_err(out[k], ref[k]) <= helpers.REL_TOL, k
    # labelling alone (aggregation_layer.py:160-183)
    lab, n = layer.batchwise_break_segmentation_mask((cat["mask"] != 0).to(DEV))
    lab_ref, n_ref = port.label_instances(cat["mask"] != 0)
    assert n == n_ref and torch.equal(lab.cpu(), lab_ref.to(torch.int32))


def _vote_inputs(frames, h, w, seed=5):
    logits = syn.render_heads(frames, h, w, seed=seed)
    cat = port.class_compression(logits, 7)
    agg = port.aggregate(cat)
    vertex = agg["xy"].permute(0, 2, 3, 1).unsqueeze(3)              # non-contiguous view, as hough_voting.py:51
    return agg, vertex


@pytest.mark.parametrize("name", list(helpers.scenes().keys()))
@pytest.mark.parametrize("hn", [64, 130])
def test_ransac_voting_layer_v3_bit_exact_votes(name, hn):
    from fastposecnn_b200 import ransac_voting_layer_v3
    frames, h, w = helpers.scenes()[name]
    agg, vertex = _vote_inputs(frames, h, w)
    det_ref = []
    ref = port.ransac_voting_layer_v3(agg["instance_masks"], vertex, hn, idx_source=port.seeded_idx_source(11), details=det_ref)
    tns = helpers.oracle_tns(agg)
    idxs = syn.presampled_idxs(tns, hn, seed=11)                     # [N,hn,1,2]
    det = []
    g_vertex = agg["xy"].to(DEV).permute(0, 2, 3, 1).unsqueeze(3)
    out = ransac_voting_layer_v3(agg["instance_masks"].to(DEV), g_vertex, hn, idxs=idxs.to(DEV), details=det)
    assert out.shape == ref.shape
    d = det[0]
    for i, r in enumerate(det_ref):
        if r["skipped"]:
            assert int(d["tn"][i]) == 0 and torch.equal(out[i].cpu(), torch.zeros(1, 2))
            continue
        assert int(d["tn"][i]) == r["tn"]
        assert torch.equal(d["hyp"][i].cpu(), r["hyp"][:, 0]), f"instance {i}: hypotheses differ"
        assert torch.equal(d["counts"][i].cpu(), r["counts"][:, 0].int()), f"instance {i}: vote counts differ"
        assert int(d["win_idx"][i]) == int(r["win_idx"][0]) and int(d["win_counts"][i]) == int(r["win_counts"][0])
        assert int(d["refine_inliers"][i]) == r["refine_inliers"]
    assert helpers.rel_err(out.reshape(out.shape[0], -1), ref.reshape(ref.shape[0], -1)) <= helpers.REL_TOL


def test_ransac_voting_layer_v3_subsampling_and_min_num():
    """max_num sub-sampling with an explicit keep mask (RNG-dependent in the reference) + min_num skip."""
    from fastposecnn_b200 import ransac_voting_layer_v3
    frames, h, w = helpers.scenes()["three_frames_one_empty"]
    agg, vertex = _vote_inputs(frames, h, w)
    n = agg["instance_masks"].shape[0]
    g = torch.Generator().manual_seed(5)
    keep = (torch.rand(agg["instance_masks"].shape, generator=g) < 0.5)
    max_num, min_num, hn = 700, 620, 32           # instances: 613 (skipped), 1009, 441 (skipped), 1257 px
    det_ref = []
    ref = port.ransac_voting_layer_v3(agg["instance_masks"], vertex, hn, min_num=min_num, max_num=max_num,
                                      idx_source=port.seeded_idx_source(3), select_masks={i: keep[i] for i in range(n)},
                                      details=det_ref)
    tns = [r["tn"] if not r["skipped"] else 0 for r in det_ref]
    idxs = syn.presampled_idxs(tns, hn, seed=3, min_num=1)
    det = []
    out = ransac_voting_layer_v3(agg["instance_masks"].to(DEV), agg["xy"].to(DEV).permute(0, 2, 3, 1).unsqueeze(3), hn,
                                 min_num=min_num, max_num=max_num, idxs=idxs.to(DEV), select_mask=keep.to(DEV), details=det)
    for i, r in enumerate(det_ref):
        if r["skipped"]:
            assert torch.equal(out[i].cpu(), torch.zeros(1, 2))
        else:
            assert int(det[0]["tn"][i]) == r["tn"]
            assert torch.equal(det[0]["counts"][i].cpu(), r["counts"][:, 0].int())
    assert helpers.rel_err(out.reshape(n, -1), ref.reshape(n, -1)) <= helpers.REL_TOL


def test_ransac_voting_layer_v1():
    from fastposecnn_b200 import ransac_voting_layer
    frames, h, w = helpers.scenes()["three_frames_one_empty"]
    logits = syn.render_heads(frames, h, w, seed=5)
    cat = port.class_compression(logits, 7)
    vertex = cat["xy"].permute(0, 2, 3, 1).unsqueeze(3).contiguous()
    hn = 48
    # replay the oracle's draws: problems in (image, class) order, drawn only when fg >= min_num
    tns = [int((cat["mask"][bi] == k + 1).sum()) for bi in range(len(frames)) for k in range(6)]
    ref = port.ransac_voting_layer(cat["mask"], vertex, 7, hn, idx_source=_problem_idx_source(tns, hn))
    idxs = syn.presampled_idxs(tns, hn, seed=21)
    out = ransac_voting_layer(cat["mask"].to(DEV), vertex.to(DEV), 7, hn, idxs=idxs.to(DEV))
    assert out.shape == ref.shape == (len(frames), 6, 1, 2)
    assert torch.equal(out.cpu(), ref)          # the winning hypothesis itself: bit-exact


def _problem_idx_source(tns, hn):
    table = syn.presampled_idxs(tns, hn, seed=21)

    def src(problem, hn_, vn, tn):
        assert tn == tns[problem]
        return table[problem]
    return src


def test_native_module_mirrors_multi_keypoint():
    """fpc_generate_hypothesis / fpc_voting_for_hypothesis behind the reference's pybind names, vn = 3, both arithmetic modes."""
    from fastposecnn_b200 import _lib
    from fastposecnn_b200.ransac_voting_gpu_layer import ransac_voting as rv
    from oracle import native
    g = torch.Generator().manual_seed(0)
    tn, vn, hn = 2500, 3, 77
    coords = torch.stack([torch.randint(0, 640, (tn,), generator=g), torch.randint(0, 480, (tn,), generator=g)], 1).float()
    centre = torch.tensor([300.0, 200.0])
    direct = (centre - coords).unsqueeze(1).repeat(1, vn, 1) + torch.randn(tn, vn, 2, generator=g) * 3
    direct = direct / direct.norm(dim=2, keepdim=True).clamp_min(1e-9)
    direct[5] = 0                                    # |n| < 1e-6 -> never an inlier
    idxs = torch.randint(0, tn, (hn, vn, 2), generator=g, dtype=torch.int32)
    idxs[3, 0, 1] = idxs[3, 0, 0]                    # degenerate pair -> (0,0)
    for arith, cpu in ((_lib.ARITH_IEEE, native.ransac_voting), (_lib.ARITH_NVCC_FMA, native.ransac_voting_fma)):
        rv.ARITH = arith
        try:
            hyp_ref = cpu.generate_hypothesis(direct, coords, idxs)
            hyp = rv.generate_hypothesis(direct.to(DEV), coords.to(DEV), idxs.to(DEV))
            assert torch.equal(hyp.cpu(), hyp_ref)
            inl_ref = torch.zeros((hn, vn, tn), dtype=torch.uint8)
            cpu.voting_for_hypothesis(direct, coords, hyp_ref, inl_ref, 0.999)
            inl = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=DEV)
            rv.voting_for_hypothesis(direct.to(DEV), coords.to(DEV), hyp, inl, 0.999)
            assert torch.equal(inl.cpu(), inl_ref)
            assert int(inl_ref.sum()) > 1000
        finally:
            rv.ARITH = _lib.ARITH_IEEE
    with pytest.raises(RuntimeError):
        rv.generate_hypothesis(direct, coords, idxs)                       # CPU tensors: no fallback
    with pytest.raises(RuntimeError, match="contiguous"):
        rv.generate_hypothesis(direct.to(DEV).transpose(0, 1), coords.to(DEV), idxs.to(DEV))


def test_samplewise_get_rt_and_staged_model():
    import fastposecnn_b200 as fp
    frames, h, w = helpers.scenes()["three_frames_one_empty"]
    logits = syn.render_heads(frames, h, w, seed=9)
    cat, agg, details = helpers.run_oracle(logits, HN, seed=77)
    inv_k = torch.inverse(syn.camera_intrinsics())
    # RT alone on oracle inputs
    probe = {k: agg[k].to(DEV) for k in ("quaternion", "xy", "z")}
    out = fp.samplewise_get_RT(probe, inv_k.to(DEV))
    for k in ("R", "T", "RT"):
        assert out[k].shape == agg[k].shape and helpers.rel_err(out[k], agg[k]) <= helpers.REL_TOL, k

    # the whole staged (drop-in) sequence through the Model mixin
    class HP:
        HV_NUM_OF_HYPOTHESES = HN
        PERFORM_AGGREGATION = True
        PERFORM_HOUGH_VOTING = True
        PERFORM_RT_CALCULATION = True
    model = fp.PoseRecovery(HP, 7, syn.camera_intrinsics())
    g_logits = _gpu(logits)
    model._sync_device(g_logits["mask"].device)
    cat_g = model.class_compression(g_logits)
    agg_g = model.aggregate(cat_g)
    idxs = syn.presampled_idxs(helpers.oracle_tns(agg), HN, seed=77)
    agg_g = model.hough_voting_layer(agg_g, idxs=idxs.to(DEV))
    agg_g = model.perform_RT_calculation(agg_g)
    assert torch.equal(agg_g["class_ids"].cpu(), agg["class_ids"].long())
    assert torch.equal(agg_g["instance_masks"].cpu(), agg["instance_masks"])
    assert agg_g["hypothesis"].shape == agg["hypothesis"].shape and agg_g["xy_mask"].shape == agg["xy_mask"].shape
    for k in ("quaternion", "scales", "z", "xy", "R", "T", "RT"):
        assert helpers.rel_err(agg_g[k], agg[k]) <= helpers.REL_TOL, k
    full = model(g_logits)                                                 # device-side sampling, no idxs
    assert set(full.keys()) == {"logits", "categorical", "aggregated"}
    assert (full["aggregated"]["xy"].cpu() - agg["xy"]).abs().max() < 0.2  # different random pairs, same centres


@pytest.mark.parametrize("hn,thresh", [(129, 0.999), (1000, 0.999), (1030, 0.999), (96, 0.5), (96, 0.99999), (64, -1.0), (64, 1.5)])
def test_voting_code_paths_bit_exact(hn, thresh):
    """Odd hn (no bulk copy of the hypotheses), hn that is not a multiple of 128, more than one hypothesis batch
    (hn > 1024), loose / tight thresholds and thresholds outside the fast test's domain (every vote settled exactly)."""
    from fastposecnn_b200 import ransac_voting_layer_v3
    frames, h, w = helpers.scenes()["wide"]
    agg, vertex = _vote_inputs(frames, h, w)
    det_ref = []
    ref = port.ransac_voting_layer_v3(agg["instance_masks"], vertex, hn, inlier_thresh=thresh,
                                      idx_source=port.seeded_idx_source(13), details=det_ref)
    idxs = syn.presampled_idxs(helpers.oracle_tns(agg), hn, seed=13)
    det = []
    out = ransac_voting_layer_v3(agg["instance_masks"].to(DEV), agg["xy"].to(DEV).permute(0, 2, 3, 1).unsqueeze(3), hn,
                                 inlier_thresh=thresh, idxs=idxs.to(DEV), details=det)
    for i, r in enumerate(det_ref):
        assert torch.equal(det[0]["hyp"][i].cpu(), r["hyp"][:, 0])
        assert torch.equal(det[0]["counts"][i].cpu(), r["counts"][:, 0].int()), f"instance {i} (hn={hn}, t={thresh})"
        assert int(det[0]["win_idx"][i]) == int(r["win_idx"][0])
        assert int(det[0]["refine_inliers"][i]) == r["refine_inliers"]
    n = out.shape[0]
    assert helpers.rel_err(out.reshape(n, -1), ref.reshape(n, -1)) <= helpers.REL_TOL


def test_voting_multi_keypoint_vn2():
    """vn > 1 (PVNet generality): every keypoint is voted independently with its own pixel pairs."""
    from fastposecnn_b200 import ransac_voting_layer_v3
    frames, h, w = helpers.scenes()["three_frames_one_empty"]
    agg, vertex1 = _vote_inputs(frames, h, w)
    n = agg["instance_masks"].shape[0]
    g = torch.Generator().manual_seed(2)
    noise = torch.randn(vertex1.shape, generator=g) * 0.01
    second = port.normalize((vertex1 + noise).squeeze(3).permute(0, 3, 1, 2), 1).permute(0, 2, 3, 1).unsqueeze(3)
    vertex = torch.cat([vertex1, second * agg["instance_masks"][..., None, None]], dim=3).contiguous()   # [N,h,w,2,2]
    hn = 40
    det_ref = []
    ref = port.ransac_voting_layer_v3(agg["instance_masks"], vertex, hn, idx_source=port.seeded_idx_source(4), details=det_ref)
    # same stream of pairs: [hn, vn, 2] per instance
    gi = torch.Generator().manual_seed(4)
    idxs = torch.stack([torch.randint(0, tn, (hn, 2, 2), generator=gi, dtype=torch.int32) for tn in helpers.oracle_tns(agg)])
    out = ransac_voting_layer_v3(agg["instance_masks"].to(DEV), vertex.to(DEV), hn, idxs=idxs.to(DEV))
    assert out.shape == ref.shape == (n, 2, 2)
    assert helpers.rel_err(out.reshape(n, -1), ref.reshape(n, -1)) <= helpers.REL_TOL


def test_native_module_mirrors_validate_shapes():
    """mismatched buffers must raise on the host instead of becoming out-of-bounds device accesses."""
    from fastposecnn_b200.ransac_voting_gpu_layer import ransac_voting as rv
    tn, vn, hn = 50, 2, 8
    direct = torch.randn(tn, vn, 2, device=DEV)
    coords = torch.rand(tn, 2, device=DEV)
    idxs = torch.randint(0, tn, (hn, vn, 2), dtype=torch.int32, device=DEV)
    hyp = rv.generate_hypothesis(direct, coords, idxs)
    with pytest.raises(RuntimeError, match="idxs must be"):
        rv.generate_hypothesis(direct, coords, idxs[:, :1].contiguous())
    with pytest.raises(RuntimeError, match="coords must be"):
        rv.generate_hypothesis(direct, coords[:-1].contiguous(), idxs)
    with pytest.raises(RuntimeError, match="inliers must be"):
        rv.voting_for_hypothesis(direct, coords, hyp, torch.zeros((hn, vn, tn - 1), dtype=torch.uint8, device=DEV), 0.999)
    with pytest.raises(RuntimeError, match="hypo_pts must be"):
        rv.voting_for_hypothesis_vanishing_point(direct, coords, hyp, torch.zeros((hn, vn, tn), dtype=torch.uint8, device=DEV), 0.999)


def test_quats_2_rotation_matrix_matches_the_reference_for_non_unit_quaternions():
    import fastposecnn_b200 as fp
    g = torch.Generator().manual_seed(0)
    q = torch.randn(64, 4, generator=g) * 1.7
    q[3] = 0
    ref = port.quats_2_rotation_matrix(q)
    got = fp.quats_2_rotation_matrix(q.to(DEV)).cpu()
    assert helpers.rel_err(got, ref) <= helpers.REL_TOL


def test_arg_max_of_raw_logits_versus_log_softmax_on_unquantised_logits():
    """cat_mask = argmax(LogSoftmax(x)) in the reference (pose_regressor.py:449); the kernels take argmax(x).  Log-softmax is
    monotone, so the two agree unless its rounding collapses two DISTINCT logits onto one value and first-index tie-breaking
    then prefers the earlier class.  On unquantised logits every disagreement must be such a near-tie (top-2 gap of a few ulp of
    the log-sum-exp), and there are hardly any (documented deviation)."""
    import fastposecnn_b200 as fp
    g = torch.Generator().manual_seed(3)
    b, h, w = 4, 120, 160
    logits = syn.render_heads([[(40, 40, 20, 2)], [], [(80, 60, 30, 5)], []], h, w, seed=7, quantize_mask=False)
    logits["mask"] += torch.randn(b, 7, h, w, generator=g) * 2.0
    # plant exact near-ties: class 4 a hair below class 1 on a few thousand pixels
    sel = torch.rand(b, h, w, generator=g) < 0.05
    top = logits["mask"].max(dim=1).values + 1.0
    logits["mask"][:, 4][sel] = top[sel]
    logits["mask"][:, 1][sel] = torch.nextafter(top[sel], torch.tensor(-1e30))
    ref = port.class_compression(logits, 7)["mask"]
    got = fp.class_compression(_gpu(logits), 7)["mask"].cpu()
    diff = got != ref
    x = logits["mask"].permute(0, 2, 3, 1)[diff]
    if x.numel():
        top2 = x.topk(2, dim=1).values
        gap = (top2[:, 0] - top2[:, 1]).abs()
        scale = torch.logsumexp(x, dim=1).abs().clamp_min(1.0)
        assert float((gap / scale).max()) <= 4 * 2.0 ** -23, "a disagreement that is not a rounding collapse"
    assert int(diff.sum()) <= int(sel.sum())          # only planted near-ties can flip
    assert int((got != logits["mask"].argmax(dim=1)).sum()) == 0


def test_direction_normalisation_equals_torch_cuda_bit_for_bit():
    """which formula does a GPU torch build of the reference use?  gtf.normalize (gpu_tensor_funcs.py:37-50) =
    torch.norm + division ON THE DEVICE.  Measured here (tools/diag_norm.py): torch-CUDA's norm is sqrt(x*x + y*y) with the two
    products and the sum rounded separately (no FMA contraction), and the division is IEEE.  The kernels use exactly that
    (torch_norm2 / torch_norm4 in fpc_common.cuh): class_compression's xy and quaternion fields, normalize() and the directions
    the fused path votes with equal torch-CUDA's bit for bit (torch-CPU's vectorised norm differs in the last ulp on ~0.7 % of the
    pixels, which is why the CPU oracle is compared at 1e-4)."""
    import fastposecnn_b200 as fp
    frames, h, w = helpers.scenes()["wide"]
    logits = syn.render_heads(frames, h, w, seed=5)
    g_logits = _gpu(logits)
    out = fp.class_compression(g_logits, 7)
    cat_mask = out["mask"]
    # the reference's ops on CUDA tensors: select the predicted class's xy channels, then gtf.normalize
    onehot = torch.zeros((len(frames), 7, h, w), device=DEV).scatter_(1, cat_mask.unsqueeze(1), 1.0)[:, 1:]
    xy = g_logits["xy"].reshape(len(frames), 6, 2, h, w)
    sel = torch.where(onehot.unsqueeze(2).bool(), xy.double(), torch.zeros((), dtype=torch.float64, device=DEV)).float().sum(dim=1)
    ref = port.normalize(sel, 1)                       # torch ops, run on the GPU
    assert ref.is_cuda
    assert torch.equal(out["xy"], ref)
    x = torch.randn(3, 4, 64, 65, device=DEV)
    assert torch.equal(fp.normalize(x, 1), port.normalize(x, 1))
    q = g_logits["quaternion"].reshape(len(frames), 6, 4, h, w)
    qsel = torch.where(onehot.unsqueeze(2).bool(), q.double(), torch.zeros((), dtype=torch.float64, device=DEV)).float().sum(dim=1)
    assert torch.equal(out["quaternion"], port.normalize(qsel, 1))
