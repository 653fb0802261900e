"""GPU: training support.  Gradients through the drop-in class_compress / class_compression and AggregationLayer.forward
against torch autograd through the oracle restatements of the reference's torch ops (oracle/port.py: the same ops the
reference differentiates), for random upstream gradients: <= 1e-4 relative per tensor."""
import types

import pytest
import torch

import helpers
from helpers import port, syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ("quaternion", "scales", "xy", "z")


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def scene(seed=3):
    frames = [[(30, 30, 14, 1), (90, 40, 18, 3), (60, 75, 12, 6)], [(40, 50, 20, 2), (52 + 20, 50, 12, 5)], []]   # two touching blobs in frame 1
    return syn.render_heads(frames, 96, 128, seed=seed)


def test_class_compress_backward():
    import fastposecnn_b200 as fp
    logits = scene()
    g = torch.Generator().manual_seed(0)
    # oracle: autograd through the reference's ops on the CPU
    ref_in = {k: v.clone().requires_grad_(k in KEYS) for k, v in logits.items()}
    ref_out = port.class_compression(ref_in, 7)
    ups = {k: torch.randn(ref_out[k].shape, generator=g) for k in KEYS}
    sum((ref_out[k] * ups[k]).sum() for k in KEYS).backward()
    # drop-in on the GPU
    gpu_in = {k: v.to(DEV).requires_grad_(k in KEYS) for k, v in logits.items()}
    out = fp.class_compression(gpu_in, 7)
    assert torch.equal(out["mask"].cpu(), ref_out["mask"]) and not out["mask"].requires_grad
    for k in KEYS:
        assert out[k].requires_grad and rel(out[k], ref_out[k]) <= 1e-6
    sum((out[k] * ups[k].to(DEV)).sum() for k in KEYS).backward()
    for k in KEYS:
        assert gpu_in[k].grad is not None and gpu_in[k].grad.shape == logits[k].shape
        assert rel(gpu_in[k].grad, ref_in[k].grad) <= helpers.REL_TOL, k
        # only the predicted class's channels of foreground pixels receive gradient
        assert int((gpu_in[k].grad != 0).sum()) <= int((ref_out["mask"] != 0).sum()) * out[k].shape[1] if out[k].dim() == 4 else True
    # explicit cat_mask variant (gtf.class_compress) and partial upstream gradients (only scales used)
    gpu_in2 = {k: v.to(DEV).requires_grad_(k in KEYS) for k, v in logits.items()}
    out2 = fp.class_compress(7, out["mask"], gpu_in2)
    (out2["scales"] * ups["scales"].to(DEV)).sum().backward()
    assert rel(gpu_in2["scales"].grad, ref_in["scales"].grad) <= helpers.REL_TOL
    assert gpu_in2["quaternion"].grad is None or float(gpu_in2["quaternion"].grad.abs().max()) == 0.0


def test_aggregation_backward_chain():
    """class_compression -> AggregationLayer, gradients all the way back to the raw head maps."""
    import fastposecnn_b200 as fp
    logits = scene(seed=5)
    g = torch.Generator().manual_seed(1)
    ref_in = {k: v.clone().requires_grad_(k in KEYS) for k, v in logits.items()}
    ref_agg = port.aggregate(port.class_compression(ref_in, 7))
    n = ref_agg["class_ids"].shape[0]
    ups = {k: torch.randn(ref_agg[k].shape, generator=g) for k in KEYS}
    sum((ref_agg[k] * ups[k]).sum() for k in KEYS).backward()

    gpu_in = {k: v.to(DEV).requires_grad_(k in KEYS) for k, v in logits.items()}
    layer = fp.AggregationLayer(types.SimpleNamespace(HV_NUM_OF_HYPOTHESES=32), 7)
    agg = layer(fp.class_compression(gpu_in, 7))
    assert agg["class_ids"].shape[0] == n and torch.equal(agg["class_ids"].cpu(), ref_agg["class_ids"].long())
    assert torch.equal(agg["instance_masks"].cpu(), ref_agg["instance_masks"])
    for k in KEYS:
        assert agg[k].requires_grad and rel(agg[k], ref_agg[k]) <= helpers.REL_TOL, k
    sum((agg[k] * ups[k].to(DEV)).sum() for k in KEYS).backward()
    for k in KEYS:
        assert rel(gpu_in[k].grad, ref_in[k].grad) <= helpers.REL_TOL, k


def test_no_graph_without_requires_grad():
    import fastposecnn_b200 as fp
    logits = {k: v.to(DEV) for k, v in scene().items()}
    out = fp.class_compression(logits, 7)
    assert all(not out[k].requires_grad for k in KEYS)
    with torch.no_grad():
        gin = {k: v.clone().requires_grad_(k in KEYS) for k, v in logits.items()}
        assert not fp.class_compression(gin, 7)["quaternion"].requires_grad


def test_fused_path_backward():
    """pose_recover as one differentiable node: gradients of (quaternion, scales, z) w.r.t. their raw head maps against
    autograd through the oracle chain class_compression -> aggregate."""
    import fastposecnn_b200 as fp
    logits = scene(seed=7)
    g = torch.Generator().manual_seed(2)
    keys = ("quaternion", "scales", "z")
    ref_in = {k: v.clone().requires_grad_(k in keys) for k, v in logits.items()}
    ref_agg = port.aggregate(port.class_compression(ref_in, 7))
    ups = {k: torch.randn(ref_agg[k].shape, generator=g) for k in keys}
    sum((ref_agg[k] * ups[k]).sum() for k in keys).backward()
    gpu_in = {k: v.to(DEV).requires_grad_(k in keys) for k, v in logits.items()}
    inv_k = torch.inverse(syn.camera_intrinsics()).to(DEV)
    out = fp.pose_recover(gpu_in, inv_k, 32, seed=1234)
    assert out["quaternion"].requires_grad
    for k in keys:
        assert rel(out[k], ref_agg[k]) <= helpers.REL_TOL, k
    sum((out[k] * ups[k].to(DEV)).sum() for k in keys).backward(retain_graph=True)
    for k in keys:
        # the fused forward normalises pixel quaternions with rsqrt (1e-4 budget), so does its gradient
        assert rel(gpu_in[k].grad, ref_in[k].grad) <= 2 * helpers.REL_TOL, k
    # rotation / transformation outputs are differentiable through q and z as well
    assert out["R"].requires_grad and out["RT"].requires_grad
    for k in keys:
        gpu_in[k].grad = None
    (out["R"].sum() + out["T"].sum()).backward()
    assert float(gpu_in["quaternion"].grad.abs().max()) > 0 and float(gpu_in["z"].grad.abs().max()) > 0
    # a second call must not disturb the first call's saved tensors
    out2 = fp.pose_recover(gpu_in, inv_k, 32, seed=1234)
    assert torch.equal(out2["class_ids"], out["class_ids"])


@pytest.mark.parametrize("vn", [1, 2])
def test_voting_refinement_backward(vn):
    """ransac_voting_layer_v3: gradient of the refined centres w.r.t. the direction field against autograd through the
    oracle driver (same fixed pixel pairs, same inlier sets)."""
    from fastposecnn_b200 import ransac_voting_layer_v3
    frames = [[(30, 30, 14, 1)], [(64, 48, 22, 3)], [(100, 60, 1.0, 2)], []]
    logits = syn.render_heads(frames, 96, 128, seed=7)
    cat = port.class_compression(logits, 7)
    mask = (cat["mask"] != 0).float()
    vertex = cat["xy"].permute(0, 2, 3, 1).unsqueeze(3)
    if vn > 1:
        vertex = torch.cat([vertex, vertex.flip(-1) * torch.tensor([1.0, -1.0])], dim=3)
    vertex = vertex.contiguous()
    hn = 40
    log = []
    src = port.seeded_idx_source(3)

    def draw(i, hn_, vn_, tn):
        t = src(i, hn_, vn_, tn)
        log.append((i, t))
        return t
    ref_v = vertex.clone().requires_grad_(True)
    ref_pts = port.ransac_voting_layer_v3(mask, ref_v, hn, idx_source=draw)
    up = torch.randn(ref_pts.shape, generator=torch.Generator().manual_seed(4))
    (ref_pts * up).sum().backward()
    idxs = torch.zeros((mask.shape[0], hn, vn, 2), dtype=torch.int32)
    for i, t in log:
        idxs[i] = t
    gpu_v = vertex.to(DEV).requires_grad_(True)
    pts = ransac_voting_layer_v3(mask.to(DEV), gpu_v, hn, idxs=idxs.to(DEV))
    assert pts.requires_grad and helpers.rel_err(pts.reshape(-1, 2), ref_pts.detach().reshape(-1, 2)) <= helpers.REL_TOL
    (pts * up.to(DEV)).sum().backward()
    assert gpu_v.grad.shape == vertex.shape
    assert rel(gpu_v.grad, ref_v.grad) <= 1e-3
    # gradient only on inlier pixels of live instances: nothing outside the masks, nothing for the 5-pixel / empty frames' background
    assert float(gpu_v.grad[mask.to(DEV) == 0].abs().max()) == 0.0
    assert float(gpu_v.grad[3].abs().max()) == 0.0


def test_full_chain_to_the_xy_head():
    """class_compression -> AggregationLayer -> HoughVotingLayer: a loss on the voted centres reaches the raw xy head map."""
    import fastposecnn_b200 as fp
    logits = {k: v.to(DEV) for k, v in scene(seed=9).items()}
    logits["xy"].requires_grad_(True)
    hp = types.SimpleNamespace(HV_NUM_OF_HYPOTHESES=32)
    agg = fp.HoughVotingLayer(hp)(fp.AggregationLayer(hp, 7)(fp.class_compression(logits, 7)))
    assert agg["xy"].shape[1:] == (2,) and agg["xy"].requires_grad
    target = agg["xy"].detach() + 1.0
    ((agg["xy"] - target) ** 2).sum().backward()
    g = logits["xy"].grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0
    fg = fp.class_compression({k: v.detach() for k, v in logits.items()}, 7)["mask"] != 0
    assert float(g.permute(0, 2, 3, 1)[~fg].abs().max()) == 0.0          # background pixels carry no gradient


def test_get_rt_backward():
    """batchwise_get_RT: gradients of (R, T, RT) w.r.t. (q, xy, z) against autograd through the oracle (two torch.inverse)."""
    import fastposecnn_b200 as fp
    g = torch.Generator().manual_seed(5)
    n = 40
    q = torch.randn(n, 4, generator=g) * 1.7                      # not normalised: the normalisation is part of the function
    xy = torch.rand(n, 2, generator=g) * torch.tensor([640.0, 480.0])
    z = 700 + 600 * torch.rand(n, 1, generator=g)
    inv_k = torch.inverse(syn.camera_intrinsics())
    ups = [torch.randn(n, 3, 3, generator=g), torch.randn(n, 3, generator=g), torch.randn(n, 4, 4, generator=g)]
    ref = [t.clone().requires_grad_(True) for t in (q, xy, z)]
    outs = port.batchwise_get_RT(*ref, inv_k)
    sum((o * u).sum() for o, u in zip(outs, ups)).backward()
    gpu = [t.to(DEV).requires_grad_(True) for t in (q, xy, z)]
    R, T, RT = fp.batchwise_get_RT(*gpu, inv_k.to(DEV))
    assert R.requires_grad and rel(RT, outs[2]) <= helpers.REL_TOL
    sum((o * u.to(DEV)).sum() for o, u in zip((R, T, RT), ups)).backward()
    for a, b, name in zip(gpu, ref, ("q", "xy", "z")):
        assert a.grad.shape == b.grad.shape and rel(a.grad, b.grad) <= 1e-3, name
    # only one output used upstream
    gpu2 = [t.to(DEV).requires_grad_(True) for t in (q, xy, z)]
    fp.batchwise_get_RT(*gpu2, inv_k.to(DEV))[1].sum().backward()
    ref2 = [t.clone().requires_grad_(True) for t in (q, xy, z)]
    port.batchwise_get_RT(*ref2, inv_k)[1].sum().backward()
    assert rel(gpu2[1].grad, ref2[1].grad) <= 1e-4 and rel(gpu2[2].grad, ref2[2].grad) <= 1e-4
    assert float(gpu2[0].grad.abs().max()) == 0.0


def test_whole_drop_in_chain_gradients_vs_reference_chain():
    """Capstone: class_compression -> AggregationLayer -> HoughVotingLayer -> samplewise_get_RT with a loss on every
    differentiable output (quaternion, scales, z, voted xy, R, T, RT); gradients w.r.t. all four raw regression head maps
    against autograd through the oracle's restatement of the same chain, on the same fixed pixel pairs."""
    import fastposecnn_b200 as fp
    logits = scene(seed=11)
    hn = 32
    inv_k = torch.inverse(syn.camera_intrinsics())
    keys = ("quaternion", "scales", "z", "xy", "R", "T", "RT")
    ref_in = {k: v.clone().requires_grad_(k in KEYS) for k, v in logits.items()}
    log = []
    src = port.seeded_idx_source(8)

    def draw(i, hn_, vn_, tn):
        t = src(i, hn_, vn_, tn)
        log.append((i, t))
        return t
    _, ref = port.pose_recover(ref_in, inv_k, hn, idx_source=draw)
    n = ref["class_ids"].shape[0]
    g = torch.Generator().manual_seed(9)
    ups = {k: torch.randn(ref[k].shape, generator=g) * (0.01 if k in ("xy", "T", "RT") else 1.0) for k in keys}
    sum((ref[k] * ups[k]).sum() for k in keys).backward()

    idxs = torch.zeros((n, hn, 1, 2), dtype=torch.int32)
    for i, t in log:
        idxs[i] = t
    gpu_in = {k: v.to(DEV).requires_grad_(k in KEYS) for k, v in logits.items()}
    hp = types.SimpleNamespace(HV_NUM_OF_HYPOTHESES=hn)
    agg = fp.AggregationLayer(hp, 7)(fp.class_compression(gpu_in, 7))
    agg = fp.HoughVotingLayer(hp)(agg, idxs=idxs.to(DEV))
    agg = fp.samplewise_get_RT(agg, inv_k.to(DEV))
    for k in keys:
        assert agg[k].requires_grad and rel(agg[k], ref[k]) <= helpers.REL_TOL, k
    sum((agg[k] * ups[k].to(DEV)).sum() for k in keys).backward()
    for k in KEYS:
        assert rel(gpu_in[k].grad, ref_in[k].grad) <= 2e-3, k


def test_fused_path_xy_backward_equals_drop_in_chain():
    """The fused path's gradient of the voted centres w.r.t. the raw xy head against the drop-in chain's (itself checked
    against autograd through the oracle): same pixel pairs -> same winners, same inlier sets."""
    import fastposecnn_b200 as fp
    base = scene(seed=13)
    hn = 32
    inv_k = torch.inverse(syn.camera_intrinsics()).to(DEV)
    hp = types.SimpleNamespace(HV_NUM_OF_HYPOTHESES=hn)
    a_in = {k: v.to(DEV).requires_grad_(k == "xy") for k, v in base.items()}
    agg = fp.AggregationLayer(hp, 7)(fp.class_compression(a_in, 7))
    n = agg["class_ids"].shape[0]
    idxs = syn.presampled_idxs(agg["instance_masks"].sum(dim=(1, 2)).long().cpu().tolist(), hn, seed=21)
    agg = fp.samplewise_get_RT(fp.HoughVotingLayer(hp)(agg, idxs=idxs.to(DEV)), inv_k)
    up = torch.randn(n, 2, generator=torch.Generator().manual_seed(6)).to(DEV)
    upT = torch.randn(n, 3, generator=torch.Generator().manual_seed(7)).to(DEV)
    ((agg["xy"] * up).sum() + (agg["T"] * upT).sum()).backward()
    b_in = {k: v.to(DEV).requires_grad_(k == "xy") for k, v in base.items()}
    out = fp.pose_recover(b_in, inv_k, hn, idxs=idxs.reshape(n, hn, 2).to(DEV))
    assert out["xy"].requires_grad and helpers.rel_err(out["xy"], agg["xy"].detach()) <= helpers.REL_TOL
    ((out["xy"] * up).sum() + (out["T"] * upT).sum()).backward()
    assert rel(b_in["xy"].grad, a_in["xy"].grad) <= 1e-3
