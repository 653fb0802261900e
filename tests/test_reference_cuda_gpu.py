"""GPU: the product's FPC_ARITH_NVCC_FMA mode against the reference's OWN CUDA kernels
(lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu, unmodified, compiled for sm_100a into oracle/_ref by
oracle/build_ref_cuda.py).  Hypotheses, the full hn x tn inlier matrix and the per-hypothesis vote counts must be
bit-identical.  Skipped when oracle/_ref was not built (it can only be built where /root/reference exists)."""
import pytest
import torch

import helpers
from helpers import port, syn
from oracle import build_ref_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not build_ref_cuda.available(), reason="oracle/_ref not built")]
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    return build_ref_cuda.load()


def test_native_mirrors_match_reference_cuda_kernels(ref):
    from fastposecnn_b200 import _lib
    from fastposecnn_b200.ransac_voting_gpu_layer import ransac_voting as rv
    g = torch.Generator().manual_seed(3)
    tn, vn, hn = 6000, 2, 200
    coords = torch.stack([torch.randint(0, 640, (tn,), generator=g), torch.randint(0, 480, (tn,), generator=g)], 1).float()
    d = torch.tensor([310.0, 205.0]) - coords
    direct = d.unsqueeze(1).repeat(1, vn, 1) + torch.randn(tn, vn, 2, generator=g) * 2.5
    direct = (direct / direct.norm(dim=2, keepdim=True).clamp_min(1e-9)).contiguous()
    idxs = torch.randint(0, tn, (hn, vn, 2), generator=g, dtype=torch.int32)
    idxs[7, 1, 1] = idxs[7, 1, 0]
    direct, coords, idxs = direct.to(DEV), coords.to(DEV), idxs.to(DEV)
    hyp_ref = ref.generate_hypothesis(direct, coords, idxs)
    inl_ref = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=DEV)
    ref.voting_for_hypothesis(direct, coords, hyp_ref, inl_ref, 0.999)
    torch.cuda.synchronize()
    rv.ARITH = _lib.ARITH_NVCC_FMA
    try:
        hyp = rv.generate_hypothesis(direct, coords, idxs)
        inl = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=DEV)
        rv.voting_for_hypothesis(direct, coords, hyp, inl, 0.999)
    finally:
        rv.ARITH = _lib.ARITH_IEEE
    assert torch.equal(hyp, hyp_ref), "hypotheses differ from the reference CUDA kernel"
    assert torch.equal(inl, inl_ref), "inlier matrix differs from the reference CUDA kernel"
    assert int(inl_ref.sum()) > 10000


@pytest.mark.parametrize("name", ["three_frames_one_empty", "wide", "odd_width"])
def test_batched_voting_matches_reference_cuda_kernels(ref, name):
    """Our batched ransac_voting_layer_v3 (fast tangent-form test + exact band settling, NVCC_FMA arithmetic) against
    per-instance calls of the reference kernels driven the way ransac_voting_gpu.py:547-567 drives them."""
    from fastposecnn_b200 import _lib, ransac_voting_layer_v3
    frames, h, w = helpers.scenes()[name]
    logits = syn.render_heads(frames, h, w, seed=5)
    cat = port.class_compression(logits, 7)
    agg = port.aggregate(cat)
    hn = 96
    tns = helpers.oracle_tns(agg)
    idxs = syn.presampled_idxs(tns, hn, seed=11).to(DEV)
    masks = agg["instance_masks"].to(DEV)
    vertex = agg["xy"].to(DEV).permute(0, 2, 3, 1).unsqueeze(3)
    det = []
    ransac_voting_layer_v3(masks, vertex, hn, idxs=idxs, details=det, arith=_lib.ARITH_NVCC_FMA)
    d = det[0]
    for i, tn in enumerate(tns):
        if tn < 5:
            continue
        cur = masks[i].bool()
        coords = torch.nonzero(cur).float()[:, [1, 0]].contiguous()
        direct = vertex[i].masked_select(cur.unsqueeze(2).unsqueeze(3)).view(tn, 1, 2).contiguous()
        hyp_ref = ref.generate_hypothesis(direct, coords, idxs[i].contiguous())
        inl = torch.zeros((hn, 1, tn), dtype=torch.uint8, device=DEV)
        ref.voting_for_hypothesis(direct, coords, hyp_ref, inl, 0.999)
        counts_ref = inl.sum(2)[:, 0].int()
        assert torch.equal(d["hyp"][i], hyp_ref[:, 0]), f"instance {i}: hypotheses"
        assert torch.equal(d["counts"][i], counts_ref), f"instance {i}: vote counts"
