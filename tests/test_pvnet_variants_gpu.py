"""GPU: the other PVNet drivers against their oracle restatements on identical fixed pixel
pairs: hypotheses and vote counts bit-exact, refined points / moments <= 1e-4."""
import pytest
import torch

import helpers
from helpers import port, syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def class_scene(vn=1, seed=2):
    frames = [[(30, 30, 14, 1), (90, 40, 18, 3), (60, 75, 12, 2)], [(40, 50, 20, 1), (100, 30, 9, 2)], [(64, 48, 22, 3)]]
    logits = syn.render_heads(frames, 96, 128, seed=seed)
    cat = port.class_compression(logits, 7)
    vertex = cat["xy"].permute(0, 2, 3, 1).unsqueeze(3)
    if vn > 1:
        vertex = torch.cat([vertex, vertex.flip(-1) * torch.tensor([1.0, -1.0])], dim=3)
    return cat["mask"], vertex.contiguous()


def recorded(seed):
    """An idx source that also records what it handed out, so the GPU call can be given the same pairs."""
    src, log = port.seeded_idx_source(seed), []

    def draw(i, hn, vn, tn):
        t = src(i, hn, vn, tn)
        log.append((i, t))
        return t
    return draw, log


@pytest.mark.parametrize("vn", [1, 2])
def test_v2(vn):
    from fastposecnn_b200.ransac_voting_gpu_layer.ransac_voting_gpu import ransac_voting_layer_v2
    mask, vertex = class_scene(vn)
    hn, class_num = 40, 4
    draw, log = recorded(9)
    want = port.ransac_voting_layer_v2(mask, vertex, class_num, hn, idx_source=draw)
    idxs = torch.zeros((mask.shape[0] * (class_num - 1), hn, vn, 2), dtype=torch.int32)
    for i, t in log:
        idxs[i] = t
    got = ransac_voting_layer_v2(mask.to(DEV), vertex.to(DEV), class_num, hn, idxs=idxs.to(DEV))
    assert got.shape == want.shape
    assert helpers.rel_err(got.reshape(-1, 2), want.reshape(-1, 2)) <= helpers.REL_TOL
    v1 = ransac_voting_layer_v2(mask.to(DEV), vertex.to(DEV), class_num, hn, refine_iter_num=0, idxs=idxs.to(DEV))
    assert torch.equal(v1.cpu(), port.ransac_voting_layer(mask, vertex, class_num, hn, idx_source=port.seeded_idx_source(9)))
    with pytest.raises(NotImplementedError):
        ransac_voting_layer_v2(mask.to(DEV), vertex.to(DEV), class_num, hn, refine_iter_num=2)


def test_hypothesis_dump():
    from fastposecnn_b200.ransac_voting_gpu_layer.ransac_voting_gpu import ransac_voting_hypothesis
    mask, vertex = class_scene()
    hn = 48
    draw, log = recorded(4)
    want_h, want_c = port.ransac_voting_hypothesis(mask, vertex, hn, idx_source=draw)
    idxs = torch.zeros((mask.shape[0], hn, 1, 2), dtype=torch.int32)
    for i, t in log:
        idxs[i] = t
    got_h, got_c = ransac_voting_hypothesis(mask.to(DEV), vertex.to(DEV), hn, idxs=idxs.to(DEV))
    assert got_c.dtype == torch.int64 and torch.equal(got_c.cpu(), want_c)
    assert torch.equal(got_h.cpu(), want_h)


def test_distribution_estimators():
    from fastposecnn_b200.ransac_voting_gpu_layer.ransac_voting_gpu import (estimate_voting_distribution,
                                                                            estimate_voting_distribution_with_mean)
    mask, vertex = class_scene()
    mask, vertex = mask[:2], vertex[:2]
    kw = dict(round_hyp_num=32, min_hyp_num=96, topk=24)
    rounds = 3

    def fixed(seed):
        draw, log = recorded(seed)
        return draw, log

    draw, log = fixed(6)
    want_mean, want_cov = port.estimate_voting_distribution(mask, vertex, idx_source=draw, **kw)
    idxs = torch.zeros((2, rounds * 32, 1, 2), dtype=torch.int32)
    for i, t in log:                                    # i = image * rounds + round
        idxs[i // rounds, (i % rounds) * 32:(i % rounds + 1) * 32] = t
    got_mean, got_cov = estimate_voting_distribution(mask.to(DEV), vertex.to(DEV), idxs=idxs.to(DEV), **kw)
    # vote ratios are small integers over tn, so many hypotheses tie at the top-k cut and torch.topk (sorted=False) is free
    # to keep different ones on the two devices: with the cut, only a loose agreement can be asked for ...
    assert helpers.rel_err(got_mean.reshape(-1, 2), want_mean.reshape(-1, 2)) <= 2e-3
    assert float((got_cov.cpu() - want_cov).abs().max()) <= 0.5 * float(want_cov.abs().max())
    # ... without it (top-k = all hypotheses) the moments must agree to the usual tolerance
    kw_all = dict(kw, topk=rounds * 32)
    draw_all, log_all = fixed(6)
    all_mean, all_cov = port.estimate_voting_distribution(mask, vertex, idx_source=draw_all, **kw_all)
    g_mean, g_cov = estimate_voting_distribution(mask.to(DEV), vertex.to(DEV), idxs=idxs.to(DEV), **kw_all)
    assert helpers.rel_err(g_mean.reshape(-1, 2), all_mean.reshape(-1, 2)) <= helpers.REL_TOL
    assert helpers.rel_err(g_cov.reshape(-1, 4), all_cov.reshape(-1, 4)) <= 1e-3
    draw, log = fixed(7)
    _, want_cov2 = port.estimate_voting_distribution_with_mean(mask, vertex, want_mean, idx_source=draw, **kw)
    for i, t in log:
        idxs[i // rounds, (i % rounds) * 32:(i % rounds + 1) * 32] = t
    _, got_cov2 = estimate_voting_distribution_with_mean(mask.to(DEV), vertex.to(DEV), want_mean.to(DEV), idxs=idxs.to(DEV), **kw)
    assert helpers.rel_err(got_cov2.reshape(-1, 4), want_cov2.reshape(-1, 4)) <= 1e-3


@pytest.mark.parametrize("vn", [1, 2])
def test_v4_v5(vn):
    from fastposecnn_b200.ransac_voting_gpu_layer.ransac_voting_gpu import ransac_voting_layer_v4, ransac_voting_layer_v5
    from test_pvnet_variants_oracle import instance_scene
    mask, vertex = instance_scene(vn)
    hn = 40

    def fixed(seed):
        draw, log = recorded(seed)
        return draw, log

    draw, log = fixed(3)
    want_pts, want_var = port.ransac_voting_layer_v4(mask, vertex, hn, idx_source=draw)
    idxs = torch.zeros((mask.shape[0], hn, vn, 2), dtype=torch.int32)
    for i, t in log:
        idxs[i] = t
    got_pts, got_var = ransac_voting_layer_v4(mask.to(DEV), vertex.to(DEV), hn, idxs=idxs.to(DEV))
    assert got_pts.shape == want_pts.shape and got_var.shape == want_var.shape
    assert helpers.rel_err(got_pts.reshape(-1, 2), want_pts.reshape(-1, 2)) <= helpers.REL_TOL
    live = torch.isfinite(want_var)
    assert torch.equal(torch.isfinite(got_var.cpu()), live)
    assert float(((got_var.cpu() - want_var)[live].abs() / want_var[live].abs().clamp_min(1e-6)).max()) <= 1e-3
    draw, log = fixed(5)
    want_pts, want_conf = port.ransac_voting_layer_v5(mask, vertex, hn, max_num=30000, idx_source=draw)
    for i, t in log:
        idxs[i] = t
    got_pts, got_conf = ransac_voting_layer_v5(mask.to(DEV), vertex.to(DEV), hn, max_num=30000, idxs=idxs.to(DEV))
    assert helpers.rel_err(got_pts.reshape(-1, 2), want_pts.reshape(-1, 2)) <= helpers.REL_TOL
    # the confidence counts votes at the refined point: a last-ulp difference of that point can flip a borderline pixel
    assert float((got_conf.cpu() - want_conf).abs().max()) <= 2.0 / 600


@pytest.mark.parametrize("vn", [1, 2])
def test_v6(vn):
    from fastposecnn_b200.ransac_voting_gpu_layer.ransac_voting_gpu import ransac_voting_layer_v6
    from test_pvnet_variants_oracle import dense_scene
    mask, vertex = dense_scene(vn)
    hn = 40
    draw, log = recorded(8)
    want_pts, want_conf = port.ransac_voting_layer_v6(mask, vertex, hn, max_num=30000, idx_source=draw)
    idxs = torch.zeros((mask.shape[0], hn, vn, 2), dtype=torch.int32)
    for i, t in log:
        idxs[i] = t
    got_pts, got_conf = ransac_voting_layer_v6(mask.to(DEV), vertex.to(DEV), hn, max_num=30000, idxs=idxs.to(DEV))
    assert got_pts.shape == want_pts.shape and got_conf.shape == want_conf.shape
    assert helpers.rel_err(got_pts.reshape(-1, 2), want_pts.reshape(-1, 2)) <= helpers.REL_TOL
    assert float((got_conf.cpu() - want_conf).abs().max()) <= 2.0 / 600
    # the whole batch's foreground count decides the sub-sampling: every image thinned by max_num / sum(mask), same uniforms
    u = torch.rand(mask.shape, generator=torch.Generator().manual_seed(5))
    draw, log = recorded(9)
    want_pts, want_conf = port.ransac_voting_layer_v6(mask, vertex, hn, max_num=900, idx_source=draw, select_u=u)
    for i, t in log:
        idxs[i] = t
    got_pts, got_conf = ransac_voting_layer_v6(mask.to(DEV), vertex.to(DEV), hn, max_num=900, idxs=idxs.to(DEV), select_u=u.to(DEV))
    assert helpers.rel_err(got_pts.reshape(-1, 2), want_pts.reshape(-1, 2)) <= helpers.REL_TOL
    assert float((got_conf.cpu() - want_conf).abs().max()) <= 2.0 / 200
    # ... and the skip: below min_num as a batch, zeros everywhere
    got_pts, got_conf = ransac_voting_layer_v6(mask.to(DEV), vertex.to(DEV), hn, min_num=10 ** 6)
    assert got_pts.shape == (3, vn, 2) and int(got_pts.abs().sum()) == 0 and int(got_conf.abs().sum()) == 0


def test_center_motion_and_hypothesis_driver():
    from fastposecnn_b200.ransac_voting_gpu_layer.ransac_voting_gpu import (generate_hypothesis, ransac_motion_voting,
                                                                           ransac_voting_center)
    from test_pvnet_variants_oracle import dense_scene, instance_scene
    mask, vertex = instance_scene(2)
    want = port.ransac_voting_center(mask, vertex[:, :, :, 0], 16, min_num=100)
    got = ransac_voting_center(mask.to(DEV), vertex[:, :, :, 0].to(DEV), 16, min_num=100)
    assert len(got) == len(want) == 2 and all(g.shape == w_.shape and int(g.abs().sum()) == 0 and g.is_cuda for g, w_ in zip(got, want))
    want = port.ransac_motion_voting(mask, vertex)
    got = ransac_motion_voting(mask.to(DEV), vertex.to(DEV))
    assert got.shape == want.shape == (4, 2, 2) and int(got[3].abs().sum()) == 0
    assert helpers.rel_err(got.reshape(-1, 2), want.reshape(-1, 2)) <= helpers.REL_TOL
    mask, vertex = dense_scene(2)
    hn = 24
    draw, log = recorded(12)
    want_h, want_c = port.generate_hypothesis(mask, vertex, hn, idx_source=draw)
    idxs = torch.zeros((mask.shape[0], hn, 2, 2), dtype=torch.int32)
    for i, t in log:
        idxs[i] = t
    got_h, got_c = generate_hypothesis(mask.to(DEV), vertex.to(DEV), hn, idxs=idxs.to(DEV))
    assert got_c.dtype == want_c.dtype and torch.equal(got_h.cpu(), want_h) and torch.equal(got_c.cpu(), want_c)
    with pytest.raises(RuntimeError, match="fewer than min_num"):
        generate_hypothesis(*[t.to(DEV) for t in instance_scene(2)], hn)
