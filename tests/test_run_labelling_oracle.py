"""CPU: the run-based labelling algorithm of csrc/fpc_runs.cu, restated in numpy and checked against scipy.ndimage.label
(what the reference calls, lib/aggregation_layer.py:160-183).  No CUDA involved: this pins the ALGORITHM's claims --

  * runs cut at 128-pixel spans of the linear pixel index (what the arg-max kernels emit) are re-joined by the union-find;
  * uniting a run with the runs of the row above whose column intervals overlap it is 4-connectivity;
  * roots = smallest run id of a component, instance id = rank of the root run in raster order = scipy's numbering
    (components numbered by their first pixel in raster order), image after image;

the CUDA implementation itself is compared with scipy on the GPU box (tests/test_scale_gpu.py, tests/test_fused_gpu.py)."""
import numpy as np
import pytest
from scipy import ndimage


def emit_runs(fg, span):
    """k_argmax_runs* + k_emit_runs: maximal foreground segments inside one image row AND one span of the linear pixel index."""
    b, h, w = fg.shape
    flat = fg.reshape(-1)
    p = np.arange(flat.size)
    x = p % w
    prev = np.concatenate(([False], flat[:-1]))
    nxt = np.concatenate((flat[1:], [False]))
    start = flat & ((x == 0) | (p % span == 0) | ~prev)
    end = flat & ((x == w - 1) | ((p + 1) % span == 0) | ~nxt)
    return np.nonzero(start)[0], np.nonzero(end)[0]


def find(parent, a):
    while parent[a] != a:
        parent[a] = parent[parent[a]]          # path halving, pointers only ever decrease
        a = parent[a]
    return a


def unite(parent, a, b):
    a, b = find(parent, a), find(parent, b)
    if a != b:
        parent[max(a, b)] = min(a, b)          # the larger root goes under the smaller one


def label_runs(fg, span=128):
    """-> label volume like the fused path's (0 = background, global instance index + 1), number of instances"""
    b, h, w = fg.shape
    start, end = emit_runs(fg, span)
    M = start.size
    row = start // w                                            # global row index = image * h + y
    x0, x1 = start - row * w, end - row * w
    rowrun = np.searchsorted(row, np.arange(b * h + 1))         # first run at or after the start of every image row
    parent = np.arange(M)
    for m in range(M):                                          # k_run_merge
        if m > 0 and x0[m] > 0 and end[m - 1] == start[m] - 1:
            unite(parent, m, m - 1)                             # the left piece of a run cut at a span border
        if row[m] % h == 0:
            continue
        for u in range(rowrun[row[m] - 1], rowrun[row[m]]):     # the row above, same image
            if x1[u] >= x0[m] and x0[u] <= x1[m]:
                unite(parent, m, u)
    root = np.array([find(parent, m) for m in range(M)])        # k_run_flatten
    is_root = root == np.arange(M)
    inst_of_root = np.cumsum(is_root) - 1                       # k_scan_roots + k_run_assign: rank in raster order
    out = np.zeros(b * h * w, np.int32)
    for m in range(M):                                          # k_relabel
        out[start[m]:end[m] + 1] = inst_of_root[root[m]] + 1
    return out.reshape(b, h, w), int(is_root.sum())


def scipy_labels(fg):
    out, offset = np.zeros(fg.shape, np.int32), 0
    for i in range(fg.shape[0]):
        lab, n = ndimage.label(fg[i])                           # default structure: 4-connectivity
        out[i] = np.where(lab > 0, lab + offset, 0)
        offset += n
    return out, offset


def blobs(rng, b, h, w, n, rmax):
    fg = np.zeros((b, h, w), bool)
    ys, xs = np.mgrid[0:h, 0:w]
    for i in range(b):
        for _ in range(n):
            cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(1, rmax)
            fg[i] |= (ys - cy) ** 2 + (xs - cx) ** 2 <= r * r
    return fg


@pytest.mark.parametrize("w", [128, 100, 257])                  # rows aligned with the spans, spans straddling rows, odd width
def test_blobs_match_scipy(w):
    rng = np.random.default_rng(w)
    fg = blobs(rng, 3, 60, w, 14, 12.0)
    got, n = label_runs(fg)
    want, n_want = scipy_labels(fg)
    assert n == n_want and np.array_equal(got, want)


def test_noise_holes_and_spirals_match_scipy():
    rng = np.random.default_rng(5)
    noise = rng.random((2, 48, 200)) < 0.45                     # thousands of specks, many touching diagonally only
    got, n = label_runs(noise)
    want, n_want = scipy_labels(noise)
    assert n == n_want > 500 and np.array_equal(got, want)
    # a U shape and a ring: components whose first run is not connected to later runs until further down
    shape = np.zeros((1, 40, 300), bool)
    shape[0, 5:35, 10:14] = shape[0, 5:35, 280:284] = True      # the two arms
    shape[0, 31:35, 10:284] = True                              # the bottom joins them, across two span borders
    shape[0, 2:4, 100:200] = True                               # a separate bar above
    yy, xx = np.mgrid[0:40, 0:300]
    ring = ((yy - 18) ** 2 + (xx - 150) ** 2 <= 100) & ((yy - 18) ** 2 + (xx - 150) ** 2 >= 49)
    shape[0] |= ring
    got, n = label_runs(shape)
    want, n_want = scipy_labels(shape)
    assert n == n_want == 3 and np.array_equal(got, want)


def test_span_cutting_does_not_change_the_result():
    rng = np.random.default_rng(9)
    fg = blobs(rng, 2, 50, 320, 10, 40.0)                       # runs far longer than a span
    a, na = label_runs(fg, span=128)
    b_, nb = label_runs(fg, span=1 << 30)                       # uncut runs (what k_cls_runs emits)
    c, nc = label_runs(fg, span=32)
    assert na == nb == nc and np.array_equal(a, b_) and np.array_equal(a, c)
    s128, _ = emit_runs(fg, 128)
    s_uncut, _ = emit_runs(fg, 1 << 30)
    assert s128.size > s_uncut.size                             # the cut really produced extra pieces


def test_empty_and_full_frames():
    fg = np.zeros((2, 8, 16), bool)
    got, n = label_runs(fg)
    assert n == 0 and not got.any()
    fg[1] = True                                                # an instance covering a whole image
    got, n = label_runs(fg)
    assert n == 1 and (got[1] == 1).all() and not got[0].any()
