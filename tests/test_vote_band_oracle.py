"""CPU: the exactness argument of the vote kernel's fast test, checked numerically (no GPU, no CUDA code involved).

``k_vote`` does not evaluate the reference's cosine test (ransac_voting_kernel.cu:112-125).  It computes, in binary32 with
fused multiply-adds and in the instance-local frame,  s = |W| - tau U  (five FMAs; fpc_voting.cu ``prepare_pixel_fast`` /
``vote_s``), takes sign(s) as the answer and hands every vote with |s| < delta(h) to ``k_vote_settle``, which evaluates the
reference expression.  The claim behind the bit-exact vote counts ("exactness"):

    |s| >= delta(h)   ==>   (s < 0)  ==  reference_inlier(h, c, n)

This file restates both sides in numpy -- the reference expression with numpy's correctly rounded binary32 operations, the
fast test with an exact binary32 FMA (float64 product and sum, exact rational arithmetic in the rare double-rounding ties),
the constants exactly as ``vote_consts`` / ``vote_frame`` / ``band_delta`` compute them -- and checks the claim on random and
on adversarial votes (hypotheses constructed to sit on the threshold cone of a pixel).  It also checks that the band is narrow
(the settle kernel's share of the votes) so that a vacuous band cannot pass."""
from fractions import Fraction

import numpy as np
import pytest

f32 = np.float32
U24 = 2.0 ** -24


# ---- exact binary32 fused multiply-add -------------------------------------------------------------------------------
def _round_fraction_to_f32(x: Fraction) -> np.float32:
    if x == 0:
        return f32(0.0)
    sign = -1 if x < 0 else 1
    x = abs(x)
    e = x.numerator.bit_length() - x.denominator.bit_length()          # 2^(e-1) <= x < 2^(e+1)
    e = max(e - 24, -149)                                              # quantum 2^e: 24 significant bits (or subnormal)
    while x / Fraction(2) ** e >= 2 ** 24:
        e += 1
    while e > -149 and x / Fraction(2) ** e < 2 ** 23:
        e -= 1
    q = x / Fraction(2) ** e
    n = q.numerator // q.denominator
    rem = q - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n % 2 == 1):
        n += 1
    return f32(sign * float(Fraction(n) * Fraction(2) ** e))


def fma32(a, b, c):
    """Correctly rounded a*b + c for binary32 arrays: the product is exact in binary64, the sum is rounded to binary64 and
    then to binary32 -- entries where that double rounding could differ from a single rounding are redone exactly."""
    a, b, c = np.broadcast_arrays(np.atleast_1d(np.asarray(a, f32)), np.atleast_1d(np.asarray(b, f32)), np.atleast_1d(np.asarray(c, f32)))
    r64 = a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)
    r32 = r64.astype(f32)
    up = np.nextafter(r32, f32(np.inf)).astype(np.float64)
    dn = np.nextafter(r32, f32(-np.inf)).astype(np.float64)
    mid_hi, mid_lo = 0.5 * (r32.astype(np.float64) + up), 0.5 * (r32.astype(np.float64) + dn)
    with np.errstate(invalid="ignore"):
        risky = np.isfinite(r64) & (np.minimum(np.abs(r64 - mid_hi), np.abs(r64 - mid_lo)) <= 8 * 2.0 ** -53 * np.abs(r64))
    if risky.any():
        r32 = r32.copy()
        for idx in zip(*np.nonzero(risky)):
            r32[idx] = _round_fraction_to_f32(Fraction(float(a[idx])) * Fraction(float(b[idx])) + Fraction(float(c[idx])))
    return r32


def test_fma32_is_a_single_rounding():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(2000).astype(f32), rng.standard_normal(2000).astype(f32)
    c = (-(a.astype(np.float64) * b.astype(np.float64)) * (1 + rng.standard_normal(2000) * 1e-7)).astype(f32)   # heavy cancellation
    got = fma32(a, b, c)
    want = np.array([_round_fraction_to_f32(Fraction(float(x)) * Fraction(float(y)) + Fraction(float(z))) for x, y, z in zip(a, b, c)])
    assert np.array_equal(got, want)
    # a constructed double-rounding trap: (1 + 2^-23) * 1.5 = 1.5 + 2^-23 + 2^-24 is a binary32 tie, c = -2^-60 puts the exact
    # result just below it (-> 1.5 + 2^-23); binary64 cannot see c, and rounding its tie to even would give 1.5 + 2^-22
    assert fma32(f32(1 + 2.0 ** -23), f32(1.5), f32(-2.0 ** -60))[0] == f32(1.5 + 2.0 ** -23)
    assert fma32(f32(1 + 2.0 ** -23), f32(1.5), f32(2.0 ** -60))[0] == f32(1.5 + 2.0 ** -22)


# ---- the two sides ---------------------------------------------------------------------------------------------------
def reference_inlier(cx, cy, nx, ny, hx, hy, t, nvcc_fma=False):
    """ransac_voting_kernel.cu:112-125.  FPC_ARITH_IEEE: every operation rounded to binary32 separately (a CPU build);
    nvcc_fma: the contraction nvcc makes of the same source, a*b + c*d -> fma(a, b, c*d) (FPC_ARITH_NVCC_FMA)."""
    def sum_prod(a, b, c, d):
        cd = (c * d).astype(f32)
        a, b, cd = np.broadcast_arrays(a, b, cd)
        return fma32(a, b, cd).reshape(cd.shape) if nvcc_fma else ((a * b).astype(f32) + cd).astype(f32)
    dx, dy = (hx - cx).astype(f32), (hy - cy).astype(f32)
    norm1 = np.sqrt(sum_prod(nx, nx, ny, ny), dtype=f32)
    norm2 = np.sqrt(sum_prod(dx, dx, dy, dy), dtype=f32)
    skip = (norm1 <= f32(1e-6)) | (norm2 <= f32(1e-6))
    with np.errstate(divide="ignore", invalid="ignore"):
        ang = (sum_prod(dx, nx, dy, ny) / (norm1 * norm2).astype(f32)).astype(f32)
    return ~skip & (ang > f32(t))


def vote_consts(t):
    """fpc_voting.cu vote_consts(): computed in double, handed to the kernel as binary32."""
    eps_r = 1.02 * (7.0 + 1.0 / t + np.sqrt(1.0 - t * t) / t) * U24
    c_hi, c_lo = t * (1.0 + eps_r), t * (1.0 - eps_r)
    T = lambda c: np.sqrt(1.0 - c * c) / c      # noqa: E731
    t_hi, t_lo = T(c_hi), T(c_lo)
    tau = f32(0.5 * (t_hi + t_lo))
    assert t_hi < float(tau) < t_lo
    hw = f32(max(t_lo - float(tau), float(tau) - t_hi))
    if float(hw) < max(t_lo - float(tau), float(tau) - t_hi):
        hw = np.nextafter(hw, f32(np.inf))
    opt = f32(1.0 + t_lo)
    if float(opt) < 1.0 + t_lo:
        opt = np.nextafter(opt, f32(np.inf))
    return f32(-tau), hw, opt


def band_delta(lx, ly, rsum, rdiag, half_w, one_plus_tlo):
    a = (np.abs(lx) + np.abs(ly)).astype(f32)
    E = (f32(4.8e-7) * (a + rsum).astype(f32)).astype(f32)
    inner = (half_w * ((f32(1.0002) * (a + rdiag).astype(f32)).astype(f32) + E).astype(f32)).astype(f32)
    return (f32(1.01) * (inner + (one_plus_tlo * E).astype(f32)).astype(f32)).astype(f32)


def scale_direction(nx, ny):
    """prepare_pixel_rare: a direction whose squared norm lies outside [0.25, 1.0002] is scaled by a power of two (exact) so
    that it lands in [0.25, 1); directions the fast prepare accepts stay as they are."""
    nn = ((nx * nx).astype(f32) + (ny * ny).astype(f32)).astype(f32)
    fast = (nn <= f32(1.0002)) & (nn >= f32(0.25))
    e2 = ((nn.view(np.uint32) >> 23) & 0xFF).astype(np.int64) - 127
    sh = (e2 + 2) >> 1
    sc = ((127 - sh).astype(np.uint32) << 23).view(f32)
    sc = np.where(fast, f32(1), sc)
    return (nx * sc).astype(f32), (ny * sc).astype(f32)


def fast_s(cx, cy, nx, ny, lx, ly, ox, oy, ntau):
    """prepare_pixel_fast / prepare_pixel_rare + vote_s (FPC_ARITH_IEEE): the same roundings in the same order."""
    nx, ny = scale_direction(nx, ny)
    ccx, ccy = (cx - ox).astype(f32), (cy - oy).astype(f32)                       # exact: integers
    pu = -fma32(ccx, nx, (ccy * ny).astype(f32))
    pw = -fma32(ccx, ny, -(ccy * nx).astype(f32))
    U = fma32(lx, nx, fma32(ly, ny, pu))
    W = fma32(lx, ny, fma32(-ly, nx, pw))
    return fma32(U, ntau, np.abs(W))


def disc_instance(rng, radius, cx0, cy0):
    ys, xs = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    keep = xs * xs + ys * ys <= radius * radius
    px, py = (xs[keep] + cx0).astype(f32), (ys[keep] + cy0).astype(f32)
    x0, x1, y0, y1 = int(px.min()), int(px.max()), int(py.min()), int(py.max())
    ox, oy = f32((x0 + x1 + 1) >> 1), f32((y0 + y1 + 1) >> 1)                      # vote_frame()
    rx, ry = f32(0.5) * f32(x1 - x0) + f32(1), f32(0.5) * f32(y1 - y0) + f32(1)
    rmax2 = f32(np.max((px - ox) ** 2 + (py - oy) ** 2))
    rdiag = f32(1.0001) * min(np.sqrt(rx * rx + ry * ry, dtype=f32), np.sqrt(rmax2, dtype=f32) + f32(1e-3))
    return px, py, ox, oy, f32(rx + ry), f32(rdiag)


def unit_directions(rng, px, py, tx, ty, noise):
    """what k_gather hands over: value / norm with IEEE sqrt and divide (|n| = 1 up to rounding)"""
    dx = (tx - px + rng.standard_normal(px.shape) * noise).astype(f32)
    dy = (ty - py + rng.standard_normal(px.shape) * noise).astype(f32)
    nrm = np.sqrt((dx * dx).astype(f32) + (dy * dy).astype(f32), dtype=f32)
    nrm = np.where(nrm == 0, f32(1), nrm)
    return (dx / nrm).astype(f32), (dy / nrm).astype(f32)


def on_fast_path(hx, hy, lx, ly):
    near_lattice = (np.abs(hx - np.rint(hx)) < f32(1e-3)) & (np.abs(hy - np.rint(hy)) < f32(1e-3))
    return ~near_lattice & ((np.abs(lx) + np.abs(ly)) < f32(1e12))


def check(t, px, py, nx, ny, hx, hy, ox, oy, rsum, rdiag, nvcc_fma=False):
    """all hypotheses x all pixels; returns (votes checked outside the band, votes inside the band)"""
    ntau, half_w, opt = vote_consts(t)
    lx, ly = (hx - ox).astype(f32), (hy - oy).astype(f32)                          # k_hypotheses: hloc
    fast = on_fast_path(hx, hy, lx, ly)
    hx, hy, lx, ly = hx[fast], hy[fast], lx[fast], ly[fast]
    # shrink delta a little: the device evaluates band_delta with nvcc's FMA contraction, a last-ulp difference
    delta = band_delta(lx, ly, rsum, rdiag, half_w, opt) * f32(1 - 1e-5)
    H, P = hx[:, None], px[None, :]
    s = fast_s(P, py[None, :], nx[None, :], ny[None, :], lx[:, None], ly[:, None], ox, oy, ntau)
    ref = reference_inlier(P, py[None, :], nx[None, :], ny[None, :], H, hy[:, None], t, nvcc_fma)
    certain = np.abs(s) >= delta[:, None]
    wrong = certain & ((s < 0) != ref)
    assert not wrong.any(), (f"t={t}: {int(wrong.sum())} votes outside the band disagree with the reference, e.g. "
                             f"s={s[wrong][:3]}, delta={np.broadcast_to(delta[:, None], s.shape)[wrong][:3]}")
    return int(certain.sum()), int((~certain).sum())


@pytest.mark.parametrize("nvcc_fma", [False, True], ids=["ieee", "nvcc_fma"])
@pytest.mark.parametrize("t", [0.999, 0.99, 0.9])
def test_sign_of_s_is_the_reference_answer_outside_the_band(t, nvcc_fma):
    rng = np.random.default_rng(int(t * 1000))
    total_certain = total_band = 0
    for radius, cx0, cy0, noise in ((35, 320, 240, 0.6), (12, 53, 460, 0.2), (60, 600, 70, 2.0)):
        px, py, ox, oy, rsum, rdiag = disc_instance(rng, radius, cx0, cy0)
        nx, ny = unit_directions(rng, px, py, cx0 + 0.37, cy0 - 0.21, noise)
        # random hypotheses: around the centre (where RANSAC puts them), across the box, and far outside
        hx = np.concatenate([cx0 + rng.standard_normal(24) * 1.5, cx0 + rng.uniform(-radius, radius, 12), cx0 + rng.uniform(-3000, 3000, 6)]).astype(f32)
        hy = np.concatenate([cy0 + rng.standard_normal(24) * 1.5, cy0 + rng.uniform(-radius, radius, 12), cy0 + rng.uniform(-3000, 3000, 6)]).astype(f32)
        c, b = check(t, px, py, nx, ny, hx, hy, ox, oy, rsum, rdiag, nvcc_fma)
        total_certain += c
        total_band += b
    # the band is what k_vote_settle pays for: it must stay a small share of random votes
    assert total_band < 0.02 * (total_certain + total_band)


@pytest.mark.parametrize("nvcc_fma", [False, True], ids=["ieee", "nvcc_fma"])
@pytest.mark.parametrize("t", [0.999, 0.99])
def test_hypotheses_on_the_threshold_cone(t, nvcc_fma):
    """Adversarial: every hypothesis is built from a pixel c and its direction n so that the angle between n and h - c is the
    threshold angle scaled by 1 +- k ulp-sized steps: those votes sit in or right next to the band.  Outside the band the
    sign must still be the reference's answer; and the construction must really exercise the band (many votes inside it)."""
    rng = np.random.default_rng(7)
    px, py, ox, oy, rsum, rdiag = disc_instance(rng, 35, 320, 240)
    nx, ny = unit_directions(rng, px, py, 320.4, 239.7, 0.5)
    theta = np.arccos(t)
    pick = rng.choice(px.shape[0], 48, replace=False)
    hx, hy = [], []
    for k, p in enumerate(pick):
        ang = theta * (1.0 + (k % 9 - 4) * 3e-7) * (1 if k % 2 else -1)           # threshold angle +- a few 1e-7 relative
        dist = rng.uniform(3.0, 60.0)
        c, s_ = np.cos(ang), np.sin(ang)
        dxr = float(nx[p]) * c - float(ny[p]) * s_
        dyr = float(nx[p]) * s_ + float(ny[p]) * c
        hx.append(float(px[p]) + dist * dxr)
        hy.append(float(py[p]) + dist * dyr)
    hx, hy = np.array(hx, f32), np.array(hy, f32)
    ntau, half_w, opt = vote_consts(t)
    lx, ly = (hx - ox).astype(f32), (hy - oy).astype(f32)
    delta = band_delta(lx, ly, rsum, rdiag, half_w, opt)
    s_own = fast_s(px[pick], py[pick], nx[pick], ny[pick], lx, ly, ox, oy, ntau)   # each hypothesis against ITS pixel
    assert (np.abs(s_own) < delta).mean() > 0.5                                    # the construction does land in the band
    check(t, px, py, nx, ny, hx, hy, ox, oy, rsum, rdiag, nvcc_fma)


@pytest.mark.parametrize("scale", [3e-4, 0.3, 7.0, 4e4])
def test_unnormalised_directions_of_the_voting_drop_in(scale):
    """ransac_voting_layer* accept any direction field: the kernel scales a direction by a power of two into [1/2, 1) before the
    fast test (the reference's cosine does not depend on |n|).  Same claim, directions of norm `scale` (x 1..2)."""
    rng = np.random.default_rng(11)
    px, py, ox, oy, rsum, rdiag = disc_instance(rng, 20, 100, 90)
    nx, ny = unit_directions(rng, px, py, 100.3, 89.6, 0.4)
    k = (scale * rng.uniform(1.0, 2.0, px.shape)).astype(f32)
    nx, ny = (nx * k).astype(f32), (ny * k).astype(f32)
    sx, sy = scale_direction(nx, ny)
    snn = sx.astype(np.float64) ** 2 + sy.astype(np.float64) ** 2
    assert snn.min() >= 0.2499 and snn.max() <= 1.0003
    hx = (100 + rng.standard_normal(40) * 2.0).astype(f32)
    hy = (90 + rng.standard_normal(40) * 2.0).astype(f32)
    certain, band = check(0.999, px, py, nx, ny, hx, hy, ox, oy, rsum, rdiag)
    assert band < 0.02 * (certain + band)


def test_lattice_and_far_hypotheses_stay_off_the_fast_path():
    hx, hy = np.array([10.0, 10.0004, 10.5, 3e12], f32), np.array([20.0, 19.9996, 20.0, 0.0], f32)
    lx, ly = hx - f32(8), hy - f32(16)
    assert on_fast_path(hx, hy, lx, ly).tolist() == [False, False, True, False]
