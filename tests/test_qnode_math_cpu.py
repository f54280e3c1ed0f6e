"""The arithmetic of the quantised 32-byte BVH nodes, emulated in numpy float32 exactly as the kernels do it
(drt_b200/csrc/bvh.cuh: grid_kernel, qpair; trace.cuh: safe_inv, ray_setup, qplane, node_step), against the EXACT
slab test of the original float32 box in float64: a box that the true ray touches within [0, tmax] must never be
rejected.  This is the conservativeness claim of DESIGN.md 3.2 / trace.cuh, checked on the CPU for millions of
random box/ray pairs incl. axis-parallel rays, tiny direction components, origins inside, on and far outside the grid."""
import numpy as np

F32 = np.float32
GRID = F32(65520.0)


def fma32(a, b, c):
    """float32 FMA: the product of two float32 is exact in float64; one rounding of the sum to float32 (the double
    rounding through float64 can differ from a true FMA by at most one float32 ulp in ~2^-29 of the cases -- far inside
    the margins under test)"""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def round32(x64, up):
    """float64 -> float32 rounded towards +inf (up) or -inf, like the __f*_ru / __f*_rd intrinsics (exact values stay)"""
    r = x64.astype(F32)
    if up:
        return np.where(r.astype(np.float64) < x64, np.nextafter(r, F32(np.inf)), r).astype(F32)
    return np.where(r.astype(np.float64) > x64, np.nextafter(r, F32(-np.inf)), r).astype(F32)


def make_grid(root_lo, root_hi):
    ext = round32(root_hi.astype(np.float64) - root_lo.astype(np.float64), True)      # __fadd_ru
    emax = float(ext.max())
    if not emax > 7.888609052210118e-31:
        emax = 1.0
    ext = np.maximum(ext, F32(emax) * F32(2.0 ** -20))
    s = round32(ext.astype(np.float64) / 65520.0, True)                                # __fdiv_ru
    seven_s = round32(7.0 * s.astype(np.float64), True)                                # __fmul_ru
    g0 = round32(root_lo.astype(np.float64) - seven_s.astype(np.float64), False)       # __fadd_rd
    return g0, s


def qpair(lo, hi, g0, s):
    inv_s = (F32(1.0) / s).astype(F32)
    ql = np.maximum(0, np.floor(((lo - g0).astype(F32) * inv_s).astype(F32)).astype(np.int64) - 3)
    qh = np.minimum(65535, np.ceil(((hi - g0).astype(F32) * inv_s).astype(F32)).astype(np.int64) + 3)
    return ql, qh


def safe_inv(d):
    tiny = np.abs(d) < F32(8.271806125530277e-25)
    with np.errstate(divide="ignore"):
        inv = (F32(1.0) / d).astype(F32)
    return np.where(tiny, np.copysign(F32(1.2089258196146292e24), d), inv).astype(F32)


def kernel_box_test(o, d, tmax, ql, qh, g0, s):
    """-> bool[n]: node_step's decision for one child box per ray"""
    inv = safe_inv(d)
    A = (s * inv).astype(F32)
    w = (g0 - o).astype(F32)
    C = (w * inv).astype(F32)
    M = (F32(8388608.0) * A).astype(F32)
    Cp = (C - M).astype(F32)
    far = ~np.all(np.abs(w) <= F32(64.0 * 65520.0) * s, axis=1)
    E = np.where(far, F32(4.76837158203125e-07) * (np.abs(C) + np.abs(M)).max(axis=1), F32(0)).astype(F32)
    near_q = np.where(inv >= 0, ql, qh)                                       # the PRMT selector picks by the sign of 1/d
    far_q = np.where(inv >= 0, qh, ql)
    magic = lambda q: (np.uint32(0x4B000000) | q.astype(np.uint32)).view(F32)  # noqa: E731  float bits 0x4B00qqqq = 2^23 + q
    tn = fma32(magic(near_q), A, Cp)
    tf = fma32(magic(far_q), A, Cp)
    n = np.maximum(tn.max(axis=1), F32(0))
    f = np.minimum(tf.min(axis=1), tmax)
    return n <= fma32(f, np.full_like(f, F32(1.00000095367431640625)), E)


def exact_touch(o, d, tmax, lo, hi):
    """the true ray o + t d, t in [0, tmax], against the closed box [lo, hi], in float64 (inputs are float32 values)"""
    o, d, lo, hi = (x.astype(np.float64) for x in (o, d, lo, hi))
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (lo - o) / d, (hi - o) / d
    tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
    par = d == 0                                                             # axis-parallel: inside the slab or never
    inside = (o >= lo) & (o <= hi)
    tn = np.where(par, np.where(inside, -np.inf, np.inf), tn)
    tf = np.where(par, np.where(inside, np.inf, -np.inf), tf)
    n = np.maximum(tn.max(axis=1), 0.0)
    f = np.minimum(tf.min(axis=1), tmax.astype(np.float64))
    return n <= f


def _cases(rng, n, scale, shift):
    root_lo = (np.array([-1.0, -0.6, -0.3]) * scale + shift).astype(F32)
    root_hi = (np.array([1.0, 0.8, 0.2]) * scale + shift).astype(F32)
    g0, s = make_grid(root_lo, root_hi)
    span = (root_hi - root_lo).astype(np.float64)
    c = root_lo + rng.random((n, 3)) * span
    half = rng.random((n, 3)) * span * rng.choice([1e-4, 3e-3, 0.05, 0.5], (n, 1))
    lo = np.maximum(c - half, root_lo).astype(F32)
    hi = np.minimum(c + half, root_hi).astype(F32)
    hi = np.maximum(hi, lo)
    # rays aimed at (or just past) the box from near, far, inside and on its faces
    kind = rng.integers(0, 5, n)
    dist = np.choose(kind, [0.5, 5.0, 300.0, 0.0, 0.01])[:, None] * span.max()
    o = (c + rng.normal(size=(n, 3)) * dist)
    o = np.where((kind == 3)[:, None], lo + rng.random((n, 3)) * (hi - lo), o).astype(F32)
    tgt = lo + rng.random((n, 3)) * (hi.astype(np.float64) - lo) + rng.normal(size=(n, 3)) * half * rng.choice([0.0, 0.0, 1.5], (n, 1))
    d = (tgt - o)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    d = d.astype(F32)
    axis = rng.random(n) < 0.15                                             # axis-parallel and nearly axis-parallel rays
    k = rng.integers(0, 3, n)
    d[axis, k[axis]] = rng.choice([0.0, 1e-30, -1e-28, 1e-12], axis.sum()).astype(F32)
    flip = rng.random(n) < 0.1                                              # un-normalised directions (silhouette rays)
    d[flip] *= F32(37.5)
    tmax = np.where(rng.random(n) < 0.5, F32(np.inf), (rng.random(n) * 4 * dist[:, 0] / np.where(flip, 37.5, 1.0) + 1e-6).astype(F32)).astype(F32)
    return o, d, tmax, lo, hi, g0, s


def test_quantised_slab_test_never_rejects_a_touched_box():
    rng = np.random.default_rng(7)
    total = touched = 0
    for scale, shift in ((1.0, 0.0), (60.0, 0.0), (1e-3, 0.0), (1e3, 0.0), (60.0, 500.0), (1.0, -40.0)):
        o, d, tmax, lo, hi, g0, s = _cases(rng, 400_000, scale, np.array([shift, -shift, 0.5 * shift]))
        ql, qh = qpair(lo, hi, g0, s)
        assert (ql >= 0).all() and (qh <= 65535).all() and (ql < qh).all()
        # the stored planes enclose the box by at least two and at most ~five grid steps (three of margin + rounding)
        plo, phi = g0.astype(np.float64) + ql * s.astype(np.float64), g0.astype(np.float64) + qh * s.astype(np.float64)
        assert (plo <= lo - 1.9 * s).all() and (phi >= hi + 1.9 * s).all()
        assert (plo >= lo - 5.1 * s).all() and (phi <= hi + 5.1 * s).all()
        need = exact_touch(o, d, tmax, lo, hi)
        got = kernel_box_test(o, d, tmax, ql, qh, g0, s)
        missed = need & ~got
        assert not missed.any(), (scale, shift, int(missed.sum()), o[missed][:3], d[missed][:3], lo[missed][:3], hi[missed][:3])
        total += len(need)
        touched += int(need.sum())
        # and it is still a useful filter: boxes far from the ray are rejected
        assert (got & ~need).mean() < 0.2
    assert touched > 0.3 * total


def test_grid_covers_the_root_box_with_spare_steps():
    for lo, hi in (([-1, -2, -3], [4, 5, 6]), ([0, 0, 0], [0, 0, 0]), ([5, 5, 5], [5, 9, 5]), ([-1e4, 3, -2e-3], [1e4, 3.5, 2e-3]),
                   ([0, 0, 0], [1e-40, 0, 0]), ([-3e37, 0, 0], [3e37, 1, 1])):
        lo, hi = np.array(lo, F32), np.array(hi, F32)
        g0, s = make_grid(lo, hi)
        assert (s > 0).all() and np.isfinite(F32(1.0) / s).all() and np.isfinite(g0).all()
        ql, qh = qpair(lo[None], hi[None], g0, s)
        assert (ql >= 0).all() and (qh <= 65535).all()
