"""GPU parity for BN254 G2: Fq2 / G2 operation pyramid and the G2 MSM through the C ABI, against
oracle/bn254_g2.py.  Bar: bit-exact Fq2 words for the field ops, equal group element for points."""
import random

import numpy as np
import pytest

import bn254 as o
import bn254_g2 as g2
import helpers as h

pytestmark = pytest.mark.gpu


def _pack_f2(vals):
    a = np.zeros((len(vals), 8), dtype=np.uint64)
    for i, (c0, c1) in enumerate(vals):
        a[i, 0:4] = h.words(o.to_mont(c0))
        a[i, 4:8] = h.words(o.to_mont(c1))
    return a


def _unpack_f2(arr):
    return [(o.from_mont(h.unwords(r[0:4])), o.from_mont(h.unwords(r[4:8]))) for r in arr]


def _pack_xyzz(pts, rng):
    """Random XYZZ representatives (X, Y, ZZ, ZZZ) = (x z^2, y z^3, z^2, z^3) of affine points; None -> infinity."""
    a = np.zeros((len(pts), 32), dtype=np.uint64)
    for i, pt in enumerate(pts):
        if pt is None:
            continue
        z = (rng.randrange(1, o.P), rng.randrange(o.P))
        zz = g2.f2_sqr(z)
        zzz = g2.f2_mul(zz, z)
        a[i] = _pack_f2([g2.f2_mul(pt[0], zz), g2.f2_mul(pt[1], zzz), zz, zzz]).reshape(-1)
    return a


def _xyzz_to_affine(rec):
    x, y, zz, zzz = _unpack_f2(np.asarray(rec).reshape(4, 8))
    if g2.f2_is_zero(zz):
        return None
    return (g2.f2_mul(x, g2.f2_inv(zz)), g2.f2_mul(y, g2.f2_inv(zzz)))


def _pack_bases(pts):
    return np.array([g2.encode_base(pt) for pt in pts], dtype=np.uint64)


def test_fq2_mul_sqr(ctx):
    rng = random.Random(21)
    edge = [(0, 0), (1, 0), (0, 1), (o.P - 1, o.P - 1), (o.P - 1, 0), (0, o.P - 1)]
    cs = [(a, b) for a in edge for b in edge] + [((rng.randrange(o.P), rng.randrange(o.P)), (rng.randrange(o.P), rng.randrange(o.P)))
                                                 for _ in range(200)]
    a = _pack_f2([x for x, _ in cs])
    b = _pack_f2([y for _, y in cs])
    assert _unpack_f2(ctx.testkit_op(30, a, b, 8)) == [g2.f2_mul(x, y) for x, y in cs]
    assert _unpack_f2(ctx.testkit_op(31, a, None, 8)) == [g2.f2_sqr(x) for x, _ in cs]


def test_g2_point_ops_complete(ctx):
    rng = random.Random(22)
    pts = g2.random_points(16, 5)
    # madd: generic, acc infinity, P + P, P + (-P)
    accs = pts[:8] + [None, pts[3], pts[4]]
    adds = pts[8:16] + [pts[0], pts[3], g2.affine_neg(pts[4])]
    out = ctx.testkit_op(32, _pack_xyzz(accs, rng), _pack_bases(adds)[:, :16].copy(), 32)
    for a, b, r in zip(accs, adds, out):
        want = g2.jac_to_affine(g2.jac_add(g2.affine_to_jac(a), g2.affine_to_jac(b)))
        assert _xyzz_to_affine(r) == want
    # add: generic, either side infinity, doubling through different representatives, cancellation
    lhs = pts[:6] + [None, pts[1], pts[2], pts[5]]
    rhs = pts[6:12] + [pts[0], None, pts[2], g2.affine_neg(pts[5])]
    out = ctx.testkit_op(33, _pack_xyzz(lhs, rng), _pack_xyzz(rhs, rng), 32)
    for a, b, r in zip(lhs, rhs, out):
        want = g2.jac_to_affine(g2.jac_add(g2.affine_to_jac(a), g2.affine_to_jac(b)))
        assert _xyzz_to_affine(r) == want
    out = ctx.testkit_op(34, _pack_xyzz(pts[:6] + [None], rng), None, 32)
    for a, r in zip(pts[:6] + [None], out):
        assert _xyzz_to_affine(r) == g2.jac_to_affine(g2.jac_dbl(g2.affine_to_jac(a)))


@pytest.mark.parametrize("n", [1, 2, 33, 300, 2050])
def test_g2_msm_matches_oracle(ctx, n):
    pts = g2.random_points(n, 100 + n)
    sc = o.random_scalars(n, 200 + n)
    if n > 10:
        pts[2] = None
        sc[3] = 0
        sc[4] = o.R_ORDER - 1
        pts[6] = pts[5]
        sc[6] = sc[5]
        pts[8] = g2.affine_neg(pts[7])
        sc[8] = sc[7]
    want = g2.jac_to_affine(g2.msm_pippenger(pts, sc, 8 if n > 64 else 4))
    bases, scal = _pack_bases(pts), h.pack_scalars(sc)
    # scalar split (phi acts through beta^2) / plain; sliced upload; cooperative K4 levels / thread-per-segment K4
    for glv, slices, coop in ((-1, 0, -1), (0, 0, -1), (-1, 2, -1), (0, 5, -1), (-1, 0, 0), (0, 3, 0)):
        for w in ((0, 5, 13) if n <= 300 else (0,)):
            ctx.set_option("window_bits", w)
            ctx.set_option("glv", glv)
            ctx.set_option("slices", slices)
            ctx.set_option("coop_reduce", coop)
            try:
                got = g2.jac_to_affine(g2.decode_jacobian(ctx.msm_g2(bases, scal)))
            finally:
                ctx.set_option("window_bits", 0)
                ctx.set_option("glv", -1)
                ctx.set_option("slices", 0)
                ctx.set_option("coop_reduce", -1)
            assert got == want, (n, w, glv, slices, coop)
    # 128-byte records without the flag word
    if n == 33:
        keep = [i for i, pt in enumerate(pts) if pt is not None]
        got = g2.jac_to_affine(g2.decode_jacobian(ctx.msm_g2(bases[keep][:, :16].copy(), scal[keep])))
        assert got == g2.jac_to_affine(g2.msm_naive([pts[i] for i in keep], [sc[i] for i in keep]))


def test_g2_msm_2_16_checksum(ctx):
    """2^16 distinct G2 points built as T1[i % 256] + T2[i // 256] of generator multiples: the expected MSM is one scalar
    multiplication of the generator by sum_i s_i (a_i + b_i)."""
    n = 1 << 16
    rng = random.Random(77)
    t1 = [rng.randrange(1, o.R_ORDER) for _ in range(256)]
    t2 = [rng.randrange(1, o.R_ORDER) for _ in range(256)]
    G = g2.affine_to_jac(g2.GEN2)
    T1 = [g2.jac_scalar_mul(k, G) for k in t1]
    T2 = [g2.jac_scalar_mul(k, G) for k in t2]
    jacs = [g2.jac_add(T1[i & 255], T2[i >> 8]) for i in range(n)]
    # batch normalisation (Montgomery's trick over Fq2)
    pref = [g2.F2_ONE]
    for j in jacs:
        pref.append(g2.f2_mul(pref[-1], j[2]))
    inv = g2.f2_inv(pref[-1])
    pts = [None] * n
    for i in reversed(range(n)):
        zi = g2.f2_mul(inv, pref[i])
        inv = g2.f2_mul(inv, jacs[i][2])
        zi2 = g2.f2_sqr(zi)
        pts[i] = (g2.f2_mul(jacs[i][0], zi2), g2.f2_mul(jacs[i][1], g2.f2_mul(zi2, zi)))
    sc = o.random_scalars(n, 78)
    dlog = sum(s * (t1[i & 255] + t2[i >> 8]) for i, s in enumerate(sc)) % o.R_ORDER
    want = g2.jac_to_affine(g2.jac_scalar_mul(dlog, G))
    bases, scal = _pack_bases(pts), h.pack_scalars(sc)
    for slices in (1, 3):        # unsliced, and the sliced upload pipeline (merged per-slice bucket arrays)
        ctx.set_option("slices", slices)
        try:
            got = g2.jac_to_affine(g2.decode_jacobian(ctx.msm_g2(bases, scal)))
        finally:
            ctx.set_option("slices", 0)
        assert got == want, slices


def test_g2_msm_skewed_scalars(ctx):
    """All-equal and 0/1 scalars put thousands of entries into single buckets (the per-CTA long-bucket fold); a narrow top
    window (c = 12: two bits) does the same for uniform scalars."""
    n = 3000
    pts = g2.random_points(n, 909)
    total = g2.JAC2_INF
    for pt in pts:
        total = g2.jac_add(total, g2.affine_to_jac(pt))
    bases = _pack_bases(pts)
    k = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
    for s, w in ((1, 0), (k, 0), (k, 12), (o.R_ORDER - 1, 8)):
        ctx.set_option("window_bits", w)
        try:
            got = g2.jac_to_affine(g2.decode_jacobian(ctx.msm_g2(bases, h.pack_scalars([s] * n))))
        finally:
            ctx.set_option("window_bits", 0)
        assert got == g2.jac_to_affine(g2.jac_scalar_mul(s, total)), (hex(s), w)
    sc = o.random_scalars(n, 910)
    ctx.set_option("window_bits", 12)
    try:
        got = g2.jac_to_affine(g2.decode_jacobian(ctx.msm_g2(bases, h.pack_scalars(sc))))
    finally:
        ctx.set_option("window_bits", 0)
    assert got == g2.jac_to_affine(g2.msm_pippenger(pts, sc, 8))


@pytest.mark.parametrize("precompute", [0, 1, 13, 20])
def test_g2_registered_bases(ctx, precompute):
    """B2 bases of a proving key stay on the GPU; with precompute the window table 2^(c w) P is built once and every MSM
    over the handle uses one bucket set.  Full, shorter, zero and skewed scalar sets; infinity records in the set."""
    n = 257
    pts = g2.random_points(n, 333)
    pts[0] = None
    pts[100] = None
    hnd = ctx.g2_register_bases(_pack_bases(pts), precompute)
    try:
        for sc in (o.random_scalars(n, 40), o.random_scalars(100, 41), [0] * n, [1] * n, [o.R_ORDER - 1] * 50):
            got = g2.jac_to_affine(g2.decode_jacobian(ctx.g2_msm_registered(hnd, h.pack_scalars(sc))))
            assert got == g2.jac_to_affine(g2.msm_pippenger(pts[: len(sc)], sc, 6)), (precompute, len(sc))
        with pytest.raises(Exception):
            ctx.g2_msm_registered(hnd, h.pack_scalars([1] * (n + 1)))      # more scalars than registered bases
    finally:
        ctx.g2_release_bases(hnd)


def test_g2_registered_2_16_table(ctx):
    """2^16 registered G2 bases with the automatic window table against the generator checksum."""
    n = 1 << 16
    rng = random.Random(79)
    t1 = [rng.randrange(1, o.R_ORDER) for _ in range(256)]
    t2 = [rng.randrange(1, o.R_ORDER) for _ in range(256)]
    G = g2.affine_to_jac(g2.GEN2)
    T1 = [g2.jac_to_affine(g2.jac_scalar_mul(k, G)) for k in t1]
    T2 = [g2.jac_scalar_mul(k, G) for k in t2]
    # points on the device path only need to be valid curve points: build them with mixed additions, normalise in batch
    jacs = [g2.jac_add(g2.affine_to_jac(T1[i & 255]), T2[i >> 8]) for i in range(n)]
    pref = [g2.F2_ONE]
    for j in jacs:
        pref.append(g2.f2_mul(pref[-1], j[2]))
    inv = g2.f2_inv(pref[-1])
    pts = [None] * n
    for i in reversed(range(n)):
        zi = g2.f2_mul(inv, pref[i])
        inv = g2.f2_mul(inv, jacs[i][2])
        zi2 = g2.f2_sqr(zi)
        pts[i] = (g2.f2_mul(jacs[i][0], zi2), g2.f2_mul(jacs[i][1], g2.f2_mul(zi2, zi)))
    sc = o.random_scalars(n, 80)
    dlog = sum(s * (t1[i & 255] + t2[i >> 8]) for i, s in enumerate(sc)) % o.R_ORDER
    want = g2.jac_to_affine(g2.jac_scalar_mul(dlog, G))
    hnd = ctx.g2_register_bases(_pack_bases(pts), 1)
    try:
        assert g2.jac_to_affine(g2.decode_jacobian(ctx.g2_msm_registered(hnd, h.pack_scalars(sc)))) == want
    finally:
        ctx.g2_release_bases(hnd)


def _xyzz_rec_to_affine(rec):
    return _xyzz_to_affine(rec)


@pytest.mark.parametrize("n,w", [(1, 6), (3, 4), (60, 5), (300, 8), (700, 11), (1500, 13)])
@pytest.mark.parametrize("glv", [0, 1])
@pytest.mark.parametrize("coop", [1, 0])
def test_g2_window_sums(ctx, n, w, glv, coop):
    """Stage 4 of the G2 pipeline -- coop = 1: the cooperative levels (k_g2_reduce_level), coop = 0: k_g2_bucket_reduce +
    k_g2_window_finish (g2_block_weighted_sum): the per-window sums
    G_w = sum_m m * bucket[w][m] against a bucket-by-bucket restatement over oracle/bn254_g2.py -- the G2 counterpart of
    test_gpu_stages.py::test_window_sums (reference semantics: smvp.metal:14-107 + pbpr.metal:33-148)."""
    pts = g2.random_points(n, 40 + n)
    sc = o.random_scalars(n, 41 + n)
    bases = _pack_bases(pts)[:, :16].copy()
    ctx.set_option("glv", glv)
    ctx.set_option("coop_reduce", coop)
    try:
        got = ctx.testkit_g2_window_sums(bases, h.pack_scalars(sc), w)
    finally:
        ctx.set_option("glv", -1)
        ctx.set_option("coop_reduce", -1)
    if glv:
        ks, pp = [], []
        halves = [o.glv_decompose(s) for s in sc]
        pp = list(pts) + [g2.glv_phi(pt) for pt in pts]
        ks = [k1 for k1, _ in halves] + [k2 for _, k2 in halves]
    else:
        pp, ks = list(pts), list(sc)
    K = o.num_windows_for(w, 127 if glv else 254)
    assert len(got) == K
    digs = [o.signed_digits_signed(k, w, K) for k in ks]
    for k in range(K):
        total = g2.JAC2_INF
        for pt, d in zip(pp, digs):
            if d[k] == 0:
                continue
            q = g2.affine_to_jac(pt if d[k] > 0 else g2.affine_neg(pt))
            total = g2.jac_add(total, g2.jac_scalar_mul_raw(abs(d[k]), q))
        assert _xyzz_rec_to_affine(got[k]) == g2.jac_to_affine(total), (n, w, k)
