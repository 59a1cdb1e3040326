"""GPU parity, bottom of the pyramid: limb/field -> Montgomery -> curve ops, through the C ABI's
test-kit entry point (which runs the PRODUCTION device functions, one thread per element).
Mirrors the reference's tests/{bigint,field,mont_backend,curve}/ levels; bar: bit-exact."""
import random

import numpy as np
import pytest

import bn254 as o
import helpers as h

pytestmark = pytest.mark.gpu

EDGE = [0, 1, 2, o.P - 1, o.P - 2, (1 << 253), (1 << 32) - 1, (1 << 64), o.P >> 1]


def _pairs(k=400):
    rng = random.Random(5)
    cs = [(a, b) for a in EDGE for b in EDGE]
    cs += [(rng.randrange(o.P), rng.randrange(o.P)) for _ in range(k)]
    return cs


def test_fq_mul_add_sub_sqr(ctx):
    cs = _pairs()
    a = h.pack_fq([x for x, _ in cs])
    b = h.pack_fq([y for _, y in cs])
    # operands are passed as Montgomery words; op results are Montgomery words
    got = h.unpack_fq(ctx.testkit_op(0, a, b, 4))
    assert got == [x * y % o.P for x, y in cs]  # montmul(aR, bR) = abR  (mont_mul_cios.rs:19-76)
    assert h.unpack_fq(ctx.testkit_op(1, a, b, 4)) == [(x + y) % o.P for x, y in cs]
    assert h.unpack_fq(ctx.testkit_op(2, a, b, 4)) == [(x - y) % o.P for x, y in cs]
    assert h.unpack_fq(ctx.testkit_op(3, a, None, 4)) == [x * x % o.P for x, _ in cs]


def test_fq_raw_words_are_canonical(ctx):
    # limb-exact: raw output words equal the oracle's Montgomery integer (fully reduced)
    cs = _pairs(100)
    a = h.pack_fq([x for x, _ in cs])
    b = h.pack_fq([y for _, y in cs])
    out = ctx.testkit_op(0, a, b, 4)
    for (x, y), r in zip(cs, out):
        assert h.unwords(r) == o.to_mont(x * y % o.P)


def test_fr_from_mont(ctx):
    rng = random.Random(8)
    vals = [0, 1, o.R_ORDER - 1, 1 << 253] + [rng.randrange(o.R_ORDER) for _ in range(300)]
    out = ctx.testkit_op(20, h.pack_scalars(vals), None, 4)
    assert [h.unwords(r) for r in out] == vals


def _rand_xyzz(pt, rng):
    if pt is None:
        return o.XYZZ_INF
    z = rng.randrange(1, o.P)
    zz, zzz = z * z % o.P, z * z * z % o.P
    return (pt[0] * zz % o.P, pt[1] * zzz % o.P, zz, zzz)


def test_xyzz_madd_complete(ctx):
    rng = random.Random(13)
    pts = o.random_points(64, 77)
    accs, adds, want = [], [], []
    for i in range(60):
        accs.append(_rand_xyzz(pts[i], rng)); adds.append(pts[(i * 7 + 1) % 64])
    # inf + P, P + P (different representative), P + (-P)
    accs += [o.XYZZ_INF, _rand_xyzz(pts[3], rng), _rand_xyzz(pts[4], rng), (pts[5][0], pts[5][1], 1, 1)]
    adds += [pts[0], pts[3], o.affine_neg(pts[4]), pts[5]]
    for a, p in zip(accs, adds):
        want.append(o.xyzz_to_affine(o.xyzz_madd(a, p)))
    badd = h.pack_bases(adds, with_inf=False)
    got = h.unpack_xyzz(ctx.testkit_op(10, h.pack_xyzz(accs), badd, 16))
    assert [o.xyzz_to_affine(g) for g in got] == want
    for g in got:  # XYZZ invariant ZZ^3 == ZZZ^2
        assert pow(g[2], 3, o.P) == pow(g[3], 2, o.P)


def test_xyzz_add_dbl_complete(ctx):
    rng = random.Random(14)
    pts = o.random_points(40, 78)
    A = [_rand_xyzz(pts[i], rng) for i in range(30)]
    B = [_rand_xyzz(pts[(i * 3 + 2) % 40], rng) for i in range(30)]
    # inf+inf, inf+P, P+inf, P+P (different reps), P+(-P)
    A += [o.XYZZ_INF, o.XYZZ_INF, _rand_xyzz(pts[1], rng), _rand_xyzz(pts[2], rng), _rand_xyzz(pts[6], rng)]
    B += [o.XYZZ_INF, _rand_xyzz(pts[0], rng), o.XYZZ_INF, _rand_xyzz(pts[2], rng), _rand_xyzz(o.affine_neg(pts[6]), rng)]
    got = h.unpack_xyzz(ctx.testkit_op(11, h.pack_xyzz(A), h.pack_xyzz(B), 16))
    assert [o.xyzz_to_affine(g) for g in got] == [o.xyzz_to_affine(o.xyzz_add(a, b)) for a, b in zip(A, B)]
    got = h.unpack_xyzz(ctx.testkit_op(12, h.pack_xyzz(A), None, 16))
    assert [o.xyzz_to_affine(g) for g in got] == [o.xyzz_to_affine(o.xyzz_dbl(a)) for a in A]


def test_xyzz_to_jacobian(ctx):
    rng = random.Random(15)
    pts = o.random_points(20, 79)
    A = [_rand_xyzz(p, rng) for p in pts] + [o.XYZZ_INF]
    out = ctx.testkit_op(13, h.pack_xyzz(A), None, 12)
    for a, r in zip(A, out):
        assert o.jac_to_affine(o.decode_jacobian(r)) == o.xyzz_to_affine(a)
    # infinity is arkworks' Projective::zero() = (R, R, 0)
    assert h.unwords(out[-1][0:4]) == o.R_MOD_P and h.unwords(out[-1][8:12]) == 0


def test_fq_neg_inv_dbl(ctx):
    rng = random.Random(12)
    vals = EDGE + [rng.randrange(o.P) for _ in range(200)]
    a = h.pack_fq(vals)
    assert h.unpack_fq(ctx.testkit_op(4, a, None, 4)) == [(-x) % o.P for x in vals]          # ff "p - a", 0 -> 0
    assert h.unpack_fq(ctx.testkit_op(6, a, None, 4)) == [2 * x % o.P for x in vals]
    assert h.unpack_fq(ctx.testkit_op(5, a, None, 4)) == [pow(x, o.P - 2, o.P) for x in vals]  # Fermat; inv(0) = 0


def test_fq_mulsub_fused(ctx):
    """fq_mulsub (a*b - c*d with one Montgomery reduction, the Y3 step of every XYZZ addition) -- bit-exact words."""
    rng = random.Random(14)
    quads = [(a, b, c, d) for a in EDGE[:5] for b in EDGE[:5] for c in EDGE[:5] for d in EDGE[:5]]
    quads += [(o.P - 1, o.P - 1, 0, 0), (o.P - 1, o.P - 1, o.P - 1, 0), (5, 7, 7, 5)]
    quads += [tuple(rng.randrange(o.P) for _ in range(4)) for _ in range(3000)]
    ac = np.concatenate([h.pack_fq([q[0] for q in quads]), h.pack_fq([q[2] for q in quads])], axis=1)
    bd = np.concatenate([h.pack_fq([q[1] for q in quads]), h.pack_fq([q[3] for q in quads])], axis=1)
    got = h.unpack_fq(ctx.testkit_op(8, np.ascontiguousarray(ac), np.ascontiguousarray(bd), 4))
    assert got == [(a * b - c * d) % o.P for a, b, c, d in quads]


def test_fq_inv_safegcd(ctx):
    """fq_inv_by (Bernstein-Yang divsteps, the inversion of the batched-affine accumulation) against Python's modular
    inverse: edge values, small values, values next to p and to powers of two, 20 000 random ones; inv(0) = 0."""
    rng = random.Random(13)
    vals = EDGE + list(range(1, 70)) + [o.P - k for k in range(1, 70)] + [1 << k for k in range(1, 254)]
    vals += [(1 << k) - 1 for k in range(2, 254)] + [rng.randrange(o.P) for _ in range(20000)]
    vals = [v % o.P for v in vals]
    got = h.unpack_fq(ctx.testkit_op(7, h.pack_fq(vals), None, 4))
    assert got == [pow(x, o.P - 2, o.P) for x in vals]


def test_jacobian_dbl_2009_l(ctx):
    """The reference's jacobian_dbl_2009_l level (tests/curve/jacobian_dbl_2009_l.rs): same point as the oracle's doubling,
    on random Jacobian representatives."""
    rng = random.Random(13)
    pts = o.random_points(40, 130)
    a = np.zeros((len(pts), 12), dtype=np.uint64)
    for i, pt in enumerate(pts):
        z = rng.randrange(1, o.P)
        j = (pt[0] * z * z % o.P, pt[1] * z * z * z % o.P, z)
        for c in range(3):
            a[i, 4 * c:4 * c + 4] = h.words(o.to_mont(j[c]))
    out = ctx.testkit_op(14, a, None, 12)
    for pt, r in zip(pts, out):
        assert o.jac_to_affine(o.decode_jacobian(r)) == o.jac_to_affine(o.jac_dbl(o.affine_to_jac(pt)))


def test_scalar_mul_u32(ctx):
    """The reference's jacobian_scalar_mul level (tests/curve/jacobian_scalar_mul.rs): k * P for 32-bit k."""
    rng = random.Random(14)
    pts = o.random_points(24, 140)
    ks = [0, 1, 2, 3, (1 << 32) - 1, 1 << 31] + [rng.randrange(1 << 32) for _ in range(18)]
    b = np.array(ks, dtype=np.uint64).reshape(-1, 1)
    out = h.unpack_xyzz(ctx.testkit_op(15, h.pack_bases(pts, with_inf=False), b, 16))
    for pt, k, r in zip(pts, ks, out):
        assert o.xyzz_to_affine(r) == o.jac_to_affine(o.jac_scalar_mul(k, o.affine_to_jac(pt)))


def test_device_math_library_example_runs():
    """cpp/example_math.cu: a third-party kernel built only from include/b200math.cuh walks k*G for k = 1..64 (madd from
    infinity, the doubling path, generic additions), normalises with the safegcd inversion, checks the curve equation and
    Fermat == safegcd, and P + (-P) = infinity."""
    import os
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "example_math")
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I",
                        os.path.join(root, "include"), "-o", exe, os.path.join(root, "gpu-acceleration_b200", "cpp", "example_math.cu")],
                       check=True, capture_output=True)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        # 64 G, x coordinate, low Montgomery limb -- against the oracle
        want = o.to_mont(o.jac_to_affine(o.jac_scalar_mul(64, o.affine_to_jac(o.GEN)))[0]) & 0xFFFFFFFF
        assert f"x64G_mont_limb0={want:08x}" in r.stdout
