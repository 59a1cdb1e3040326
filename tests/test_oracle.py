"""CPU suite: the oracle against the reference's known-answer literals, public BN254 vectors,
its own two MSM formulations, the stage restatements and the committed golden fixtures."""
import os
import random

import numpy as np

import bn254 as o
import helpers as h

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "msm_cases.npz")


def test_reference_literals():
    # p, R mod p, n0, R^-1, Barrett mu, r: constants.metal:30-47,229-246; mont_params.rs:9,116-122; barrett_params.rs:25-28
    o.self_check()


def test_public_curve_vectors():
    # alt_bn128 (EIP-196) ECADD/ECMUL known answers: 2G and 3G
    g2 = o.jac_to_affine(o.jac_dbl(o.affine_to_jac(o.GEN)))
    assert g2 == (1368015179489954701390400359078579693043519447331113978918064868415326638035,
                  9918110051302171585080402603319702774565515993150576347155970296011118125764)
    g3 = o.jac_to_affine(o.jac_add(o.affine_to_jac(g2), o.affine_to_jac(o.GEN)))
    assert g3 == (3353031288059533942658390886683067124040920775575537747144343083137631628272,
                  19321533766552368860946552437480515441416830039777911637913418824951667761761)
    assert o.jac_to_affine(o.jac_scalar_mul(3, o.affine_to_jac(o.GEN))) == g3


def test_montgomery_forms_agree():
    rng = random.Random(7)
    for _ in range(200):
        a, b = rng.randrange(o.P), rng.randrange(o.P)
        assert o.mont_mul_cios32(a, b) == o.mont_mul(a, b)
        assert o.from_mont(o.to_mont(a)) == a
    for _ in range(50):
        a = rng.randrange(o.R_ORDER)
        assert o.mont_mul_cios32(o.to_mont(a, o.R_ORDER), 1, o.R_ORDER) == a  # Fr Montgomery -> canonical


def test_group_law_edge_cases():
    # the cases tests/curve/jacobian_add_2007_b1.rs:123-157 lists (inf+G, inf+inf, G+G, P+P) plus P+(-P)
    G = o.affine_to_jac(o.GEN)
    assert o.jac_eq(o.jac_add(o.JAC_INF, G), G) and o.jac_eq(o.jac_add(G, o.JAC_INF), G)
    assert o.jac_is_inf(o.jac_add(o.JAC_INF, o.JAC_INF))
    assert o.jac_eq(o.jac_add(G, G), o.jac_dbl(G))
    P = o.jac_scalar_mul(12345, G)
    P2 = (P[0] * 9 % o.P, P[1] * 27 % o.P, P[2] * 3 % o.P)  # same point, different representative
    assert o.jac_eq(P, P2) and o.jac_eq(o.jac_add(P, P2), o.jac_dbl(P))
    assert o.jac_is_inf(o.jac_add(P, o.jac_neg(P2)))
    # XYZZ formulas agree with Jacobian ones
    pa = o.jac_to_affine(P)
    x = o.xyzz_madd(o.XYZZ_INF, pa)
    x = o.xyzz_madd(x, pa)  # doubling path
    assert o.xyzz_to_affine(x) == o.jac_to_affine(o.jac_dbl(P))
    x = o.xyzz_madd(x, o.affine_neg(o.jac_to_affine(o.jac_dbl(P))))
    assert o.xyzz_to_affine(x) is None
    q = o.xyzz_add(o.xyzz_madd(o.XYZZ_INF, pa), o.xyzz_dbl_affine(o.GEN))
    assert o.xyzz_to_affine(q) == o.jac_to_affine(o.jac_add(P, o.jac_dbl(G)))
    assert o.jac_to_affine(o.xyzz_to_jac(q)) == o.xyzz_to_affine(q)


def test_signed_digits_recompose():
    rng = random.Random(3)
    for w in (4, 5, 8, 11, 13, 15, 16, 17, 20, 22):
        K = o.num_windows_for(w)
        for s in [0, 1, o.R_ORDER - 1, (1 << 253), (1 << (w - 1)), (1 << w) - 1] + [rng.randrange(o.R_ORDER) for _ in range(50)]:
            d = o.signed_digits(s, w, K)
            assert all(-(1 << (w - 1)) <= x < (1 << (w - 1)) for x in d)
            assert sum(x << (w * k) for k, x in enumerate(d)) == s


def test_pippenger_matches_definition():
    pts = o.random_points(33, 11)
    sc = o.random_scalars(33, 12)
    want = o.jac_to_affine(o.msm_naive(pts, sc))
    for w in (3, 4, 8, 13, 16):
        assert o.jac_to_affine(o.msm_pippenger(pts, sc, w)) == want
    assert o.jac_to_affine(o.msm_pippenger(pts, sc)) == want  # arkworks window rule
    assert o.ark_window_size(1 << 20) == 15 and o.ark_window_size(1 << 24) == 18 and o.ark_window_size(10) == 3


def test_stage_restatements_compose():
    """digits -> CSC -> bucket sums -> running-sum reduce -> Horner == MSM."""
    n, w = 40, 5
    pts = o.random_points(n, 21)
    sc = o.random_scalars(n, 22)
    half = 1 << (w - 1)
    rows = o.stage_rows(sc, w)
    acc = o.JAC_INF
    for k in reversed(range(len(rows))):
        col_ptr, val_idx = o.stage_csc(rows[k], 1 << w)
        assert col_ptr[-1] == n and sorted(val_idx) == list(range(n))
        for d in range(1 << w):  # stable: indices ascending inside a column
            seg = val_idx[col_ptr[d]:col_ptr[d + 1]]
            assert seg == sorted(seg) and all(rows[k][i] == d for i in seg)
        digits = [r - half for r in rows[k]]
        buckets = o.stage_bucket_sums(pts, digits, half)
        gk = o.stage_bucket_reduce(buckets)
        for _ in range(w):
            acc = o.jac_dbl(acc)
        acc = o.jac_add(acc, o.xyzz_to_jac(gk))
    assert o.jac_to_affine(acc) == o.jac_to_affine(o.msm_naive(pts, sc))


def test_golden_fixtures():
    z = np.load(GOLDEN)
    for name in z["names"]:
        bases, scalars, exp = z[f"{name}/bases"], z[f"{name}/scalars"], z[f"{name}/expected"]
        pts = [None if int(b[8]) else (o.from_mont(h.unwords(b[0:4])), o.from_mont(h.unwords(b[4:8]))) for b in bases]
        assert all(o.is_on_curve(p) for p in pts)
        sc = [o.from_mont(h.unwords(s), o.R_ORDER) for s in scalars]
        got = o.jac_to_affine(o.msm_pippenger(pts, sc, 6))
        want = None if int(exp[8]) else (h.unwords(exp[0:4]), h.unwords(exp[4:8]))
        assert got == want, name


def test_encode_roundtrip():
    pts = o.random_points(5, 5) + [None]
    raw = o.encode_bases(pts, stride=72, inf_off=64)
    arr = np.frombuffer(raw, dtype=np.uint64).reshape(-1, 9)
    assert np.array_equal(arr, h.pack_bases(pts))
    assert np.frombuffer(o.encode_scalars([1, 2]), dtype=np.uint64).reshape(-1, 4).tolist() == h.pack_scalars([1, 2]).tolist()


def test_glv_restatement():
    """Engine-internal GLV split: constants, identity k1 + k2*lambda = s, the 127-bit bound, and the
    equivalence of the expanded 2n-term MSM with the original."""
    o.glv_self_check()
    rng = random.Random(17)
    r = o.R_ORDER
    for s in [0, 1, 2, r - 1, r // 2, r // 2 + 1, o.GLV_LAMBDA, r - o.GLV_LAMBDA, 1 << 253] + [rng.randrange(r) for _ in range(20000)]:
        k1, k2 = o.glv_decompose(s)
        assert (k1 + k2 * o.GLV_LAMBDA - s) % r == 0 and abs(k1) < (1 << 127) and abs(k2) < (1 << 127)
        for k in (k1, k2):
            d = o.signed_digits_signed(k, 16, o.num_windows_for(16, 127))
            assert sum(x << (16 * i) for i, x in enumerate(d)) == k and all(-32768 <= x <= 32767 for x in d)
    # the int16 corner: a negative value whose plain recoding would contain -32768
    d = o.signed_digits_signed(-(0x8000 + (0x8000 << 16)), 16, 8)
    assert sum(x << (16 * i) for i, x in enumerate(d)) == -(0x8000 + (0x8000 << 16)) and all(-32768 <= x <= 32767 for x in d)
    if True:
        pass
    pts = o.random_points(12, 61) + [None]
    sc = o.random_scalars(13, 62)
    p2, k2s = o.glv_expand(pts, sc)
    acc = o.JAC_INF
    for pt, k in zip(p2, k2s):
        if pt is not None and k:
            acc = o.jac_add(acc, o.jac_scalar_mul(abs(k), o.affine_to_jac(pt if k > 0 else o.affine_neg(pt))))
    assert o.jac_to_affine(acc) == o.jac_to_affine(o.msm_naive(pts, sc))


def test_reference_srs_points_pin_memory_layout():
    """Externally produced BN254 G1 points found in the reference tree (example-app/ios/*_srs.bin; fixture made
    by tests/golden/make_ref_srs_fixture.py): raw Montgomery x || y words.  Every one must decode to a curve
    point under the oracle's conventions (R = 2^256, LE limbs), and the first of each file is the generator."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_srs_g1.npz"))
    total = 0
    for key in ("plonk", "gemini", "hyperplonk"):
        pts = z[key]
        for row in pts:
            pt = (o.from_mont(h.unwords(row[0:4])), o.from_mont(h.unwords(row[4:8])))
            assert o.is_on_curve(pt)
            total += 1
        assert (o.from_mont(h.unwords(pts[0][0:4])), o.from_mont(h.unwords(pts[0][4:8]))) == o.GEN
    assert total == 80


def test_table_mode_restatement():
    """The precomputed-table formulation (one bucket set, no Horner) is the same group element as the definition."""
    pts = o.random_points(24, 5150)
    pts[3] = None
    sc = o.random_scalars(24, 5151)
    sc[5] = 0
    sc[6] = o.R_ORDER - 1
    want = o.jac_to_affine(o.msm_naive(pts, sc))
    for w in (8, 13, 16):
        tab = o.table_expand(pts[:4], w)
        assert len(tab) == o.num_windows_for(w) and tab[0] == pts[:4] and tab[1][3] is None
        assert tab[2][0] == o.jac_to_affine(o.jac_scalar_mul(1 << (2 * w), o.affine_to_jac(pts[0])))
    for w in (5, 8):
        assert o.jac_to_affine(o.msm_table_mode(pts, sc, w)) == want
