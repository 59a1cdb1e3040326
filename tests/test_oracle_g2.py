"""CPU suite: the G2 oracle against the public alt_bn128 generator and against itself (definition vs bucket method)."""
import bn254 as o
import bn254_g2 as g2


def test_generator_curve_and_order():
    g2.self_check()
    assert g2.f2_mul(g2.B2, (9, 1)) == (3, 0)            # b' (9 + u) = 3


def test_field_and_group_laws():
    a, b = (5, 7), (o.P - 3, 11)
    assert g2.f2_mul(a, g2.f2_inv(a)) == g2.F2_ONE
    assert g2.f2_mul((0, 1), (0, 1)) == (o.P - 1, 0)     # u^2 = -1
    assert g2.f2_sub(g2.f2_add(a, b), b) == a
    G = g2.affine_to_jac(g2.GEN2)
    p5 = g2.jac_scalar_mul(5, G)
    assert g2.jac_to_affine(g2.jac_add(g2.jac_scalar_mul(2, G), g2.jac_scalar_mul(3, G))) == g2.jac_to_affine(p5)
    assert g2.jac_is_inf(g2.jac_add(p5, g2.affine_to_jac(g2.affine_neg(g2.jac_to_affine(p5)))))
    assert g2.jac_to_affine(g2.jac_add(p5, p5)) == g2.jac_to_affine(g2.jac_dbl(p5))
    assert g2.is_on_curve(g2.jac_to_affine(p5))


def test_pippenger_matches_definition():
    pts = g2.random_points(20, 7)
    pts[3] = None
    sc = o.random_scalars(20, 8)
    sc[4] = 0
    sc[5] = o.R_ORDER - 1
    want = g2.jac_to_affine(g2.msm_naive(pts, sc))
    for w in (4, 7):
        assert g2.jac_to_affine(g2.msm_pippenger(pts, sc, w)) == want


def test_memory_encoding_roundtrip():
    pt = g2.jac_to_affine(g2.jac_scalar_mul(12345, g2.affine_to_jac(g2.GEN2)))
    w = g2.encode_base(pt)
    assert len(w) == 17 and w[16] == 0 and g2.encode_base(None)[16] == 1
    jac = g2.decode_jacobian(w[:16] + g2.words(o.to_mont(1)) + [0, 0, 0, 0])   # Z = 1
    assert g2.jac_to_affine(jac) == pt


def test_glv_endomorphism_on_g2():
    """phi(x, y) = (beta^2 x, y) is multiplication by lambda on G2 (with beta itself it is lambda^2), so the G1 scalar split
    s = k1 + k2 * lambda carries over: s P = k1 P + k2 phi(P)."""
    P = g2.jac_to_affine(g2.jac_scalar_mul(987654321, g2.affine_to_jac(g2.GEN2)))
    lam = o.GLV_LAMBDA
    assert g2.glv_phi(P) == g2.jac_to_affine(g2.jac_scalar_mul(lam, g2.affine_to_jac(P)))
    wrong = (g2.f2_scale(P[0], o.GLV_BETA), P[1])
    assert wrong == g2.jac_to_affine(g2.jac_scalar_mul(lam * lam % o.R_ORDER, g2.affine_to_jac(P)))
    s = o.random_scalars(1, 31)[0]
    k1, k2 = o.glv_decompose(s)
    lhs = g2.jac_scalar_mul(s, g2.affine_to_jac(P))
    rhs = g2.jac_add(g2.jac_scalar_mul(k1 % o.R_ORDER, g2.affine_to_jac(P)),
                     g2.jac_scalar_mul(k2 % o.R_ORDER, g2.affine_to_jac(g2.glv_phi(P))))
    assert g2.jac_to_affine(lhs) == g2.jac_to_affine(rhs)
