"""Fixtures that neither the oracle nor the CUDA path produced:

* tests/golden/msm_sympy.npz -- MSM cases computed by sympy's elliptic-curve arithmetic (make_golden_sympy.py shares
  no code with oracle/): the MSM-level pin of oracle/bn254.py, oracle/cpu_msm.c and -- in the `gpu` tests -- of the
  CUDA path through the C ABI.
* tests/golden/ref_zkey_points.npz -- the G1 and G2 points of the Groth16 proving key in the reference tree
  (example-app/test-vectors/circom/multiplier2_final.zkey): pins the Fq2 / G2Affine memory layout.

The reference's own MSM test is relational (metal == arkworks on random inputs, tests/cuzk/e2e.rs:14-63) and stores
no vectors; these files are the stored vectors this repo tests against instead.  Bar: equal group element."""
import os
import random

import numpy as np
import pytest

import bn254 as o
import bn254_g2 as g2
import cpu_msm
import helpers as h

HERE = os.path.dirname(os.path.abspath(__file__))
SYMPY = os.path.join(HERE, "golden", "msm_sympy.npz")
ZKEY = os.path.join(HERE, "golden", "ref_zkey_points.npz")


def _cases():
    z = np.load(SYMPY)
    for name in z["names"]:
        exp = z[f"{name}/expected"]
        want = None if int(exp[8]) else (h.unwords(exp[0:4]), h.unwords(exp[4:8]))
        yield str(name), z[f"{name}/bases"], z[f"{name}/scalars"], want


def _decode_inputs(bases, scalars):
    pts = [None if int(b[8]) else (o.from_mont(h.unwords(b[0:4])), o.from_mont(h.unwords(b[4:8]))) for b in bases]
    sc = [o.from_mont(h.unwords(s), o.R_ORDER) for s in scalars]
    return pts, sc


def _aff(words):
    return o.jac_to_affine(o.decode_jacobian(words))


def test_sympy_fixture_is_complete():
    names = [c[0] for c in _cases()]
    assert len(names) == 13 and "sy_rand_1024" in names and "sy_inf_bases" in names and "sy_pairs_p_minus_p" in names


def test_python_oracle_matches_sympy():
    for name, bases, scalars, want in _cases():
        pts, sc = _decode_inputs(bases, scalars)
        assert all(pt is None or o.is_on_curve(pt) for pt in pts), name
        w = 10 if len(pts) > 256 else 5
        assert o.jac_to_affine(o.msm_pippenger(pts, sc, w)) == want, name
        if len(pts) <= 40:
            assert o.jac_to_affine(o.msm_naive(pts, sc)) == want, name


def test_c_oracle_matches_sympy():
    for name, bases, scalars, want in _cases():
        for threads, w in ((1, 0), (8, 0), (3, 6)):
            out, _ = cpu_msm.msm(bases, scalars, threads, w)
            assert _aff(out) == want, (name, threads, w)


def _zkey():
    z = np.load(ZKEY)
    return z["g1_names"], z["g1"], z["g2_names"], z["g2"]


def _g2_point(rec):
    if not rec.any():
        return None
    f = [o.from_mont(h.unwords(rec[4 * k:4 * k + 4])) for k in range(4)]
    return ((f[0], f[1]), (f[2], f[3]))


def test_zkey_points_decode_under_the_oracle_conventions():
    """Reference-side bytes: every G1 record is on y^2 = x^3 + 3, every G2 record on the twist y^2 = x^3 + 3/(9+u) and in
    the order-r subgroup, with Fq2 = c0 | c1 in Montgomery words.  Any other reading of the layout (c1 first, canonical
    instead of Montgomery words, y before x) fails these checks."""
    n1, p1, n2, p2 = _zkey()
    assert len(p1) == 19 and len(p2) == 7
    for nm, rec in zip(n1, p1):
        if rec.any():
            assert o.is_on_curve((o.from_mont(h.unwords(rec[0:4])), o.from_mont(h.unwords(rec[4:8])))), nm
    finite = 0
    for nm, rec in zip(n2, p2):
        pt = _g2_point(rec)
        if pt is None:
            continue
        finite += 1
        assert g2.is_on_curve(pt), nm
        assert g2.jac_is_inf(g2.jac_scalar_mul_raw(o.R_ORDER, g2.affine_to_jac(pt))), nm
        swapped = ((pt[0][1], pt[0][0]), (pt[1][1], pt[1][0]))
        assert not g2.is_on_curve(swapped), nm
    assert finite == 4
    # B1 and B2 hold the same query polynomial in G1 and G2: the same wires are absent (infinity) in both
    b1 = [rec.any() for nm, rec in zip(n1, p1) if str(nm).startswith("b1")]
    b2 = [rec.any() for nm, rec in zip(n2, p2) if str(nm).startswith("b2_")]
    assert b1 == b2


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("glv", [-1, 0, 1])
def test_cuda_msm_matches_sympy(ctx, glv):
    """The CUDA path through the C ABI against the independent fixtures (auto policy, plain windows, GLV split)."""
    from b200msm import cuda_variable_base_msm
    ctx.set_option("glv", glv)
    try:
        for name, bases, scalars, want in _cases():
            for wb in ((0,) if glv < 0 else (0, 7, 13)):
                ctx.set_option("window_bits", wb)
                res = cuda_variable_base_msm(bases, scalars, ctx)
                assert _aff(res.words) == want, (name, glv, wb)
    finally:
        ctx.set_option("glv", -1)
        ctx.set_option("window_bits", 0)


@pytest.mark.gpu
def test_cuda_registered_and_table_match_sympy(ctx):
    for name, bases, scalars, want in _cases():
        if len(bases) < 8:
            continue
        for pre in (0, 1, 8):
            key = ctx.register_bases(bases, precompute=pre)
            try:
                assert _aff(ctx.msm_registered(key, scalars).words) == want, (name, pre)
            finally:
                key.release()


@pytest.mark.gpu
def test_cuda_g1_msm_over_zkey_points(ctx):
    from b200msm import cuda_variable_base_msm
    _, p1, _, _ = _zkey()
    rng = random.Random(77)
    bases = np.zeros((len(p1), 9), dtype=np.uint64)
    bases[:, :8] = p1
    bases[:, 8] = [0 if rec.any() else 1 for rec in p1]       # arkworks sets `infinity` for the all-zero records
    sc = [rng.randrange(o.R_ORDER) for _ in p1]
    pts = [None if int(b[8]) else (o.from_mont(h.unwords(b[0:4])), o.from_mont(h.unwords(b[4:8]))) for b in bases]
    want = o.jac_to_affine(o.msm_naive(pts, sc))
    assert _aff(cuda_variable_base_msm(bases, h.pack_scalars(sc), ctx).words) == want
    # the device format's own infinity marker: the (0,0) record without a flag word
    assert _aff(cuda_variable_base_msm(np.ascontiguousarray(bases[:, :8]), h.pack_scalars(sc), ctx).words) == want


@pytest.mark.gpu
@pytest.mark.parametrize("glv", [-1, 0])
def test_cuda_g2_msm_over_zkey_points(ctx, glv):
    """G2 MSM over the proving key's own G2 points (beta2, gamma2, delta2, B2 query), repeated to 56 terms."""
    _, _, _, p2 = _zkey()
    rng = random.Random(78)
    reps = 8
    bases = np.zeros((len(p2) * reps, 17), dtype=np.uint64)
    bases[:, :16] = np.tile(p2, (reps, 1))
    bases[:, 16] = [0 if rec.any() else 1 for rec in bases[:, :16]]
    sc = [rng.randrange(o.R_ORDER) for _ in range(len(bases))]
    sc[3] = o.R_ORDER - 1
    sc[4] = 1
    pts = [_g2_point(rec[:16]) for rec in bases]
    want = g2.jac_to_affine(g2.msm_naive(pts, sc))
    ctx.set_option("glv", glv)
    try:
        out = ctx.msm_g2(bases, h.pack_scalars(sc))
    finally:
        ctx.set_option("glv", -1)
    assert g2.jac_to_affine(g2.decode_jacobian(out)) == want
