"""CPU oracle for BN254 G2 (the Fq2 twist) -- TEST INFRASTRUCTURE ONLY, same rules as oracle/bn254.py.

Scope: SURVEY.md 8(f) rank 4 ("G2 / other curves ... needed for B2 (G2) MSM in Groth16").  The reference has no G2
code at all (its roadmap names it, README.md:199-200), so there is nothing of the reference to restate or to pin
against: PARITY UNPINNED by reference data.  What this module is pinned against instead (checked by `self_check()`):
  * the public alt_bn128 / EIP-197 G2 generator, which must satisfy the twist equation y^2 = x^3 + 3/(9+u) over
    Fq2 = Fq[u]/(u^2+1) and have order r;
  * arkworks `ark-bn254 0.4.0` conventions (not vendored under /root/reference): G2Affine {x: Fq2, y: Fq2, infinity},
    Fq2 = QuadExtField {c0, c1} with each Fq the Montgomery integer a*R mod p in four LE u64; G2Projective is Jacobian.
Arithmetic is plain Python ints; Fq2 elements are (c0, c1) tuples meaning c0 + c1*u.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import bn254 as o

P = o.P
R_ORDER = o.R_ORDER
Fq2 = Tuple[int, int]
Affine2 = Optional[Tuple[Fq2, Fq2]]
Jac2 = Tuple[Fq2, Fq2, Fq2]

F2_ZERO: Fq2 = (0, 0)
F2_ONE: Fq2 = (1, 0)


def f2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a: Fq2) -> Fq2:
    return ((-a[0]) % P, (-a[1]) % P)


def f2_mul(a: Fq2, b: Fq2) -> Fq2:
    # u^2 = -1
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a: Fq2) -> Fq2:
    return f2_mul(a, a)


def f2_scale(a: Fq2, k: int) -> Fq2:
    return (a[0] * k % P, a[1] * k % P)


def f2_inv(a: Fq2) -> Fq2:
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def f2_is_zero(a: Fq2) -> bool:
    return a[0] == 0 and a[1] == 0


# twist coefficient b' = 3 / (9 + u)
B2: Fq2 = f2_mul((3, 0), f2_inv((9, 1)))

# public alt_bn128 G2 generator (EIP-197 / py_ecc `G2`): coefficients (c0, c1)
GEN2: Affine2 = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)

JAC2_INF: Jac2 = (F2_ONE, F2_ONE, F2_ZERO)

# GLV on G2 (engine-internal): (x, y) -> (GLV_BETA_G2 * x, y) is multiplication by o.GLV_LAMBDA on the order-r subgroup.
# It is G1's beta SQUARED: with beta itself the map acts as lambda^2 (checked in tests/test_oracle_g2.py).
GLV_BETA_G2 = o.GLV_BETA * o.GLV_BETA % P


def glv_phi(pt: "Affine2") -> "Affine2":
    return None if pt is None else (f2_scale(pt[0], GLV_BETA_G2), pt[1])


def is_on_curve(pt: Affine2) -> bool:
    if pt is None:
        return True
    x, y = pt
    return f2_sqr(y) == f2_add(f2_mul(f2_sqr(x), x), B2)


def affine_neg(pt: Affine2) -> Affine2:
    return None if pt is None else (pt[0], f2_neg(pt[1]))


def affine_to_jac(pt: Affine2) -> Jac2:
    return JAC2_INF if pt is None else (pt[0], pt[1], F2_ONE)


def jac_is_inf(a: Jac2) -> bool:
    return f2_is_zero(a[2])


def jac_to_affine(a: Jac2) -> Affine2:
    if jac_is_inf(a):
        return None
    zi = f2_inv(a[2])
    zi2 = f2_sqr(zi)
    return (f2_mul(a[0], zi2), f2_mul(a[1], f2_mul(zi2, zi)))


def jac_dbl(a: Jac2) -> Jac2:
    if jac_is_inf(a):
        return a
    X, Y, Z = a
    A = f2_sqr(X)
    B = f2_sqr(Y)
    C = f2_sqr(B)
    D = f2_scale(f2_sub(f2_sub(f2_sqr(f2_add(X, B)), A), C), 2)
    E = f2_scale(A, 3)
    F = f2_sqr(E)
    X3 = f2_sub(F, f2_scale(D, 2))
    Y3 = f2_sub(f2_mul(E, f2_sub(D, X3)), f2_scale(C, 8))
    Z3 = f2_scale(f2_mul(Y, Z), 2)
    return (X3, Y3, Z3)


def jac_add(a: Jac2, b: Jac2) -> Jac2:
    if jac_is_inf(a):
        return b
    if jac_is_inf(b):
        return a
    X1, Y1, Z1 = a
    X2, Y2, Z2 = b
    Z1Z1 = f2_sqr(Z1)
    Z2Z2 = f2_sqr(Z2)
    U1 = f2_mul(X1, Z2Z2)
    U2 = f2_mul(X2, Z1Z1)
    S1 = f2_mul(Y1, f2_mul(Z2, Z2Z2))
    S2 = f2_mul(Y2, f2_mul(Z1, Z1Z1))
    if U1 == U2:
        return jac_dbl(a) if S1 == S2 else JAC2_INF
    H = f2_sub(U2, U1)
    R = f2_sub(S2, S1)
    HH = f2_sqr(H)
    HHH = f2_mul(H, HH)
    V = f2_mul(U1, HH)
    X3 = f2_sub(f2_sub(f2_sqr(R), HHH), f2_scale(V, 2))
    Y3 = f2_sub(f2_mul(R, f2_sub(V, X3)), f2_mul(S1, HHH))
    Z3 = f2_mul(f2_mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def jac_scalar_mul(k: int, a: Jac2) -> Jac2:
    k %= R_ORDER
    acc = JAC2_INF
    for bit in reversed(range(k.bit_length())):
        acc = jac_dbl(acc)
        if (k >> bit) & 1:
            acc = jac_add(acc, a)
    return acc


def self_check() -> None:
    assert is_on_curve(GEN2), "EIP-197 G2 generator is not on y^2 = x^3 + 3/(9+u)"
    assert jac_is_inf(jac_scalar_mul_raw(R_ORDER, affine_to_jac(GEN2))), "G2 generator does not have order r"
    two = jac_to_affine(jac_dbl(affine_to_jac(GEN2)))
    assert is_on_curve(two) and jac_to_affine(jac_add(affine_to_jac(GEN2), affine_to_jac(GEN2))) == two


def jac_scalar_mul_raw(k: int, a: Jac2) -> Jac2:
    """k*a without reducing k mod r (used to check the group order itself)."""
    acc = JAC2_INF
    for bit in reversed(range(k.bit_length())):
        acc = jac_dbl(acc)
        if (k >> bit) & 1:
            acc = jac_add(acc, a)
    return acc


# ----------------------------------------------------------------------------- MSM
def msm_naive(bases: Sequence[Affine2], scalars: Sequence[int]) -> Jac2:
    acc = JAC2_INF
    for pt, s in zip(bases, scalars):
        if pt is not None and s % R_ORDER:
            acc = jac_add(acc, jac_scalar_mul(s, affine_to_jac(pt)))
    return acc


def msm_pippenger(bases: Sequence[Affine2], scalars: Sequence[int], w: int = 8) -> Jac2:
    """Signed-digit bucket method with the same digit rule as the G1 engine (oracle/bn254.py::signed_digits)."""
    K = o.num_windows_for(w)
    half = 1 << (w - 1)
    digs = [o.signed_digits(s % R_ORDER, w, K) for s in scalars]
    result = JAC2_INF
    for k in reversed(range(K)):
        for _ in range(w):
            result = jac_dbl(result)
        buckets = [JAC2_INF] * (half + 1)
        for pt, d in zip(bases, digs):
            if pt is None or d[k] == 0:
                continue
            q = affine_to_jac(pt if d[k] > 0 else affine_neg(pt))
            buckets[abs(d[k])] = jac_add(buckets[abs(d[k])], q)
        running = JAC2_INF
        total = JAC2_INF
        for m in range(half, 0, -1):
            running = jac_add(running, buckets[m])
            total = jac_add(total, running)
        result = jac_add(result, total)
    return result


def table_expand(bases: Sequence[Affine2], w: int) -> List[List[Affine2]]:
    """Window table of a registered G2 base set: table[k][i] = 2^(w*k) * P_i (affine); infinity stays infinity."""
    K = o.num_windows_for(w)
    rows = [list(bases)]
    for _ in range(1, K):
        nxt = []
        for pt in rows[-1]:
            if pt is None:
                nxt.append(None)
                continue
            j = affine_to_jac(pt)
            for _ in range(w):
                j = jac_dbl(j)
            nxt.append(jac_to_affine(j))
        rows.append(nxt)
    return rows


# ----------------------------------------------------------------------------- inputs & arkworks memory
def random_points(n: int, seed: int) -> List[Affine2]:
    """n G2 points as sums of two small tables of generator multiples (cheap: 2*sqrt(n) scalar multiplications)."""
    import random
    rng = random.Random(seed)
    side = 1
    while side * side < n:
        side += 1
    G = affine_to_jac(GEN2)
    t1 = [jac_scalar_mul(rng.randrange(1, R_ORDER), G) for _ in range(side)]
    t2 = [jac_scalar_mul(rng.randrange(1, R_ORDER), G) for _ in range(side)]
    return [jac_to_affine(jac_add(t1[i % side], t2[i // side])) for i in range(n)]


def words(v: int) -> List[int]:
    return [(v >> (64 * j)) & ((1 << 64) - 1) for j in range(4)]


def encode_base(pt: Affine2) -> List[int]:
    """arkworks G2Affine payload: x.c0, x.c1, y.c0, y.c1 (Montgomery, 4 LE u64 each) + infinity flag word = 17 u64."""
    if pt is None:
        return [0] * 16 + [1]
    out: List[int] = []
    for c in (pt[0][0], pt[0][1], pt[1][0], pt[1][1]):
        out += words(o.to_mont(c))
    return out + [0]


def decode_jacobian(w: Sequence[int]) -> Jac2:
    """24 u64 (X.c0, X.c1, Y.c0, Y.c1, Z.c0, Z.c1; Montgomery) -> canonical Jacobian point."""
    vals = [o.from_mont(sum(int(w[4 * k + j]) << (64 * j) for j in range(4))) for k in range(6)]
    assert all(v < P for v in vals)
    return ((vals[0], vals[1]), (vals[2], vals[3]), (vals[4], vals[5]))
