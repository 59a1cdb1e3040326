#!/usr/bin/env python3
"""Generate tests/golden/msm_sympy.npz with an INDEPENDENT implementation: sympy's elliptic-curve arithmetic.

    python tests/golden/make_golden_sympy.py          (about 8 minutes; sympy needs ~0.3 s per scalar multiplication)

This script shares NO code with oracle/ (it does not import it): points, scalar multiplications and sums come from
`sympy.ntheory.elliptic_curve.EllipticCurve(0, 3, modulus=p)` (affine chord-and-tangent over Python integers), square
roots from `sympy.ntheory.sqrt_mod`, and the Montgomery encoding of the arkworks memory words (a * 2^256 mod m) is
spelled out here.  The reference stores no MSM result vectors (its e2e test is relational: tests/cuzk/e2e.rs:14-63), so
these fixtures are what pins oracle/bn254.py, oracle/cpu_msm.c and the CUDA path at the MSM level against something
none of them produced.  Same record format as msm_cases.npz: bases (n, 9) u64 = arkworks G1Affine {x, y, infinity}
Montgomery words, scalars (n, 4) u64 = Fr Montgomery words, expected = [x(4), y(4), is_inf] canonical words.
"""
import hashlib
import os

import numpy as np
from sympy.ntheory import sqrt_mod
from sympy.ntheory.elliptic_curve import EllipticCurve

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
HERE = os.path.dirname(os.path.abspath(__file__))
E = EllipticCurve(0, 3, modulus=P)
O = E(1, 2) - E(1, 2)
M64 = (1 << 64) - 1


def words(v):
    return [(v >> (64 * j)) & M64 for j in range(4)]


def prng(tag, count, mod):
    out, ctr = [], 0
    while len(out) < count:
        v = int.from_bytes(hashlib.sha256(b"%s/%d" % (tag, ctr)).digest() + hashlib.sha256(b"%s/%d/b" % (tag, ctr)).digest()[:8], "big")
        ctr += 1
        out.append(v % mod)
    return out


def random_points(tag, count):
    """try-and-increment on x; y = the square root sympy returns, sign from the hash stream"""
    pts, ctr = [], 0
    xs = iter(prng(tag, 8 * count + 64, P))
    while len(pts) < count:
        x = next(xs)
        y = sqrt_mod((x * x * x + 3) % P, P)
        if y is None:
            continue
        y = int(y)
        if (x ^ ctr) & 1:
            y = P - y
        ctr += 1
        pts.append((x, y))
    return pts


def neg(pt):
    return (pt[0], (P - pt[1]) % P)


def msm(points, scalars):
    acc = O
    for pt, s in zip(points, scalars):
        if pt is None or s % R == 0:
            continue
        acc = acc + (s % R) * E(pt[0], pt[1])
    if acc == O:
        return None
    return (int(acc.x), int(acc.y))


def cases():
    pts = random_points(b"sympy-golden-points", 1100)
    sc = prng(b"sympy-golden-scalars", 1100, R)
    G = (1, 2)
    yield "sy_n1", pts[:1], sc[:1]
    yield "sy_n2", pts[:2], sc[:2]
    yield "sy_rand_33", pts[2:35], sc[2:35]
    yield "sy_inf_bases", [pts[40], None, pts[41], None, None, pts[42]], sc[40:46]
    yield "sy_pairs_p_minus_p", [pts[50], neg(pts[50]), pts[51], neg(pts[51]), pts[52]], [sc[50], sc[50], 9, 9, sc[52]]
    yield "sy_repeated_base", [pts[60]] * 24, sc[60:80] + [1, 1, R - 1, 2]
    yield "sy_r_minus_1", pts[80:88], [R - 1] * 8
    yield "sy_zero_and_one", pts[90:106], [0, 1] * 8
    yield "sy_all_zero", pts[106:110], [0] * 4
    yield "sy_small_scalars", pts[110:142], [s & 0xFFFFFFFF for s in sc[110:142]]
    yield "sy_top_bits", pts[142:150], [(1 << 253) + k for k in range(8)]
    yield "sy_generator", [G, G, G, neg(G)], [1, 2, R - 3, 5]
    yield "sy_rand_1024", pts[-1024:], sc[-1024:]


def main():
    assert (2 * E(1, 2)).x == 1368015179489954701390400359078579693043519447331113978918064868415326638035  # EIP-196 2G
    out, names = {}, []
    for name, pts, sc in cases():
        n = len(pts)
        b = np.zeros((n, 9), dtype=np.uint64)
        s = np.zeros((n, 4), dtype=np.uint64)
        for i, (pt, k) in enumerate(zip(pts, sc)):
            if pt is None:
                b[i, 8] = 1
            else:
                assert (pt[1] * pt[1] - pt[0] ** 3 - 3) % P == 0
                b[i, 0:4] = words(pt[0] * (1 << 256) % P)
                b[i, 4:8] = words(pt[1] * (1 << 256) % P)
            s[i] = words((k % R) * (1 << 256) % R)
        exp = msm(pts, sc)
        e = np.zeros(9, dtype=np.uint64)
        if exp is None:
            e[8] = 1
        else:
            e[0:4] = words(exp[0])
            e[4:8] = words(exp[1])
        out[name + "/bases"], out[name + "/scalars"], out[name + "/expected"] = b, s, e
        names.append(name)
        print(name, n, "inf" if exp is None else hex(exp[0])[:18], flush=True)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "msm_sympy.npz"), **out)


if __name__ == "__main__":
    main()
