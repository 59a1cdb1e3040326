"""Conversions between the oracle's Python ints and the arkworks-memory numpy arrays the C ABI takes."""
from __future__ import annotations

import numpy as np

import bn254 as o

M64 = (1 << 64) - 1


def words(v: int):
    return [(v >> (64 * j)) & M64 for j in range(4)]


def unwords(w) -> int:
    return sum(int(w[j]) << (64 * j) for j in range(4))


def pack_bases(points, with_inf: bool = True) -> np.ndarray:
    """(n, 9) uint64 = arkworks G1Affine {x, y, infinity} records (72 B); (n, 8) if not with_inf."""
    n = len(points)
    a = np.zeros((n, 9 if with_inf else 8), dtype=np.uint64)
    for i, pt in enumerate(points):
        if pt is None:
            assert with_inf
            a[i, 8] = 1
            continue
        a[i, 0:4] = words(o.to_mont(pt[0]))
        a[i, 4:8] = words(o.to_mont(pt[1]))
    return a


def pack_scalars(scalars) -> np.ndarray:
    a = np.zeros((len(scalars), 4), dtype=np.uint64)
    for i, s in enumerate(scalars):
        a[i] = words(o.to_mont(s % o.R_ORDER, o.R_ORDER))
    return a


def pack_fq(vals) -> np.ndarray:
    """canonical ints -> (n, 4) Montgomery words"""
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        a[i] = words(o.to_mont(v))
    return a


def unpack_fq(arr):
    return [o.from_mont(unwords(r)) for r in arr]


def pack_xyzz(pts) -> np.ndarray:
    a = np.zeros((len(pts), 16), dtype=np.uint64)
    for i, p in enumerate(pts):
        for c in range(4):
            a[i, 4 * c:4 * c + 4] = words(o.to_mont(p[c]))
    return a


def unpack_xyzz(arr):
    out = []
    for r in arr:
        vals = [unwords(r[4 * c:4 * c + 4]) for c in range(4)]
        assert all(v < o.P for v in vals)
        out.append(tuple(o.from_mont(v) for v in vals))
    return out


def result_affine(res) -> "o.Affine":
    """b200msm.G1Projective -> canonical affine via the ORACLE's arithmetic (independent of the binding's)."""
    return o.jac_to_affine(o.decode_jacobian(res.words))
