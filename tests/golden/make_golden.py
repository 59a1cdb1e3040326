#!/usr/bin/env python3
"""Generate tests/golden/msm_cases.npz with the Python big-int oracle (oracle/bn254.py).

The reference holds no stored MSM vectors and cannot be built here (Rust + Apple Metal +
un-vendored arkworks), so these fixtures are produced by the oracle's DEFINITION
path (`msm_naive`: double-and-add per term) and cross-checked against its bucket-method path
before being written.  Each case stores raw arkworks memory (bases (n,9) u64, scalars (n,4) u64)
and the expected affine result (2 x 4 u64 canonical words + an `is_inf` flag).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import bn254 as o  # noqa: E402
import helpers as h  # noqa: E402


def cases():
    r = o.R_ORDER
    pts = o.random_points(160, 0xB2000001)
    sc = o.random_scalars(160, 0xB2000002)
    yield "rand_100", pts[:100], sc[:100]
    yield "n1", pts[:1], sc[:1]
    yield "n3", pts[:3], sc[:3]
    # every edge the reference's pipeline mishandles or its tests exercise (§4)
    P0, P1, P2 = pts[100], pts[101], pts[102]
    edge_pts = [P0, None, P1, P1, o.affine_neg(P1), P2, P2, P0, None, o.GEN, o.GEN, P2]
    edge_sc = [0, 5, 1, r - 1, 7, 12345, 12345, 1 << 253, 0, 2, r - 2, (1 << 128) + 1]
    yield "edge_mixed", edge_pts, edge_sc
    yield "all_same_base", [P0] * 64, sc[100:164 - 0][:60] + [1, 1, 2, r - 1]
    yield "cancel", [P0, o.affine_neg(P0), P1, o.affine_neg(P1)], [sc[5], sc[5], sc[6], sc[6]]
    yield "all_zero_scalars", pts[:16], [0] * 16
    yield "all_one_scalars", pts[:48], [1] * 48
    yield "all_equal_scalars", pts[:48], [sc[7]] * 48
    yield "small_scalars", pts[:64], [s & 0xFFFFFFFF for s in sc[:64]]
    yield "generator_multiples", [o.GEN] * 3, [1, 1, 1]  # = 3G, known: see oracle self-check (EIP-196 vectors)


def main():
    o.self_check()
    out = {}
    names = []
    for name, pts, sc in cases():
        assert len(pts) == len(sc), name
        exp = o.jac_to_affine(o.msm_naive(pts, sc))
        for w in (4, 7, 13):
            assert o.jac_to_affine(o.msm_pippenger(pts, sc, w)) == exp, (name, w)
        out[name + "/bases"] = h.pack_bases(pts)
        out[name + "/scalars"] = h.pack_scalars(sc)
        e = np.zeros(9, dtype=np.uint64)
        if exp is None:
            e[8] = 1
        else:
            e[0:4] = h.words(exp[0])
            e[4:8] = h.words(exp[1])
        out[name + "/expected"] = e
        names.append(name)
        print(name, len(pts), "inf" if exp is None else hex(exp[0])[:18])
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "msm_cases.npz"), **out)


if __name__ == "__main__":
    main()
