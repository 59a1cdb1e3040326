"""Small G2 MSMs through the cooperative bucket reduce (k_g2_reduce_level, one and two levels), the Fq-granular Horner
engine (k_g2_combine) and the sliced host call with the sort-ahead stream, for compute-sanitizer.
usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_g2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gpu-acceleration_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b200msm  # noqa: E402
import bn254 as o  # noqa: E402
import bn254_g2 as g2  # noqa: E402
import helpers as h  # noqa: E402

ctx = b200msm.Context([0])
n = 700
pts = g2.random_points(n, 77)
sc = o.random_scalars(n, 78)
pts[3] = None
sc[5] = 0
want = g2.jac_to_affine(g2.msm_pippenger(pts, sc, 8))
bases = np.array([g2.encode_base(pt) for pt in pts], dtype=np.uint64)
scal = h.pack_scalars(sc)
for w, glv, coop, slices in ((0, -1, 1, 0), (5, 0, 1, 0), (11, 1, 1, 2), (13, 0, 1, 3), (8, -1, 0, 0)):
    ctx.set_option("window_bits", w)
    ctx.set_option("glv", glv)
    ctx.set_option("coop_reduce", coop)
    ctx.set_option("slices", slices)
    got = g2.jac_to_affine(g2.decode_jacobian(ctx.msm_g2(bases, scal)))
    assert got == want, (w, glv, coop, slices)
    print("ok g2", w, glv, coop, slices, flush=True)
for k, v in (("window_bits", 0), ("glv", -1), ("coop_reduce", -1), ("slices", 0)):
    ctx.set_option(k, v)
# G1 host call in three slices with the later sorts on the sort stream, and on the main stream
g1p = o.random_points(3000, 5)
g1s = o.random_scalars(3000, 6)
w1 = o.jac_to_affine(o.msm_pippenger(g1p, g1s, 8)) if hasattr(o, "msm_pippenger") else None
hb, hs = h.pack_bases(g1p), h.pack_scalars(g1s)
res = []
for ov in (1, 0):
    ctx.set_option("sort_overlap", ov)
    ctx.set_option("slices", 3)
    res.append(h.result_affine(ctx.msm(hb, hs)))
    print("ok g1 sliced, sort_overlap", ov, flush=True)
ctx.set_option("sort_overlap", -1)
ctx.set_option("slices", 0)
assert res[0] == res[1] and (w1 is None or res[0] == w1)
print("done", flush=True)
