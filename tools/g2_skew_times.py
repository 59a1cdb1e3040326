import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gpu-acceleration_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b200msm, bn254 as o, bn254_g2 as g2, helpers as h
ctx = b200msm.Context()
base_pts = g2.random_points(4096, 1)
rec = np.array([g2.encode_base(pt) for pt in base_pts], dtype=np.uint64)
g1_rec = h.pack_bases(o.random_points(4096, 2))
one = np.array(h.words(o.R_MOD_R), dtype=np.uint64)
for lg in (20, 22):
    n = 1 << lg
    reps = -(-n // 4096)
    bases = np.tile(rec, (reps, 1))[:n].copy(); g1b = np.tile(g1_rec, (reps, 1))[:n].copy()
    rng = np.random.default_rng(lg)
    for kind in ("uniform", "witness_45_45", "all_one"):
        scal = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        u = rng.random(n)
        if kind == "witness_45_45":
            scal[u < 0.45] = 0; scal[(u >= 0.45) & (u < 0.9)] = one
        elif kind == "all_one":
            scal[:] = one
        out = {}
        for name, fn in (("g2", lambda: ctx.msm_g2(bases, scal)), ("g1", lambda: ctx.msm(g1b, scal))):
            ts = []
            for it in range(5):
                t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
            out[name + "_ms"] = round(sorted(ts[1:])[len(ts[1:]) // 2], 3)
        print(json.dumps({"log_n": lg, "scalars": kind, **out}), flush=True)
