"""Wall time of the G2 MSM (host buffers in, point out) at a few sizes, with the G1 time beside it.
Bases are generator multiples built on the host with the oracle (checker code used here only to MAKE inputs).
usage: python tools/g2_times.py [LOGN ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gpu-acceleration_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b200msm  # noqa: E402
import bn254 as o  # noqa: E402
import bn254_g2 as g2  # noqa: E402
import helpers as h  # noqa: E402


def main():
    ctx = b200msm.Context()
    sizes = [int(a) for a in sys.argv[1:]] or [10, 14, 16, 18, 20]
    base_pts = g2.random_points(4096, 1)                      # 4096 distinct points, tiled to the size (timing only)
    rec = np.array([g2.encode_base(pt) for pt in base_pts], dtype=np.uint64)
    g1_rec = h.pack_bases(o.random_points(4096, 2))
    for lg in sizes:
        n = 1 << lg
        reps = -(-n // 4096)
        bases = np.tile(rec, (reps, 1))[:n].copy()
        g1b = np.tile(g1_rec, (reps, 1))[:n].copy()
        rng = np.random.default_rng(lg)
        scal = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)   # < r as Montgomery words (top limb < 2^62)
        out = {}
        for name, fn in (("g2", lambda: ctx.msm_g2(bases, scal)), ("g1", lambda: ctx.msm(g1b, scal))):
            ts = []
            for it in range(6):
                t0 = time.perf_counter()
                fn()
                ts.append((time.perf_counter() - t0) * 1e3)
            out[name + "_ms"] = round(sorted(ts[1:])[len(ts[1:]) // 2], 3)
        t = ctx.timings()
        print(json.dumps({"log_n": lg, **out, "ratio": round(out["g2_ms"] / out["g1_ms"], 2)}), flush=True)


if __name__ == "__main__":
    main()
