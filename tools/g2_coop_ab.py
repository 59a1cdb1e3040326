"""Interleaved A/B of the G2 bucket reduce engines: cooperative levels (coop_reduce -1/1) vs thread-per-segment (0).
usage: python tools/g2_coop_ab.py [LOGN ...]  -> one JSON line per size (host-buffer G2 MSM wall time, median of 7)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gpu-acceleration_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b200msm  # noqa: E402
import bn254_g2 as g2  # noqa: E402


def main():
    ctx = b200msm.Context()
    sizes = [int(a) for a in sys.argv[1:]] or [10, 14, 16, 18, 20]
    rec = np.array([g2.encode_base(pt) for pt in g2.random_points(4096, 1)], dtype=np.uint64)
    for lg in sizes:
        n = 1 << lg
        bases = np.tile(rec, (-(-n // 4096), 1))[:n].copy()
        scal = np.random.default_rng(lg).integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        ts = {0: [], 1: []}
        res = {}
        for it in range(9):
            for coop in (0, 1):
                ctx.set_option("coop_reduce", coop)
                t0 = time.perf_counter()
                r = ctx.msm_g2(bases, scal)
                dt = (time.perf_counter() - t0) * 1e3
                res[coop] = r
                if it >= 2:
                    ts[coop].append(dt)
        ctx.set_option("coop_reduce", -1)
        med = lambda v: round(sorted(v)[len(v) // 2], 3)
        same = bool(g2.jac_to_affine(g2.decode_jacobian(res[0])) == g2.jac_to_affine(g2.decode_jacobian(res[1])))
        print(json.dumps({"log_n": lg, "g2_thread_per_segment_ms": med(ts[0]), "g2_cooperative_ms": med(ts[1]), "same_point": same}), flush=True)


if __name__ == "__main__":
    main()
