"""Device-resident MSM stage times under arbitrary option sets -- tuning aid.
usage: python tools/time_opts.py LOGN[,LOGN...] "key=val,key=val" ["key=val,..." ...]
Prints one JSON line per (size, option set): best of 3 timed runs after one warm-up, result checked against the first
option set's result for that size."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-acceleration_b200"))
import b200msm  # noqa: E402

DEFAULTS = {"window_bits": 0, "glv": -1, "chunk": 0, "batch_affine": -1, "ba_chunk": 0, "ba_min_pairs": 0, "coop_reduce": -1,
            "reduce_log2": -1, "ranked_sort": -1, "groups": 0, "fix_chunks": -1, "rowcol_reduce": -1}


def main():
    logs = [int(x) for x in sys.argv[1].split(",")]
    sets = [dict(kv.split("=") for kv in s.split(",") if kv) for s in sys.argv[2:]] or [{}]
    ctx = b200msm.Context([0])
    ctx.set_option("timing", 1)
    nmax = 1 << max(logs)
    d_bases = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.testkit_generate(1, nmax, d_bases, d_scalars)
    for lg in logs:
        n = 1 << lg
        ref = None
        for opts in sets:
            for k, v in DEFAULTS.items():
                ctx.set_option(k, int(opts.get(k, v)))
            best = None
            for rep in range(4):
                ctx.msm_device(d_bases, d_scalars, n, d_out)
                t = ctx.timings()
                if rep and (best is None or t["total_ms"] < best["total_ms"]):
                    best = t
            res = d_out.cpu().numpy().view(np.uint64).copy()
            pr = b200msm.G1Projective(res)
            if ref is None:
                ref = pr
            best = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in best.items()}
            best.update({"log_n": lg, "opts": opts, "same_as_first": bool(pr == ref)})
            print(json.dumps(best), flush=True)


if __name__ == "__main__":
    main()
