"""Stage timings of the device-resident MSM for a list of (log2 n, window_bits) — tuning aid.
usage: python tools/stage_times.py LOGN:C[:CHUNK[:REDUCE_LOG2[:GROUPS[:GLV[:RANKED_SORT]]]]] ...   (window 0 = auto)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-acceleration_b200"))
import b200msm  # noqa: E402


def main():
    ctx = b200msm.Context([0])
    ctx.set_option("timing", 1)
    specs = [a for a in sys.argv[1:] if ":" in a]
    maxlog = max(int(s.split(":")[0]) for s in specs)
    nmax = 1 << maxlog
    d_bases = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.testkit_generate(1, nmax, d_bases, d_scalars)
    for s in specs:
        parts = s.split(":")
        lg, w = int(parts[0]), int(parts[1])
        chunk = int(parts[2]) if len(parts) > 2 else 0
        rlog = int(parts[3]) if len(parts) > 3 else -1
        ctx.set_option("reduce_log2", rlog)
        groups = int(parts[4]) if len(parts) > 4 else 0
        ctx.set_option("groups", groups)
        glv = int(parts[5]) if len(parts) > 5 else -1
        ctx.set_option("glv", glv)
        ranked = int(parts[6]) if len(parts) > 6 else -1
        ctx.set_option("ranked_sort", ranked)
        n = 1 << lg
        ctx.set_option("window_bits", w)
        ctx.set_option("chunk", chunk)
        best = None
        for rep in range(4):
            ctx.msm_device(d_bases, d_scalars, n, d_out)
            t = ctx.timings()
            if rep and (best is None or t["total_ms"] < best["total_ms"]):
                best = t
        best["log_n"] = lg
        best["chunk"] = chunk
        best["reduce_log2"] = rlog
        best["groups"] = groups
        best["glv"] = glv
        best["ranked_sort"] = ranked
        best = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in best.items()}
        print(json.dumps(best), flush=True)


if __name__ == "__main__":
    main()
