"""Stage times for witness-like scalar distributions (many zeros / ones / small values) at 2^LOGN, device-resident.
usage: python tools/skew_times.py [LOGN=20]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gpu-acceleration_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b200msm  # noqa: E402
import bn254 as o  # noqa: E402


def words(v):
    return [(v >> (64 * j)) & ((1 << 64) - 1) for j in range(4)]


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << lg
    ctx = b200msm.Context([0])
    ctx.set_option("timing", 1)
    d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.testkit_generate(0x5CE3, n, d_bases, d_scalars)
    base = d_scalars.clone()
    one = torch.tensor(words(o.R_MOD_R), dtype=torch.uint64).view(torch.int64).cuda()
    small = lambda k: torch.tensor(words(k * o.R_MOD_R % o.R_ORDER), dtype=torch.uint64).view(torch.int64).cuda()
    cases = {
        "uniform": lambda sc, u: None,
        "45% zero, 45% one": lambda sc, u: (sc.__setitem__(u < 0.45, 0), sc.__setitem__((u >= 0.45) & (u < 0.9), one)),
        "90% zero": lambda sc, u: sc.__setitem__(u < 0.9, 0),
        "all one": lambda sc, u: sc.__setitem__(u >= 0, one),
        "all equal (random value)": lambda sc, u: sc.__setitem__(u >= 0, sc[0].clone()),
        "50% of 16 small values": lambda sc, u: [sc.__setitem__((u >= k / 32) & (u < (k + 1) / 32), small(k + 2)) for k in range(16)],
    }
    for name, fn in cases.items():
        d_scalars.copy_(base)
        sc = d_scalars.view(torch.int64).reshape(n, 4)
        u = torch.rand(n, device="cuda")
        fn(sc, u)
        torch.cuda.synchronize()
        best = None
        for rep in range(4):
            ctx.msm_device(d_bases, d_scalars, n, d_out)
            t = ctx.timings()
            if rep and (best is None or t["total_ms"] < best["total_ms"]):
                best = t
        print(json.dumps({"log_n": lg, "scalars": name, **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items()}}),
              flush=True)


if __name__ == "__main__":
    main()
