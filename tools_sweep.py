"""Window-size sweep: total device time per (log2 n, c).  usage: python tools_sweep.py 12 26"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gpu-acceleration_b200"))
import b200msm

lo, hi = int(sys.argv[1]), int(sys.argv[2])
ctx = b200msm.Context([0]); ctx.set_option("timing", 1)
nmax = 1 << hi
d_bases = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
d_scalars = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ctx.testkit_generate(1, nmax, d_bases, d_scalars)
for lg in range(lo, hi + 1):
    n = 1 << lg
    row = {"log_n": lg}
    cands = [c for c in range(max(6, lg - 8), min(22, lg + 1) + 1) if (254 + c - 1) // c * n < (1 << 32)]
    best = None
    for c in cands:
        ctx.set_option("window_bits", c)
        ts = []
        try:
            for rep in range(3):
                ctx.msm_device(d_bases, d_scalars, n, d_out)
                ts.append(ctx.timings())
        except b200msm.MsmError as e:
            row[str(c)] = "err"
            continue
        t = min(ts[1:], key=lambda x: x["total_ms"])
        row[str(c)] = round(t["total_ms"], 3)
        if best is None or t["total_ms"] < best[1]:
            best = (c, t["total_ms"], t)
    row["best_c"] = best[0]
    row["best"] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in best[2].items()}
    print(json.dumps(row), flush=True)
