"""(glv, window) sweep: total device time per log2 n.  usage: python tools_sweep.py 10 26"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
    best = None
    for glv in (1, 0):
        ctx.set_option("glv", glv)
        bits = 127 if glv else 254
        neff = 2 * n if glv else n
        cands = [c for c in range(max(6, lg - 8), min(22, lg + 2) + 1) if -(-bits // c) * neff < (1 << 32)]
        if glv:  # keep only windows whose top digit has >= 6 bits (a narrow top window = few huge buckets)
            cands = [c for c in cands if bits - c * (-(-bits // c) - 1) >= 6]
        for c in cands:
            ctx.set_option("window_bits", c)
            ts = []
            try:
                for rep in range(3):
                    ctx.msm_device(d_bases, d_scalars, n, d_out)
                    ts.append(ctx.timings())
            except b200msm.MsmError:
                continue
            t = min(ts[1:], key=lambda x: x["total_ms"])
            row[f"g{glv}c{c}"] = round(t["total_ms"], 3)
            if best is None or t["total_ms"] < best[2]:
                best = (glv, c, t["total_ms"], t)
    row["best"] = {"glv": best[0], "c": best[1], **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in best[3].items()}}
    print(json.dumps(row), flush=True)
