"""Wall time of a lone MSM over a registered handle (plain and with the window table) against the slice count of its
scalar upload, at 2^20 / 2^22 / 2^24.  usage: python tools/registered_slices.py   -> one JSON line per case"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpu-acceleration_b200"))
import b200msm
ctx = b200msm.Context()
OVERLAP = int(os.environ.get("SORT_OVERLAP", "-1"))
ctx.set_option("sort_overlap", OVERLAP)
PRE = tuple(int(x) for x in os.environ.get("PRE", "0").split(","))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
SPECS = [(1, 0), (2, 400), (3, 300), (3, 200), (4, 200), (4, 250), (5, 200)]
for lg in (int(a) for a in (sys.argv[1:] or ['20', '22', '24'])):
    n = 1 << lg
    d_b = torch.empty(n * 64, dtype=torch.uint8, device="cuda"); d_s = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.testkit_generate(0x99 + lg, n, d_b, d_s)
    hb = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
    hs = d_s.cpu().pin_memory().numpy().view(np.uint64).reshape(n, 4)
    del d_b, d_s
    for pre in PRE:
        ctx.set_option("precompute", pre); h = ctx.register_bases(hb); ctx.set_option("precompute", 0)
        ref = None
        for S, ratio in SPECS:
            ctx.set_option("slices", S)
            ctx.set_option("slice_ratio", ratio)
            ts = []
            for it in range(9):
                flush.fill_(it); torch.cuda.synchronize()
                t0 = time.perf_counter(); r = ctx.msm_registered(h, hs); ts.append((time.perf_counter() - t0) * 1e3)
            ts = sorted(ts[2:])
            ref = ref or r
            print(json.dumps({"log_n": lg, "table": bool(pre), "slices": S, "ratio_pct": ratio, "sort_overlap": OVERLAP, "ms_median": round(ts[len(ts)//2], 3), "same": bool(r == ref)}), flush=True)
        ctx.set_option("slices", 0)
        ctx.set_option("slice_ratio", 0)
        h.release()
