"""Small MSMs through the round-2 code paths (partitioned sort, per-chunk fix-up, table mode, skew) for compute-sanitizer.
usage: compute-sanitizer --tool memcheck python tools/sanitize_paths.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-acceleration_b200"))
import b200msm  # noqa: E402

ctx = b200msm.Context([0])
n = (1 << 14) + 333
d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ctx.testkit_generate(5, n, d_bases, d_scalars)


def run(**opts):
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.msm_device(d_bases, d_scalars, n, d_out)
    torch.cuda.synchronize()
    for k in opts:
        ctx.set_option(k, -1 if k not in ("window_bits", "chunk") else 0)
    return b200msm.G1Projective(d_out.cpu().numpy().view(np.uint64).copy())


for skew in (False, True):
    if skew:
        sc = d_scalars.view(torch.int64).reshape(n, 4)
        sc[: n // 2] = sc[0:1].clone()
        torch.cuda.synchronize()
    ref = run(ranked_sort=1, fix_chunks=0)
    for opts in ({"ranked_sort": 2, "fix_chunks": 1}, {"ranked_sort": 2, "fix_chunks": 1, "glv": 0, "window_bits": 13},
                 {"ranked_sort": 2, "fix_chunks": 1, "glv": 0, "window_bits": 20}, {"ranked_sort": 2, "glv": 1, "window_bits": 8, "chunk": 16},
                 {"ranked_sort": 0, "fix_chunks": 1}):
        assert run(**opts) == ref, (skew, opts)
        print("ok", skew, opts, flush=True)
# host path (slices) and registered bases with the window table through the same engines
hb = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
hs = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4)
ctx.set_option("ranked_sort", 2)
ctx.set_option("fix_chunks", 1)
a = ctx.msm(hb, hs)
ctx.set_option("precompute", 13)
h = ctx.register_bases(hb)
ctx.set_option("precompute", 0)
b = ctx.msm_registered(h, hs)
h.release()
assert a == ref and b == ref
print("ok host + table", flush=True)
ctx.set_option("ranked_sort", -1)
ctx.set_option("fix_chunks", -1)
# heavy partitions (k_place_heavy) and giant buckets (k_fixup_giant): every scalar equal / witness-like zeros and ones at 2^17
n2 = (1 << 17) + 77
d_b2 = torch.empty(n2 * 64, dtype=torch.uint8, device="cuda")
d_s2 = torch.empty(n2 * 32, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ctx.testkit_generate(9, n2, d_b2, d_s2)
sc2 = d_s2.view(torch.int64).reshape(n2, 4)
for kind in ("all_equal", "zeros_ones"):
    if kind == "all_equal":
        sc2[:] = sc2[0:1].clone()
    else:
        u = torch.rand(n2, device="cuda")
        sc2[u < 0.5] = 0
    torch.cuda.synchronize()
    res = {}
    for name, opts in (("ranked", {"ranked_sort": 1, "fix_chunks": 0}), ("partitioned", {"ranked_sort": 2, "fix_chunks": 1})):
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.msm_device(d_b2, d_s2, n2, d_out)
        torch.cuda.synchronize()
        res[name] = b200msm.G1Projective(d_out.cpu().numpy().view(np.uint64).copy())
        for k in opts:
            ctx.set_option(k, -1)
    assert res["ranked"] == res["partitioned"], kind
    print("ok skew", kind, flush=True)
ctx.close()
