"""Device-resident MSM at large sizes, verified with the discrete-log checksum (C oracle = checker only).
usage: python tools/verify_large.py LOGN [LOGN ...]   -> one JSON line per size (time, window, verified)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gpu-acceleration_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b200msm  # noqa: E402
import bn254 as o  # noqa: E402
import cpu_msm  # noqa: E402


def main():
    ctx = b200msm.Context([0])
    ctx.set_option("timing", 1)
    for lg in [int(a) for a in sys.argv[1:]]:
        n = 1 << lg
        d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        t1, t2 = ctx.testkit_generate(0x1A26E + lg, n, d_bases, d_scalars, want_dlogs=True)
        ctx.msm_device(d_bases, d_scalars, n, d_out)
        t0 = time.perf_counter()
        ctx.msm_device(d_bases, d_scalars, n, d_out)
        wall = (time.perf_counter() - t0) * 1e3
        t = ctx.timings()
        hs = d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4)
        kw = cpu_msm.dlog_checksum(hs, t1, t2)
        want = o.jac_to_affine(o.decode_jacobian(cpu_msm.scalar_mul_gen(kw)))
        got = o.jac_to_affine(o.decode_jacobian(d_out.cpu().numpy().view(np.uint64)))
        print(json.dumps({"log_n": lg, "verified_vs_oracle": bool(got == want), "wall_ms": round(wall, 3),
                          "device_ms": round(t["total_ms"], 3), "window_bits": t["window_bits"], "num_windows": t["num_windows"],
                          "entries": t["entries"], "mem_allocated_gb": round(torch.cuda.memory_allocated() / 2**30, 2)}), flush=True)
        del d_bases, d_scalars, hs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
