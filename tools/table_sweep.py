"""Registered-bases MSM time against the window size of the precomputed table (and against the plain registered path).
usage: python tools/table_sweep.py LOGN:C,C,... ...   (C = 0 means no table)  -> one JSON line per (log_n, c)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "gpu-acceleration_b200"))
import b200msm  # noqa: E402


def main():
    ctx = b200msm.Context()
    ctx.set_option("timing", 1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for spec in sys.argv[1:]:
        lg, cs = spec.split(":")
        lg = int(lg)
        n = 1 << lg
        d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.testkit_generate(0x7AB + lg, n, d_bases, d_scalars)
        hb = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
        hs = d_scalars.cpu().pin_memory().numpy().view(np.uint64).reshape(n, 4)
        del d_bases
        ref = None
        for c in [int(x) for x in cs.split(",")]:
            ctx.set_option("precompute", c)
            t0 = time.perf_counter()
            try:
                handle = ctx.register_bases(hb)
            except b200msm.MsmError as e:
                print(json.dumps({"log_n": lg, "table_c": c, "error": str(e)}), flush=True)
                continue
            finally:
                ctx.set_option("precompute", 0)
            reg_ms = (time.perf_counter() - t0) * 1e3
            dev, wall = [], []
            for it in range(8):
                flush.fill_(it)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = ctx.msm_registered(handle, hs)
                wall.append((time.perf_counter() - t0) * 1e3)
                dev.append(ctx.timings())
            handle.release()
            if ref is None:
                ref = res
            t = dev[-1]
            med = lambda k: sorted(x[k] for x in dev[2:])[len(dev[2:]) // 2]
            print(json.dumps({"log_n": lg, "table_c": c, "windows": t["num_windows"], "register_ms": round(reg_ms, 2),
                              "device_ms": round(med("total_ms"), 4), "sort_ms": round(med("decompose_ms") + med("sort_ms"), 4),
                              "accumulate_ms": round(med("accumulate_ms"), 4), "reduce_ms": round(med("reduce_ms"), 4),
                              "wall_ms": round(sorted(wall[2:])[len(wall[2:]) // 2], 4), "same_result": bool(res == ref)}), flush=True)


if __name__ == "__main__":
    main()
