"""Interleaved A/B on one box: host-buffer MSM and registered-handle MSM with / without the sort-ahead stream and with
different slice counts.  usage: python tools/ab_sort_overlap.py [LOGN=20] [ROUNDS=15]   -> one JSON line per variant
(median / min over the rounds; the variants alternate inside every round, L2 flushed before each call)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpu-acceleration_b200"))
import b200msm  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    n = 1 << lg
    ctx = b200msm.Context()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    d_s = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.testkit_generate(0xAB + lg, n, d_b, d_s)
    hb8 = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
    hb = np.zeros((n, 9), dtype=np.uint64)
    hb[:, :8] = hb8
    h_bases = torch.from_numpy(hb).pin_memory()
    h_scalars = d_s.cpu().pin_memory()
    hs = h_scalars.numpy().view(np.uint64).reshape(n, 4)
    plain = ctx.register_bases(hb8)
    ctx.set_option("precompute", 1)
    table = ctx.register_bases(hb8)
    ctx.set_option("precompute", 0)
    del d_b, d_s

    def host():
        return ctx.msm_raw(h_bases.data_ptr(), 72, 0, 32, 64, h_scalars.data_ptr(), 32, n)

    variants = []
    for ov in (0, 1):
        variants.append((f"host_auto_slices_overlap{ov}", host, {"sort_overlap": ov, "slices": 0}))
    for name, hnd in (("registered", plain), ("registered_table", table)):
        for S, ov in ((1, 0), (2, 0), (2, 1), (3, 1)):
            variants.append((f"{name}_slices{S}_overlap{ov}", (lambda hnd=hnd: ctx.msm_registered(hnd, hs)), {"sort_overlap": ov, "slices": S}))
    times = {v[0]: [] for v in variants}
    ref = {}
    for it in range(rounds + 2):
        for name, fn, opts in variants:
            for k, v in opts.items():
                ctx.set_option(k, v)
            flush.fill_(it & 255)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            dt = (time.perf_counter() - t0) * 1e3
            key = name.split("_slices")[0].split("_auto")[0]
            ref.setdefault(key, r)
            assert r == ref[key], name
            if it >= 2:
                times[name].append(dt)
    for name, _, _ in variants:
        t = sorted(times[name])
        print(json.dumps({"log_n": lg, "variant": name, "ms_median": round(t[len(t) // 2], 3), "ms_min": round(t[0], 3)}), flush=True)


if __name__ == "__main__":
    main()
