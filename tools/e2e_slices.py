"""Host-buffer (reference-facing) MSM latency against the slice count of the upload/accumulate pipeline.
usage: python tools/e2e_slices.py LOGN[:SLICES,SLICES,...[:RATIO_PCT]] ...   -> one JSON line per (log_n, slices)
Pinned 72-byte arkworks records + 32-byte scalars, wall clock around b200msm_bn254_g1_msm, L2 flushed between calls."""
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
    overlap = int(os.environ.get("SORT_OVERLAP", "-1"))   # -1 auto (on), 0 = every slice's sort on the main stream
    ctx.set_option("sort_overlap", overlap)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for spec in sys.argv[1:]:
        parts = spec.split(":")
        lg = int(parts[0])
        slist = [int(x) for x in parts[1].split(",")] if len(parts) > 1 else [1, 2, 3, 4, 6, 8]
        ratio = int(parts[2]) if len(parts) > 2 else 0
        ctx.set_option("slice_ratio", ratio)
        n = 1 << lg
        d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.testkit_generate(0xE2E + lg, n, d_bases, d_scalars)
        hb = np.zeros((n, 9), dtype=np.uint64)
        hb[:, :8] = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
        h_bases = torch.from_numpy(hb).pin_memory()
        h_scalars = d_scalars.cpu().pin_memory()
        ref = None
        for S in slist:
            ctx.set_option("slices", S)
            ms = []
            for it in range(13):
                flush.fill_(it & 255)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = ctx.msm_raw(h_bases.data_ptr(), 72, 0, 32, 64, h_scalars.data_ptr(), 32, n)
                dt = (time.perf_counter() - t0) * 1e3
                if it >= 3:
                    ms.append(dt)
            if ref is None:
                ref = res
            ms.sort()
            print(json.dumps({"log_n": lg, "slices": S, "ratio_pct": ratio, "sort_overlap": overlap, "ms_median": ms[len(ms) // 2], "ms_min": ms[0], "ms_max": ms[-1],
                              "same_result": bool(res == ref)}), flush=True)
        ctx.set_option("slices", 0)


if __name__ == "__main__":
    main()
