"""Host-buffer MSM from PAGEABLE memory (what an arkworks Vec is) against the number of staging copy threads.
usage: python tools/pageable_threads.py LOGN [LOGN ...]   -> one JSON line per (log_n, threads); threads = 0 is pinned input"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpu-acceleration_b200"))
import b200msm  # noqa: E402


def main():
    ctx = b200msm.Context()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for lg in [int(a) for a in sys.argv[1:]]:
        n = 1 << lg
        d_b = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        d_s = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.testkit_generate(0xFA6E + lg, n, d_b, d_s)
        hb = np.zeros((n, 9), dtype=np.uint64)
        hb[:, :8] = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
        hs = d_s.cpu().numpy().view(np.uint64).reshape(n, 4).copy()
        pb = torch.from_numpy(hb).pin_memory()
        ps = torch.from_numpy(hs).pin_memory()
        del d_b, d_s
        ref = None
        for threads in (0, 1, 2, 4, 6, 8, 12, 16):
            if threads:
                ctx.set_option("copy_threads", threads)
            ms = []
            for it in range(9):
                flush.fill_(it)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if threads:
                    r = ctx.msm_raw(hb.ctypes.data, 72, 0, 32, 64, hs.ctypes.data, 32, n)
                else:
                    r = ctx.msm_raw(pb.data_ptr(), 72, 0, 32, 64, ps.data_ptr(), 32, n)
                ms.append((time.perf_counter() - t0) * 1e3)
            ms = sorted(ms[2:])
            ref = ref or r
            print(json.dumps({"log_n": lg, "copy_threads": threads if threads else "pinned input", "ms_median": round(ms[len(ms) // 2], 3),
                              "same_result": bool(r == ref)}), flush=True)
        ctx.set_option("copy_threads", 6)


if __name__ == "__main__":
    main()
