"""Host-buffer MSM latency against the slice plan with N ranks uploading at once (one process per GPU under torchrun): the
regime where the host's aggregate H2D bandwidth, not the arithmetic, bounds the call.
usage: python -m torch.distributed.run --nproc-per-node N tools/e2e_slices_multi.py LOGN SLICES:RATIO[,SLICES:RATIO...]
(0:0 = the library's automatic plan).  Rank 0 prints one JSON line per plan: max over ranks of each rank's median."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-acceleration_b200"))
sys.path.insert(0, ROOT)
import b200msm  # noqa: E402
from bench import bind_to_gpu_numa  # noqa: E402


def main():
    lg = int(sys.argv[1])
    plans = [tuple(int(x) for x in p.split(":")) for p in sys.argv[2].split(",")]
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b200msm.Context([local])
    n = 1 << lg
    d_b = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    d_s = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.testkit_generate(0x5EED + rank, n, d_b, d_s)
    hb = np.zeros((n, 9), dtype=np.uint64)
    hb[:, :8] = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
    h_bases = torch.from_numpy(hb).pin_memory()
    h_scalars = d_s.cpu().pin_memory()
    del d_b, d_s
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t_red = torch.zeros(1, dtype=torch.float64, device="cuda")
    ref = None
    for S, ratio in plans:
        ctx.set_option("slices", S)
        ctx.set_option("slice_ratio", ratio)
        ts = []
        for it in range(10):
            flush.fill_(it)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            r = ctx.msm_raw(h_bases.data_ptr(), 72, 0, 32, 64, h_scalars.data_ptr(), 32, n)
            dt = (time.perf_counter() - t0) * 1e3
            if it >= 3:
                ts.append(dt)
        ref = ref or r
        assert r == ref
        t_red[0] = sorted(ts)[len(ts) // 2]
        if world > 1:
            dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"log_n_per_gpu": lg, "n_gpus": world, "slices": S, "ratio_pct": ratio, "ms_max_over_ranks": round(float(t_red[0]), 3)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
