"""Two device-resident MSMs per size (warm-up + the one to read) for an `ncu --set full` capture.
Only the second MSM of every size lies between cudaProfilerStart/Stop.
usage: ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof python tools/ncu_two_sizes.py 20 24"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-acceleration_b200"))
import b200msm  # noqa: E402

ctx = b200msm.Context([0])
for lg in (int(a) for a in sys.argv[1:]):
    n = 1 << lg
    d_bases = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.testkit_generate(7, n, d_bases, d_scalars)
    ctx.msm_device(d_bases, d_scalars, n, d_out)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ctx.msm_device(d_bases, d_scalars, n, d_out)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    del d_bases, d_scalars
ctx.close()
