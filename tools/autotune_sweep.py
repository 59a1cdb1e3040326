"""Auto-tuner evidence: total device time of a resident MSM for every candidate (scalar split, window size) per input size,
on the full chip and with part of the SMs taken away (b200msm_testkit_occupy_sms: the situation of a MIG slice / green
context / shared GPU).  Feeds the cost model of auto_policy() in csrc/b200msm.cu.

usage: python tools/autotune_sweep.py OUT.jsonl [--sms 148,74] [--logs 12,14,...]
One JSON line per (SM count, size, split, window): best of 3 timed runs after one warm-up, stage split included."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-acceleration_b200"))
import b200msm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--sms", default="148,74")
    ap.add_argument("--logs", default="12,14,16,18,20,22,24")
    ap.add_argument("--policy-only", action="store_true", help="time only the configuration the engine picks itself")
    args = ap.parse_args()
    logs = [int(x) for x in args.logs.split(",")]
    ctx = b200msm.Context([0])
    ctx.set_option("timing", 1)
    hw_sms = torch.cuda.get_device_properties(0).multi_processor_count
    nmax = 1 << max(logs)
    d_bases = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
    d_scalars = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.testkit_generate(1, nmax, d_bases, d_scalars)
    out = open(args.out, "w")
    for sms in [int(x) for x in args.sms.split(",")]:
        ctx.set_option("sm_count", 0 if sms >= hw_sms else sms)
        for lg in logs:
            n = 1 << lg
            if args.policy_only:
                cands = [(-1, 0)]
            else:
                cands = [(glv, c) for glv in (0, 1) for c in range(max(6, lg - 7), min(22, lg + 1) + 1)
                         if not (glv and c in (9, 14, 18, 21))]
            if sms < hw_sms:
                ctx.occupy_sms(hw_sms - sms, 100.0)
            try:
                for glv, c in cands:
                    ctx.set_option("glv", glv)
                    ctx.set_option("window_bits", c)
                    best = None
                    for rep in range(4):
                        ctx.msm_device(d_bases, d_scalars, n, d_out)
                        t = ctx.timings()
                        if rep and (best is None or t["total_ms"] < best["total_ms"]):
                            best = t
                    row = {"sms": sms, "log_n": lg, "glv": glv, "c": best["window_bits"], "W": best["num_windows"],
                           **{k: round(best[k], 4) for k in ("decompose_ms", "sort_ms", "accumulate_ms", "reduce_ms", "total_ms")},
                           "entries": best["entries"]}
                    out.write(json.dumps(row) + "\n")
                    out.flush()
            finally:
                if sms < hw_sms:
                    ctx.release_sms()
    ctx.close()


if __name__ == "__main__":
    main()
