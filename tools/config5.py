"""BASELINE config #5: four concurrent 2^22 MSMs (A, B1, C, H of a Groth16 prover) over four REGISTERED
(device-resident) base sets on one 8xB200 box, through the single-process multi-device C ABI
(b200msm_register_bases_on + b200msm_msm_batch).  Each MSM is sharded over a pair of GPUs by point range.
Reports the batch makespan (host scalars in, four points out) and each MSM's stand-alone latency; results
are verified with the discrete-log checksum (oracle = checker only).

usage: python tools/config5.py [log_n=22] [reps=5] [precompute=0] [layout=pairs|all]
  precompute=1 registers the base sets with the precomputed window table; layout=all shards every MSM over all GPUs
  (four pipelined items per device) instead of giving each MSM its own group of GPUs."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gpu-acceleration_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b200msm, bn254 as o, cpu_msm

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
precompute = int(sys.argv[3]) if len(sys.argv) > 3 else 0
layout = sys.argv[4] if len(sys.argv) > 4 else "pairs"
n = 1 << log_n
ngpu = torch.cuda.device_count()
ctx = b200msm.Context(list(range(ngpu)))
per = ngpu if layout == "all" else max(1, ngpu // 4)
ctx.set_option("precompute", precompute)
bases, scal, want = [], [], []
for m in range(4):
    d_b = torch.empty(n * 64, dtype=torch.uint8, device="cuda:0")
    d_s = torch.empty(n * 32, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    t1, t2 = ctx.testkit_generate(0xC5 + m, n, d_b, d_s, want_dlogs=True)
    hb = np.zeros((n, 9), dtype=np.uint64)
    hb[:, :8] = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
    hs = d_s.cpu().numpy().view(np.uint64).reshape(n, 4).copy()
    devs = [(m * per + k) % ngpu for k in range(per)]
    bases.append(ctx.register_bases(hb, devs))
    scal.append(torch.from_numpy(hs).pin_memory().numpy())
    k = sum(int(w) << (64 * j) for j, w in enumerate(cpu_msm.dlog_checksum(hs, t1, t2)))
    kw = np.array([(k >> (64 * j)) & ((1 << 64) - 1) for j in range(4)], dtype=np.uint64)
    want.append(o.jac_to_affine(o.decode_jacobian(cpu_msm.scalar_mul_gen(kw))))
    del d_b, d_s, hb
ok = True
for _ in range(2):
    outs = ctx.msm_batch(bases, scal)
for r, w in zip(outs, want):
    ok &= o.jac_to_affine(o.decode_jacobian(r.words)) == w
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); ctx.msm_batch(bases, scal); ts.append((time.perf_counter() - t0) * 1e3)
single = []
for m in range(4):
    ctx.msm_registered(bases[m], scal[m])
    t0 = time.perf_counter(); ctx.msm_registered(bases[m], scal[m]); single.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"config": f"4 concurrent MSMs of 2^{log_n} over registered bases, {ngpu} GPUs, {per} GPU(s) per MSM",
                  "precomputed_table": bool(precompute), "layout": layout, "sum_of_single_latencies_ms": float(sum(single)),
                  "verified_vs_oracle": bool(ok), "batch_makespan_ms_median": float(np.median(ts)), "batch_makespan_ms_min": min(ts),
                  "points_per_s": 4 * n / (np.median(ts) * 1e-3), "single_msm_latency_ms": single,
                  "h2d_scalar_bytes": 4 * n * 32, "api": "b200msm_register_bases_on + b200msm_msm_batch (host scalars, pinned)"}))
