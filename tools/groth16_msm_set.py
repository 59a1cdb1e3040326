"""The MSM phase of one Groth16 proof at 2^LOGN constraints on one B200: four G1 MSMs (A, B1, C, H) over REGISTERED bases
with the precomputed window table, submitted as one batch, plus the G2 MSM (B2) over registered G2 bases (window table
too) -- the G1 batch and the G2 call run on two contexts from two host threads, so they overlap on the same GPU.  Inputs: device-generated G1 bases,
generator-multiple G2 bases tiled from 4096 distinct points (timing tool; parity lives in tests/).
usage: python tools/groth16_msm_set.py [LOGN=20] [REPS=5]"""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gpu-acceleration_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b200msm  # noqa: E402
import bn254_g2 as g2  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    n = 1 << lg
    c1, c2 = b200msm.Context([0]), b200msm.Context([0])
    handles, scal = [], []
    for m in range(4):
        d_b = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        d_s = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        c1.testkit_generate(0x616 + m, n, d_b, d_s)
        hb = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
        handles.append(c1.register_bases(hb, precompute=1))
        scal.append(d_s.cpu().pin_memory().numpy().view(np.uint64).reshape(n, 4))
        del d_b, d_s
    rec = np.array([g2.encode_base(pt) for pt in g2.random_points(4096, 1)], dtype=np.uint64)
    g2_bases = np.tile(rec, (-(-n // 4096), 1))[:n].copy()
    g2_handle = c2.g2_register_bases(g2_bases, 1)
    out = {}

    def g1_batch():
        t0 = time.perf_counter()
        c1.msm_batch(handles, scal)
        out["g1_batch_ms"] = (time.perf_counter() - t0) * 1e3

    def g2_msm():
        t0 = time.perf_counter()
        c2.g2_msm_registered(g2_handle, scal[1])
        out["g2_ms"] = (time.perf_counter() - t0) * 1e3

    rows = []
    for it in range(reps + 2):
        torch.cuda.synchronize()
        # alone
        g1_batch()
        g2_msm()
        alone = dict(out)
        # together
        t0 = time.perf_counter()
        th = [threading.Thread(target=g1_batch), threading.Thread(target=g2_msm)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        both = (time.perf_counter() - t0) * 1e3
        if it >= 2:
            rows.append((alone["g1_batch_ms"], alone["g2_ms"], both))
    med = lambda k: float(np.median([r[k] for r in rows]))
    print(json.dumps({"log_n": lg, "g1_batch_4_msms_table_ms": round(med(0), 3), "g2_msm_ms": round(med(1), 3),
                      "sequential_sum_ms": round(med(0) + med(1), 3), "overlapped_two_contexts_ms": round(med(2), 3),
                      "points_per_s_g1_equiv": round(5 * n / (med(2) * 1e-3))}))


if __name__ == "__main__":
    main()
