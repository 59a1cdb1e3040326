"""A/B of two library builds on the same box: run once plain (the checkout's library) and once with B200MSM_LIB=/path/to/other/libb200msm.so.
Prints e2e (pinned host), registered, registered + table and resident medians at 2^20."""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpu-acceleration_b200"))
import b200msm
ctx = b200msm.Context()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
n = 1 << 20
d_b = torch.empty(n*64, dtype=torch.uint8, device="cuda"); d_s = torch.empty(n*32, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); ctx.testkit_generate(11, n, d_b, d_s)
hb = np.zeros((n, 9), dtype=np.uint64); hb[:, :8] = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
h_b = torch.from_numpy(hb).pin_memory(); h_s = d_s.cpu().pin_memory()
hs_np = h_s.numpy().view(np.uint64).reshape(n, 4)
def med(f, reps=15):
    ts = []
    for it in range(reps):
        flush.fill_(it); torch.cuda.synchronize()
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts = sorted(ts[3:]); return round(ts[len(ts)//2], 4)
out = {"lib": os.environ.get("B200MSM_LIB", "HEAD")}
out["e2e_ms"] = med(lambda: ctx.msm_raw(h_b.data_ptr(), 72, 0, 32, 64, h_s.data_ptr(), 32, n))
for pre in (0, 1):
    ctx.set_option("precompute", pre); h = ctx.register_bases(hb[:, :8].copy()); ctx.set_option("precompute", 0)
    out["registered_table_ms" if pre else "registered_ms"] = med(lambda: ctx.msm_registered(h, hs_np))
    h.release()
d_o = torch.zeros(96, dtype=torch.uint8, device="cuda")
out["resident_ms"] = med(lambda: (ctx.msm_device(d_b, d_s, n, d_o), torch.cuda.synchronize()))
print(json.dumps(out), flush=True)
