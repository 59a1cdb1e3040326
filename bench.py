#!/usr/bin/env python3
"""bench.py -- BN254 G1 variable-base MSM throughput on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 20] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one MSM over one batch of synthetic input: 2^log_n points PER GPU (weak scaling:
N ranks compute one MSM of N*2^log_n points sharded by contiguous point range; the only exchange
is an all-gather of the 96-byte per-rank partial sums, then every rank adds them on device, all
stream-ordered behind the MSM -- one host synchronisation per step).
Default workload = BASELINE.json configs[1]: "BN254 G1 MSM 2^20 on 1xB200".

  value : points/s with bases + scalars resident in HBM (CUDA events on the launch stream,
          L2 flushed between steps, max over ranks)
  e2e   : the same metric through the reference-facing C-ABI call b200msm_bn254_g1_msm with
          PINNED HOST buffers in arkworks layout (72-byte G1Affine records, 32-byte Fr), H2D
          of bases+scalars and D2H of the result inside the timed region (wall clock around
          the blocking call, max over ranks); e2e.h2d_ceiling = what N concurrent plain pinned
          uploads of the same bytes achieve on this box
  roofline : dominant kernel k_accumulate against the measured IMAD.WIDE issue rate
  cpu_baseline : oracle/cpu_msm.c (a C port of arkworks' msm_bigint_wnaf; arkworks itself
          cannot be built: no Rust toolchain) on the box's host cores, same inputs
  north_star : the other BASELINE.json configs under the same clock, every result checked against
          the oracle: configs[2] ONE 2^24-point MSM sharded over the N GPUs (strong scaling; resident,
          e2e and registered-bases e2e), configs[3] the 2^12..2^26 size sweep with the auto-tuned
          window (N = 1 only), configs[4] the Groth16-style batch of four 2^22 MSMs over registered
          bases (every MSM sharded over the N GPUs)

--impl reference times that CPU port with all host threads (rank 0 only).
Nothing here reads /root/reference.  oracle/ is used only as checker / CPU baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "gpu-acceleration_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "bn254_g1_msm_points_per_s"
UNIT = "points/s"
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML during the timed region."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.004)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ helpers
def _words_to_int(w) -> int:
    return sum(int(w[j]) << (64 * j) for j in range(4))


def expected_dlog(scalars_mont: np.ndarray, t1: np.ndarray, t2: np.ndarray) -> int:
    """sum_i s_i*(a[i%4096]+b[i//4096]) mod r by the C oracle (checker only)."""
    import cpu_msm
    return _words_to_int(cpu_msm.dlog_checksum(scalars_mont, t1, t2))


def point_of_dlog(k: int):
    import bn254 as o
    import cpu_msm
    kw = np.array([(k >> (64 * j)) & ((1 << 64) - 1) for j in range(4)], dtype=np.uint64)
    return o.jac_to_affine(o.decode_jacobian(cpu_msm.scalar_mul_gen(kw)))


def result_affine(words: np.ndarray):
    import bn254 as o
    return o.jac_to_affine(o.decode_jacobian(words))


def traffic_from_profiles():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def bind_to_gpu_numa(index: int):
    """One process per GPU: run this rank on the CPU cores NVML reports as local to its GPU, so that the pinned host
    buffers it allocates (first touch) live on that socket and its H2D copies do not cross the inter-socket link.
    Returns the number of cores bound, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------ B200 arm
def workload_config(log_n: int) -> dict:
    """The `config` object both arms print (the driver compares them)."""
    return {"workload": f"BN254 G1 MSM, 2^{log_n} random bases/scalars per GPU, one MSM per step, result bit-exact vs oracle",
            "log_n_per_gpu": log_n}


class Rig:
    """One rank's engine context + torch plumbing."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import b200msm
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
        self.numa_cores = bind_to_gpu_numa(self.local_rank) if self.world > 1 else None
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = b200msm.Context([self.local_rank])
        self.stream = torch.cuda.current_stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        self.ctx.set_option("timing", 1)
        self.ctx.set_option("window_bits", args.window_bits)
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.d_part = torch.zeros(96, dtype=torch.uint8, device=self.dev)
        self.d_out = torch.zeros(96, dtype=torch.uint8, device=self.dev)
        self.d_gather = torch.zeros(96 * max(1, self.world), dtype=torch.uint8, device=self.dev)
        self.launches = 0

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def generate(self, n: int, seed: int):
        torch = self.torch
        d_bases = torch.empty(n * 64, dtype=torch.uint8, device=self.dev)
        d_scalars = torch.empty(n * 32, dtype=torch.uint8, device=self.dev)
        torch.cuda.synchronize()
        t1, t2 = self.ctx.testkit_generate(seed, n, d_bases, d_scalars, want_dlogs=True)
        return d_bases, d_scalars, t1, t2

    def expected_point(self, d_scalars, n, t1, t2):
        """Checker only: point predicted from the discrete logs (sum over ALL ranks' shards)."""
        my = expected_dlog(d_scalars[:n * 32].cpu().numpy().view(np.uint64).reshape(n, 4), t1, t2)
        if self.world > 1:
            dl = [None] * self.world
            self.dist.all_gather_object(dl, my)
            my = sum(dl) % R_ORDER
        return point_of_dlog(my) if self.rank == 0 else None

    def combine_async(self):
        """d_part (this rank's partial) -> d_out on every rank; stream-ordered, no host synchronisation."""
        self.dist.all_gather_into_tensor(self.d_gather, self.d_part)
        self.ctx.sum_partials_device(self.d_gather, self.world, self.d_out, sync=False)
        self.launches += 1

    def step_resident(self, d_bases, d_scalars, n):
        """One MSM step, everything enqueued on torch's current stream; the caller synchronises once."""
        self.ctx.msm_device(d_bases, d_scalars, n, self.d_part if self.world > 1 else self.d_out, sync=False)
        if self.world > 1:
            self.combine_async()

    def time_resident(self, d_bases, d_scalars, n, steps, warmup, sampler=None):
        torch = self.torch
        for _ in range(warmup):
            self.flush.zero_()
            self.step_resident(d_bases, d_scalars, n)
        self.ctx.sync()
        self.launches = 0
        self.barrier()
        if sampler:
            sampler.start()
        wall0 = time.perf_counter()
        step_ms, stage = [], []
        for _ in range(steps):
            self.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            self.step_resident(d_bases, d_scalars, n)
            e1.record(self.stream)
            e1.synchronize()
            self.ctx.sync()                 # collects the stage events of the asynchronous MSM
            t = self.ctx.timings()
            t["sort_engine"] = self.ctx.last_sort_engine()
            self.launches += t["kernel_launches"]
            stage.append(t)
            step_ms.append(e0.elapsed_time(e1))
        self.barrier()
        wall_ms = (time.perf_counter() - wall0) * 1e3
        clocks = sampler.stop() if sampler else None
        total = self.max_over_ranks(sum(step_ms))
        return total / steps, stage, wall_ms, clocks

    def host_copies(self, d_bases, d_scalars, n):
        """Pinned host buffers in arkworks layout: (n, 9) u64 G1Affine records + (n, 4) u64 Fr."""
        torch = self.torch
        hb = np.zeros((n, 9), dtype=np.uint64)
        hb[:, :8] = d_bases[:n * 64].cpu().numpy().view(np.uint64).reshape(n, 8)
        return hb, torch.from_numpy(hb).pin_memory(), d_scalars[:n * 32].cpu().pin_memory()

    def time_e2e(self, call, steps, warmup):
        """call() -> 12 result words (host); wall clock around the blocking C-ABI call + combine, max over ranks."""
        torch = self.torch
        ms, final = [], None
        for it in range(warmup + steps):
            self.flush.zero_()
            self.barrier()
            t0 = time.perf_counter()
            words = call()
            if self.world > 1:
                self.d_part.copy_(torch.from_numpy(words.view(np.uint8)))
                self.combine_async()
                final = self.d_out.cpu().numpy().view(np.uint64)  # D2H read of the combined result
            else:
                final = words
            dt = (time.perf_counter() - t0) * 1e3
            if it >= warmup:
                ms.append(dt)
        return self.max_over_ranks(sum(ms)) / steps, final

    def h2d_ceiling(self, nbytes):
        """Plain pinned host->device copies of `nbytes`, all ranks at once: what the box's PCIe fabric gives N concurrent
        uploads (the e2e number cannot scale past it)."""
        torch = self.torch
        src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dst = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        dst.copy_(src, non_blocking=True)
        best = 1e30
        for _ in range(3):
            self.barrier()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        slowest = self.max_over_ranks(best)
        return {"bytes_per_rank": nbytes, "per_rank_gbs": nbytes / slowest / 1e9, "aggregate_gbs": self.world * nbytes / slowest / 1e9,
                "note": "N concurrent cudaMemcpyAsync from pinned memory, slowest rank; e2e moves this many bytes per step per rank"}


def stage_summary(stage, n, hbm_peak, sort_engine=1):
    dec_ms = statistics.mean(s["decompose_ms"] for s in stage)
    sort_ms = statistics.mean(s["sort_ms"] for s in stage)
    acc_ms = statistics.mean(s["accumulate_ms"] for s in stage)
    entries = stage[-1]["entries"]
    W, c = stage[-1]["num_windows"], stage[-1]["window_bits"]
    dbytes = 2 if c <= 16 else 4
    # Algorithmic bytes.  `entries` = non-zero digits ~ W * pseudo-points.
    #   K1: 32 B of scalar read per point + one digit written per entry (+ 4 B of rank per entry in the ranked engine)
    #   K2 partitioned: digit read, (entry 4 B + key 2 B) staged and read back, entry written  = dbytes + 16 B per entry
    #   K2 ranked:      digit + rank read, bucket end read, entry written                      = dbytes + 12 B per entry
    dec_bytes = n * 32 + entries * (dbytes + (0 if sort_engine == 2 else 4))
    sort_bytes = entries * (dbytes + (16 if sort_engine == 2 else 12))
    return {"decompose_ms": dec_ms, "sort_ms": sort_ms, "accumulate_ms": acc_ms,
            "reduce_ms": statistics.mean(s["reduce_ms"] for s in stage),
            "sort_engine": {0: "cursor atomics", 1: "ranked (global histogram atomics with rank + atomic-free scatter)",
                            2: "partitioned (shared-memory radix partition)"}.get(sort_engine),
            "decompose_hbm_gbs": dec_bytes / (dec_ms * 1e-3) / 1e9 if dec_ms else None,
            "sort_hbm_gbs": sort_bytes / (sort_ms * 1e-3) / 1e9 if sort_ms else None,
            "decompose_hbm_frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / hbm_peak if dec_ms and hbm_peak else None,
            "sort_hbm_frac": sort_bytes / (sort_ms * 1e-3) / 1e9 / hbm_peak if sort_ms and hbm_peak else None,
            "hbm_peak_gbs": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json" if hbm_peak else None,
            "window_bits": c, "num_windows": W, "entries": entries}


def north_star_block(rig, args):
    """BASELINE configs[2]: ONE MSM of 2^24 points sharded by contiguous point range over the N ranks (2^24 / N each)."""
    world, ctx = rig.world, rig.ctx
    log_total = args.north_star_log_n
    lg = log_total - int(np.log2(world))
    n = 1 << lg
    d_bases, d_scalars, t1, t2 = rig.generate(n, 0xB2240000 + rig.rank)
    want = rig.expected_point(d_scalars, n, t1, t2)
    ms, stage, _, _ = rig.time_resident(d_bases, d_scalars, n, steps=5, warmup=2)
    ok = True
    if rig.rank == 0:
        ok = result_affine(rig.d_out.cpu().numpy().view(np.uint64)) == want
    hb, h_bases, h_scalars = rig.host_copies(d_bases, d_scalars, n)
    e2e_ms, final = rig.time_e2e(lambda: ctx.msm_raw(h_bases.data_ptr(), 72, 0, 32, 64, h_scalars.data_ptr(), 32, n).words, steps=3, warmup=1)
    if rig.rank == 0:
        ok = ok and result_affine(final) == want
    handle = ctx.register_bases(hb, precompute=0)
    try:
        hs_np = h_scalars.numpy().view(np.uint64).reshape(n, 4)
        reg_ms, final = rig.time_e2e(lambda: ctx.msm_registered(handle, hs_np).words, steps=3, warmup=1)
    finally:
        handle.release()
    if rig.rank == 0:
        ok = ok and result_affine(final) == want
        if not ok:
            raise SystemExit("bench.py: north_star MSM result does not match the oracle -- number invalid")
    st = stage_summary(stage, n, None, stage[-1].get("sort_engine", 1))
    del d_bases, d_scalars, h_bases, h_scalars, hb
    rig.torch.cuda.empty_cache()
    return {"workload": f"ONE BN254 G1 MSM of 2^{log_total} points, sharded by contiguous point range over {world} GPU(s) "
                        f"(2^{lg} per GPU), 96-byte partial all-gather + device add", "log_n_total": log_total, "log_n_per_gpu": lg,
            "scaling": "strong", "resident_ms": ms, "resident_points_per_s": (1 << log_total) / (ms * 1e-3),
            "e2e_ms": e2e_ms, "e2e_points_per_s": (1 << log_total) / (e2e_ms * 1e-3),
            "e2e_h2d_bytes_per_gpu": n * 104, "e2e_registered_bases_ms": reg_ms, "e2e_registered_h2d_bytes_per_gpu": n * 32,
            "target_ms": 20.0, "verified_vs_oracle": bool(ok),
            "stages_rank0": {k: st[k] for k in ("decompose_ms", "sort_ms", "accumulate_ms", "reduce_ms", "window_bits", "num_windows")},
            "timing": "resident: CUDA events on the launch stream, 2 warm-up + 5 timed steps, L2 flushed, max over ranks; "
                      "e2e: wall clock around b200msm_bn254_g1_msm (pinned host buffers) / b200msm_msm_registered, 1 + 3 steps"}


def sweep_block(rig, args):
    """BASELINE configs[3]: input-size sweep with the auto-tuned window on ONE GPU (resident inputs)."""
    top = args.sweep_max_log_n
    nmax = 1 << top
    d_bases, d_scalars, t1, t2 = rig.generate(nmax, 0xB2260000)
    rows = []
    for lg in range(12, top + 1, 2):
        n = 1 << lg
        ms, stage, _, _ = rig.time_resident(d_bases, d_scalars, n, steps=3, warmup=1)
        want = point_of_dlog(expected_dlog(d_scalars[:n * 32].cpu().numpy().view(np.uint64).reshape(n, 4), t1, t2[:max(1, n >> 12)]))
        ok = result_affine(rig.d_out.cpu().numpy().view(np.uint64)) == want
        if not ok:
            raise SystemExit(f"bench.py: sweep 2^{lg} result does not match the oracle")
        rows.append({"log_n": lg, "ms": round(ms, 4), "points_per_s": n / (ms * 1e-3), "window_bits": stage[-1]["window_bits"],
                     "num_windows": stage[-1]["num_windows"], "verified": True})
    del d_bases, d_scalars
    rig.torch.cuda.empty_cache()
    return {"workload": f"2^12..2^{top} on one GPU, auto-tuned (scalar split, window) per size, resident inputs, 1 + 3 steps each", "rows": rows}


def batch_block(rig, args):
    """BASELINE configs[4]: Groth16-style batch -- 4 independent MSMs of 2^22 points over FIXED (registered) bases, every MSM
    sharded over all N ranks; per rank ONE b200msm_msm_batch call (scalars from pinned host memory) that pipelines its four
    shard-MSMs, then one all-gather of the 4 x 96-byte partials."""
    torch, ctx, world = rig.torch, rig.ctx, rig.world
    lg = args.batch_log_n - int(np.log2(world))
    n = 1 << lg
    sets, wants = [], []
    for k in range(4):
        d_b, d_s, t1, t2 = rig.generate(n, 0xB2500000 + 16 * k + rig.rank)
        wants.append(rig.expected_point(d_s, n, t1, t2))
        hb = d_b.cpu().numpy().view(np.uint64).reshape(n, 8)
        hs = d_s.cpu().pin_memory()
        sets.append((hb, hs))
        del d_b, d_s
    d_parts = torch.zeros(4 * 96, dtype=torch.uint8, device=rig.dev)
    d_all = torch.zeros(4 * 96 * world, dtype=torch.uint8, device=rig.dev)
    d_res = torch.zeros(4 * 96, dtype=torch.uint8, device=rig.dev)
    out = {}
    for mode, pre in (("plain", 0), ("window_table", 1)):
        handles = [ctx.register_bases(hb, precompute=pre) for hb, _ in sets]
        try:
            ms = []
            for it in range(1 + 3):
                rig.flush.zero_()
                rig.barrier()
                t0 = time.perf_counter()
                res = ctx.msm_batch(handles, [hs.numpy().view(np.uint64).reshape(n, 4) for _, hs in sets])
                if world > 1:
                    d_parts.copy_(torch.from_numpy(np.concatenate([r.words for r in res]).view(np.uint8)))
                    rig.dist.all_gather_into_tensor(d_all, d_parts)
                    v = d_all.view(world, 4, 96)
                    for k in range(4):
                        ctx.sum_partials_device(v[:, k, :].contiguous(), world, d_res[96 * k:96 * k + 96], sync=False)
                    final = d_res.cpu().numpy().view(np.uint64).reshape(4, 12)
                else:
                    final = np.stack([r.words for r in res])
                dt = (time.perf_counter() - t0) * 1e3
                if it >= 1:
                    ms.append(dt)
            makespan = rig.max_over_ranks(sum(ms)) / 3
        finally:
            for h_ in handles:
                h_.release()
        if rig.rank == 0:
            for k in range(4):
                if result_affine(final[k]) != wants[k]:
                    raise SystemExit("bench.py: batch MSM result does not match the oracle -- number invalid")
        out[mode + "_makespan_ms"] = makespan
    out.update({"workload": f"4 MSMs of 2^{args.batch_log_n} points over registered bases, each sharded over {world} GPU(s) (2^{lg} per GPU and "
                            "MSM); scalars from pinned host memory inside the timed region", "points_per_s": 4 * (1 << args.batch_log_n) /
                (min(out["plain_makespan_ms"], out["window_table_makespan_ms"]) * 1e-3), "verified_vs_oracle": True,
                "h2d_bytes_per_gpu": 4 * n * 32, "timing": "wall clock, barrier + synchronize on both sides, 1 warm-up + 3 timed batches, max over ranks"})
    rig.torch.cuda.empty_cache()
    return out


def run_b200(args):
    rig = Rig(args)
    torch, ctx, world, rank = rig.torch, rig.ctx, rig.world, rig.rank
    n = 1 << args.log_n
    d_bases, d_scalars, t1, t2 = rig.generate(n, 0xB2000000 + rank)
    ctx.msm_device(d_bases, d_scalars, n, rig.d_out)   # first touch of every buffer
    peak_macs = ctx.imad_peak()  # measured IMAD.WIDE rate on this device, this run
    sampler = ClockSampler(rig.local_rank)
    ms_per_step, stage, wall_ms, clocks = rig.time_resident(d_bases, d_scalars, n, args.steps, args.warmup, sampler)
    gpu_launches = rig.launches
    value = world * n / (ms_per_step * 1e-3)

    # ---- verification of the resident result (outside the timed region; oracle = checker only)
    want = rig.expected_point(d_scalars, n, t1, t2)
    verified = True
    if rank == 0:
        verified = result_affine(rig.d_out.cpu().numpy().view(np.uint64)) == want
        if not verified:
            raise SystemExit("bench.py: resident MSM result does not match the oracle -- number invalid")

    # ---- e2e: host buffers (arkworks layout, pinned) through the drop-in C-ABI call
    hb, h_bases, h_scalars = rig.host_copies(d_bases, d_scalars, n)
    e2e_ms, final = rig.time_e2e(lambda: ctx.msm_raw(h_bases.data_ptr(), 72, 0, 32, 64, h_scalars.data_ptr(), 32, n).words,
                                 args.steps, args.warmup)
    if rank == 0 and result_affine(final) != want:
        raise SystemExit("bench.py: e2e MSM result does not match the oracle -- number invalid")
    e2e_value = world * n / (e2e_ms * 1e-3)
    h2d = rig.h2d_ceiling(n * 104)

    # ---- the two other timings SURVEY 8(d) names, N = 1 only: (ii) registered (device-resident) bases with the scalars
    # coming from pinned host memory, (iii) the cold drop-in call from PAGEABLE host memory.  Same result check.
    variants = None
    if world == 1 and not args.no_variants:
        hs_np = h_scalars.numpy().view(np.uint64).reshape(n, 4)
        reg = {}
        for mode in (0, 1):           # plain registered bases, then with the precomputed window table
            t0 = time.perf_counter()
            handle = ctx.register_bases(hb, precompute=mode)
            reg[("register_ms", mode)] = (time.perf_counter() - t0) * 1e3
            try:
                ms = []
                for it in range(3 + 5):
                    rig.flush.zero_()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    r2 = ctx.msm_registered(handle, hs_np)
                    if it >= 3:
                        ms.append((time.perf_counter() - t0) * 1e3)
                reg[("ms", mode)] = statistics.median(ms)
                reg[("c", mode)] = ctx.timings()["window_bits"]
            finally:
                handle.release()
            if result_affine(r2.words) != want:
                raise SystemExit("bench.py: registered MSM result does not match the oracle -- number invalid")
        hb_pageable, hs_pageable = hb.copy(), hs_np.copy()
        page_ms = []
        for it in range(2 + 4):
            rig.flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r3 = ctx.msm_raw(hb_pageable.ctypes.data, 72, 0, 32, 64, hs_pageable.ctypes.data, 32, n)
            if it >= 2:
                page_ms.append((time.perf_counter() - t0) * 1e3)
        if result_affine(r3.words) != want:
            raise SystemExit("bench.py: pageable MSM result does not match the oracle -- number invalid")
        variants = {"registered_bases_ms": reg[("ms", 0)], "registered_h2d_bytes": n * 32,
                    "registered_table_ms": reg[("ms", 1)], "table_window_bits": reg[("c", 1)],
                    "table_build_ms": reg[("register_ms", 1)],
                    "pageable_host_ms": statistics.median(page_ms), "pageable_h2d_bytes": n * (72 + 32),
                    "note": "wall clock around the C-ABI call; registered = b200msm_msm_registered (scalars from pinned host "
                            "memory, bases resident; _table_ = with the one-time precomputed 2^(c*w)*P table); "
                            "pageable = b200msm_bn254_g1_msm on plain malloc'd numpy arrays"}

    # ---- roofline of the dominant kernel (k_accumulate alone: the stage events bracket exactly its launches)
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    stages = stage_summary(stage, n, hbm_peak, stage[-1].get("sort_engine", 1))
    acc_ms, entries, W, c = stages["accumulate_ms"], stages["entries"], stages["num_windows"], stages["window_bits"]
    W_plain = -(-254 // c)          # window count of the plain (non-GLV) algorithm the BASELINE.md formula assumes
    glv = W < W_plain               # the engine split the scalars (127-bit halves over 2n pseudo-points)
    alg_macs = entries * 10 * 136  # mixed XYZZ additions x (8M+2S) x 136 MAC32
    achieved = alg_macs / (acc_ms * 1e-3)
    traffic = traffic_from_profiles()
    roofline = {"bound": "imad", "kernel": "k_accumulate", "achieved": achieved / 1e12, "peak": peak_macs / 1e12,
                "unit": "TMAC32/s", "frac": achieved / peak_macs, "peak_source": "measured in this run (IMAD.WIDE.U32 issue rate)",
                "traffic": (traffic.get("k_accumulate_dram_bytes_per_launch")
                            if traffic.get("log_n") == args.log_n and traffic.get("window_bits") == c else None),
                "kernel_ms": acc_ms, "algorithmic_macs_per_launch": alg_macs, "executed_macs_per_addition": 8 * 136 + 200,
                "whole_msm_frac": (W_plain * (10 * n + 28 * (1 << (c - 1))) + 9 * W_plain * c) * 136 / (ms_per_step * 1e-3) / peak_macs,
                "note": "achieved = mixed additions actually executed (entries) x 10 mul x 136 MAC32 / time between the CUDA events that "
                        "bracket the k_accumulate launch(es) on the launch stream; peak = plain IMAD issue rate (64/clk/SM). A 32x32->64 MAC "
                        "with carry costs two passes of that pipe on sm_100 (profiles/r01_pipe_bench4_instruction_forms.jsonl), so 0.5 "
                        "is the practical ceiling; ncu fmaheavy pipe-busy for this kernel: 86.0% at 2^20, 89.3% at 2^24 (profiles/r02x_ncu_full_summary_2e20_2e24.json). The kernel "
                        "executes FEWER MACs than the formula charges: Y3 = R(Q-X3) - Y1*PPP is one fused product pair with a single "
                        "Montgomery reduction (fq_mulsub, 200 MACs instead of 272), i.e. 1288 per addition"}

    cfg = workload_config(args.log_n)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": cfg,
            "engine": {"points_total": world * n, "window_bits": c, "num_windows": W, "glv_split": glv,
                       "sharding": "contiguous point ranges, 96-byte partial all-gather + device add, stream-ordered (no host "
                                   "synchronisation between the MSM and the collective)" if world > 1 else "single GPU",
                       "l2": "512 MiB buffer written between timed steps (L2 flush)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (72 + 32), "d2h_bytes_per_step": 96,
                    "ms_per_step": e2e_ms, "api": "b200msm_bn254_g1_msm (host buffers, pinned, arkworks layout)",
                    "host_numa_binding": (f"rank bound to the {rig.numa_cores} cores local to its GPU" if rig.numa_cores else "none"),
                    "h2d_ceiling": h2d},
            "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline, "stages": stages,
            "verified_vs_oracle": verified, "wall_ms_timed_region": wall_ms}
    if variants:
        line["e2e_variants"] = variants
    del d_bases, d_scalars, h_bases, h_scalars, hb
    torch.cuda.empty_cache()
    if not args.no_north_star:
        line["north_star"] = north_star_block(rig, args)
        if world == 1:
            line["north_star"]["sweep"] = sweep_block(rig, args)
        line["north_star"]["groth16_batch"] = batch_block(rig, args)

    # ---- CPU baseline beside it (rank 0, N = 1 only): the C port of arkworks' MSM on the same inputs
    if rank == 0 and world == 1 and not args.no_cpu:
        import cpu_msm
        sample = min(n, 1 << args.cpu_log_n)
        b8, hs_s, _, _ = cpu_msm.testkit_generate(0xB2000000, sample)
        hb_s = np.zeros((sample, 9), dtype=np.uint64)
        hb_s[:, :8] = b8
        t0 = time.perf_counter()
        out, used = cpu_msm.msm(hb_s, hs_s, os.cpu_count() or 1)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": used, "kind": "port",
                                "host_cpus": os.cpu_count(), "seconds": dt,
                                "sample": f"first 2^{int(np.log2(sample))} points of the same workload, one MSM, "
                                          "oracle/cpu_msm.c (arkworks msm_bigint_wnaf restated; windows run in parallel, "
                                          "so threads used = min(cores, windows))"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        rig.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import cpu_msm
    n = 1 << min(args.log_n, args.cpu_log_n)
    # Inputs come from the CPU restatement of the test kit's generator (same bytes as the B200 arm's rank 0;
    # tests/test_gpu_large.py::test_device_generator_equals_cpu_generator): this arm never loads libb200msm.so.
    b8, hs, _, _ = cpu_msm.testkit_generate(0xB2000000, n)
    hb = np.zeros((n, 9), dtype=np.uint64)
    hb[:, :8] = b8
    threads = os.cpu_count() or 1
    times, used = [], 1
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, used = cpu_msm.msm(hb, hs, threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = statistics.mean(times) * 1e3
    value = n / (ms * 1e-3)
    sample = (f"one MSM of 2^{int(np.log2(n))} points per step (same generator/seed as the B200 arm's rank 0), "
              "oracle/cpu_msm.c = C port of arkworks 0.4 msm_bigint_wnaf; arkworks itself needs Rust, absent here")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": workload_config(args.log_n),
            "engine": {"implementation": "CPU port of arkworks msm_bigint_wnaf on the host cores (oracle/cpu_msm.c)", "points_total": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "host_cpus": threads, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ instance replay
def run_replay(args):
    """The reference's file-backed benchmark harness (`run_benchmark`, src/msm/arkworks_pippenger.rs:45-75: open the
    `points` / `scalars` instance files of a directory -- generate them first if they are missing --, run the MSM over
    every instance, report the average) with the CUDA engine in place of `G::msm`, and the CPU port beside it on the same
    decoded inputs (bench.py's cpu_baseline leg).  Prints the reference's BenchmarkResult fields as JSON."""
    import b200msm
    import cpu_msm
    import torch
    directory = args.replay
    ctx = b200msm.Context([0])
    if not (os.path.exists(os.path.join(directory, "points")) and os.path.exists(os.path.join(directory, "scalars"))):
        size, count = (int(x) for x in args.replay_gen.split(","))
        insts = []
        for k in range(count):   # gen_vectors (preprocess.rs:181-225): random instances; here from the device test kit
            d_b = torch.empty(size * 64, dtype=torch.uint8, device="cuda")
            d_s = torch.empty(size * 32, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            ctx.testkit_generate(0xF11E0000 + k, size, d_b, d_s)
            insts.append((d_b.cpu().numpy().view(np.uint64).reshape(size, 8), d_s.cpu().numpy().view(np.uint64).reshape(size, 4)))
        b200msm.write_instance_files(directory, insts)
    gpu_ms, cpu_ms, sizes = [], [], []
    for comp, canon in b200msm.read_instance_files(directory):
        b200msm.msm_from_instance(ctx, comp, canon)          # warm-up (buffers, first-touch)
        t0 = time.perf_counter()
        res = b200msm.msm_from_instance(ctx, comp, canon)    # decode on the GPU + MSM, as benchmark_msm times it
        gpu_ms.append((time.perf_counter() - t0) * 1e3)
        bases, bad = ctx.decompress_g1(comp)
        scal = ctx.fr_to_montgomery(canon)
        # the decoder marks the identity as (0, 0); the CPU port takes arkworks' explicit `infinity` flag (9th word)
        rec = np.zeros((len(bases), 9), dtype=np.uint64)
        rec[:, :8] = bases
        rec[:, 8] = (~bases.any(axis=1)).astype(np.uint64)
        t0 = time.perf_counter()
        out, used = cpu_msm.msm(rec, scal, os.cpu_count() or 1)
        cpu_ms.append((time.perf_counter() - t0) * 1e3)
        if result_affine(res.words) != result_affine(out):
            raise SystemExit("bench.py --replay: CUDA result differs from the CPU port")
        sizes.append(len(canon))
    if not sizes:
        raise SystemExit(f"bench.py --replay: no instances in {directory}")
    print(json.dumps({"impl": "replay", "instance_size": sizes[0], "num_instance": len(sizes),
                      "avg_processing_time": statistics.mean(gpu_ms), "unit": "ms",
                      "cpu_port_avg_processing_time": statistics.mean(cpu_ms), "cpu_cores": used,
                      "verified_equal": True, "directory": directory,
                      "mirrors": "arkworks_pippenger.rs:45-75 run_benchmark -> BenchmarkResult {instance_size, num_instance, avg_processing_time}"}))
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=20, help="log2 points per GPU (BASELINE configs[1] = 20)")
    ap.add_argument("--cpu-log-n", type=int, default=20, help="largest CPU-baseline sample (log2 points)")
    ap.add_argument("--window-bits", type=int, default=0, help="0 = auto-tuned")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-variants", action="store_true", help="skip the registered-bases / pageable timings (large sizes)")
    ap.add_argument("--no-north-star", action="store_true", help="skip the north_star block (2^24 sharded, sweep, Groth16 batch)")
    ap.add_argument("--north-star-log-n", type=int, default=24, help="log2 of the TOTAL points of the sharded north-star MSM")
    ap.add_argument("--sweep-max-log-n", type=int, default=26, help="largest size of the single-GPU sweep (log2)")
    ap.add_argument("--batch-log-n", type=int, default=22, help="log2 points of each of the four MSMs of the Groth16-style batch")
    ap.add_argument("--replay", default=None, metavar="DIR", help="replay the reference's instance files (DIR/points, DIR/scalars) "
                    "through the CUDA engine and the CPU port; generates them first when missing")
    ap.add_argument("--replay-gen", default="65536,4", help="instance_size,num_instance for --replay when DIR has no files")
    args = ap.parse_args()
    if args.replay:
        return run_replay(args)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
