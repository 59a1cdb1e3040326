#!/usr/bin/env python3
"""bench.py -- BN254 G1 variable-base MSM throughput on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 20] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one MSM over one batch of synthetic input: 2^log_n points PER GPU (weak scaling:
N ranks compute one MSM of N*2^log_n points sharded by contiguous point range; the only exchange
is an all-gather of the 96-byte per-rank partial sums, then every rank adds them on device).
Default workload = BASELINE.json configs[1]: "BN254 G1 MSM 2^20 on 1xB200".

  value : points/s with bases + scalars resident in HBM (CUDA events on the launch stream,
          L2 flushed between steps, max over ranks)
  e2e   : the same metric through the reference-facing C-ABI call b200msm_bn254_g1_msm with
          PINNED HOST buffers in arkworks layout (72-byte G1Affine records, 32-byte Fr), H2D
          of bases+scalars and D2H of the result inside the timed region (wall clock around
          the blocking call, max over ranks)
  roofline : dominant kernel k_accumulate(+fix-up) against the measured IMAD.WIDE issue rate
  cpu_baseline : oracle/cpu_msm.c (a C port of arkworks' msm_bigint_wnaf; arkworks itself
          cannot be built: no Rust toolchain) on the box's host cores, same inputs

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
def run_b200(args):
    import torch
    import torch.distributed as dist
    import b200msm
    import msm_dist as msmdist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    numa_cores = bind_to_gpu_numa(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ctx = b200msm.Context([local_rank])
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_option("timing", 1)
    ctx.set_option("window_bits", args.window_bits)

    n = 1 << args.log_n
    d_bases = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    d_scalars = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_part = torch.zeros(96, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(96, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    t1, t2 = ctx.testkit_generate(0xB2000000 + rank, n, d_bases, d_scalars, want_dlogs=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def sum_fn(gathered, cnt):
        ctx.sum_partials_device(gathered, cnt, d_out, sync=False)
        return d_out

    launches = [0]

    def step_resident():
        ctx.msm_device(d_bases, d_scalars, n, d_part if world > 1 else d_out, sync=True)
        launches[0] += ctx.timings()["kernel_launches"]
        if world > 1:
            msmdist.combine(d_part, sum_fn)
            launches[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step_resident()
    peak_macs = ctx.imad_peak()  # measured IMAD.WIDE rate on this device, this run
    launches[0] = 0
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    step_ms, stage = [], []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_resident()
        stage.append(ctx.timings())
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    gpu_launches = launches[0]
    total_ms = sum(step_ms)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- verification of the resident result (outside the timed region; oracle = checker only)
    my_dlog = expected_dlog(d_scalars.cpu().numpy().view(np.uint64).reshape(n, 4), t1, t2)
    if world > 1:
        dl = [None] * world
        dist.all_gather_object(dl, my_dlog)
        tot = sum(dl) % R_ORDER
    else:
        tot = my_dlog
    got = result_affine(d_out.cpu().numpy().view(np.uint64))
    verified = True
    if rank == 0:
        verified = got == point_of_dlog(tot)
        if not verified:
            raise SystemExit("bench.py: resident MSM result does not match the oracle -- number invalid")

    # ---- e2e: host buffers (arkworks layout, pinned) through the drop-in C-ABI call
    hb = np.zeros((n, 9), dtype=np.uint64)
    hb[:, :8] = d_bases.cpu().numpy().view(np.uint64).reshape(n, 8)
    h_bases = torch.from_numpy(hb).pin_memory()
    h_scalars = d_scalars.cpu().pin_memory()
    e2e_ctx = ctx
    e2e_ms = []
    for it in range(args.warmup + args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        res = e2e_ctx.msm_raw(h_bases.data_ptr(), 72, 0, 32, 64, h_scalars.data_ptr(), 32, n)
        if world > 1:
            d_part.copy_(torch.from_numpy(res.words.view(np.uint8)))
            msmdist.combine(d_part, sum_fn)
            final = d_out.cpu().numpy().view(np.uint64)  # D2H read of the combined result
        else:
            final = res.words
        dt = (time.perf_counter() - t0) * 1e3
        if it >= args.warmup:
            e2e_ms.append(dt)
    e2e_total = sum(e2e_ms)
    if world > 1:
        t = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    if rank == 0 and result_affine(final) != point_of_dlog(tot):
        raise SystemExit("bench.py: e2e MSM result does not match the oracle -- number invalid")
    e2e_value = world * n / (e2e_total / args.steps * 1e-3)

    # ---- the two other timings SURVEY 8(d) names, N = 1 only: (ii) registered (device-resident) bases with the scalars
    # coming from pinned host memory, (iii) the cold drop-in call from PAGEABLE host memory.  Same result check.
    variants = None
    if world == 1 and not args.no_variants:
        want_pt = point_of_dlog(tot)
        hs_np = h_scalars.numpy().view(np.uint64).reshape(n, 4)
        reg = {}
        for mode in (0, 1):           # plain registered bases, then with the precomputed window table
            ctx.set_option("precompute", mode)
            try:
                t0 = time.perf_counter()
                handle = ctx.register_bases(hb)
                reg[("register_ms", mode)] = (time.perf_counter() - t0) * 1e3
            finally:
                ctx.set_option("precompute", 0)
            try:
                ms = []
                for it in range(3 + 5):
                    flush.zero_()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    r2 = ctx.msm_registered(handle, hs_np)
                    if it >= 3:
                        ms.append((time.perf_counter() - t0) * 1e3)
                reg[("ms", mode)] = statistics.median(ms)
                reg[("c", mode)] = ctx.timings()["window_bits"]
            finally:
                handle.release()
            if result_affine(r2.words) != want_pt:
                raise SystemExit("bench.py: registered MSM result does not match the oracle -- number invalid")
        hb_pageable, hs_pageable = hb.copy(), hs_np.copy()
        page_ms = []
        for it in range(2 + 4):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r3 = ctx.msm_raw(hb_pageable.ctypes.data, 72, 0, 32, 64, hs_pageable.ctypes.data, 32, n)
            if it >= 2:
                page_ms.append((time.perf_counter() - t0) * 1e3)
        if result_affine(r3.words) != want_pt:
            raise SystemExit("bench.py: pageable MSM result does not match the oracle -- number invalid")
        variants = {"registered_bases_ms": reg[("ms", 0)], "registered_h2d_bytes": n * 32,
                    "registered_table_ms": reg[("ms", 1)], "table_window_bits": reg[("c", 1)],
                    "table_build_ms": reg[("register_ms", 1)],
                    "pageable_host_ms": statistics.median(page_ms), "pageable_h2d_bytes": n * (72 + 32),
                    "note": "wall clock around the C-ABI call; registered = b200msm_msm_registered (scalars from pinned host "
                            "memory, bases resident; _table_ = with the one-time precomputed 2^(c*w)*P table, option 'precompute'); "
                            "pageable = b200msm_bn254_g1_msm on plain malloc'd numpy arrays"}

    # ---- roofline of the dominant kernel (k_accumulate + boundary fix-up), live stage events
    acc_ms = statistics.mean(s["accumulate_ms"] for s in stage)
    entries = stage[-1]["entries"]
    W, c = stage[-1]["num_windows"], stage[-1]["window_bits"]
    W_plain = -(-254 // c)          # window count of the plain (non-GLV) algorithm the BASELINE.md formula assumes
    glv = W < W_plain               # the engine split the scalars (127-bit halves over 2n pseudo-points)
    alg_macs = entries * 10 * 136  # mixed XYZZ additions x (8M+2S) x 136 MAC32
    achieved = alg_macs / (acc_ms * 1e-3)
    traffic = traffic_from_profiles()
    roofline = {"bound": "imad", "kernel": "k_accumulate(+k_fixup)", "achieved": achieved / 1e12, "peak": peak_macs / 1e12,
                "unit": "TMAC32/s", "frac": achieved / peak_macs, "peak_source": "measured in this run (IMAD.WIDE.U32 issue rate)",
                "traffic": (traffic.get("k_accumulate_dram_bytes_per_launch")
                            if traffic.get("log_n") == args.log_n and traffic.get("window_bits") == c else None),
                "kernel_ms": acc_ms, "algorithmic_macs_per_launch": alg_macs,
                "whole_msm_frac": (W_plain * (10 * n + 28 * (1 << (c - 1))) + 9 * W_plain * c) * 136 / (ms_per_step * 1e-3) / peak_macs,
                "note": "achieved = mixed additions actually executed (entries) x 10 mul x 136 MAC32 / kernel time; peak = plain "
                        "IMAD issue rate (64/clk/SM). A 32x32->64 MAC with carry costs two passes of that pipe on sm_100 "
                        "(profiles/r01_pipe_bench4_instruction_forms.jsonl), so 0.5 is the practical ceiling; ncu fmaheavy "
                        "pipe-busy for this kernel: 85.4% (profiles/r01h_ncu_full_summary.json)"}
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    dec_ms = statistics.mean(s["decompose_ms"] for s in stage)
    sort_ms = statistics.mean(s["sort_ms"] for s in stage)
    dbytes = 2 if c <= 16 else 4
    stages = {"decompose_ms": dec_ms, "sort_ms": sort_ms, "accumulate_ms": acc_ms,
              "reduce_ms": statistics.mean(s["reduce_ms"] for s in stage),
              "decompose_hbm_gbs": n * (32 + W * dbytes) / (dec_ms * 1e-3) / 1e9,
              "sort_hbm_gbs": (n * W * dbytes + entries * 4) / (sort_ms * 1e-3) / 1e9,
              "hbm_peak_gbs": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json" if hbm_peak else None}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"BN254 G1 MSM, 2^{args.log_n} random bases/scalars per GPU ({world} x 2^{args.log_n} points total), "
                                   "bit-exact vs oracle", "log_n_per_gpu": args.log_n, "window_bits": c, "num_windows": W, "glv_split": glv,
                       "sharding": "contiguous point ranges, 96-byte partial all-gather + device add" if world > 1 else "single GPU",
                       "l2": "512 MiB buffer written between timed steps (L2 flush)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (72 + 32), "d2h_bytes_per_step": 96,
                    "ms_per_step": e2e_total / args.steps, "api": "b200msm_bn254_g1_msm (host buffers, pinned, arkworks layout)",
                    "host_numa_binding": (f"rank bound to the {numa_cores} cores local to its GPU" if numa_cores else "none")},
            "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline, "stages": stages,
            "verified_vs_oracle": verified, "wall_ms_timed_region": wall_ms}
    if variants:
        line["e2e_variants"] = variants

    # ---- CPU baseline beside it (rank 0, N = 1 only): the C port of arkworks' MSM on the same inputs
    if rank == 0 and world == 1 and not args.no_cpu:
        import cpu_msm
        sample = min(n, 1 << args.cpu_log_n)
        hb_s, hs_s = hb[:sample], h_scalars.numpy().view(np.uint64).reshape(n, 4)[:sample]
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
        dist.destroy_process_group()


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
            "config": {"workload": f"BN254 G1 MSM, 2^{args.log_n} random bases/scalars per GPU, CPU port of arkworks on host cores",
                       "log_n_per_gpu": args.log_n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "host_cpus": threads, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
