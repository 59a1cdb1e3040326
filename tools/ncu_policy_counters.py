"""Condense `ncu --csv --metrics ...` output of tools/autotune_sweep.py --policy-only into one record per (size, kernel):
the ncu evidence for each auto-tuner choice (integer-pipe utilisation of the point-arithmetic kernels, achieved HBM GB/s of
the decomposition / sort kernels).
usage: python tools/ncu_policy_counters.py ncu.csv sweep.jsonl > profiles/rNN_autotune_ncu.json"""
import csv
import json
import sys

HBM_PEAK = 6548.2   # MEASURED_PEAKS.json of this pool


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    launches = {}
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        key = int(r[col["ID"]])
        d = launches.setdefault(key, {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "")})
        d[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
    sweep = [json.loads(l) for l in open(sys.argv[2])]
    # the sweep runs 4 MSMs per size (1 warm-up + 3): group the launches in order, keep the last MSM of every size
    names = ("k_decompose_count", "k_decompose", "k_scatter_ranked", "k_partition", "k_place", "k_accumulate", "k_fixup_chunks", "k_fixup",
             "k_reduce_level", "k_bucket_reduce", "k_window_combine")
    seq = [launches[k] for k in sorted(launches)]
    out, pos = [], 0
    for row in sweep:
        per_msm = []
        for rep in range(4):
            cur, seen_acc = [], False
            while pos < len(seq):
                k = seq[pos]
                if k["kernel"].startswith("k_decompose") and cur:
                    break
                cur.append(k)
                pos += 1
            per_msm.append(cur)
        rec = {"log_n": row["log_n"], "glv": row["glv"], "window_bits": row["c"], "num_windows": row["W"], "kernels": {}}
        for k in per_msm[-1]:
            nm = next((n for n in names if k["kernel"].startswith(n)), None)
            if nm is None:
                continue
            t_ns = k.get("gpu__time_duration.sum", 0.0)
            byts = k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
            e = rec["kernels"].setdefault(nm, {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0, "fmaheavy_pct_time_weighted": 0.0})
            e["launches"] += 1
            e["time_us"] += t_ns / 1e3
            e["dram_bytes"] += byts
            e["fmaheavy_pct_time_weighted"] += k.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t_ns
        for nm, e in rec["kernels"].items():
            if e["time_us"]:
                e["fmaheavy_pct"] = round(e.pop("fmaheavy_pct_time_weighted") / (e["time_us"] * 1e3), 1)
                e["dram_gbs"] = round(e["dram_bytes"] / (e["time_us"] * 1e-6) / 1e9, 1)
                e["dram_frac_of_measured_peak"] = round(e["dram_gbs"] / HBM_PEAK, 3)
                e["time_us"] = round(e["time_us"], 1)
        out.append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
