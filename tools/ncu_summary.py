#!/usr/bin/env python
"""Digest an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed]` launch list of bench.py into per-kernel shares for ONE
denoising step (the launches between two `increment_kernel`s), and — when the counters were collected — the per-launch DRAM
traffic of the dominant kernel (profiles/rNN_gemm_traffic.json, read by bench.py) and the time-weighted tensor-pipe activity.

sm__pipe_tensor_cycles_active (NOT the `_realtime` / `hmma_*_realtime` variants, which pre-date UMMA and read garbage on sm_100)
counts UTCHMMA activity: calibrated on a convolution of known FLOPs, profiles/r02_tensor_counter_calibration.md.

    python tools/ncu_summary.py gpurun_out/launches.csv [--out profiles/r02_launches_summary.json] [--traffic profiles/r02_gemm_traffic.json] [--how "..."]
"""
import argparse
import collections
import csv
import json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--out", default="")
    ap.add_argument("--traffic", default="")
    ap.add_argument("--how", default="", help="how the pass was taken (cache control, graph / eager), recorded in the outputs")
    a = ap.parse_args()
    launches = collections.OrderedDict()
    with open(a.csv) as f:
        for r in csv.reader(f):
            if len(r) < 15 or not r[0].isdigit():
                continue
            lid, name, metric, unit, val = int(r[0]), r[4].split("(")[0], r[12], r[13], r[14].replace(",", "")
            d = launches.setdefault(lid, {"name": name})
            try:
                v = float(val)
            except ValueError:
                continue
            scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d[metric] = v * scale
    ids = list(launches)
    marks = [i for i in ids if "increment" in launches[i]["name"]]
    if len(marks) >= 2:
        ids = [i for i in ids if marks[-2] <= i < marks[-1]]
    TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
    per = collections.defaultdict(lambda: {"launches": 0, "ns": 0.0, "dram": 0.0, "tp_ns": 0.0})
    for i in ids:
        d = launches[i]
        k = per[d["name"].replace("mvd::", "").split("<")[0]]
        k["launches"] += 1
        k["ns"] += d.get("gpu__time_duration.sum", 0.0)
        k["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        k["tp_ns"] += d.get(TP, 0.0) * d.get("gpu__time_duration.sum", 0.0)
    total = sum(k["ns"] for k in per.values())
    have_tp = any(TP in launches[i] for i in ids)
    rows = [{"kernel": n, "launches": k["launches"], "ms": round(k["ns"] / 1e6, 4), "share": round(k["ns"] / total, 4),
             "dram_mb_per_launch": round(k["dram"] / k["launches"] / 1e6, 3) if k["dram"] else None,
             "tensor_pipe_active_pct": round(k["tp_ns"] / k["ns"], 2) if have_tp and k["ns"] else None}
            for n, k in sorted(per.items(), key=lambda kv: -kv[1]["ns"])]
    out = {"source": a.csv, "how": a.how, "note": "one denoising step; ncu per-launch times are serialised: compare SHARES",
           "launches_in_step": len(ids), "sum_ms": round(total / 1e6, 4),
           "tensor_pipe_active_pct_of_step": round(sum(k["tp_ns"] for k in per.values()) / total, 2) if have_tp and total else None,
           "kernels": rows}
    for r in rows[:12]:
        print(r)
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)
    if a.traffic and rows and rows[0]["dram_mb_per_launch"] is not None:
        json.dump({"kernel": rows[0]["kernel"], "dram_bytes_per_launch": rows[0]["dram_mb_per_launch"] * 1e6, "launches": rows[0]["launches"],
                   "tensor_pipe_active_pct": rows[0]["tensor_pipe_active_pct"], "tensor_pipe_active_pct_of_step": out["tensor_pipe_active_pct_of_step"],
                   "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the kernel's launches of one step ({a.csv}; {a.how})"},
                  open(a.traffic, "w"), indent=1)


if __name__ == "__main__":
    main()
