#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/launches.csv, gpurun_out/*.ncu-rep) into the small summaries committed under profiles/.

    python scripts/summarize_ncu.py <round-tag>          # e.g. r01
Writes profiles/<tag>_launches_summary.csv (per-kernel device time and share of one bench step, from the
`--metrics gpu__time_duration.sum` launch list), profiles/<tag>_ncu_full_summary.csv (the `--set full` capture: duration,
DRAM bytes, tensor-pipe activity, registers per launch) and profiles/ncu_traffic.json (DRAM bytes per launch by bench tag,
read by bench.py for roofline.traffic)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
EPI_TAG = {"1": "gemm_fwd1_bias_gelu", "2": "gemm_fwd2_bias_ssq", "3": "gemm_dh0_dgelu"}


def short(name):
    return name.replace("void ", "").replace("td::", "").split("(")[0]


def launches():
    path = os.path.join(ROOT, "gpurun_out", f"{TAG}_launches.csv")
    if not os.path.isfile(path):
        path = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.isfile(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ci = {h: i for i, h in enumerate(rows[0])}
    seq = [(short(r[ci["Kernel Name"]]), float(r[ci["Metric Value"]]) / 1e3, r[ci["Grid Size"]], r[ci["Block Size"]])
           for r in rows[1:] if r[ci["Metric Name"]] == "gpu__time_duration.sum"]
    starts = [i for i, s in enumerate(seq) if "cu_seqlens" in s[0]]
    a, b = starts[-2], starts[-1]  # one full step of the last (profiled) pass
    step = seq[a:b]
    total = sum(s[1] for s in step)
    agg = {}
    for name, us, grid, block in step:
        e = agg.setdefault(name, [0, 0.0, grid, block])
        e[0] += 1
        e[1] += us
    with open(os.path.join(OUT, f"{TAG}_launches_summary.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches_per_step", "us_per_step", "share_of_step_kernel_time", "grid", "block"])
        for name, (n, us, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, f"{us:.1f}", f"{us / total:.4f}", grid, block])
        w.writerow(["TOTAL (ncu: cold-cache, serialised launches)", len(step), f"{total:.1f}", "1.0", "", ""])
    print(f"launch list: {len(step)} launches/step, {total:.0f} us of kernel time")


def full():
    reps = [f for f in os.listdir(os.path.join(ROOT, "gpurun_out")) if f.endswith(".ncu-rep")]
    traffic = {}
    out_rows = []
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg",
            "sm__warps_active.avg.pct_of_peak_sustained_active"]
    for rep in sorted(reps):
        raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", rep), "--page", "raw", "--csv"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ci = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = short(r[ci["Kernel Name"]])
            rec = {"report": rep, "kernel": name}
            for m in want:
                if m in ci:
                    rec[m + " [" + units[ci[m]] + "]"] = r[ci[m]]
            out_rows.append(rec)
            # map to bench tags
            def mb(key):
                v, u = float(r[ci[key]]), units[ci[key]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            total = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
            tag = None
            if name.startswith("gemm_bf16_kernel<"):
                args = [a.strip() for a in name[name.index("<") + 1: name.index(">")].split(",")]
                if args[3] in EPI_TAG:
                    tag = EPI_TAG[args[3]]
                elif args[3] == "4":
                    tag = "gemm_dW2" if "gemm_dW2" not in traffic else "gemm_dW1"
            elif "pack_rows_kernel" in name:
                tag = "pack_varlen"
            elif "norm_mse_bwd" in name:
                tag = "norm_mse_bwd_fused"
            elif "adamw" in name:
                tag = "adamw_bf16"
            elif "rmsnorm_fwd" in name:
                tag = "rmsnorm_fwd"
            elif "rmsnorm_bwd" in name:
                tag = "rmsnorm_bwd"
            elif "masked_mse" in name:
                tag = "masked_mse"
            elif "masked_ce" in name:
                tag = "masked_ce"
            if tag and tag not in traffic:
                traffic[tag] = total
    if not out_rows:
        return
    keys = list(out_rows[0].keys())
    with open(os.path.join(OUT, f"{TAG}_ncu_full_summary.csv"), "w") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for rec in out_rows:
            w.writerow({k: rec.get(k, "") for k in keys})
    tpath = os.path.join(OUT, "ncu_traffic.json")
    old = json.load(open(tpath)) if os.path.isfile(tpath) else {}
    old.update(traffic)
    json.dump(old, open(tpath, "w"), indent=1)
    print("full capture:", len(out_rows), "launches;", "traffic tags:", sorted(traffic))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    launches()
    full()
