#!/usr/bin/env python
"""Compact view of `ncu -i report.ncu-rep --page source --csv` for one kernel launch: the SASS instructions that collected the most
warp-stall samples and the totals per stall reason (the per-instruction page itself is several MB). No GPU needed.
    ncu -i r.ncu-rep --page source --csv --kernel-name regex:NAME --launch-skip K --launch-count 1 | python scripts/summarize_ncu_source.py [--top 30]"""
import argparse
import csv
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=30)
    a = ap.parse_args()
    rows = list(csv.reader(sys.stdin))
    name = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else ""
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body, seen = [], set()
    for r in rows[hdr_i + 1 :]:  # (the page lists every instruction once per source view: keep the first)
        if len(r) == len(hdr) and r[0] not in seen:
            seen.add(r[0])
            body.append(r)
    col = {h: i for i, h in enumerate(hdr)}
    samples = col["# Samples"] if "# Samples" in col else col["Warp Stall Sampling (All Samples)"]
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0

    total = sum(num(r[samples]) for r in body) or 1.0
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", name])
    w.writerow(["instructions", len(body), "stall samples", int(total)])
    w.writerow([])
    w.writerow(["stall reason", "samples", "share"])
    for h in sorted(reasons, key=lambda h: -sum(num(r[col[h]]) for r in body)):
        s = sum(num(r[col[h]]) for r in body)
        if s:
            w.writerow([h, int(s), f"{s / total:.3f}"])
    w.writerow([])
    w.writerow(["rank", "offset", "SASS", "samples", "share", "top reason", "instructions executed"])
    base = int(body[0][0], 16)
    ranked = sorted(body, key=lambda r: -num(r[samples]))[: a.top]
    for k, r in enumerate(ranked):
        top = max(reasons, key=lambda h: num(r[col[h]]))
        w.writerow([k + 1, hex(int(r[0], 16) - base), " ".join(r[col["Source"]].split()), int(num(r[samples])), f"{num(r[samples]) / total:.3f}", top,
                    r[col["Instructions Executed"]]])


if __name__ == "__main__":
    main()
