#!/usr/bin/env python3
"""Per-source-line stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: srcstalls.py <report.ncu-rep> <kernel regex> [top N]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, hdr, data = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":     # line-aggregated row
        d = dict(zip(hdr, r))
        try:
            s, i = int(d["# Samples"]), int(d["Instructions Executed"])
        except ValueError:
            continue
        st = {k[6:]: int(d[k]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and d[k].isdigit()}
        data.append((s, i, fname, r[0], r[1].strip(), st, d.get("L1 Wavefronts Shared", "0"), d.get("L1 Wavefronts Shared Ideal", "0")))
ts, ti = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"# kernel regex {rx}: {ts} stall samples, {ti} warp instructions")
for s, i, f, ln, src, st, wf, wfi in sorted(data, key=lambda x: -x[0])[:top]:
    tops = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100 * s / ts:5.1f}% samp {100 * i / ti:5.1f}% inst {f}:{ln:>4s} smem_wf {wf}/{wfi} | {src[:90]} | {tops}")
