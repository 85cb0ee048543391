#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel of an .ncu-rep (captured with --import-source on):
python tools/ncu_lines.py report.ncu-rep kernel_name [min_share_percent]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
cur, agg, tot, tsamp = None, {}, 0, 0
for r in csv.reader(out.splitlines()):
    if len(r) < 8:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        continue
    if r[0] in ("Line No", ""):
        continue
    try:
        ln, inst, samp = int(r[0]), int(r[7]), int(r[4])
    except ValueError:
        continue
    a = agg.setdefault((cur, ln), [r[1], 0, 0])
    a[1] += inst
    a[2] += samp
    tot += inst
    tsamp += samp
print("warp instructions", tot, "samples", tsamp)
for k, (src, inst, samp) in sorted(agg.items()):
    if inst > tot * min_share / 100 or samp > tsamp * min_share / 100:
        print("%-13s %4d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0][:13], k[1], 100 * inst / tot, 100 * samp / max(tsamp, 1), src.strip()[:105]))
