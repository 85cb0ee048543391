#!/usr/bin/env python
"""One markdown table row block per kernel of an .ncu-rep (`ncu --set full`): duration, DRAM bytes, throughputs, occupancy,
registers, issue statistics.  python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/summary.md"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_active", "executed IPC (per SM, active cycles)"),
    ("smsp__issue_active.avg.pct", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps per scheduler"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("launch__registers_per_thread", "registers per thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory per CTA"),
    ("launch__shared_mem_per_block_static", "static shared memory per CTA"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "CTA limit: registers"),
    ("launch__occupancy_limit_shared_mem", "CTA limit: shared memory"),
    ("launch__occupancy_limit_warps", "CTA limit: warps"),
]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("## %s: no kernels" % rep)
        continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("## %s\n" % rep.split("/")[-1])
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        print("### `%s`\n" % name)
        print("| metric | value |\n|---|---|")
        for m, label in METRICS:
            if m in col:
                v, u = r[col[m]], units[col[m]]
                print("| %s | %s %s |" % (label, v, u))
        print()
