#!/usr/bin/env python
"""Key raw metrics of EVERY launch of an ncu capture, one block per launch (profiles/*_kernels_summary.txt).
usage: ncu_kernels.py x.ncu-rep out.txt"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write(f"# {rep}: {len(rows) - 2} launches (ncu --set full --clock-control none; durations are cold-cache and serialised)\n")
    for k, r in enumerate(rows[2:]):
        if len(r) != len(hdr): continue
        f.write(f"\n## launch {k}: {r[hdr.index('Kernel Name')].split('(')[0]}\n")
        for i, h in enumerate(hdr):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i].replace(',', '') or 0) >= 0.2):
                f.write(f"{h} = {r[i]} {units[i]}\n")
print(open(out).read()[:1500])
