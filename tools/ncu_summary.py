#!/usr/bin/env python
"""Summarise ncu captures for profiles/: key raw metrics of one .ncu-rep (and optionally the hottest source lines).
usage: tools/ncu_summary.py gpurun_out/<x>.ncu-rep profiles/<x>_raw.txt [--source]"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# {rep}: kernel {vals[hdr.index('Kernel Name')]}"]
    for i, h in enumerate(hdr):
        if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            lines.append(f"{h} = {vals[i]} {units[i]}")
    if "--source" in sys.argv:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        if srows:
            sh = srows[0]
            def col(name):
                return next((i for i, h in enumerate(sh) if h.strip() == name), None)
            c_src, c_samp, c_inst = col("Source"), col("# Samples"), col("Instructions Executed")
            if c_src is not None and c_samp is not None:
                body = []
                for r in srows[1:]:
                    try:
                        body.append((int(float(r[c_samp] or 0)), r[c_src].strip()[:140]))
                    except (ValueError, IndexError):
                        pass
                tot = sum(b[0] for b in body) or 1
                lines.append("# hottest source lines by stall samples")
                for n, t in sorted(body, reverse=True)[:25]:
                    lines.append(f"{100.0 * n / tot:5.1f}%  {t}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
