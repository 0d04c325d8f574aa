#!/usr/bin/env python
"""profiles/traffic.json from one ncu --set full capture of ALL launches of one bench step:
dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum per kernel name, summed over the launches of that
kernel in the step (k_seed = fast + slow 0 + slow 1, k_chain = 3 launches).  usage: ncu_traffic.py x.ncu-rep out.json reads_per_step [config key of bench.py: workload|index|reads]"""
import csv, io, json, subprocess, sys

rep, out, rps = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def to_ms(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(u, 1))


res = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    d = res.setdefault(name, {"dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "dram_bytes": 0.0, "ncu_duration_ms": 0.0, "launches": 0})
    rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    d["dram_bytes_read"] += rd; d["dram_bytes_write"] += wr; d["dram_bytes"] += rd + wr
    d["ncu_duration_ms"] += to_ms(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]]); d["launches"] += 1
res["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum per kernel, summed over the launches of one bench step "
                f"(ncu --set full --clock-control none, {rep.split('/')[-1]}), bench.py default workload")
res["_reads_per_step"] = rps
res["config"] = sys.argv[4] if len(sys.argv) > 4 else f"long|viral-gs|{rps}"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
