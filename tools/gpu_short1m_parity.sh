#!/usr/bin/env bash
# Developer tool (gpurun): 1 M x 150 bp simulated reads -- driver text against `deSAMBA_zero classify -t 1` (the zero-initialised
# build of the unmodified reference), compared read by read.
set -uo pipefail
cd "$(dirname "$0")/.."
python - <<'PY'
import os, sys, subprocess
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_binding as ob
ob.ensure_demo_index()
p = "/dev/shm/dsb_short1m.fq"
if not os.path.exists(p):
    subprocess.run([ob.SIMREADS, "short", ob.DEMO_FA, "1000000", "0.01", "20261031", p], check=True)
PY
IDX=oracle/_ref/demo/idx
oracle/_ref/deSAMBA_zero classify -t 1 -f SAM -o /dev/shm/z1.sam $IDX /dev/shm/dsb_short1m.fq 2> /dev/null &
desamba_b200/bin/deSAMBA-b200 classify -g 1 -f SAM -o /dev/shm/g.sam $IDX /dev/shm/dsb_short1m.fq 2> /dev/null
desamba_b200/bin/deSAMBA-b200 classify -g 1 -B 50000 -f SAM -o /dev/shm/g2.sam $IDX /dev/shm/dsb_short1m.fq 2> /dev/null
wait
python - <<'PY'
import collections
def blocks(path):
    d = collections.OrderedDict()
    for l in open(path, "rb"):
        d.setdefault(l.split(b"\t", 1)[0], []).append(l)
    return d
z = blocks("/dev/shm/z1.sam")
for name in ("g", "g2"):
    g = blocks(f"/dev/shm/{name}.sam")
    bad = [k for k in z if g.get(k) != z[k]]
    print(f"{name}: reads {len(z)} / {len(g)}, same order {list(z) == list(g)}, reads whose lines differ: {len(bad)}")
    for k in bad[:4]:
        print("  ref:", b" | ".join(x.strip()[:160] for x in z[k]).decode())
        print("  gpu:", b" | ".join(x.strip()[:160] for x in g.get(k, [])).decode())
PY
rm -f /dev/shm/z1.sam /dev/shm/g.sam /dev/shm/g2.sam
