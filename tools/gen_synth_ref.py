#!/usr/bin/env python
"""Synthetic multi-genome reference of BASELINE.json configs[3] (SURVEY.md 8d cfg 4), scalable:
n_species "species" x n_strains "strains" x genome_bp, strains derived from a random ancestor at 1-3 % divergence (so that
unitigs have several REF_POS and secondaries exist), headers `>tid|<taxid>|ref|SYN_<i>.1`, 2 kb decoy contigs first and last.
usage: gen_synth_ref.py out.fa [n_species=15] [n_strains=4] [genome_bp=5000000] [seed=20261023]"""
import sys
import numpy as np

out = sys.argv[1]
n_species = int(sys.argv[2]) if len(sys.argv) > 2 else 15
n_strains = int(sys.argv[3]) if len(sys.argv) > 3 else 4
genome_bp = int(sys.argv[4]) if len(sys.argv) > 4 else 5_000_000
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 20261023
rng = np.random.default_rng(seed)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def write(f, name, codes):
    f.write(b">" + name.encode() + b"\n")
    s = ACGT[codes].tobytes()
    for i in range(0, len(s), 80):
        f.write(s[i:i + 80]); f.write(b"\n")


def mutate(anc, div):
    g = anc.copy()
    n_sub = int(len(g) * div * 0.9)
    pos = rng.integers(0, len(g), n_sub)
    g[pos] = (g[pos] + rng.integers(1, 4, n_sub)) % 4
    n_indel = int(len(g) * div * 0.1)
    keep = np.ones(len(g), dtype=bool)
    keep[rng.integers(0, len(g), n_indel // 2)] = False          # deletions
    g = g[keep]
    ins_pos = np.sort(rng.integers(0, len(g), n_indel // 2))
    g = np.insert(g, ins_pos, rng.integers(0, 4, len(ins_pos)).astype(np.uint8))
    return g


with open(out, "wb") as f:
    write(f, "tid|1|ref|DECOY_0.1", rng.integers(0, 4, 2000).astype(np.uint8))
    k = 0
    for sp in range(n_species):
        anc = rng.integers(0, 4, genome_bp).astype(np.uint8)
        for st in range(n_strains):
            g = anc if st == 0 else mutate(anc, rng.uniform(0.01, 0.03))
            write(f, f"tid|{1000 + sp * 10 + st}|ref|SYN_{k}.1", g)
            k += 1
    write(f, "tid|2|ref|DECOY_1.1", rng.integers(0, 4, 2000).astype(np.uint8))
print(f"{out}: {n_species} x {n_strains} x {genome_bp} bp")
