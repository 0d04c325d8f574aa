"""CPU tests: the oracle (oracle/desamba_oracle.c, a restatement of the reference's classify path) against golden outputs
of the UNMODIFIED reference (tests/golden/, produced by oracle/make_golden.sh from oracle/_ref/deSAMBA_zero -t 1 and
deSAMBA_stock -t 4), plus its primitives against independent Python restatements."""
import gzip
import hashlib
import os
import subprocess

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETS = {"long10": ("long", 300, 0.10, 20261020), "long30": ("long", 300, 0.30, 20261021),
        "short1": ("short", 5000, 0.01, 20261022), "mixed": ("mixed", (150, 1500), 0, 20261024)}


def _orc_cli(ob, args):
    exe = os.path.join(ob.ORACLE_DIR, "orc_classify")
    return subprocess.run([exe] + args, check=True, capture_output=True).stdout


def _md5_table(name):
    return dict((l.split()[1], l.split()[0]) for l in open(os.path.join(GOLD, name)) if len(l.split()) == 2)


@pytest.mark.parametrize("fmt", ["SAM", "SAM_FULL", "DES", "DES_FULL"])
def test_demo_matches_reference_md5(ob, demo_index, fmt):
    # config #1 of BASELINE.json: the reference's own demo; md5s of `deSAMBA classify -t 4 -f <fmt>`
    out = _orc_cli(ob, ["-f", fmt, demo_index, ob.DEMO_FQ])
    assert hashlib.md5(out).hexdigest() == _md5_table("demo.md5")["demo." + fmt]


def test_demo_sam_text(ob, demo_index):
    out = _orc_cli(ob, ["-f", "SAM", demo_index, ob.DEMO_FQ])
    assert out == gzip.open(os.path.join(GOLD, "demo.SAM.gz")).read()


def test_simreads_is_deterministic(ob, demo_index):
    want = _md5_table("inputs.md5")
    for name, spec in SETS.items():
        path = ob.sim_set(name, *spec)
        assert hashlib.md5(open(path, "rb").read()).hexdigest() == want[f"sets/{name}.fq"], name


@pytest.mark.parametrize("name", list(SETS))
@pytest.mark.parametrize("fmt", ["SAM", "DES_FULL"])
def test_synthetic_sets_match_reference(ob, demo_index, name, fmt):
    path = ob.sim_set(name, *SETS[name])
    out = _orc_cli(ob, ["-f", fmt, demo_index, path])
    assert out == gzip.open(os.path.join(GOLD, f"{name}.{fmt}.gz")).read()


@pytest.mark.parametrize("name,spec", [("syn_long10", ("long", 400, 0.10, 20261025)), ("syn_short1", ("short", 4000, 0.01, 20261026))])
def test_second_index_multi_strain(ob, name, spec):
    # synthetic multi-strain reference: unitigs with several reference positions, thousands of secondary hits
    try:
        idx = ob.ensure_syn_index()
    except FileNotFoundError as e:
        pytest.skip(str(e))
    path = ob.sim_set(name, *spec, fasta=ob.SYN_FA)
    out = _orc_cli(ob, ["-f", "DES_FULL", idx, path])
    gold = gzip.open(os.path.join(GOLD, f"{name}.DES_FULL.gz")).read()
    assert out == gold and gold.count(b" SEC ") > 1000


@pytest.mark.parametrize("l_ek,name", [(17, "syn_long10"), (17, "syn_short1"), (18, "syn_short1")])
def test_larger_exist_kmer_table_classes(ob, l_ek, name):
    # l_ek = 17 / 31-bit hash mask and l_ek = 18 / 33-bit mask (set_ekmer_par, idx.c:966-982): the reference picks them for
    # references of 0.3 / 1.3 Gbp and more; here the small multi-strain index carries tables of that class (oracle/rebuild_exk.c)
    try:
        idx = ob.ensure_ek_index(l_ek)
    except FileNotFoundError as e:
        pytest.skip(str(e))
    spec = ("long", 400, 0.10, 20261025) if name == "syn_long10" else ("short", 4000, 0.01, 20261026)
    path = ob.sim_set(name, *spec, fasta=ob.SYN_FA)
    out = _orc_cli(ob, ["-f", "DES_FULL", idx, path])
    gold = gzip.open(os.path.join(GOLD, f"{name}.ek{l_ek}.DES_FULL.gz")).read()
    assert out == gold
    assert gold != gzip.open(os.path.join(GOLD, f"{name}.DES_FULL.gz")).read()      # the class changes the result: the path is exercised
    if l_ek == 18:
        ob.drop_ek_index(18)                                                        # 2 GiB of shared memory


def test_rebuilt_tables_of_the_original_class_are_the_builders(ob, tmp_path):
    # the table rule of oracle/rebuild_exk.c is the index builder's: for the class the builder chose it reproduces its files
    idx = ob.ensure_syn_index()
    exe = os.path.join(ob.ORACLE_DIR, "rebuild_exk")
    subprocess.run(["gcc", "-O2", "-w", "-o", exe, exe + ".c"], check=True)
    subprocess.run([exe, idx, str(tmp_path / "c27"), "27"], check=True, capture_output=True)
    for f in ("deSAMBA.exk0", "deSAMBA.exk1", "deSAMBA.exki"):
        assert open(tmp_path / "c27" / f, "rb").read() == open(os.path.join(idx, f), "rb").read(), f


def test_options_l_s_r(ob, demo_index):
    path = ob.sim_set("long10", *SETS["long10"])
    out = _orc_cli(ob, ["-l", "100", "-s", "40", "-r", "2", demo_index, path])
    assert out == gzip.open(os.path.join(GOLD, "long10.l100s40r2.SAM.gz")).read()


def test_capi_equals_cli(ob, oracle, demo_index):
    # the flat-array API used by the GPU parity tests yields the same records the CLI prints
    names, seqs, _ = ob.read_fastq(ob.DEMO_FQ, 200)
    rr, hits, mx = oracle.classify(*ob.pack(seqs))
    gold = gzip.open(os.path.join(GOLD, "demo.DES_FULL.gz")).read().decode().split("\n\n")
    for i in range(200):
        head = gold[i].split("\n")[0].split("\t")
        assert head[0] == names[i].decode()
        assert head[1] == ("CLASSIFY" if rr["n_hit"][i] else "UNCLASSIFY")
        assert head[2] == ("FAST" if rr["fast_classify"][i] else "SLOW")
        assert head[4] == f"n_rst:[{rr['n_hit'][i]}]" and head[5] == f"n_anc:[{rr['n_anchor'][i]}]"
    assert mx == max(len(s) for s, n in zip(seqs, rr["entered_final"]) if n)


# ---- primitives
def _py_hash64_1(k):
    M = (1 << 64) - 1
    k = (~k + (k << 21)) & M; k ^= k >> 24; k = (k + (k << 3) + (k << 8)) & M; k ^= k >> 14
    k = (k + (k << 2) + (k << 4)) & M; k ^= k >> 28; k = (k + (k << 31)) & M
    return k


def _py_hash64_2(k):
    M = (1 << 64) - 1
    k = (k + (~(k << 32) & M)) & M; k ^= k >> 22; k = (k + (~(k << 13) & M)) & M; k ^= k >> 8
    k = (k + (k << 3)) & M; k ^= k >> 15; k = (k + (~(k << 27) & M)) & M; k ^= k >> 31
    return k


def test_hashes(ob):
    import ctypes as C
    L = ob.lib()
    L.orc_hash64_1.restype = L.orc_hash64_2.restype = C.c_uint64
    L.orc_hash64_1.argtypes = L.orc_hash64_2.argtypes = [C.c_uint64]
    rng = np.random.default_rng(1)
    for k in [0, 1, 0xFFFFFFFF, (1 << 40) - 1] + [int(x) for x in rng.integers(0, 1 << 40, 200)]:
        assert L.orc_hash64_1(k) == _py_hash64_1(k)
        assert L.orc_hash64_2(k) == _py_hash64_2(k)


def test_msort_is_glibc_merge_order(ob):
    # glibc 2.39 qsort = top-down merge sort (n1 = n/2, take left while cmp <= 0); checked against the
    # libc of this machine with the reference's asymmetric comparator (ties: a.sum_score % 2)
    import ctypes as C
    L = ob.lib()
    libc = C.CDLL(None)
    CMP = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32))

    def cmp(a, b):
        sa, sb = a[0] >> 3, b[0] >> 3
        if sa < sb: return 1
        if sa > sb: return -1
        return a[0] % 2
    cb = CMP(cmp)
    rng = np.random.default_rng(7)
    for n in [2, 3, 5, 8, 31, 64, 100]:
        for _ in range(20):
            v = rng.integers(0, 64, n).astype(np.int32)
            a, b = v.copy(), v.copy()
            libc.qsort(a.ctypes.data, n, 4, cb)
            L.orc_msort(b.ctypes.data, C.c_size_t(n), C.c_size_t(4), cb)
            assert a.tolist() == b.tolist()


def test_lv_extd_exact_and_single_errors(ob):
    import ctypes as C
    L = ob.lib()
    L.orc_lv_extd.restype = C.c_int32
    L.orc_lv_extd.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    rng = np.random.default_rng(3)

    def lv(r, q):
        rb = np.zeros(32, dtype=np.uint8); qb = np.zeros(32, dtype=np.uint8)
        rb[:len(r)] = r; qb[:len(q)] = q
        return L.orc_lv_extd(rb.ctypes.data, len(r), qb.ctypes.data, len(q))
    for _ in range(200):
        s = rng.integers(0, 4, 12).astype(np.uint8)
        assert lv(s, s) == 0
        t = s.copy(); p = int(rng.integers(2, 9)); t[p] = (t[p] + 1) % 4
        assert lv(s, t) == 1                     # one substitution in the middle
        assert lv(s[:0], s[:0]) == 0
    # more than 4 errors: returns the flank length (cly.c:521)
    a = np.zeros(12, dtype=np.uint8); b = np.full(12, 3, dtype=np.uint8)
    assert lv(a, b) == 12


def test_exist_kmer_tables(ob, oracle, demo_index):
    # every 16-mer of an indexed genome must "exist" unless masked as low complexity; random 16-mers almost never do
    import ctypes as C
    L = ob.lib()
    L.orc_exist_kmer.restype = C.c_int
    L.orc_exist_kmer.argtypes = [C.c_void_p, C.c_uint64]
    seq = []
    with open(ob.DEMO_FA) as f:
        f.readline()
        for line in f:
            if line.startswith(">") or len(seq) > 3000: break
            seq.extend("ACGT".index(c) if c in "ACGT" else 0 for c in line.strip().upper())
    hits = 0
    for i in range(1000, 2000):
        k = 0
        for c in seq[i:i + 16]: k = (k << 2) | c
        hits += L.orc_exist_kmer(oracle._h, k)
    assert hits == 1000
    rng = np.random.default_rng(5)
    assert sum(L.orc_exist_kmer(oracle._h, int(k)) for k in rng.integers(1, 1 << 32, 2000)) < 20
