"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI with HOST buffers, against the
oracle on the same inputs (bit-exact: all arithmetic on the path is integer), against the golden text of the unmodified
reference through the `deSAMBA-b200 classify` driver, on edge cases, and through size-independent properties."""
import gzip
import hashlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
DRIVER = os.path.join(ROOT, "desamba_b200", "bin", "deSAMBA-b200")
SETS = {"demo": None, "long10": ("long", 300, 0.10, 20261020), "long30": ("long", 300, 0.30, 20261021),
        "short1": ("short", 5000, 0.01, 20261022), "mixed": ("mixed", (150, 1500), 0, 20261024)}
CNT = ("n_bit0", "n_bit1", "n_prefix", "n_occ", "n_locate", "n_getref", "n_getref_bytes")


def _set_path(ob, name):
    return ob.DEMO_FQ if name == "demo" else ob.sim_set(name, *SETS[name])


def _assert_same(ob, res, rr_o, hits_o, names=None):
    bad = ob.compare_results(res.rr, res.hits, rr_o, hits_o, names, max_report=5)
    assert not bad, "\n".join(bad)


def test_native_library_is_loaded(gpu):
    dsb, ix, ctx = gpu
    assert any("libdesamba_b200.so" in l for l in open("/proc/self/maps"))
    assert ix.hbm_bytes > 500e6 and ix.l_ek == 16


@pytest.mark.parametrize("name", list(SETS))
def test_set_parity_against_oracle(gpu, ob, oracle, name):
    dsb, ix, ctx = gpu
    names, seqs, _ = ob.read_fastq(_set_path(ob, name))
    cat, offs = ob.pack(seqs)
    oracle.counters(reset=True)
    rr_o, hits_o, mx_o = oracle.classify(cat, offs)
    cnt_o = oracle.counters()
    res = ctx.classify(cat, offs)
    _assert_same(ob, res, rr_o, hits_o, names)
    assert res.max_read_l == mx_o
    # kernel-level parity: island seeds of both strands, and the algorithmic counters of every stage
    for i in range(0, len(seqs), max(1, len(seqs) // 100)):
        for s in (0, 1):
            sg, tg = ctx.seeds(i, s)
            so, to = oracle.seeds(seqs[i], s)
            assert tg == to and sg.tobytes() == so.tobytes(), (name, i, s)
    cnt_g = ctx.counters()
    assert {k: cnt_g[k] for k in CNT} == {k: cnt_o[k] for k in CNT}
    assert ctx.launches() == 11      # probe, islands, 3 x (seed, chain), score + score_heavy, finalize


def _run_driver(args):
    r = subprocess.run([DRIVER, "classify"] + args, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    return r.stdout


@pytest.mark.parametrize("fmt", ["SAM", "SAM_FULL", "DES", "DES_FULL"])
def test_driver_demo_md5(gpu, ob, demo_index, fmt):
    # config #1: identical text to `deSAMBA classify -t 4 -f <fmt>` of the unmodified reference
    want = dict((l.split()[1], l.split()[0]) for l in open(os.path.join(GOLD, "demo.md5")) if len(l.split()) == 2)
    out = _run_driver(["-t", "4", "-f", fmt, demo_index, ob.DEMO_FQ])
    assert hashlib.md5(out).hexdigest() == want["demo." + fmt]


@pytest.mark.parametrize("name", ["long10", "long30", "short1", "mixed"])
@pytest.mark.parametrize("fmt,batch", [("SAM", "65536"), ("DES_FULL", "97")])
def test_driver_sets_text(gpu, ob, demo_index, name, fmt, batch):
    # small batches (-B 97) exercise the carry of max_read_l between batches (mixed set: short reads after long ones)
    out = _run_driver(["-f", fmt, "-B", batch, demo_index, _set_path(ob, name)])
    assert out == gzip.open(os.path.join(GOLD, f"{name}.{fmt}.gz")).read()


@pytest.mark.parametrize("extra", [["-P", "0"], ["-P", "3", "-B", "50"], ["-f", "SAM_FULL", "-P", "5"]])
def test_driver_reader_modes(gpu, ob, demo_index, tmp_path, extra):
    # serial FASTQ reader (-P 0) and the parallel indexer with other thread counts / batch sizes: same text; SAM_FULL prints
    # the bases and qualities the reader handed on; a file that is not 4-line FASTQ half way falls back to the serial reader
    path = _set_path(ob, "mixed")
    want = _run_driver(["-f", "SAM_FULL" if "SAM_FULL" in extra else "SAM", "-P", "0", demo_index, path])
    if "SAM_FULL" not in extra:
        assert want == gzip.open(os.path.join(GOLD, "mixed.SAM.gz")).read()
    assert _run_driver(extra + [demo_index, path]) == want
    if extra == ["-P", "0"]:
        txt = open(path, "rb").read()
        recs = txt.split(b"\n@")
        k = len(recs) // 2
        seq = recs[k].split(b"\n")                       # wrap one record's sequence and quality over two lines each
        seq[1] = seq[1][:7] + b"\n" + seq[1][7:]; seq[3] = seq[3][:7] + b"\n" + seq[3][7:]
        recs[k] = b"\n".join(seq)
        wrapped = tmp_path / "wrapped.fq"
        wrapped.write_bytes(b"\n@".join(recs))
        assert _run_driver(["-P", "4", demo_index, str(wrapped)]) == want


def test_driver_reader_blocks_and_read_ahead(gpu, ob, demo_index):
    # the parallel FASTQ reader works on blocks of the file (256 MB; 1 MB here) and reads the next block ahead while the records
    # of the current one go into batches: same text as with the serial reader
    path = _set_path(ob, "mixed")
    want = _run_driver(["-f", "DES_FULL", "-P", "0", demo_index, path, path])          # the serial reader on the same two files
    assert want.startswith(gzip.open(os.path.join(GOLD, "mixed.DES_FULL.gz")).read())
    for extra in (["-P", "4"], ["-P", "7", "-B", "333"]):
        r = subprocess.run([DRIVER, "classify", "-f", "DES_FULL"] + extra + [demo_index, path, path], capture_output=True, env=dict(os.environ, DSB_FQ_BLOCK_MB="1"))
        assert r.returncode == 0, r.stderr.decode()[-1000:]
        assert r.stdout == want


def test_driver_options_and_gz_input(gpu, ob, demo_index, tmp_path):
    path = _set_path(ob, "long10")
    gz = tmp_path / "long10.fq.gz"
    with open(path, "rb") as f, gzip.open(gz, "wb", compresslevel=1) as g:
        g.write(f.read())
    out = _run_driver(["-l", "100", "-s", "40", "-r", "2", "-o", str(tmp_path / "o.sam"), demo_index, str(gz)])
    assert out == b""
    assert open(tmp_path / "o.sam", "rb").read() == gzip.open(os.path.join(GOLD, "long10.l100s40r2.SAM.gz")).read()


@pytest.mark.parametrize("name,spec", [("syn_long10", ("long", 400, 0.10, 20261025)), ("syn_short1", ("short", 4000, 0.01, 20261026))])
def test_second_index_multi_strain(gpu, ob, name, spec):
    # another index (synthetic, 4 strains per species: several REF_POS per unitig, many secondaries): records against the
    # oracle, driver text against the golden output of the unmodified reference
    dsb, _, _ = gpu
    idx = ob.ensure_syn_index()
    path = ob.sim_set(name, *spec, fasta=ob.SYN_FA)
    names, seqs, _ = ob.read_fastq(path)
    cat, offs = ob.pack(seqs)
    orc = ob.Oracle(idx)
    orc.counters(reset=True)
    rr_o, hits_o, mx_o = orc.classify(cat, offs)
    cnt_o = orc.counters()
    ix2 = dsb.Index(idx, 0)
    ctx2 = dsb.Context(ix2)
    try:
        res = ctx2.classify(cat, offs)
        _assert_same(ob, res, rr_o, hits_o, names)
        cnt_g = ctx2.counters()
        assert {k: cnt_g[k] for k in CNT} == {k: cnt_o[k] for k in CNT}
        assert int((res.hits["primary"] == 2).sum()) > 500           # secondaries exist on this index
    finally:
        ctx2.close(); ix2.close(); orc.close()
    out = _run_driver(["-f", "DES_FULL", "-B", "1000", idx, path])
    assert out == gzip.open(os.path.join(GOLD, f"{name}.DES_FULL.gz")).read()


@pytest.mark.parametrize("l_ek", [17, 18, 19, 20])
def test_larger_exist_kmer_table_classes(gpu, ob, l_ek):
    # l_ek = 17..20 with hash masks of 31..37 bits (set_ekmer_par, idx.c:966-982; the classes a 0.3 .. 40 Gbp reference gets):
    # tables of 2 x 256 MiB .. 2 x 16 GiB (no L2 summary above 256 MiB), single_base_max 13..16.  Records and every algorithmic
    # counter against the oracle, driver text against the unmodified reference's output on the same index files.
    dsb, _, _ = gpu
    try:
        idx = ob.ensure_ek_index(l_ek)
    except FileNotFoundError as e:
        pytest.skip(str(e))
    try:
        orc = ob.Oracle(idx)
        ix2 = dsb.Index(idx, 0)
        ctx2 = dsb.Context(ix2)
        assert ix2.l_ek == l_ek
        try:
            for name, spec in (("syn_long10", ("long", 400, 0.10, 20261025)), ("syn_short1", ("short", 4000, 0.01, 20261026))):
                path = ob.sim_set(name, *spec, fasta=ob.SYN_FA)
                names, seqs, _ = ob.read_fastq(path)
                cat, offs = ob.pack(seqs)
                orc.counters(reset=True)
                rr_o, hits_o, mx_o = orc.classify(cat, offs)
                cnt_o = orc.counters()
                res = ctx2.classify(cat, offs)
                _assert_same(ob, res, rr_o, hits_o, names)
                cnt_g = ctx2.counters()
                assert {k: cnt_g[k] for k in CNT} == {k: cnt_o[k] for k in CNT}
                gold = os.path.join(GOLD, f"{name}.ek{l_ek}.DES_FULL.gz")
                if os.path.exists(gold):
                    assert _run_driver(["-f", "DES_FULL", "-B", "1000", idx, path]) == gzip.open(gold).read()
        finally:
            ctx2.close(); ix2.close(); orc.close()
    finally:
        if l_ek >= 18:
            ob.drop_ek_index(l_ek)


def test_edge_cases(gpu, ob, oracle):
    dsb, ix, ctx = gpu
    _, demo, _ = ob.read_fastq(ob.DEMO_FQ, 40)
    long_read = b"".join(demo)[:5000]
    rng = np.random.default_rng(11)
    rnd = lambda n: bytes(rng.choice(list(b"ACGT"), n).tolist())
    reads = [b"", b"A", rnd(39), rnd(40), rnd(41), b"N" * 200, b"A" * 300, b"ACGT" * 100, demo[0].lower(), demo[1][:55],
             long_read[:1023], long_read[:1024], long_read[:1025], long_read[:2048], long_read[:2049], long_read[:1039], long_read[:1040],
             demo[2].replace(b"A", b"N"), demo[3][::-1], rnd(3000)] + demo[4:12]
    cat, offs = ob.pack(reads)
    rr_o, hits_o, mx_o = oracle.classify(cat, offs)
    res = ctx.classify(cat, offs)
    _assert_same(ob, res, rr_o, hits_o)
    assert res.max_read_l == mx_o
    assert int(res.rr["n_hit"][0]) == 0 and int(res.rr["fast_classify"][0]) == 1      # < 40 bp: untouched result (cly.c:3089)
    # empty batch and a single read
    e = ctx.classify(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    assert len(e.rr) == 0 and len(e.hits) == 0
    one = ctx.classify(*ob.pack([demo[0]]))
    r1, h1, _ = oracle.classify(*ob.pack([demo[0]]))
    _assert_same(ob, one, r1, h1)


def test_very_long_reads_many_anchors(gpu, ob, oracle):
    # reads of 150-400 kb built from many simulated reads: thousands of anchors (merge-sort path of the anchor sort, many
    # chain groups, 18-bit 9-mer index) -- the reference handles these through realloc'd vectors
    dsb, ix, ctx = gpu
    _, seqs, _ = ob.read_fastq(_set_path(ob, "long10"), 120)
    reads = [b"".join(seqs[0:20]), b"".join(seqs[20:70]), b"".join(seqs[70:120])[:400000], seqs[3]]
    cat, offs = ob.pack(reads)
    rr_o, hits_o, mx_o = oracle.classify(cat, offs)
    assert int(rr_o["n_anchor"].max()) > 2500
    res = ctx.classify(cat, offs)
    _assert_same(ob, res, rr_o, hits_o)
    assert res.max_read_l == mx_o


def test_repeat_rich_reads_heavy_path(gpu, ob, oracle):
    # tandem copies of a read segment: every scanned 9-mer of a reference window matches many read positions, so sdp_match
    # collects thousands of candidates (candidate flushes, merge-sort ordering of > 256 matches) and the read is deferred to
    # the CTA-per-read kernel (> 1024 matches in one extension); plus low-complexity reads that yield no seeds at all
    dsb, ix, ctx = gpu
    _, seqs, _ = ob.read_fastq(_set_path(ob, "long10"), 12)
    long_ = [s for s in seqs if len(s) > 3000]
    reads = [long_[0][200:500] * 25, long_[1][100:160] * 120, long_[2][300:1300] * 6, long_[0][:2000] + long_[0][200:500] * 15 + long_[0][2000:3000],
             b"AC" * 3000, b"ACG" * 2000, b"A" * 5000 + long_[1][:1500], long_[3]]
    # reads from the (TAACCC)n telomeric repeats of HHV-7 (NC_001716.2: ~143-146 kb and the terminal repeat at 1-3 kb): repeats in
    # the reference AND the read -- the reads the bench workload defers to k_score_heavy
    c, on = [], False
    for l in open(ob.DEMO_FA):
        if l.startswith(">"):
            if on:
                break
            on = l[1:].split()[0] == "tid|10372|ref|NC_001716.2"
        elif on:
            c.append(l.strip().upper())
    c = "".join(c).encode()
    assert len(c) == 153080
    rng = np.random.default_rng(5)

    def mutate(seq, rate):
        a = np.frombuffer(seq, dtype=np.uint8).copy()
        m = rng.random(len(a)) < rate
        a[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(m.sum()))
        return a.tobytes()

    rnd = lambda n: bytes(rng.choice(list(b"ACGT"), n).tolist())
    reads += [rnd(80) + mutate(c[142500:147500], 0.05) + rnd(80), rnd(80) + mutate(c[800:3800], 0.08) + rnd(80), rnd(80) + mutate(c[143000:146000], 0.02) + rnd(80)]
    cat, offs = ob.pack(reads)
    rr_o, hits_o, mx_o = oracle.classify(cat, offs)
    res = ctx.classify(cat, offs)
    _assert_same(ob, res, rr_o, hits_o)
    assert res.max_read_l == mx_o
    assert int(rr_o["n_hit"].sum()) > 0
    ms = ctx.kernel_ms()
    assert ms[9] > 0.05, ms                     # k_score_heavy had work (its empty launch takes a few microseconds)


def test_max_read_l_state(gpu, ob, oracle):
    # Classify_buff_pool.max_read_l (cly.c:2958): short reads are filtered differently once a read >= 510 bp has reached the
    # filter -- inside a batch (input order) and across batches (max_read_l_in / max_read_l_out)
    dsb, ix, ctx = gpu
    _, short, _ = ob.read_fastq(_set_path(ob, "short1"), 300)
    _, long_, _ = ob.read_fastq(_set_path(ob, "long10"), 3)
    reads = short[:100] + long_[:1] + short[100:200] + long_[1:] + short[200:]
    cat, offs = ob.pack(reads)
    for mx_in in (0, 509, 510, 30000):
        rr_o, hits_o, mx_o = oracle.classify(cat, offs, mx_in)
        res = ctx.classify(cat, offs, mx_in)
        _assert_same(ob, res, rr_o, hits_o)
        assert res.max_read_l == mx_o
    # chained batches == one batch
    whole_rr, whole_hits, _ = oracle.classify(cat, offs, 0)
    mx, cap, got = 0, 0, []
    for lo in range(0, len(reads), 64):
        res = ctx.classify(*ob.pack(reads[lo:lo + 64]), mx, cap)
        mx, cap = res.max_read_l, ctx.bin_capacity()
        got += [res.read_hits(i).tobytes() for i in range(len(res.rr))]
    want = [whole_hits[int(o):int(o) + int(n)].tobytes() for o, n in zip(whole_rr["hit_off"], whole_rr["n_hit"])]
    assert got == want


def test_upload_run_download_equals_classify_batch(gpu, ob):
    dsb, ix, ctx = gpu
    _, seqs, _ = ob.read_fastq(_set_path(ob, "long30"), 120)
    cat, offs = ob.pack(seqs)
    a = ctx.classify(cat, offs)
    ctx.upload(cat, offs)
    for _ in range(2):                        # re-running on inputs resident in HBM is idempotent
        ctx.run(0)
        b = ctx.download()
        assert b.rr.tobytes() != b"" and [b.read_hits(i).tobytes() for i in range(len(seqs))] == [a.read_hits(i).tobytes() for i in range(len(seqs))]
        assert b.rr["n_anchor"].tolist() == a.rr["n_anchor"].tolist()
    ms = ctx.kernel_ms()
    assert len(ms) == 11 and all(m >= 0 for m in ms) and ms[2] > 0 and ms[8] > 0


def test_batch_composition_invariance_large(gpu, ob):
    # property at a larger size (no oracle): per-read results do not depend on batch composition, order or repetition
    dsb, ix, ctx = gpu
    path = ob.sim_set("long10_4k", "long", 4000, 0.10, 20261030)
    _, seqs, _ = ob.read_fastq(path)
    cat, offs = ob.pack(seqs)
    a = ctx.classify(cat, offs, 10**6)
    ha = [a.read_hits(i).tobytes() for i in range(len(seqs))]
    assert sum(int(x) for x in a.rr["n_hit"]) > 3000 and not a.rr["error"].any()
    perm = np.random.default_rng(2).permutation(len(seqs))
    b = ctx.classify(*ob.pack([seqs[i] for i in perm]), 10**6)
    assert [b.read_hits(k).tobytes() for k in range(len(seqs))] == [ha[i] for i in perm]
    assert b.rr["n_anchor"].tolist() == a.rr["n_anchor"][perm].tolist()
    c = ctx.classify(*ob.pack(seqs[1000:1500]), 10**6)
    assert [c.read_hits(k).tobytes() for k in range(500)] == ha[1000:1500]


def test_sample_of_large_batch_against_oracle(gpu, ob, oracle):
    dsb, ix, ctx = gpu
    path = ob.sim_set("short1_50k", "short", 50000, 0.01, 20261031)
    names, seqs, _ = ob.read_fastq(path)
    res = ctx.classify(*ob.pack(seqs))
    idx = list(range(len(seqs)))                                     # every read (a few seconds of the oracle)
    rr_o, hits_o, _ = oracle.classify(*ob.pack([seqs[i] for i in idx]))
    for k, i in enumerate(idx):
        assert res.read_hits(i).tobytes() == hits_o[int(rr_o["hit_off"][k]):int(rr_o["hit_off"][k]) + int(rr_o["n_hit"][k])].tobytes(), i
        assert int(res.rr["n_anchor"][i]) == int(rr_o["n_anchor"][k])


def test_capacity_is_reported_not_fatal(gpu, ob, oracle, demo_index, tmp_path):
    # a read built from 50 simulated reads has > 2500 anchors (test_very_long_reads_many_anchors): with max_anchors = 1024 it
    # must come back with dsb_read_result.error = 1 and DSB_E_CAPACITY, its neighbours of the batch untouched; the driver
    # (-A 1024) writes it as unclassified, warns, and finishes the run
    dsb, ix, _ = gpu
    _, seqs, _ = ob.read_fastq(_set_path(ob, "long10"), 120)
    reads = [seqs[0], b"".join(seqs[20:70]), seqs[1], seqs[2]]
    cat, offs = ob.pack(reads)
    rr_o, hits_o, _ = oracle.classify(cat, offs)
    assert int(rr_o["n_anchor"][1]) > 2500
    small = dsb.Context(ix, max_anchors=1024)
    try:
        with pytest.raises(dsb.DsbError) as ei:
            small.classify(cat, offs)
        e = ei.value
        assert e.code == -5 and "capacity" in str(e)
        res = e.result
        assert res.rr["error"].tolist() == [0, 1, 0, 0] and int(res.rr["n_hit"][1]) == 0
        for i in (0, 2, 3):
            assert res.read_hits(i).tobytes() == hits_o[int(rr_o["hit_off"][i]):int(rr_o["hit_off"][i]) + int(rr_o["n_hit"][i])].tobytes()
    finally:
        small.close()
    fq = tmp_path / "cap.fq"
    fq.write_bytes(b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads)))
    r = subprocess.run([DRIVER, "classify", "-A", "1024", "-f", "DES", demo_index, str(fq)], capture_output=True)
    assert r.returncode == 0 and b"exceeded a per-read capacity" in r.stderr
    heads = [l.split(b"\t") for l in r.stdout.split(b"\n") if l.startswith(b"r")]
    assert [h[1] for h in heads] == [b"CLASSIFY" if int(rr_o["n_hit"][i]) and i != 1 else b"UNCLASSIFY" for i in range(4)]


def test_pool_overflow_grows_and_reruns(gpu, ob, oracle):
    # the per-batch device pools (seed tasks, staging chunks, anchors, chains, hits) are sized from the batch; a batch richer than
    # the sizing assumes must not fail: the pools are doubled and the batch is run again inside the call (dsb_batch_retries)
    dsb, ix, _ = gpu
    _, seqs, _ = ob.read_fastq(_set_path(ob, "long10"), 300)
    cat, offs = ob.pack(seqs)
    rr_o, hits_o, _ = oracle.classify(cat, offs)
    tiny = dsb.Context(ix, pool_scale_pct=3)
    try:
        res = tiny.classify(cat, offs)
        assert tiny.retries() >= 1
        _assert_same(ob, res, rr_o, hits_o)
        res = tiny.classify(cat, offs)                   # the grown pools are kept
        assert tiny.retries() == 0
        _assert_same(ob, res, rr_o, hits_o)
    finally:
        tiny.close()


def test_driver_two_gpus_small_batches(gpu, ob, demo_index):
    # -g 2 (index loaded once and copied device to device, batches dealt to 2 x 3 contexts, records merged in input order,
    # max_read_l carried across GPUs): the mixed set in batches of 97 reads against the unmodified reference's text
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    for fmt in ("SAM", "DES_FULL"):
        out = _run_driver(["-g", "2", "-B", "97", "-f", fmt, demo_index, _set_path(ob, "mixed")])
        assert out == gzip.open(os.path.join(GOLD, f"mixed.{fmt}.gz")).read()


def test_index_clone_equals_load(gpu, ob, oracle):
    # dsb_index_clone (the copy the driver gives to GPUs 1..N-1): same results as the loaded index (here onto the same device)
    dsb, ix, _ = gpu
    clone = ix.clone(0)
    ctx2 = dsb.Context(clone)
    try:
        assert clone.l_ek == ix.l_ek and clone.hbm_bytes == ix.hbm_bytes
        _, seqs, _ = ob.read_fastq(_set_path(ob, "long30"), 60)
        cat, offs = ob.pack(seqs)
        rr_o, hits_o, _ = oracle.classify(cat, offs)
        _assert_same(ob, ctx2.classify(cat, offs), rr_o, hits_o)
    finally:
        ctx2.close(); clone.close()


def test_gather_microbenchmark(gpu):
    dsb, _, _ = gpu
    gbs, ms = dsb.gather_bench(0, 2 << 30, 1 << 26, 1)
    assert 100 < gbs < 20000 and ms > 0


def test_next_batch_uploaded_while_the_current_one_runs(gpu, ob):
    """dsb_batch_upload of batch k + 1 before dsb_batch_download of batch k (two input sets per context, copy stream): the
    records equal those of the blocking call, batch after batch, also when a batch is run twice or comes empty"""
    dsb, ix, ctx = gpu
    sets = []
    for name in ("long10", "short1", "long30"):
        _, seqs, _ = ob.read_fastq(_set_path(ob, name))
        cat, offs = ob.pack(seqs[:2000])
        sets.append((np.ascontiguousarray(cat, dtype=np.uint8), np.ascontiguousarray(offs, dtype=np.uint64)))
    want = [ctx.classify(cat, offs, 10**6) for cat, offs in sets]
    c2 = dsb.Context(ix)
    order = [0, 1, 2, 1, 0, 0, 2]
    got = []
    cur = None
    for k in order + [None]:
        if k is not None:
            c2.upload_async(sets[k][0], sets[k][1])             # (while batch `cur` is still on the GPU)
        if cur is not None:
            n = len(sets[cur][1]) - 1
            rr = np.zeros(n, dtype=dsb.RR_DTYPE); hits = np.zeros(24 * n + 4096, dtype=dsb.HIT_DTYPE)
            used, _ = c2.download_into(rr, hits)
            got.append((cur, rr, hits[:used]))
        if k is not None:
            c2.run(10**6)
        cur = k
    assert [g[0] for g in got] == order
    for k, rr, hits in got:
        assert ob.compare_results(rr, hits, want[k].rr, want[k].hits, None, max_report=3) == [], k
    # the blocking call on the same context afterwards, and a second upload without a run in between is refused
    res = c2.classify(sets[1][0], sets[1][1], 10**6)
    assert ob.compare_results(res.rr, res.hits, want[1].rr, want[1].hits, None, max_report=3) == []
    c2.upload_async(sets[0][0], sets[0][1])
    with pytest.raises(dsb.DsbError):
        c2.upload_async(sets[2][0], sets[2][1])
    c2.run(10**6)
    n = len(sets[0][1]) - 1
    rr = np.zeros(n, dtype=dsb.RR_DTYPE); hits = np.zeros(24 * n + 4096, dtype=dsb.HIT_DTYPE)
    used, _ = c2.download_into(rr, hits)
    assert ob.compare_results(rr, hits[:used], want[0].rr, want[0].hits, None, max_report=3) == []
    c2.close()


def test_driver_reads_longer_than_the_limit_are_unclassified_not_fatal(gpu, ob, demo_index):
    """-L below the longest read of the set: those reads come out unclassified with their true length, every other read as in
    the run without the limit, and the run ends normally with a warning"""
    path = _set_path(ob, "long10")
    _, seqs, _ = ob.read_fastq(path)
    lens = sorted(len(s) for s in seqs)
    limit = lens[len(lens) * 2 // 3]
    n_over = sum(1 for l in lens if l > limit)
    assert 0 < n_over < len(lens)

    def blocks(text):
        return [b for b in text.split(b"\n\n") if b]
    full = blocks(_run_driver(["-f", "DES", demo_index, path]))
    r = subprocess.run([DRIVER, "classify", "-f", "DES", "-L", str(limit), "-B", "64", demo_index, path], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-1000:]
    assert ("%d read(s) longer than %d bases" % (n_over, limit)).encode() in r.stderr
    cut = blocks(r.stdout)
    assert len(full) == len(cut) == len(seqs)
    for a, b, s in zip(full, cut, seqs):
        if len(s) > limit:
            name = a.split(b"\t", 1)[0]
            assert b == name + b"\tUNCLASSIFY\tFAST\t%d\tn_rst:[0]\tn_anc:[0]\t" % len(s)      # (the record of a read classify_seq does not look at, like one below 40 bp)
        else:
            assert a == b
