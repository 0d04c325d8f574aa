"""The driver's FASTQ input (desamba_b200/csrc/fastq_reader.h): the parallel indexer for plain 4-line FASTQ hands on exactly the
records of the serial kseq-style reader (kseq_read, utils.c:939-977) -- for every block size and thread count, and by falling
back to the serial reader on anything that is not strict 4-line FASTQ."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "desamba_b200", "bin", "fastq_dump")


def _dump(mode, path, *args):
    r = subprocess.run([DUMP, mode, path] + [str(a) for a in args], capture_output=True)
    assert r.returncode in (0, 3), r.stderr
    return r.stdout, r.returncode


def _rand_fastq(rng, n, lo, hi, qual_at=False, crlf=False, comment=False, last_newline=True):
    out = []
    for i in range(n):
        L = int(rng.integers(lo, hi + 1))
        seq = bytes(rng.choice(list(b"ACGTN"), L).tolist())
        qual = bytes(rng.choice(list(b"@+>I5#"), L).tolist()) if qual_at else b"I" * L
        name = b"@r%d" % i + (b" some comment\twith tabs" if comment and i % 2 else b"")
        nl = b"\r\n" if crlf else b"\n"
        out.append(name + nl + seq + nl + (b"+r%d" % i if i % 3 == 0 else b"+") + nl + qual + nl)
    txt = b"".join(out)
    return txt if last_newline else txt.rstrip(b"\r\n")


@pytest.fixture(scope="module")
def ob():
    import oracle_binding
    oracle_binding.build()                       # builds the product too (make -C desamba_b200/csrc)
    assert os.path.exists(DUMP)
    return oracle_binding


@pytest.mark.parametrize("kw", [dict(), dict(qual_at=True), dict(crlf=True), dict(comment=True), dict(last_newline=False),
                                dict(qual_at=True, crlf=True, comment=True, last_newline=False)])
def test_parallel_equals_serial_on_4_line_fastq(ob, tmp_path, kw):
    rng = np.random.default_rng(7)
    p = str(tmp_path / "a.fq")
    open(p, "wb").write(_rand_fastq(rng, 400, 1, 700, **kw))
    want, _ = _dump("serial", p)
    assert want.count(b"\n") == 400
    for block, thr, margin in ((1 << 26, 4, 1 << 24), (5000, 7, 6000), (997, 3, 6000), (200, 16, 1 << 20), (64, 2, 6000), (1, 1, 6000)):
        got, rc = _dump("parallel", p, block, thr, margin)
        assert got == want, (kw, block, thr, margin)
        assert rc == 0, (kw, block, thr, margin)         # handled by the parallel indexer, no fallback (margin > 3 records)
    for block, thr, margin in ((5000, 4, 100), (300, 3, 0), (64, 2, 700)):
        got, rc = _dump("parallel", p, block, thr, margin)   # margin smaller than a record: the block fails, the serial reader takes over
        assert got == want, (kw, block, thr, margin)


def test_long_records_span_several_shares(ob, tmp_path):
    rng = np.random.default_rng(8)
    p = str(tmp_path / "long.fq")
    open(p, "wb").write(_rand_fastq(rng, 30, 20000, 90000, qual_at=True))
    want, _ = _dump("serial", p)
    for block, thr, margin in ((1 << 20, 8, 1 << 20), (30000, 8, 1 << 20), (100000, 64, 1 << 20)):
        got, rc = _dump("parallel", p, block, thr, margin)
        assert got == want and rc == 0


@pytest.mark.parametrize("name,text", [
    ("fasta", b">a\nACGT\nACGT\n>b desc\nGGGG\n"),
    ("wrapped", b"@a\nACGT\nACGT\n+\nIIII\nIIII\n@b\nAC\n+\nII\n"),
    ("truncated_quality", b"@a\nACGTACGT\n+\nIIII\n"),
    ("cut_off", b"@a\nACGT\n+\nIIII\n@b\nACG"),
    ("junk_between", b"@a\nACGT\n+\nIIII\n\n\n@b\nAC\n+\nII\n"),
    ("leading_junk", b"\n\n@a\nACGT\n+\nIIII\n"),
    ("empty_sequence", b"@a\n\n+\n\n@b\nAC\n+\nII\n"),
    ("fasta_after_fastq", b"@a\nACGT\n+\nIIII\n>b\nACGT\n"),
    ("empty", b""),
])
def test_everything_else_falls_back_to_the_serial_reader(ob, tmp_path, name, text):
    p = str(tmp_path / (name + ".fq"))
    open(p, "wb").write(text)
    want, _ = _dump("serial", p)
    for block, thr in ((1 << 20, 4), (8, 3)):
        got, rc = _dump("parallel", p, block, thr)
        assert got == want, (name, block, thr)


def test_demo_and_simulated_sets(ob):
    ob.ensure_demo_index()
    for path in (ob.DEMO_FQ, ob.sim_set("short1", "short", 5000, 0.01, 20261022), ob.sim_set("long10", "long", 300, 0.10, 20261020)):
        want, _ = _dump("serial", path)
        got, rc = _dump("parallel", path, 1 << 18, 8)
        assert got == want and rc == 0


# ---------------------------------------------------------------- the driver's whole host pipeline, without a GPU
DRIVER = os.path.join(ROOT, "desamba_b200", "bin", "deSAMBA-b200")


def _host_only(tmp_path, files, fmt, env=None, *opts):
    """`deSAMBA-b200 classify` with DSB_HOST_ONLY=1: reader -> batches -> writer, no device opened, every read unclassified"""
    out = str(tmp_path / "out.txt")
    e = dict(os.environ, DSB_HOST_ONLY="1", **(env or {}))
    r = subprocess.run([DRIVER, "classify", "-f", fmt, "-o", out] + [str(o) for o in opts] + ["no_index_needed"] + files, capture_output=True, env=e)
    assert r.returncode == 0, r.stderr.decode()[-600:]
    return open(out, "rb").read(), r.stderr.decode()


def _records(rng, n, lo, hi, tag):
    recs = []
    for i in range(n):
        L = int(rng.integers(lo, hi + 1))
        recs.append((b"%s%d" % (tag, i), bytes(rng.choice(list(b"ACGTN"), L).tolist()), bytes(rng.choice(list(b"@+>I5#"), L).tolist())))
    return recs


def _fastq(recs, crlf=False, comment=False):
    nl = b"\r\n" if crlf else b"\n"
    return b"".join(b"@" + n + (b" a comment" if comment and i % 2 else b"") + nl + s + nl + b"+" + nl + q + nl for i, (n, s, q) in enumerate(recs))


def test_driver_host_pipeline_hands_on_every_record_in_order(ob, tmp_path):
    """several files (strict FASTQ through the parallel indexer, mapped or read; CRLF; gzip and a FASTA tail through the serial
    reader), small blocks and small batches: the text holds every record once, in input order, with its bases and qualities"""
    import gzip
    rng = np.random.default_rng(11)
    sets = [_records(rng, 300, 1, 6000, b"a"), _records(rng, 200, 50, 300, b"b"), _records(rng, 40, 20000, 60000, b"c"), _records(rng, 150, 1, 900, b"d")]
    files = [str(tmp_path / f"f{i}.fq") for i in range(4)]
    open(files[0], "wb").write(_fastq(sets[0]))
    open(files[1], "wb").write(_fastq(sets[1], crlf=True, comment=True))
    open(files[2], "wb").write(_fastq(sets[2]))
    files[3] += ".gz"
    with gzip.open(files[3], "wb") as f:
        f.write(_fastq(sets[3]))
    allrec = [r for s in sets for r in s]
    want_full = b"".join(n + b"\t4\t*\t0\t0\t*\t*\t0\t0\t" + s + b"\t" + q + b"\t\n" for n, s, q in allrec)
    want_des = b"".join(n + b"\tUNCLASSIFY\tSLOW\t%d\tn_rst:[0]\tn_anc:[0]\t\n\n" % len(s) for n, s, q in allrec)
    for env, opts in (({}, ()), ({"DSB_FQ_BLOCK_KB": "64"}, ("-B", 37, "-P", 3)), ({"DSB_FQ_BLOCK_KB": "7", "DSB_FQ_MMAP": "0"}, ("-B", 1000, "-M", 1, "-P", 5)),
                      ({"DSB_FQ_BLOCK_KB": "300"}, ("-B", 64, "-P", 0)), ({"DSB_FQ_BLOCK_MB": "1", "DSB_READ_AHEAD": "1"}, ("-B", 500, "-P", 8))):
        got, err = _host_only(tmp_path, files, "SAM_FULL", env, *opts)
        assert got == want_full, (env, opts)
        assert "%d sequences processed" % len(allrec) in err
    got, _ = _host_only(tmp_path, files, "DES", {"DSB_FQ_BLOCK_KB": "16"}, "-B", 100, "-P", 4)
    assert got == want_des


def test_driver_host_pipeline_falls_back_inside_a_file(ob, tmp_path):
    """a file that stops being strict 4-line FASTQ after some blocks: the records before the break come from the indexer, the
    rest (wrapped sequence lines) from the serial reader, nothing lost or doubled"""
    rng = np.random.default_rng(12)
    head = _records(rng, 400, 100, 400, b"h")
    p = str(tmp_path / "mixed.fq")
    wrapped = b"@w1\nACGT\nACGT\n+\nIIII\nIIII\n@w2\nAC\n+\nII\n"
    open(p, "wb").write(_fastq(head) + wrapped + _fastq(_records(rng, 50, 10, 50, b"t")))
    want, _ = _dump("serial", p)
    want_des = b"".join(l.split(b"\t")[0] + b"\tUNCLASSIFY\tSLOW\t%d\tn_rst:[0]\tn_anc:[0]\t\n\n" % len(l.split(b"\t")[1]) for l in want.splitlines())
    assert want_des.count(b"UNCLASSIFY") == 452
    for env in ({"DSB_FQ_BLOCK_KB": "32"}, {"DSB_FQ_BLOCK_KB": "32", "DSB_FQ_MMAP": "0"}, {}):
        got, _ = _host_only(tmp_path, [p, p], "DES", env, "-B", 90, "-P", 4)
        assert got == want_des + want_des, env


def test_driver_text_appended_to_a_file_stays_in_order(ob, tmp_path):
    """stdout redirected with >> or > behind existing bytes of a regular file: the text lands behind them, in order"""
    rng = np.random.default_rng(14)
    recs = _records(rng, 12000, 30, 120, b"q")
    p = str(tmp_path / "app.fq")
    open(p, "wb").write(_fastq(recs))
    want = b"".join(n + b"\t4\t*\t0\t0\t*\t*\t0\t0\t" + s + b"\t" + q + b"\t\n" for n, s, q in recs)
    out = str(tmp_path / "appended.sam")
    open(out, "wb").write(b"header line\n")
    with open(out, "ab") as f:
        r = subprocess.run([DRIVER, "classify", "-f", "SAM_FULL", "-B", "5000", "-P", "4", "no_index_needed", p], stdout=f, stderr=subprocess.PIPE,
                           env=dict(os.environ, DSB_HOST_ONLY="1"))
    assert r.returncode == 0, r.stderr.decode()[-300:]
    assert open(out, "rb").read() == b"header line\n" + want
    # redirected with > at a non-zero offset of a regular file
    with open(out, "wb") as f:
        f.write(b"header line\n"); f.flush()
        r = subprocess.run([DRIVER, "classify", "-f", "SAM_FULL", "-B", "5000", "-P", "4", "no_index_needed", p], stdout=f, stderr=subprocess.PIPE,
                           env=dict(os.environ, DSB_HOST_ONLY="1"))
    assert r.returncode == 0
    assert open(out, "rb").read() == b"header line\n" + want


def test_driver_text_of_large_batches_formatted_by_helper_threads(ob, tmp_path):
    """batches of >= 4096 reads are formatted in shares by helper threads; the text is that of the serial writer, batch after
    batch, to a file and to a pipe, also after a small (serially formatted) batch"""
    rng = np.random.default_rng(13)
    recs = _records(rng, 23000, 30, 200, b"p")
    p = str(tmp_path / "many.fq")
    open(p, "wb").write(_fastq(recs))
    want = b"".join(n + b"\t4\t*\t0\t0\t*\t*\t0\t0\t" + s + b"\t" + q + b"\t\n" for n, s, q in recs)
    small = str(tmp_path / "few.fq")
    open(small, "wb").write(_fastq(recs[:100]))
    want_small = b"".join(n + b"\t4\t*\t0\t0\t*\t*\t0\t0\t" + s + b"\t" + q + b"\t\n" for n, s, q in recs[:100])
    for opts in (("-B", 5000, "-P", 4), ("-B", 9000, "-P", 16), ("-B", 4096, "-P", 2)):
        got, _ = _host_only(tmp_path, [small, p, small, p], "SAM_FULL", {}, *opts)
        assert got == want_small + want + want_small + want, opts
    r = subprocess.run([DRIVER, "classify", "-f", "SAM_FULL", "-B", "5000", "-P", "4", "no_index_needed", p], capture_output=True,
                       env=dict(os.environ, DSB_HOST_ONLY="1"))
    assert r.returncode == 0 and r.stdout == want


def test_gz_files_are_inflated_ahead_of_the_parser(ob, tmp_path):
    """.gz inputs: the next few files of the command line are inflated by threads of their own into bounded queues; the text is
    that of the one-stream reader (DSB_GZ_AHEAD=0), whatever the mix of plain and compressed files, chunk boundaries included"""
    import gzip
    rng = np.random.default_rng(15)
    files, allrec = [], []
    for i in range(9):
        recs = _records(rng, 40 if i % 3 else 3000, 50, 3000 if i != 4 else 9000, b"g%d_" % i)
        path = str(tmp_path / f"part{i}.fq")
        if i in (2, 6):                                   # plain files between the compressed ones
            open(path, "wb").write(_fastq(recs))
        else:
            path += ".gz"
            with gzip.open(path, "wb", compresslevel=1) as f:
                f.write(_fastq(recs, crlf=(i == 5)))
        files.append(path); allrec += recs
    empty = str(tmp_path / "empty.fq.gz")
    with gzip.open(empty, "wb") as f:
        pass
    files.insert(3, empty)
    want = b"".join(n + b"\t4\t*\t0\t0\t*\t*\t0\t0\t" + s + b"\t" + q + b"\t\n" for n, s, q in allrec)
    assert sum(len(s) for _, s, _ in allrec) > 12 << 20          # several 4-MB chunks per large file
    for env, opts in (({"DSB_GZ_AHEAD": "0"}, ("-B", 700)), ({}, ("-B", 700)), ({"DSB_GZ_AHEAD": "1"}, ("-B", 5000, "-P", 3)),
                      ({"DSB_GZ_AHEAD": "3"}, ("-B", 97, "-P", 0)), ({"DSB_GZ_AHEAD": "16"}, ("-B", 100000, "-P", 8))):
        got, err = _host_only(tmp_path, files, "SAM_FULL", env, *opts)
        assert got == want, (env, opts)
        assert "%d sequences processed" % len(allrec) in err
    # a damaged stream ends its file where zlib gives up, like the one-stream reader; the files behind it are read
    bad = str(tmp_path / "cut.fq.gz")
    blob = open(files[0], "rb").read()
    open(bad, "wb").write(blob[: len(blob) // 2])
    a, _ = _host_only(tmp_path, [bad, files[1]], "DES", {"DSB_GZ_AHEAD": "0"}, "-B", 500)
    b, _ = _host_only(tmp_path, [bad, files[1]], "DES", {}, "-B", 500)
    assert a == b and a.count(b"UNCLASSIFY") > 40


def test_reads_from_pipes_plain_or_gzip(ob, tmp_path):
    """inputs that cannot be looked into without consuming them (stdin, process substitution) go through zlib, compressed or not
    -- the reference reads everything with gzread (utils.c:835-977)"""
    import gzip
    rng = np.random.default_rng(16)
    recs = _records(rng, 300, 40, 400, b"s")
    plain = _fastq(recs)
    want = b"".join(n + b"\tUNCLASSIFY\tSLOW\t%d\tn_rst:[0]\tn_anc:[0]\t\n\n" % len(s) for n, s, q in recs)
    e = dict(os.environ, DSB_HOST_ONLY="1")
    for blob in (plain, gzip.compress(plain, 1)):
        r = subprocess.run([DRIVER, "classify", "-f", "DES", "no_index_needed", "/dev/stdin"], input=blob, capture_output=True, env=e)
        assert r.returncode == 0 and r.stdout == want


def test_reads_longer_than_the_limit_do_not_end_the_run(ob, tmp_path):
    """a read longer than -L travels without bases, is written as unclassified with its true length (no bases in SAM_FULL: '*')
    and counted for a warning -- through the block indexer, the one-stream reader and gzip"""
    import gzip
    rng = np.random.default_rng(17)
    recs = _records(rng, 700, 20, 900, b"L")
    n_over = sum(1 for _, s, _ in recs if len(s) > 300)
    assert 100 < n_over < 650
    plain = str(tmp_path / "l.fq"); open(plain, "wb").write(_fastq(recs))
    gz = str(tmp_path / "l.fq.gz"); open(gz, "wb").write(gzip.compress(_fastq(recs), 1))
    want_des = b"".join(n + b"\tUNCLASSIFY\tSLOW\t%d\tn_rst:[0]\tn_anc:[0]\t\n\n" % len(s) for n, s, q in recs)
    want_full = b"".join(n + b"\t4\t*\t0\t0\t*\t*\t0\t0\t" + ((s + b"\t" + q) if len(s) <= 300 else b"*\t*") + b"\t\n" for n, s, q in recs)
    for files, env, opts in (([plain], {"DSB_FQ_BLOCK_KB": "64"}, ("-B", 90, "-P", 4)), ([plain], {}, ("-B", 5000, "-P", 0)), ([gz, plain], {}, ("-B", 64, "-P", 3))):
        got, err = _host_only(tmp_path, files, "DES", env, "-L", 300, *opts)
        assert got == want_des * len(files), (files, env, opts)
        assert "%d read(s) longer than 300 bases" % (n_over * len(files)) in err
        got, _ = _host_only(tmp_path, files, "SAM_FULL", env, "-L", 300, *opts)
        assert got == want_full * len(files), (files, env, opts)
    got, err = _host_only(tmp_path, [plain], "DES", {}, "-B", 200)        # the default limit is 1 Mbase: nothing is held back
    assert got == want_des and "longer than" not in err
