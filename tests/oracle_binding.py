"""ctypes binding of oracle/liborc.so (TEST INFRASTRUCTURE: the CPU restatement of the reference's classify_seq).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.  Records come back in the numpy dtypes of
desamba_b200.api (same byte layout as include/desamba_b200.h), so parity checks are array comparisons.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liborc.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
DEMO_IDX = os.path.join(REF_DIR, "demo", "idx")
DEMO_FQ = os.path.join(REF_DIR, "demo", "ERR1050068.fastq")
DEMO_FA = os.path.join(REF_DIR, "demo", "viral-gs.fa")
SETS_DIR = os.path.join(REF_DIR, "sets")
SYN_IDX = os.path.join(REF_DIR, "syn", "idx")      # second index: synthetic 3 species x 4 strains x 300 kb (oracle/make_golden.sh)
SYN_FA = os.path.join(REF_DIR, "syn", "syn.fa")
SIMREADS = os.path.join(ROOT, "desamba_b200", "bin", "simreads")

HIT_DTYPE = np.dtype([("ref_ID", "<u4"), ("t_st", "<u4"), ("t_ed", "<u4"), ("q_st", "<u4"), ("q_ed", "<u4"), ("sum_score", "<u4"),
                      ("indel", "<u4"), ("direction", "u1"), ("primary", "u1"), ("pri_index", "u1"), ("pad", "u1")])
RR_DTYPE = np.dtype([("hit_off", "<u8"), ("n_hit", "<u4"), ("n_anchor", "<u4"), ("fast_classify", "u1"), ("entered_final", "u1"),
                     ("error", "<u2"), ("read_len", "<u4")])
SEED_DTYPE = np.dtype([("offset", "<u4"), ("len", "<u2"), ("top", "u1"), ("pad", "u1")])


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def ensure_demo_index():
    """The demo index (built by the unmodified reference, oracle/build_index.sh) travels to the GPU box as
    oracle/_ref/demo/idx.tgz (the raw directory is 0.8 GB, mostly the fixed 512 MiB prefix table); unpack on first use."""
    if os.path.exists(os.path.join(DEMO_IDX, "deSAMBA.bwt")):
        return DEMO_IDX
    tgz = os.path.join(REF_DIR, "demo", "idx.tgz")
    if not os.path.exists(tgz):
        raise FileNotFoundError(f"{DEMO_IDX} and {tgz} are both missing: run oracle/build_ref.sh + oracle/build_index.sh where /root/reference exists")
    subprocess.run(["tar", "xzf", tgz, "-C", os.path.join(REF_DIR, "demo")], check=True)
    return DEMO_IDX


def ensure_syn_index():
    """the synthetic multi-strain index (built by the unmodified reference), shipped as oracle/_ref/syn/idx.tgz"""
    if os.path.exists(os.path.join(SYN_IDX, "deSAMBA.bwt")):
        return SYN_IDX
    tgz = os.path.join(REF_DIR, "syn", "idx.tgz")
    if not os.path.exists(tgz):
        raise FileNotFoundError(f"{SYN_IDX} and {tgz} are both missing: run oracle/make_golden.sh where /root/reference exists")
    subprocess.run(["tar", "xzf", tgz, "-C", os.path.join(REF_DIR, "syn")], check=True)
    return SYN_IDX


EK_CLASSES = {17: 28, 18: 30, 19: 32, 20: 34}      # l_ek -> log2 of the exist-k-mer table size in bytes (idx.c:966-996)


def ensure_ek_index(l_ek):
    """the synthetic multi-strain index with its exist-k-mer tables re-written for the size class of `l_ek` (oracle/rebuild_exk.c:
    the index builder's own rule applied to the unitigs recovered from the index; the unmodified reference made the goldens
    tests/golden/*.ek<l_ek>.* on exactly these files).  Tables of 2 x 2^28 .. 2 x 2^34 bytes: built on first use in shared
    memory / a temporary directory, never shipped."""
    src = ensure_syn_index()
    exe = os.path.join(ORACLE_DIR, "rebuild_exk")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(exe + ".c"):
        subprocess.run(["gcc", "-O2", "-w", "-o", exe, exe + ".c"], check=True)
    lg = EK_CLASSES[l_ek]
    need = 2 * (1 << lg) + (1 << 20)
    base = None
    for cand in (os.environ.get("DSB_EK_DIR"), "/dev/shm", "/tmp"):
        if cand and os.path.isdir(cand):
            st = os.statvfs(cand)
            if st.f_bavail * st.f_frsize > need + (2 << 30):
                base = cand
                break
    if base is None:
        raise FileNotFoundError(f"no room for an l_ek {l_ek} index ({need >> 20} MiB)")
    dst = os.path.join(base, "dsb_ek", f"c{lg}")
    if not (os.path.exists(os.path.join(dst, "deSAMBA.exk1")) and os.path.getsize(os.path.join(dst, "deSAMBA.exk1")) == (1 << lg)):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        subprocess.run([exe, src, dst, str(lg)], check=True, capture_output=True)
    return dst


def drop_ek_index(l_ek):
    import shutil
    for cand in (os.environ.get("DSB_EK_DIR"), "/dev/shm", "/tmp"):
        if cand:
            shutil.rmtree(os.path.join(cand, "dsb_ek", f"c{EK_CLASSES[l_ek]}"), ignore_errors=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_capi_open.restype = C.c_void_p
        _lib.orc_capi_open.argtypes = [C.c_char_p, C.c_int, C.c_int]
        _lib.orc_capi_close.argtypes = [C.c_void_p]
        _lib.orc_capi_classify.restype = C.c_int64
        _lib.orc_capi_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_uint64]
        _lib.orc_capi_seeds.restype = C.c_int
        _lib.orc_capi_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        _lib.orc_capi_counters.argtypes = [C.POINTER(C.c_uint64 * 16), C.c_int]
        _lib.orc_capi_classify_mt.restype = C.c_uint64
        _lib.orc_capi_classify_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        _lib.orc_capi_l_ek.argtypes = [C.c_void_p]
        _lib.orc_capi_set_m_bin_read.argtypes = [C.c_uint32]
        _lib.orc_capi_get_m_bin_read.restype = C.c_uint32
    return _lib


COUNTER_NAMES = ["n_hits", "n_reads", "_2", "_3", "n_bit0", "n_bit1", "n_prefix", "n_occ", "n_locate", "n_getref", "n_getref_bytes", "_11", "n_bases"]


class Oracle:
    def __init__(self, index_dir=DEMO_IDX, l_min_match=170, min_score=64):
        if index_dir == DEMO_IDX:
            ensure_demo_index()
        self._h = lib().orc_capi_open(os.fsencode(index_dir), l_min_match, min_score)
        if not self._h:
            raise RuntimeError(f"oracle: cannot load index {index_dir}")

    def close(self):
        if self._h:
            lib().orc_capi_close(self._h)
            self._h = None

    def classify(self, cat, offs, max_read_l_in=0, m_bin_read_in=0):
        """one batch with a fresh scratch buffer (both cross-read states of the reference given explicitly)"""
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        rr = np.zeros(n, dtype=RR_DTYPE)
        cap = max(4096, 8 * n)
        while True:
            hits = np.zeros(cap, dtype=HIT_DTYPE)
            mx = C.c_int32(0)
            lib().orc_capi_set_m_bin_read(m_bin_read_in)
            used = lib().orc_capi_classify(self._h, cat.ctypes.data, offs.ctypes.data, n, max_read_l_in, C.byref(mx), rr.ctypes.data, hits.ctypes.data, cap)
            if used < 0:
                cap *= 4
                continue
            return rr, hits[:used], mx.value

    def seeds(self, seq, strand):
        b = np.frombuffer(seq if isinstance(seq, (bytes, bytearray)) else bytes(seq), dtype=np.uint8)
        out = np.zeros(len(b) // 2 + 4, dtype=SEED_DTYPE)
        ts = C.c_uint32(0)
        n = lib().orc_capi_seeds(self._h, b.ctypes.data, len(b), strand, out.ctypes.data, len(out), C.byref(ts))
        return out[:n], ts.value

    def counters(self, reset=False):
        out = (C.c_uint64 * 16)()
        lib().orc_capi_counters(C.byref(out), 1 if reset else 0)
        return dict(zip(COUNTER_NAMES, list(out)))

    def classify_mt(self, cat, offs, n_threads):
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        return lib().orc_capi_classify_mt(self._h, cat.ctypes.data, offs.ctypes.data, len(offs) - 1, n_threads)


def read_fastq(path, limit=None):
    """-> (names, seqs, quals) as lists of bytes; 4-line FASTQ"""
    names, seqs, quals = [], [], []
    with open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip(b"\r\n")
            f.readline()
            q = f.readline().rstrip(b"\r\n")
            names.append(h[1:].split()[0])
            seqs.append(s)
            quals.append(q)
            if limit and len(seqs) >= limit:
                break
    return names, seqs, quals


def pack(seqs):
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    cat = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return cat, offs


def sim_set(name, mode, n, err, seed, fasta=DEMO_FA):
    """deterministic synthetic read set (desamba_b200/bin/simreads) cached under oracle/_ref/sets/"""
    os.makedirs(SETS_DIR, exist_ok=True)
    path = os.path.join(SETS_DIR, f"{name}.fq")
    if not os.path.exists(path) or os.path.getsize(path) == 0:
        if mode == "mixed":
            n_long, n_short = n
            subprocess.run([SIMREADS, "mixed", fasta, str(n_long), str(n_short), str(seed), path], check=True)
        else:
            subprocess.run([SIMREADS, mode, fasta, str(n), str(err), str(seed), path], check=True)
    return path


def compare_results(rr_g, hits_g, rr_o, hits_o, names=None, max_report=10):
    """field-by-field comparison of GPU and oracle batch results; returns a list of human-readable mismatches"""
    bad = []
    n = len(rr_o)
    for i in range(n):
        g, o = rr_g[i], rr_o[i]
        tag = f"read {i}" + (f" ({names[i].decode()})" if names else "") + f" len={int(o['read_len'])}"
        if int(g["error"]):
            bad.append(f"{tag}: GPU capacity error {int(g['error'])}")
        elif (int(g["n_hit"]), int(g["n_anchor"]), int(g["fast_classify"]), int(g["entered_final"])) != \
                (int(o["n_hit"]), int(o["n_anchor"]), int(o["fast_classify"]), int(o["entered_final"])):
            bad.append(f"{tag}: header gpu(n_hit={g['n_hit']},n_anc={g['n_anchor']},fast={g['fast_classify']},fin={g['entered_final']}) "
                       f"oracle(n_hit={o['n_hit']},n_anc={o['n_anchor']},fast={o['fast_classify']},fin={o['entered_final']})")
        else:
            hg = hits_g[int(g["hit_off"]):int(g["hit_off"]) + int(g["n_hit"])]
            ho = hits_o[int(o["hit_off"]):int(o["hit_off"]) + int(o["n_hit"])]
            if hg.tobytes() != ho.tobytes():
                for k in range(len(ho)):
                    if hg[k].tobytes() != ho[k].tobytes():
                        bad.append(f"{tag}: hit {k} gpu={hg[k]} oracle={ho[k]}")
                        break
        if len(bad) >= max_report:
            break
    return bad
