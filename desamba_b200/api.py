"""ctypes binding of include/desamba_b200.h (the C ABI that replaces the reference's classify_seq call, cly.c:3064).

Nothing here computes: every call goes to libdesamba_b200.so.  A missing library is an ImportError, a missing GPU is a
DsbError(DSB_E_CUDA) from dsb_index_load -- there is no CPU fallback.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.environ.get("DSB_LIB") or os.path.join(_HERE, "lib", "libdesamba_b200.so")   # DSB_LIB: developer override (kernel variants)
if not os.path.exists(lib_path):
    raise ImportError(
        f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C desamba_b200/csrc` (nvcc, sm_100a).  There is no CPU fallback.")
lib = C.CDLL(lib_path)

# dsb_hit / dsb_read_result / dsb_seed exactly as declared in the header
HIT_DTYPE = np.dtype([("ref_ID", "<u4"), ("t_st", "<u4"), ("t_ed", "<u4"), ("q_st", "<u4"), ("q_ed", "<u4"), ("sum_score", "<u4"),
                      ("indel", "<u4"), ("direction", "u1"), ("primary", "u1"), ("pri_index", "u1"), ("pad", "u1")])
RR_DTYPE = np.dtype([("hit_off", "<u8"), ("n_hit", "<u4"), ("n_anchor", "<u4"), ("fast_classify", "u1"), ("entered_final", "u1"),
                     ("error", "<u2"), ("read_len", "<u4")])
SEED_DTYPE = np.dtype([("offset", "<u4"), ("len", "<u2"), ("top", "u1"), ("pad", "u1")])
assert HIT_DTYPE.itemsize == 32 and RR_DTYPE.itemsize == 24 and SEED_DTYPE.itemsize == 8


class Opts(C.Structure):
    _fields_ = [("l_min_match", C.c_int32), ("min_score", C.c_int32), ("max_anchors", C.c_uint32), ("max_matches", C.c_uint32),
                ("max_read_len", C.c_uint32), ("warps_per_sm", C.c_uint32), ("pool_scale_pct", C.c_uint32)]


class RefInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("seq_l", C.c_uint64), ("seq_offset", C.c_uint64)]


def _sig(name, res, *args):
    f = getattr(lib, name)
    f.restype = res
    f.argtypes = list(args)
    return f


_vp, _u64p, _i32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
_sig("dsb_last_error", C.c_char_p)
_sig("dsb_version", C.c_char_p)
_sig("dsb_index_load", C.c_int, C.c_char_p, C.c_int, C.POINTER(_vp))
_sig("dsb_index_free", None, _vp)
_sig("dsb_index_n_ref", C.c_uint64, _vp)
_sig("dsb_index_ref_info", C.POINTER(RefInfo), _vp)
_sig("dsb_index_hbm_bytes", C.c_uint64, _vp)
_sig("dsb_index_l_ek", C.c_int, _vp)
_sig("dsb_opts_default", None, C.POINTER(Opts))
_sig("dsb_ctx_create", C.c_int, _vp, C.POINTER(Opts), C.POINTER(_vp))
_sig("dsb_ctx_free", None, _vp)
_sig("dsb_classify_batch", C.c_int, _vp, _vp, _vp, C.c_uint32, C.c_int32, _i32p, _vp, _vp, C.c_uint64, _u64p)
_sig("dsb_batch_upload", C.c_int, _vp, _vp, _vp, C.c_uint32)
_sig("dsb_batch_run", C.c_int, _vp, C.c_int32)
_sig("dsb_batch_download", C.c_int, _vp, _i32p, _vp, _vp, C.c_uint64, _u64p)
_sig("dsb_batch_sync", C.c_int, _vp)
_sig("dsb_batch_get_seeds", C.c_int, _vp, C.c_uint32, C.c_int, _vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32))
_sig("dsb_batch_kernel_ms", C.c_int, _vp, C.POINTER(C.c_float * 11), C.c_int)
_sig("dsb_ctx_mark", C.c_int, _vp, C.c_int)
_sig("dsb_ctx_elapsed_ms", C.c_int, _vp, C.c_int, _vp, C.c_int, C.POINTER(C.c_float))
_sig("dsb_batch_launches", C.c_int, _vp)
_sig("dsb_batch_retries", C.c_int, _vp)
_sig("dsb_batch_timeline", C.c_int, _vp, C.c_int, _vp, C.POINTER(C.c_float * 13), C.c_int)
_sig("dsb_ctx_host_seconds", C.c_int, _vp, C.POINTER(C.c_double * 5), C.c_int)
_sig("dsb_set_sync_mode", C.c_int, C.c_int)
_sig("dsb_batch_work", C.c_int, _vp, C.POINTER(C.c_uint32 * 12))
_sig("dsb_index_clone", C.c_int, _vp, C.c_int, C.POINTER(_vp))
_sig("dsb_batch_counters", C.c_int, _vp, C.POINTER(C.c_uint64 * 16))
_sig("dsb_batch_profile", C.c_int, _vp, _vp)
_sig("dsb_ctx_set_bin_capacity", C.c_int, _vp, C.c_uint32)
_sig("dsb_ctx_bin_capacity", C.c_uint32, _vp)
_sig("dsb_ctx_stream", _vp, _vp)
_sig("dsb_host_alloc", C.c_int, C.c_size_t, C.POINTER(_vp))
_sig("dsb_host_free", None, _vp)
_sig("dsb_gather_bench", C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))

DSB_E_CAPACITY = -5
KERNEL_NAMES = ["k_encode_probe", "k_islands", "k_seed(fast)", "k_chain(fast)", "k_seed(slow0)", "k_chain(slow0)", "k_seed(slow1)", "k_chain(slow1)",
                "k_score", "k_score_heavy", "k_finalize"]
COUNTER_NAMES = ["hit_slots", "reads_taken", "first_long", "max_read_l", "n_bit0", "n_bit1", "n_prefix", "n_occ", "n_locate",
                 "n_getref_seed", "n_getref_bytes_seed", "n_errors", "n_getref_score", "n_getref_bytes_score"]


class DsbError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        super().__init__(f"{where}: error {code}: {lib.dsb_last_error().decode(errors='replace')}")


def _check(rc, where):
    if rc != 0:
        raise DsbError(rc, where)


class Index:
    """The on-disk deSAMBA index resident in one GPU's HBM (load_idx, idx.c:1103-1160)."""

    def __init__(self, index_dir, device=0, _clone_of=None):
        self._h = _vp()
        if _clone_of is None:
            _check(lib.dsb_index_load(os.fsencode(index_dir), device, C.byref(self._h)), "dsb_index_load")
        else:
            _check(lib.dsb_index_clone(_clone_of._h, device, C.byref(self._h)), "dsb_index_clone")
        self.device = device
        n = lib.dsb_index_n_ref(self._h)
        ri = lib.dsb_index_ref_info(self._h)
        self.ref_names = [ri[i].name.decode(errors="replace") for i in range(n)]
        self.ref_len = [ri[i].seq_l for i in range(n)]
        self.hbm_bytes = lib.dsb_index_hbm_bytes(self._h)
        self.l_ek = lib.dsb_index_l_ek(self._h)

    def clone(self, device):
        """the same index on another GPU, copied device to device (dsb_index_clone)"""
        return Index(None, device, _clone_of=self)

    def close(self):
        if self._h:
            lib.dsb_index_free(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """Pinned host memory exposed as a numpy uint8 array (dsb_host_alloc)."""

    def __init__(self, nbytes):
        self._p = _vp()
        self.nbytes = max(int(nbytes), 1)
        _check(lib.dsb_host_alloc(self.nbytes, C.byref(self._p)), "dsb_host_alloc")
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self._p.value))

    def close(self):
        if self._p:
            self.array = None
            lib.dsb_host_free(self._p)
            self._p = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchResult:
    """Per-read results of one batch: rr (RR_DTYPE[n_reads]) and hits (HIT_DTYPE[...], indexed by rr.hit_off)."""

    def __init__(self, rr, hits, max_read_l):
        self.rr, self.hits, self.max_read_l = rr, hits, max_read_l

    def read_hits(self, i):
        o, n = int(self.rr["hit_off"][i]), int(self.rr["n_hit"][i])
        return self.hits[o:o + n]


def pack_reads(seqs):
    """list of bytes/str -> (uint8 concatenation, uint64 offsets[n+1])"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    cat = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, dtype=np.uint8)
    return cat, offs


class Context:
    """Stream + scratch for batches on one GPU (replaces Classify_buff_pool, cly.h:137-158)."""

    def __init__(self, index, l_min_match=170, min_score=64, max_anchors=None, max_matches=None, warps_per_sm=None, max_read_len=None, pool_scale_pct=None):
        self.index = index
        o = Opts()
        lib.dsb_opts_default(C.byref(o))
        o.l_min_match, o.min_score = l_min_match, min_score
        if max_anchors:
            o.max_anchors = max_anchors
        if max_matches:
            o.max_matches = max_matches
        if warps_per_sm:
            o.warps_per_sm = warps_per_sm
        if max_read_len:
            o.max_read_len = max_read_len
        if pool_scale_pct:
            o.pool_scale_pct = pool_scale_pct
        self._h = _vp()
        _check(lib.dsb_ctx_create(index._h, C.byref(o), C.byref(self._h)), "dsb_ctx_create")
        self.n_reads = 0

    def close(self):
        if self._h:
            lib.dsb_ctx_free(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def work(self):
        """work-list sizes of the last run (dsb_batch_work)"""
        out = (C.c_uint32 * 12)()
        _check(lib.dsb_batch_work(self._h, C.byref(out)), "dsb_batch_work")
        names = ["slow0_reads", "slow1_reads", "scored", "scored_heavy", "tasks_fast", "tasks_slow0", "tasks_slow1", "chunks_fast", "chunks_slow0", "chunks_slow1", "anchors", "chains"]
        return dict(zip(names, list(out)))

    def retries(self):
        """re-runs of the last batch after a pool overflow (dsb_batch_retries)"""
        return int(lib.dsb_batch_retries(self._h))

    def timeline(self, ref, ref_mark=0):
        """ms from `ref`'s mark to: upload start, first kernel start, end of each of the 11 kernel groups (dsb_batch_timeline)"""
        a = (C.c_float * 13)()
        _check(lib.dsb_batch_timeline(ref._h, ref_mark, self._h, C.byref(a), 13), "dsb_batch_timeline")
        return list(a)

    def host_seconds(self):
        """host seconds `classify` spent on this context so far (dsb_ctx_host_seconds): upload, launches, waiting + results; calls; re-runs"""
        a = (C.c_double * 5)()
        _check(lib.dsb_ctx_host_seconds(self._h, C.byref(a), 5), "dsb_ctx_host_seconds")
        return dict(zip(("upload", "run", "download", "calls", "retries"), list(a)))

    # -- the end-to-end call with host buffers
    def classify(self, cat, offs, max_read_l_in=0, m_bin_read_in=0):
        """one batch; both cross-read states of the reference are given explicitly (0, 0 = first batch of a run)"""
        lib.dsb_ctx_set_bin_capacity(self._h, m_bin_read_in)
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        rr = np.zeros(n, dtype=RR_DTYPE)
        cap = max(4096, 24 * n)
        while True:
            hits = np.zeros(cap, dtype=HIT_DTYPE)
            mx, used = C.c_int32(0), C.c_uint64(0)
            rc = lib.dsb_classify_batch(self._h, cat.ctypes.data, offs.ctypes.data, n, max_read_l_in, C.byref(mx), rr.ctypes.data,
                                        hits.ctypes.data, cap, C.byref(used))
            if rc == DSB_E_CAPACITY and used.value > cap:
                cap = int(used.value)
                continue
            self.n_reads = n
            if rc != 0:
                e = DsbError(rc, "dsb_classify_batch")
                if rc == DSB_E_CAPACITY:                 # per-read capacity: the other reads' records are valid (dsb_read_result.error)
                    e.result = BatchResult(rr, hits[:min(used.value, cap)], mx.value)
                raise e
            return BatchResult(rr, hits[:used.value], mx.value)

    def classify_into(self, cat, offs, rr, hits, max_read_l_in=0):
        """dsb_classify_batch with caller-owned (ideally pinned) input and output arrays; returns (n_hit_slots, max_read_l)"""
        n = len(offs) - 1
        mx, used = C.c_int32(0), C.c_uint64(0)
        _check(lib.dsb_classify_batch(self._h, cat.ctypes.data, offs.ctypes.data, n, max_read_l_in, C.byref(mx), rr.ctypes.data,
                                      hits.ctypes.data, len(hits), C.byref(used)), "dsb_classify_batch")
        self.n_reads = n
        return used.value, mx.value

    def classify_reads(self, seqs, max_read_l_in=0, m_bin_read_in=0):
        cat, offs = pack_reads(seqs)
        return self.classify(cat, offs, max_read_l_in, m_bin_read_in)

    # -- the three steps separately (run works on inputs resident in HBM)
    def upload(self, cat, offs, m_bin_read_in=0):
        lib.dsb_ctx_set_bin_capacity(self._h, m_bin_read_in)
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        self.n_reads = len(offs) - 1
        _check(lib.dsb_batch_upload(self._h, cat.ctypes.data, offs.ctypes.data, self.n_reads), "dsb_batch_upload")
        _check(lib.dsb_batch_sync(self._h), "dsb_batch_sync")

    def upload_async(self, cat, offs, m_bin_read_in=0):
        """dsb_batch_upload without waiting: may be called for the NEXT batch while the one run before is still on the GPU
        (order per context: upload(k+1); download(k); run(k+1)); cat / offs must stay alive until the batch has been run"""
        lib.dsb_ctx_set_bin_capacity(self._h, m_bin_read_in)
        self.n_reads_next = len(offs) - 1
        _check(lib.dsb_batch_upload(self._h, cat.ctypes.data, offs.ctypes.data, self.n_reads_next), "dsb_batch_upload")

    def download_into(self, rr, hits):
        """dsb_batch_download into caller-owned arrays; returns (n_hit_slots, max_read_l)"""
        mx, used = C.c_int32(0), C.c_uint64(0)
        _check(lib.dsb_batch_download(self._h, C.byref(mx), rr.ctypes.data, hits.ctypes.data, len(hits), C.byref(used)), "dsb_batch_download")
        return used.value, mx.value

    def run(self, max_read_l_in=0):
        _check(lib.dsb_batch_run(self._h, max_read_l_in), "dsb_batch_run")

    def sync(self):
        _check(lib.dsb_batch_sync(self._h), "dsb_batch_sync")

    def download(self):
        n = self.n_reads
        rr = np.zeros(n, dtype=RR_DTYPE)
        cap = max(4096, 24 * n)
        hits = np.zeros(cap, dtype=HIT_DTYPE)
        mx, used = C.c_int32(0), C.c_uint64(0)
        _check(lib.dsb_batch_download(self._h, C.byref(mx), rr.ctypes.data, hits.ctypes.data, cap, C.byref(used)), "dsb_batch_download")
        return BatchResult(rr, hits[:used.value], mx.value)

    def seeds(self, read, strand):
        n, ts = C.c_uint32(0), C.c_uint32(0)
        _check(lib.dsb_batch_get_seeds(self._h, read, strand, None, 0, C.byref(n), C.byref(ts)), "dsb_batch_get_seeds")
        out = np.zeros(n.value, dtype=SEED_DTYPE)
        if n.value:
            _check(lib.dsb_batch_get_seeds(self._h, read, strand, out.ctypes.data, n.value, C.byref(n), C.byref(ts)), "dsb_batch_get_seeds")
        return out, ts.value

    def kernel_ms(self):
        """device ms of the 11 kernel groups of the last run (KERNEL_NAMES)"""
        ms = (C.c_float * 11)()
        _check(lib.dsb_batch_kernel_ms(self._h, C.byref(ms), 11), "dsb_batch_kernel_ms")
        return list(ms)

    def mark(self, which):
        _check(lib.dsb_ctx_mark(self._h, which), "dsb_ctx_mark")

    def elapsed_ms(self, mark_a, other, mark_b):
        ms = C.c_float(0)
        _check(lib.dsb_ctx_elapsed_ms(self._h, mark_a, other._h, mark_b, C.byref(ms)), "dsb_ctx_elapsed_ms")
        return ms.value

    def launches(self):
        return lib.dsb_batch_launches(self._h)

    def counters(self):
        out = (C.c_uint64 * 16)()
        _check(lib.dsb_batch_counters(self._h, C.byref(out)), "dsb_batch_counters")
        d = dict(zip(COUNTER_NAMES, list(out)))
        d["n_getref"] = d["n_getref_seed"] + d["n_getref_score"]
        d["n_getref_bytes"] = d["n_getref_bytes_seed"] + d["n_getref_bytes_score"]
        return d

    def profile(self):
        """per-read phase times [n_reads, 8] in units of 1024 SM cycles (fast, chain, slow, kidx, middle, right, left, total)"""
        out = np.zeros((self.n_reads, 8), dtype=np.uint32)
        _check(lib.dsb_batch_profile(self._h, out.ctypes.data), "dsb_batch_profile")
        return out

    def set_bin_capacity(self, m):
        """cross-read state #2 of the reference (capacity of its bin_read buffer); see include/desamba_b200.h"""
        _check(lib.dsb_ctx_set_bin_capacity(self._h, m), "dsb_ctx_set_bin_capacity")

    def bin_capacity(self):
        return lib.dsb_ctx_bin_capacity(self._h)

    def stream(self):
        return lib.dsb_ctx_stream(self._h)


def set_sync_mode(blocking):
    """host threads sleep (True) or let the CUDA runtime decide (False) while they wait for a batch; before the first Index"""
    _check(lib.dsb_set_sync_mode(1 if blocking else 0), "dsb_set_sync_mode")


def gather_bench(device=0, table_bytes=8 << 30, n_gathers=1 << 28, bytes_each=1):
    gbs, ms = C.c_double(0), C.c_double(0)
    _check(lib.dsb_gather_bench(device, table_bytes, n_gathers, bytes_each, C.byref(gbs), C.byref(ms)), "dsb_gather_bench")
    return gbs.value, ms.value
