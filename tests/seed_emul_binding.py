"""ctypes binding of tests/emul/libseedemul.so (TEST INFRASTRUCTURE): the per-lane handlers of the seeding engine
(desamba_b200/csrc/dsb_seedcore.h) compiled for the CPU and driven by emulated 32-lane warps.  Used by
tests/test_seed_engine.py to check the engine against the oracle at anchor level without a GPU, and to read the
lanes-per-turn statistics of the state vote."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "emul")
LIB = os.path.join(EMUL_DIR, "libseedemul.so")

TASK_DTYPE = np.dtype([("read", "<u4"), ("sk", "<u4")])
REC_DTYPE = np.dtype([("first_chunk", "<u4"), ("count", "<u4"), ("top_score", "<i4"), ("flag512", "<u4"), ("c_occ", "<u4"),
                      ("c_getref", "<u4"), ("c_getref_bytes", "<u4"), ("c_pl", "<u4")])
STAGED_DTYPE = np.dtype([("ref_ID", "<u4"), ("ref_offset", "<u4"), ("index_in_read", "<u4"), ("len_score", "<u4")])
STATES = ["FETCH", "CTRL", "OCC", "LOCATE", "FLANK", "LV", "RP", "DEAD"]


def build():
    subprocess.run(["make", "-C", EMUL_DIR, "-s"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.emul_open.restype = C.c_void_p
        _lib.emul_open.argtypes = [C.c_char_p]
        _lib.emul_close.argtypes = [C.c_void_p]
        _lib.emul_l_ek.argtypes = [C.c_void_p]
        _lib.emul_seed_pass.restype = C.c_int64
        _lib.emul_seed_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        _lib.emul_lv_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.emul_lv_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    return _lib


class Emul:
    def __init__(self, index_dir):
        self.h = lib().emul_open(index_dir.encode())
        if not self.h:
            raise RuntimeError(f"emul_open({index_dir}) failed")
        self.l_ek = lib().emul_l_ek(self.h)

    def close(self):
        if self.h:
            lib().emul_close(self.h)
            self.h = None

    def seed_pass(self, cat, offs, seeds0, seeds1, seed_off, tasks, slow, n_warps=4, big_rows=0, anc_cap=1 << 22):
        """-> (recs, anc_off, staged anchors, stats dict state -> (turns, busy lanes))"""
        n_reads = len(offs) - 1
        recs = np.zeros(len(tasks), REC_DTYPE)
        anc_off = np.zeros(len(tasks) + 1, np.uint64)
        anc = np.zeros(anc_cap, STAGED_DTYPE)
        stats = np.zeros(2 * len(STATES), np.uint64)
        n = lib().emul_seed_pass(self.h, cat.ctypes.data, offs.ctypes.data, n_reads, seeds0.ctypes.data, seeds1.ctypes.data,
                                 seed_off.ctypes.data, tasks.ctypes.data, len(tasks), int(slow), n_warps, big_rows,
                                 recs.ctypes.data, anc_off.ctypes.data, anc.ctypes.data, anc_cap, stats.ctypes.data)
        if n < 0:
            raise RuntimeError("emul_seed_pass: anchor capacity exceeded")
        anc_off[len(tasks)] = n
        st = {STATES[s]: (int(stats[2 * s]), int(stats[2 * s + 1])) for s in range(len(STATES))}
        return recs, anc_off, anc[:n], st
