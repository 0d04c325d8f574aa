"""The seeding engine (desamba_b200/csrc/dsb_seedcore.h: the per-lane state machine k_seed runs) against the oracle, on
the CPU: the handlers are compiled for the host and driven by emulated 32-lane warps (tests/emul), the tasks of many
reads mixed in one list as on the GPU.  Anchors of every strand pass must equal the oracle's fast_classify /
slow_classify (cly.c:1476-1611) anchor by anchor, the algorithmic counters too."""
import ctypes as C
import os
import numpy as np
import pytest

import oracle_binding as ob
import seed_emul_binding as eb

ANCHOR_DTYPE = np.dtype([("ref_ID", "<u4"), ("ref_offset", "<u4"), ("index_in_read", "<u4"), ("len_score", "<u4"), ("dir_useless", "<u4")])


def oracle_seed_pass(orc, seq, d, slow):
    ob.lib().orc_capi_seed_pass.restype = C.c_int
    ob.lib().orc_capi_seed_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_uint32,
                                            C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
    b = np.frombuffer(seq, dtype=np.uint8)
    cap = 1 << 16
    out = np.zeros(cap, ANCHOR_DTYPE)
    strand, both = C.c_int(0), C.c_int(0)
    cnt = np.zeros(5, np.uint64)
    n = ob.lib().orc_capi_seed_pass(orc._h, b.ctypes.data, len(b), d, int(slow), out.ctypes.data, cap, C.byref(strand), C.byref(both), cnt.ctypes.data)
    assert n <= cap
    return out[:n], strand.value, both.value, cnt


def run_set(index_dir, seqs, slow, n_warps=3, big_rows=0):
    """every strand pass (both search directions) of every read through the emulated engine and through the oracle"""
    orc = ob.Oracle(index_dir)
    em = eb.Emul(index_dir)
    cat, offs = ob.pack(seqs)
    n = len(seqs)
    seeds = [[None, None] for _ in range(n)]
    seed_off = np.zeros(n + 1, np.uint32)
    tot = 0
    for r, s in enumerate(seqs):
        seed_off[r] = tot
        if len(s) >= 40:
            for st in range(2):
                seeds[r][st] = orc.seeds(s, st)[0]
            tot += len(s) // 2 + 2
    seed_off[n] = tot
    arr = [np.zeros(tot + 1, ob.SEED_DTYPE), np.zeros(tot + 1, ob.SEED_DTYPE)]
    tasks = []
    for r, s in enumerate(seqs):
        if len(s) < 40:
            continue
        for st in range(2):
            sv = seeds[r][st]
            arr[st][seed_off[r]:seed_off[r] + len(sv)] = sv
            for k in range(len(sv)):
                if slow:
                    ok = not (int(sv[k]["len"]) < 3 and int(sv[0]["top"]) == 0)     # sv_f->top: seed 0's flag, as written (cly.c:1564)
                else:
                    ok = int(sv[k]["top"]) != 0
                if ok:
                    tasks.append((r, (st << 31) | k))
    tasks = np.array(tasks, eb.TASK_DTYPE) if tasks else np.zeros(0, eb.TASK_DTYPE)
    recs, anc_off, anc, stats = em.seed_pass(cat, offs, arr[0], arr[1], seed_off, tasks, slow, n_warps=n_warps, big_rows=big_rows)
    # ordered gather per (read, strand) as phase_chain does it: a seed that scored > 512 drops the NEXT seed of the list (fast only)
    by = {}
    for t in range(len(tasks)):
        by.setdefault((int(tasks[t]["read"]), int(tasks[t]["sk"]) >> 31), []).append(t)
    bad = []
    n_anchors = 0
    for r, s in enumerate(seqs):
        if len(s) < 40:
            continue
        for d in range(2):
            exp, strand, both, cnt_o = oracle_seed_pass(orc, s, d, slow)
            got = []
            cnt_g = np.zeros(5, np.uint64)
            prev_k, prev_flag = -2, False
            for t in by.get((r, strand), []):
                k = int(tasks[t]["sk"]) & 0x7fffffff
                rec = recs[t]
                dropped = prev_flag and prev_k + 1 == k
                if not dropped:
                    for a in anc[int(anc_off[t]):int(anc_off[t]) + int(rec["count"])]:
                        score = np.int16(int(a["len_score"]) >> 16)
                        got.append((int(a["ref_ID"]), int(a["ref_offset"]), int(a["index_in_read"]), int(a["len_score"]),
                                    (1 if strand == 0 else 0) | ((1 if int(score) < int(rec["top_score"]) else 0) << 8)))
                    cnt_g += np.array([int(rec["c_pl"]) & 0xffff, int(rec["c_occ"]), int(rec["c_pl"]) >> 16, int(rec["c_getref"]), int(rec["c_getref_bytes"])], np.uint64)
                prev_flag = (int(rec["flag512"]) & 1) != 0 and not dropped
                prev_k = k
                assert (int(rec["flag512"]) >> 8) == 0, f"lane error {int(rec['flag512']) >> 8}"
            # a non-eligible seed in between absorbs the skip: the next task then has k > prev_k + 1 (handled by the k test above)
            exp_t = [tuple(int(x) for x in a) for a in exp]
            n_anchors += len(exp_t)
            if got != exp_t:
                first = next((i for i in range(min(len(got), len(exp_t))) if got[i] != exp_t[i]), min(len(got), len(exp_t)))
                bad.append(f"read {r} dir {d} strand {strand}: {len(got)} vs {len(exp_t)} anchors, first difference at {first}: "
                           f"got {got[first] if first < len(got) else None} expected {exp_t[first] if first < len(exp_t) else None}")
            elif list(cnt_g) != list(cnt_o):
                bad.append(f"read {r} dir {d}: counters {list(cnt_g)} vs oracle {list(cnt_o)}")
    em.close(); orc.close()
    return bad, n_anchors, stats


def load(path, limit):
    return ob.read_fastq(path, limit)[1]


@pytest.mark.parametrize("slow", [0, 1])
def test_engine_demo_reads(slow):
    ob.ensure_demo_index()
    seqs = load(ob.DEMO_FQ, 120)
    bad, n_anchors, stats = run_set(ob.DEMO_IDX, seqs, slow)
    assert not bad, "\n".join(bad[:10])
    assert n_anchors > 100


@pytest.mark.parametrize("name,slow", [("long10", 0), ("long10", 1), ("long30", 1), ("short1", 0), ("short1", 1)])
def test_engine_synthetic_sets(name, slow):
    ob.ensure_demo_index()
    seqs = load(os.path.join(ob.SETS_DIR, f"{name}.fq"), 400 if name == "short1" else 40)
    bad, n_anchors, stats = run_set(ob.DEMO_IDX, seqs, slow)
    assert not bad, "\n".join(bad[:10])
    assert n_anchors > 50


@pytest.mark.parametrize("name,slow", [("syn_long10", 0), ("syn_long10", 1), ("syn_short1", 0), ("syn_short1", 1)])
def test_engine_multi_strain_index(name, slow):
    """second index: 3 species x 4 strains (several reference positions per unitig, short unitigs -> short flanks, the
    per-reference re-extension of get_new_ed)"""
    ob.ensure_syn_index()
    seqs = load(os.path.join(ob.SETS_DIR, f"{name}.fq"), 400 if "short" in name else 30)
    bad, n_anchors, stats = run_set(ob.SYN_IDX, seqs, slow)
    assert not bad, "\n".join(bad[:10])
    assert n_anchors > 50


def test_engine_full_row_confirmation_path():
    """indexes beyond 2^32 BWT rows confirm a tier-1 tag match of the visited-row set against the full row: forcing that
    path on a small index must not change anything"""
    ob.ensure_demo_index()
    seqs = load(os.path.join(ob.SETS_DIR, "long10.fq"), 12)
    bad, n_anchors, _ = run_set(ob.DEMO_IDX, seqs, 0, big_rows=1)
    assert not bad, "\n".join(bad[:10])


def test_packed_landau_vishkin_equals_byte_version():
    """lv_packed (2-bit windows, mn/ed rows in two registers) == lv_bytes (the plain statement) on random strings of every
    length 0..12 with random bytes in front (the reference reads up to 5 bytes before a short flank)"""
    rng = np.random.default_rng(20261018)
    L = eb.lib()
    n_diff = 0
    for it in range(60000):
        ln = int(rng.integers(0, 13))
        ref = rng.integers(0, 4, 18).astype(np.uint8)
        q = ref.copy() if it % 3 else rng.integers(0, 4, 18).astype(np.uint8)
        for _ in range(int(rng.integers(0, 5))):                    # a few edits
            p = int(rng.integers(0, 18))
            if rng.integers(0, 2):
                q[p] = rng.integers(0, 4)
            else:
                q[p:] = np.roll(q[p:], 1)
        a = L.emul_lv_packed(ref.ctypes.data, q.ctypes.data, ln)
        b = L.emul_lv_bytes(ref.ctypes.data, q.ctypes.data, ln)
        n_diff += a != b
    assert n_diff == 0
