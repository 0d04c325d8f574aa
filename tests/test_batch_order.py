"""The driver's only cross-batch dependency (desamba_b200/csrc/batch_order.h: Classify_buff_pool.max_read_l, cly.c:2958) on
the CPU: batches finish out of order on several contexts / GPUs, yet every batch must get the max_read_l of the batches
BEFORE it in input order (the reference's -t 1 semantics) as far as the `< 510` test can tell."""
import ctypes as C
import os
import random
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


class Slot(C.Structure):
    _fields_ = [("state", C.c_int), ("has_long", C.c_int), ("has_short", C.c_int), ("max_out", C.c_int32), ("seq_no", C.c_uint64)]


FREE, READY, BUSY, DONE = 0, 1, 2, 3


def _lib():
    subprocess.run(["make", "-C", os.path.join(HERE, "emul"), "-s", "libbatchorder.so"], check=True)
    lib = C.CDLL(os.path.join(HERE, "emul", "libbatchorder.so"))
    lib.capi_bo_may_start.argtypes = [C.POINTER(Slot), C.c_int, C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_int32)]
    lib.capi_bo_finished.argtypes = [C.POINTER(Slot), C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    return lib


def simulate(lib, batches, n_workers, n_slots, rng):
    """batches: list of (max read length that reaches the filter or 0, has_long, has_short).  Random completion order among the
    running batches.  Returns the max_in every batch was started with."""
    slots = (Slot * n_slots)()
    done_upto, prefix_max = C.c_uint64(0), C.c_int32(0)
    n_filled = n_claimed = n_written = 0
    running, waiting, got = [], [], {}
    while n_written < len(batches):
        # reader: fill free slots
        while n_filled < len(batches) and n_filled - n_written < n_slots and slots[n_filled % n_slots].state == FREE:
            s = slots[n_filled % n_slots]
            s.state, s.has_long, s.has_short, s.max_out, s.seq_no = READY, batches[n_filled][1], batches[n_filled][2], 0, n_filled
            n_filled += 1
        # workers claim in input order
        while n_claimed < n_filled and len(running) + len(waiting) < n_workers:
            slots[n_claimed % n_slots].state = BUSY
            waiting.append(n_claimed); n_claimed += 1
        for my in list(waiting):
            mi = C.c_int32(-1)
            if lib.capi_bo_may_start(slots, n_slots, my, done_upto.value, prefix_max.value, C.byref(mi)):
                got[my] = mi.value; waiting.remove(my); running.append(my)
        assert running, "deadlock: every claimed batch waits"
        my = running.pop(rng.randrange(len(running)))                      # any running batch may finish next
        s = slots[my % n_slots]
        s.max_out = max(got[my], batches[my][0]); s.state = DONE
        lib.capi_bo_finished(slots, n_slots, n_claimed, C.byref(done_upto), C.byref(prefix_max))
        while n_written < n_claimed and slots[n_written % n_slots].state == DONE and slots[n_written % n_slots].seq_no == n_written:
            assert n_written < done_upto.value
            slots[n_written % n_slots].state = FREE; n_written += 1
    return got


def test_value_comes_from_predecessors_only():
    lib = _lib()
    rng = random.Random(7)
    for trial in range(300):
        n = rng.randrange(1, 60)
        batches = []
        for _ in range(n):
            kind = rng.choice(["short", "long_nohit", "long", "mixed", "mixed_nohit"])
            has_long, has_short = kind != "short", kind in ("short", "mixed", "mixed_nohit")
            reached = {"short": rng.choice([0, 150]), "long_nohit": 0, "long": rng.choice([600, 9000]), "mixed": 700, "mixed_nohit": rng.choice([0, 150])}[kind]
            batches.append((reached, int(has_long), int(has_short)))
        got = simulate(lib, batches, rng.randrange(1, 9), rng.randrange(2, 20) + 8, rng)
        prefix = 0
        for k, (reached, has_long, has_short) in enumerate(batches):
            if has_short:                                                  # the only batches whose result depends on it
                assert (got[k] >= 510) == (prefix >= 510), (trial, k, got[k], prefix, batches)
            assert got[k] <= prefix                                        # never a value from a later batch
            prefix = max(prefix, reached)


def test_the_race_of_the_round_1_driver():
    """an earlier batch of long reads without hits still runs, a LATER batch pushes the maximum over 510: the short reads of
    the batch in between must still see < 510 (the old driver read a global that later batches updated)"""
    lib = _lib()
    n_slots = 8
    slots = (Slot * n_slots)()
    for k, (hl, hs) in enumerate([(1, 0), (0, 1), (1, 0)]):
        slots[k].state, slots[k].has_long, slots[k].has_short, slots[k].seq_no = BUSY, hl, hs, k
    done_upto, prefix_max, mi = C.c_uint64(0), C.c_int32(0), C.c_int32(-1)
    assert lib.capi_bo_may_start(slots, n_slots, 1, 0, 0, C.byref(mi)) == 0          # batch 0 (long reads) is unfinished: wait
    slots[2].state, slots[2].max_out = DONE, 9000                                    # the LATER batch finishes first
    lib.capi_bo_finished(slots, n_slots, 3, C.byref(done_upto), C.byref(prefix_max))
    assert done_upto.value == 0 and prefix_max.value == 0
    assert lib.capi_bo_may_start(slots, n_slots, 1, 0, 0, C.byref(mi)) == 0          # still waiting for batch 0, not fooled by batch 2
    slots[0].state, slots[0].max_out = DONE, 0                                       # batch 0 ends: none of its reads reached the filter
    lib.capi_bo_finished(slots, n_slots, 3, C.byref(done_upto), C.byref(prefix_max))
    assert done_upto.value == 1 and prefix_max.value == 0
    assert lib.capi_bo_may_start(slots, n_slots, 1, done_upto.value, prefix_max.value, C.byref(mi)) == 1 and mi.value == 0
