"""CPU tests of the multi-GPU host logic (desamba_b200/shard.py): batch splitting, round-robin dealing, ordered merge and
the max-over-ranks reduction, including a world_size-2 run over gloo where each rank classifies its batches (the oracle
stands in for the GPU) and rank 0 reassembles the records in input order."""
import os
import sys

import numpy as np
import pytest

from desamba_b200 import shard


def test_split_batches_like_read_reads():
    lens = [100] * 12
    assert shard.split_batches(lens, 5, 10**9) == [(0, 5), (5, 10), (10, 12)]
    assert shard.split_batches(lens, 100, 250) == [(0, 3), (3, 6), (6, 9), (9, 12)]      # closed once >= max_bases
    assert shard.split_batches([], 5, 5) == []
    assert shard.split_batches([1000], 5, 10) == [(0, 1)]


def test_deal_and_merge():
    per_rank = [[(b, f"p{b}") for b in shard.deal(7, 3, r)] for r in range(3)]
    assert sorted(b for lst in per_rank for b, _ in lst) == list(range(7))
    assert shard.merge_in_order(per_rank) == [f"p{b}" for b in range(7)]
    with pytest.raises(AssertionError):
        shard.merge_in_order([[(0, "a")], [(2, "c")]])


def test_needs_predecessors():
    assert shard.needs_predecessors(0, True, True)
    assert not shard.needs_predecessors(600, True, True)
    assert not shard.needs_predecessors(0, False, True)
    assert not shard.needs_predecessors(0, True, False)


def _worker(rank, world_size, port, q, use_gpu=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import oracle_binding as ob
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    names, seqs, _ = ob.read_fastq(ob.DEMO_FQ, 60)
    batches = shard.split_batches([len(s) for s in seqs], 8, 10**9)
    if use_gpu:                                   # the real thing: one process per GPU, the index replicated (both ranks share GPU 0 on a 1-GPU box)
        import torch
        import desamba_b200 as dsb
        ix = dsb.Index(ob.ensure_demo_index(), rank % torch.cuda.device_count())
        eng = dsb.Context(ix)
        def classify(cat, offs):
            r = eng.classify(cat, offs, 10**6)                  # the hit array of the library has merge-sort scratch between the reads: compact it
            hits = np.concatenate([r.read_hits(i) for i in range(len(r.rr))]) if len(r.rr) else r.hits[:0]
            return r.rr, hits
    else:                                         # CPU box: the oracle stands in for the GPU, the host logic is what is tested
        orc = ob.Oracle()
        classify = lambda cat, offs: orc.classify(cat, offs, max_read_l_in=10**6)[:2]
    mine = []
    for b in shard.deal(len(batches), world_size, rank):
        lo, hi = batches[b]
        rr, hits = classify(*ob.pack(seqs[lo:hi]))                                   # long-read state: batches are independent
        mine.append((b, (rr.tobytes(), hits.tobytes())))
    t = shard.dist_max(float(rank + 1))
    n = shard.dist_sum(float(sum(batches[b][1] - batches[b][0] for b in shard.deal(len(batches), world_size, rank))))
    gathered = [None] * world_size
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((t, n, shard.merge_in_order(gathered)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_world_size_2_on_the_gpu(ob, oracle, demo_index):
    _run_world_size_2(ob, oracle, True)


def test_world_size_2_gloo(ob, oracle, demo_index):
    _run_world_size_2(ob, oracle, False)


def _run_world_size_2(ob, oracle, use_gpu):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000 + (500 if use_gpu else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, use_gpu)) for r in range(2)]
    for p in procs: p.start()
    t, n, merged = q.get(timeout=300)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert t == 2.0 and n == 60.0
    names, seqs, _ = ob.read_fastq(ob.DEMO_FQ, 60)
    rr, hits, _ = oracle.classify(*ob.pack(seqs), max_read_l_in=10**6)
    # per-read records of the sharded run, in input order, equal the single-process run
    got_hits = b"".join(h for _, h in merged)
    assert got_hits == hits.tobytes()
    got_rr = np.concatenate([np.frombuffer(r, dtype=ob.RR_DTYPE) for r, _ in merged])
    assert got_rr["n_hit"].tolist() == rr["n_hit"].tolist() and got_rr["n_anchor"].tolist() == rr["n_anchor"].tolist()
