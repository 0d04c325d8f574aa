"""Multi-GPU host logic: reads are sharded by batch across ranks (one process per GPU, index replicated in each GPU's HBM),
no data-path collective; records are merged back in input order.  The C driver does the same with one
thread per GPU (csrc/classify_main.c); this module is what bench.py and the world_size-2 tests use under torchrun."""
import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)"""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def split_batches(lengths, max_reads, max_bases):
    """cut reads (given their lengths, in input order) into consecutive batches of <= max_reads reads and about max_bases
    bases, like read_reads (cly_mt.c:42-56): a batch is closed once it holds max_reads reads or >= max_bases bases.
    -> list of (first_read, end_read)"""
    out, lo, nb = [], 0, 0
    for i, l in enumerate(lengths):
        nb += int(l)
        if i + 1 - lo >= max_reads or nb >= max_bases:
            out.append((lo, i + 1)); lo, nb = i + 1, 0
    if lo < len(lengths):
        out.append((lo, len(lengths)))
    return out


def deal(n_batches, world_size, rank):
    """batch indices owned by `rank`: batches are dealt round-robin in input order"""
    return list(range(rank, n_batches, world_size))


def merge_in_order(per_rank):
    """per_rank[r] = list of (batch_index, payload) -> payloads sorted by batch index (= input order); checks completeness"""
    allb = sorted((b, p) for lst in per_rank for b, p in lst)
    assert [b for b, _ in allb] == list(range(len(allb))), "missing or duplicated batch"
    return [p for _, p in allb]


def needs_predecessors(known_max, batch_has_short, earlier_unfinished_has_long):
    """The only cross-read state of the reference is Classify_buff_pool.max_read_l (cly.c:2958), used as `max_read_l < 510`.
    A batch must wait for its predecessors only if the state is still < 510, it holds a read < 510 bp, and an unfinished
    earlier batch holds a read >= 510 bp (which may flip the state)."""
    return known_max < 510 and batch_has_short and earlier_unfinished_has_long


def dist_max(x, device=None):
    """max over ranks of a python float (timing is always the max over ranks); identity when not distributed"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dist_sum(x, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
