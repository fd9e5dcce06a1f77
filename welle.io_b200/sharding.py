"""Multi-GPU host logic: independent ensemble streams are partitioned across ranks (one process per GPU, no data-path
collective — the path is embarrassingly parallel).  The only communication is a broadcast of the small
work descriptor from rank 0 and an optional all-reduce of result counters / the max-over-ranks step time.
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def partition(n_total, world, rank):
    """streams e with e % world == rank (ensemble e -> GPU e mod G)"""
    return list(range(rank, n_total, world))


def broadcast_descriptor(values, device, src=0):
    """values: list of ints known on rank `src` (e.g. [n_streams_total, frames, subch start, subch size, bitrate]);
    returns the list on every rank"""
    n = torch.tensor([len(values) if dist.get_rank() == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    t = torch.tensor(values if dist.get_rank() == src else [0] * int(n.item()), dtype=torch.int64, device=device)
    dist.broadcast(t, src)
    return [int(v) for v in t.tolist()]


def reduce_counters(counters, device):
    """element-wise sum over ranks of a list of integer counters (frames, fib_ok, fib_total, rs_uncorrectable, ...)"""
    t = torch.tensor(counters, dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]


def max_over_ranks(seconds, device):
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
