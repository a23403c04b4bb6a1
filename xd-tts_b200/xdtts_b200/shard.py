"""Multi-GPU plumbing of the hot path: utterances are independent (the reference vocodes one at a
time, /root/reference src/lib.rs:83-104, 110-159), so a job is sharded across ranks -- one process
and one CUDA context per GPU -- with NO data-path collective.  The only exchange is the reduction
of the throughput counters at the end of a run: frames summed, time maxed.

Pure host logic (torch.distributed is only touched when a process group exists), exercised on CPU
with the gloo backend in tests/test_shard.py and with NCCL by bench.py under torchrun.
"""
from __future__ import annotations


def shard_utterances(frame_counts, world_size, rank):
    """Indices of the utterances rank `rank` vocodes.

    Longest-first greedy assignment to the least-loaded rank (load = frames), ties to the lowest
    rank: balanced for ragged batches, contiguous-equivalent for uniform ones, deterministic on
    every rank without communication.  Returned indices are ascending."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("rank %d of world size %d" % (rank, world_size))
    order = sorted(range(len(frame_counts)), key=lambda i: (-int(frame_counts[i]), i))
    load = [0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        load[r] += int(frame_counts[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def reduce_counters(frames, launches, times_ms, device=None):
    """(sum of frames over ranks, sum of launches, elementwise max of times_ms) -- the job's single
    collective.  Without an initialised process group the inputs are returned unchanged."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(frames), int(launches), [float(x) for x in times_ms]
    cnt = torch.tensor([float(frames), float(launches)], dtype=torch.float64, device=device)
    tms = torch.tensor([float(x) for x in times_ms], dtype=torch.float64, device=device)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    return float(cnt[0]), int(cnt[1]), [float(x) for x in tms.tolist()]


def rank_seed(base_seed, rank):
    """Seed of a rank's phase generator: ranks must not draw the same initial phases."""
    return int(base_seed) * 1000003 + int(rank)
