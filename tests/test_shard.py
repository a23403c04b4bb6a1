"""CPU tests of the N > 1 path's host logic: sharding of utterances over ranks and the counter
reduction, with a real 2-process gloo group (the GPU job uses the same code over NCCL)."""
import os
import socket
import subprocess
import sys

import pytest

from xdtts_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_is_a_partition_and_balanced():
    ts = [1000] * 32
    for world in (1, 2, 4, 8):
        parts = [shard.shard_utterances(ts, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(32))
        assert {len(p) for p in parts} == {32 // world}
    ragged = [4, 900, 17, 333, 1000, 250, 64, 64, 511, 8000, 12]
    for world in (2, 3, 8):
        parts = [shard.shard_utterances(ragged, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(len(ragged)))
        loads = [sum(ragged[i] for i in p) for p in parts]
        assert max(loads) <= max(max(ragged), -(-sum(ragged) // world) + max(sorted(ragged)[:-1]))   # LPT bound
    assert shard.shard_utterances([], 2, 1) == []
    assert shard.shard_utterances([5], 4, 3) == []
    with pytest.raises(ValueError):
        shard.shard_utterances(ts, 2, 2)
    assert shard.rank_seed(3, 0) != shard.rank_seed(3, 1)


def test_reduce_without_process_group():
    assert shard.reduce_counters(10, 3, [1.5, 2.5]) == (10.0, 3, [1.5, 2.5])


_WORKER = r"""
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "xd-tts_b200"))
import torch.distributed as dist
from xdtts_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ts = [1000, 10, 500, 500, 990, 20]
mine = shard.shard_utterances(ts, world, rank)
frames = sum(ts[i] for i in mine)
total, launches, times = shard.reduce_counters(frames, 63 * len(mine), [10.0 + rank, 5.0 - rank])
assert total == float(sum(ts)), total
assert launches == 63 * len(ts), launches
assert times == [10.0 + world - 1, 5.0], times
gathered = [None] * world
dist.all_gather_object(gathered, mine)
assert sorted(i for p in gathered for i in p) == list(range(len(ts)))
dist.barrier()
dist.destroy_process_group()
print("rank %%d ok" %% rank)
"""


def test_two_rank_gloo_job(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_c_abi_assignment_matches_the_python_rule(built):
    """xdtts_shard_assign (what xdtts_pool_infer_batch applies across the GPUs of one process) == shard.shard_utterances
    (what bench.py applies across ranks): one rule, two hosts."""
    import ctypes

    import numpy as np

    from xdtts_b200 import _ffi, shard

    lib = _ffi.load_library()
    rng = np.random.default_rng(5)
    for trial in range(40):
        b = int(rng.integers(1, 70))
        n = int(rng.integers(1, 9))
        ts = [int(x) for x in rng.integers(4, 3000, b)]
        if trial % 5 == 0:
            ts = [1000] * b                                     # the benchmark's uniform batches: ties everywhere
        t_arr = (ctypes.c_int * b)(*ts)
        out = (ctypes.c_int * b)()
        _ffi.check(lib.xdtts_shard_assign(t_arr, b, n, out))
        for r in range(n):
            assert [i for i in range(b) if out[i] == r] == shard.shard_utterances(ts, n, r)
    assert lib.xdtts_shard_assign(None, 3, 2, None) == _ffi.ERR_BAD_ARG
