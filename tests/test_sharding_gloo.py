"""CPU, world_size 2 over gloo: the host logic of the sharded (N > 1 GPU) path."""
from __future__ import annotations

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from allsteps_isaaclab_b200 import _cabi
from allsteps_isaaclab_b200.sharding import STAT_NAMES, StatsReducer, all_reduce_exchange, shard_range


def test_shard_ranges_partition_the_env_ids():
    for n, w in [(1 << 20, 8), (4096, 2), (1001, 4), (64, 8), (7, 2)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(s % 4 == 0 for s, _ in spans)
    assert shard_range(1 << 20, 3, 8) == (3 * 131072, 4 * 131072)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import philox

        n_global = 1000
        lo, hi = shard_range(n_global, rank, world)
        # per-shard statistics as the step kernel would fold them (AsStats layout)
        idx = torch.from_numpy(np.random.RandomState(0).randint(1, 20, n_global))[lo:hi]
        local = torch.zeros(len(STAT_NAMES), dtype=torch.int64)
        local[STAT_NAMES.index("n_envs")] = hi - lo
        local[STAT_NAMES.index("n_reset")] = 1 if rank == 1 else 0
        local[STAT_NAMES.index("sum_target_index")] = int(idx.sum())
        local[STAT_NAMES.index("level")] = 7  # not additive: must not be summed
        red = StatsReducer("cpu").start(local)
        g = red.wait()
        # Philox draws of a shard are the matching rows of the global table
        m_all, n_all = philox.reset_tables(9, 5, np.arange(n_global))
        m_loc, n_loc = philox.reset_tables(9, 5, np.arange(lo, hi))
        ok_philox = np.array_equal(m_loc, m_all[lo:hi]) and np.array_equal(n_loc, n_all[lo:hi])
        # the exchange record of a sharded step closed through the all-reduce route (statistics + grid outcomes)
        rec = torch.zeros(_cabi.EXCHANGE_INT64_WORDS, dtype=torch.int64)
        rec[:len(STAT_NAMES)] = local
        grid = rec[_cabi.STATS_INT64_WORDS:].view(torch.int32)          # attempts[256] then successes[256], uint32
        grid[3] = 5 + rank
        grid[256 + 3] = 2 + rank
        grid[255] = 70000 * (rank + 1)                                   # a high half next to a low half: no carry
        both = all_reduce_exchange(rec.clone(), with_grid=True)
        head_only = all_reduce_exchange(rec.clone(), with_grid=False)
        g2 = both[_cabi.STATS_INT64_WORDS:].view(torch.int32)
        ok_record = (int(g2[3]) == 11 and int(g2[256 + 3]) == 5 and int(g2[255]) == 210000 and int(g2[254]) == 0
                     and int(both[STAT_NAMES.index("level")]) == 7
                     and int(both[STAT_NAMES.index("n_envs")]) == n_global
                     and torch.equal(head_only[_cabi.STATS_INT64_WORDS:], rec[_cabi.STATS_INT64_WORDS:])
                     and torch.equal(head_only[:10], both[:10]))
        q.put((rank, g, red.promotes(g), ok_philox and ok_record, int(idx.sum())))
    finally:
        dist.destroy_process_group()


def test_stats_allreduce_and_global_promotion_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    idx = np.random.RandomState(0).randint(1, 20, 1000)
    for rank, g, promotes, ok_philox, _ in out:
        assert g["n_envs"] == 1000 and g["n_reset"] == 1
        assert g["sum_target_index"] == int(idx.sum())
        assert "level" not in g
        assert promotes == (np.float32(idx.sum()) / np.float32(1000) > 12.0)
        assert ok_philox
    assert out[0][4] + out[1][4] == int(idx.sum())


def test_promotion_rule_needs_a_reset():
    r = StatsReducer("cpu")
    assert not r.promotes({"n_reset": 0, "n_envs": 10, "sum_target_index": 190})
    assert r.promotes({"n_reset": 2, "n_envs": 10, "sum_target_index": 121})
    assert not r.promotes({"n_reset": 2, "n_envs": 10, "sum_target_index": 120})
