"""The N > 1 path on CPU: two processes over gloo exercise the partitioning, the per-step
position exchange of the i-block-sharded cloud and the counter reduction (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyqmd_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chunk = sharding.cloud_chunk(n, world)
        lo, hi = sharding.shard_range(n, rank, world)
        # every rank starts from the same replica; each advances only its own i-block
        g = torch.Generator().manual_seed(5)
        replica = torch.zeros(chunk * world, 2)
        replica[:n] = torch.rand(n, 2, generator=g)
        nxt = torch.full_like(replica, -1.0)
        for step in range(3):
            nxt[lo:hi] = replica[lo:hi] * 2.0 + (step + 1)           # "integrate" own block
            sharding.allgather_positions(nxt, rank, world, chunk)
            replica, nxt = nxt, replica
        # single-process reference of the same 3 steps
        g = torch.Generator().manual_seed(5)
        ref = torch.rand(n, 2, generator=g)
        for step in range(3):
            ref = ref * 2.0 + (step + 1)
        ok = torch.equal(replica[:n], ref)
        # counters: decays per mode summed over ranks; device time = max over ranks
        c = torch.tensor([rank + 1, 10 * (rank + 1)], dtype=torch.int64)
        sharding.sum_counters(c)
        t = sharding.max_over_ranks(0.5 + rank, "cpu")
        out.put((rank, bool(ok), c.tolist(), t, (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11, 4097])
def test_cloud_exchange_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    spans = []
    for rank, ok, c, t, span in res:
        assert ok, f"rank {rank}: replica differs from the single-process result"
        assert c == [3, 30] and t == 1.5
        spans.append(span)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == n


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 8, 1000, 65536, 1_000_000):
        for w in (1, 2, 4, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert sharding.cloud_chunk(n, w) * w >= n


def test_global_ids_make_sharding_invisible_to_the_rng():
    """Philox counters use the global nucleus id, so a nucleus draws the same uniforms whichever
    rank owns it (checked with the oracle's Philox; the GPU suite checks the kernels)."""
    from oracle import oracle as orc
    n, world = 1000, 4
    whole = orc.philox_uniforms(99, 0, n, 3, 0)
    parts = []
    for r in range(world):
        lo, hi = sharding.shard_range(n, r, world)
        parts.append(orc.philox_uniforms(99, lo, hi - lo, 3, 0))
    assert np.array_equal(np.concatenate(parts), whole)


def test_symmetric_cloud_rows_partition_and_balance():
    """Rows of the symmetric cloud scheme: every row owned by exactly one part, work balanced."""
    from pyqmd_b200.sharding import sym_row_work, sym_rows_of
    for n in (1, 1000, 4096, 200_000, 1_000_000):
        nb = (n + 1023) // 1024
        for parts in (1, 2, 3, 4, 8):
            owned = [sym_rows_of(p, parts, n) for p in range(parts)]
            flat = sorted(b for rows in owned for b in rows)
            assert flat == list(range(nb)), (n, parts)
            if n >= 200_000:
                work = [sum(sym_row_work(b, n) for b in rows) for rows in owned]
                assert max(work) <= 1.02 * (sum(work) / parts), (n, parts, work)


def _pair(i, j):
    """An antisymmetric integer 'pair force' (stands in for the fixed-point force of nucleon j on i)."""
    h = (np.minimum(i, j) * 2654435761 + np.maximum(i, j) * 40503) % 2001 - 1000
    return np.where(i < j, h, -h) * (i != j)


def _sym_worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chunk = sharding.cloud_chunk(n, world)
        acc = torch.zeros(chunk * world, 2, dtype=torch.int64)
        a = acc.numpy()
        # this rank's share of the symmetric scheme, restated on the host (csrc/cloud_sym.cu):
        # row b = i in [1024 b, 1024 b + 1024); diagonal tiles ordered, later tiles with reaction
        for b in sharding.sym_rows_of(rank, world, n):
            i = np.arange(1024 * b, min(1024 * b + 1024, n))
            diag, later = sharding.sym_row_tiles(b, n)
            for t in diag:
                j = np.arange(256 * t, min(256 * t + 256, n))
                f = _pair(i[:, None], j[None, :])
                a[i, 0] += f.sum(1)
            for t in later:
                j = np.arange(256 * t, min(256 * t + 256, n))
                f = _pair(i[:, None], j[None, :])
                a[i, 0] += f.sum(1)
                a[j, 0] -= f.sum(0)
        a[:, 1] = 3 * a[:, 0]
        mine = torch.zeros(chunk, 2, dtype=torch.int64)
        sharding.reduce_scatter_forces(acc, mine, rank, world)
        lo, hi = sharding.shard_range(n, rank, world)
        idx = np.arange(n)
        want = _pair(idx[lo:hi, None], idx[None, :]).sum(1)
        ok = np.array_equal(mine.numpy()[: hi - lo, 0], want) and \
            np.array_equal(mine.numpy()[: hi - lo, 1], 3 * want)
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [700, 2500, 4096 + 300])
def test_symmetric_cloud_force_exchange_world2(n):
    """Rows dealt to two ranks + exact int64 reduce-scatter reproduce the all-pairs sums: every
    unordered pair is counted exactly once across the ranks (gloo, no GPU)."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sym_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
