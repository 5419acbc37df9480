"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed).

The reference (OtsoBear/PyQMD) is single-device; this is new work:

* ensembles / decay populations shard by nucleus -- contiguous blocks, global nucleus ids key
  the RNG, no data-path collective; only counters are summed at reporting time;
* one big nucleon cloud shards by i-block: every rank holds a full replica of the positions,
  advances its own block [rank*chunk, (rank+1)*chunk) and the new positions are all-gathered
  once per step (8 B per nucleon).

These helpers are device-agnostic (they run on CPU tensors over gloo in the CPU test-suite and
on CUDA tensors over NCCL/NVLink in production).
"""
from __future__ import annotations

import torch


def shard_range(n: int, rank: int, world: int):
    """Contiguous block of units owned by ``rank``: [lo, hi)."""
    chunk = (n + world - 1) // world
    lo = min(rank * chunk, n)
    return lo, min(lo + chunk, n)


def cloud_chunk(n: int, world: int) -> int:
    """Nucleons per rank; replicas are padded to chunk * world rows so the all-gather is regular."""
    return (n + world - 1) // world


CLOUD_ROW = 1024     # nucleons per i-block row of the symmetric cloud scheme (csrc/cloud.cuh kIBlock)


def sym_rows_of(part: int, n_parts: int, n: int):
    """i-block rows of the symmetric cloud scheme that ``part`` evaluates (csrc/cloud_sym.cu): rows
    are dealt boustrophedon-wise -- row g*n_parts + (part if g even else n_parts-1-part) -- so that
    the triangular work (row b meets the tiles at and after its diagonal) is balanced."""
    nb = (n + CLOUD_ROW - 1) // CLOUD_ROW
    rows = []
    g = 0
    while True:
        b = g * n_parts + (n_parts - 1 - part if g & 1 else part)
        if g * n_parts >= nb:
            break
        if b < nb:
            rows.append(b)
        g += 1
    return rows


def sym_row_work(b: int, n: int) -> int:
    """j-tiles (256 nucleons) row ``b`` visits: its 4 diagonal tiles and every tile after them."""
    nt = (n + 255) // 256
    return max(nt - 4 * b, 0)


def sym_row_tiles(b: int, n: int):
    """(diagonal tiles, off-diagonal tiles) of row ``b``: the row's own 4 tiles are evaluated as
    ordered pairs without reaction, every later tile once per unordered pair with the reaction
    added to the j side (csrc/cloud_sym.cu)."""
    nt = (n + 255) // 256
    first = 4 * b
    return list(range(first, min(first + 4, nt))), list(range(first + 4, nt))


def reduce_scatter_forces(acc: torch.Tensor, mine: torch.Tensor, rank: int, world: int, group=None):
    """Exact sum over ranks of the int64 fixed-point force accumulators; rank r receives the rows
    [r*chunk, (r+1)*chunk) in ``mine``.  NCCL: reduce_scatter_tensor; backends without it (gloo in
    the CPU test-suite): all_reduce + slice."""
    import torch.distributed as dist
    chunk = mine.shape[0]
    if world == 1:
        mine.copy_(acc[:chunk])
        return mine
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(mine, acc, op=dist.ReduceOp.SUM, group=group)
    else:
        tmp = acc.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        mine.copy_(tmp[rank * chunk:(rank + 1) * chunk])
    return mine


def allgather_positions(replica: torch.Tensor, rank: int, world: int, chunk: int, group=None):
    """In-place all-gather of the rows each rank owns in its [chunk * world, 2] replica."""
    if world == 1:
        return replica
    import torch.distributed as dist
    mine = replica[rank * chunk:(rank + 1) * chunk]
    dist.all_gather_into_tensor(replica, mine, group=group)
    return replica


def sum_counters(t: torch.Tensor, group=None):
    """Sum of per-rank statistics (decays by mode, survivors ...) at reporting time."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(seconds: float, device, group=None) -> float:
    """Device-timed duration as the max over ranks (never wall clock)."""
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
