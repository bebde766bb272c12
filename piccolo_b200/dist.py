"""Multi-GPU sharding (one process per GPU, torch.distributed).  The path shards naturally: every
(pose, point) term is independent, so the only exchange is a tiny all-gather of per-pose losses after
scoring and of per-candidate (loss, pose) rows before the arg-min.  Backend: NCCL over
NVLink on the GPU box; gloo in the CPU tests, where the local compute is injected."""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int):
    """Contiguous slice [lo, hi) of n units for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _all_gather_stack(local: torch.Tensor) -> torch.Tensor:
    """(…) per rank -> (world, …) on every rank: ONE collective into one preallocated tensor (the list API of
    dist.all_gather allocates and copies one tensor per rank).  The output is passed in its concatenated form,
    which both NCCL and gloo accept."""
    ws = dist.get_world_size()
    flat = local.contiguous().reshape(-1)
    out = torch.empty(ws * flat.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat)
    return out.reshape((ws,) + tuple(local.shape))


def _all_gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather variable-length row blocks (contiguous shards) into the full (n_total, ...) tensor."""
    rank, ws = world()
    if ws == 1:
        return local
    rows = max(shard_bounds(n_total, r, ws)[1] - shard_bounds(n_total, r, ws)[0] for r in range(ws))
    pad = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = _all_gather_stack(pad)
    parts = []
    for r in range(ws):
        lo, hi = shard_bounds(n_total, r, ws)
        parts.append(out[r][: hi - lo])
    return torch.cat(parts, dim=0)


def score_sharded(score_fn: Callable[[torch.Tensor], torch.Tensor], poses: torch.Tensor, width: int = 1) -> torch.Tensor:
    """Every rank scores its contiguous slice of the flattened pose grid (index i*R+j) and all ranks end up
    with the full loss vector, so the deterministic top-K (ties -> lower index) is identical everywhere.
    score_fn maps n units to n*width losses: poses (n,6) with width 1, or the translations (n,3) of a structured
    grid with width R (whole rows of the loss table per rank).  Returns the flat (len(poses)*width,) vector."""
    rank, ws = world()
    lo, hi = shard_bounds(poses.shape[0], rank, ws)
    local = score_fn(poses[lo:hi]).reshape(hi - lo, width) if hi > lo else poses.new_zeros((0, width))
    return _all_gather_rows(local, poses.shape[0]).reshape(-1)


def rerank_sharded(blocks_fn: Callable[[torch.Tensor], tuple], finish_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                   poses: torch.Tensor) -> torch.Tensor:
    """Histogram re-rank with the K candidates sharded: every rank renders and histograms a contiguous slice
    (blocks_fn: (k,6) -> (rows (k, 2*nblk), ngt (nblk,))), the per-block rows are all-gathered, and every rank replays the
    reference's candidate loop over ALL rows in order (finish_fn) — its table persists from one candidate to the next
    (utils.py:547-579), so that part cannot be split.  Returns the (K,) scores on every rank."""
    rank, ws = world()
    lo, hi = shard_bounds(poses.shape[0], rank, ws)
    rows, ngt = blocks_fn(poses[lo:hi])
    return finish_fn(_all_gather_rows(rows, poses.shape[0]), ngt)


def refine_sharded(refine_fn: Callable[[torch.Tensor], torch.Tensor], starts: torch.Tensor) -> torch.Tensor:
    """Candidates are dealt round-robin; each rank refines its own with zero communication, then one
    all-gather of the (loss, pose) rows.  refine_fn maps (b,6) start poses -> (b,7) rows [loss, pose].
    Returns the (B,7) table in candidate order on every rank."""
    rank, ws = world()
    B = starts.shape[0]
    mine = list(range(rank, B, ws))
    local = refine_fn(starts[mine]) if mine else starts.new_zeros((0, 7))
    if ws == 1:
        return local
    rows = (B + ws - 1) // ws
    pad = torch.full((rows, 7), float("nan"), dtype=starts.dtype, device=starts.device)
    pad[: local.shape[0]] = local
    out = _all_gather_stack(pad)
    table = torch.empty((B, 7), dtype=starts.dtype, device=starts.device)
    for r in range(ws):
        idx = list(range(r, B, ws))
        if idx:
            table[idx] = out[r][: len(idx)]
    return table


_PEER_COMM = None


def peer_comm():
    """The process group's PeerComm (created on first use; the 64-byte IPC handles travel through
    torch.distributed.all_gather_object, i.e. the host side of NCCL/gloo — set-up only)."""
    global _PEER_COMM
    if _PEER_COMM is None:
        from . import engine
        rank, ws = world()

        def exchange(mine: bytes):
            out = [None] * ws
            dist.all_gather_object(out, mine)
            return out
        _PEER_COMM = engine.PeerComm(rank, ws, exchange)
    return _PEER_COMM


def argmin_candidate(table: torch.Tensor):
    """Arg-min over the gathered (loss, pose) rows; NaN losses lose; ties -> lower index."""
    loss = torch.where(torch.isnan(table[:, 0]), torch.full_like(table[:, 0], float("inf")), table[:, 0])
    k = int(torch.argmin(loss))
    return k, table[k, 1:], table[k, 0]


def gather_results(row: torch.Tensor) -> torch.Tensor:
    """Multi-query mode: each rank contributes one result row; returns (world, len(row))."""
    rank, ws = world()
    if ws == 1:
        return row[None]
    return _all_gather_stack(row)
