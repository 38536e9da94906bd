"""Multi-GPU sharding of the batch dimension.

The unit of work is an independent (stream, lane): rank r owns lanes {l : l mod R == r}.  There is NO data-path
collective: every rank runs the same single-GPU kernels on its own lanes and the results stay GPU-resident and
sharded (their consumer is per-stream).  torch.distributed is plumbing only:

  * scatter_lanes     — optional ingest: rank `src` holds all PCM and sends every rank its lanes (NCCL send/recv over
                        NVLink on GPUs, gloo on CPU tests);
  * allgather_summary — per-rank frame counts / point counts / checksums, so any rank can prove the whole job ran;
  * gather_columns    — parity-only gather of outputs to one rank (never in a timed path: at kernel rates 8 GPUs
                        emit more bytes/s than one GPU's NVLink ingest can take).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def lanes_for_rank(n_lanes: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_lanes, world))


def owner_of(lane: int, world: int) -> int:
    return lane % world


def scatter_lanes(all_lanes, n_lanes: int, samples: int, src: int = 0, device=None):
    """Rank `src` passes a (n_lanes, samples) float32 tensor (others pass None). Returns this rank's (k, samples) tensor,
    k = len(lanes_for_rank(...)). Uses point-to-point sends so no rank ever holds more than its own share (+ src)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    mine = lanes_for_rank(n_lanes, rank, world)
    out = torch.empty((len(mine), samples), dtype=torch.float32, device=device)
    # NCCL: one batched group of sends, so the transfers to all peers run concurrently over NVLink / NVSwitch instead of
    # being serialised as independent collectives (what unbatched P2P ops are in eager-init mode).
    batched = dist.get_backend() == "nccl"
    if rank == src:
        reqs, ops, keep = [], [], []
        for r in range(world):
            idx = lanes_for_rank(n_lanes, r, world)
            if not idx:
                continue
            part = all_lanes[idx].contiguous()
            if r == src:
                out.copy_(part)
            elif batched:
                keep.append(part)
                ops.append(dist.P2POp(dist.isend, part, r))
            else:
                reqs.append(dist.isend(part, dst=r))
        if ops:
            reqs = dist.batch_isend_irecv(ops)
        for q in reqs:
            q.wait()
    elif mine:
        if batched:
            for q in dist.batch_isend_irecv([dist.P2POp(dist.irecv, out, src)]):
                q.wait()
        else:
            dist.recv(out, src=src)
    return out, mine


def allgather_summary(local: Sequence[int]):
    """All-gathers a small int64 vector per rank (e.g. [frames, points, checksum]). Returns (world, len) numpy array."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(local), dtype=torch.int64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy() for o in outs])


def gather_columns(local_counts: np.ndarray, mine: Sequence[int], n_lanes: int, dst: int = 0):
    """Parity-only: gather per-lane frame point-counts (rows) to `dst` in global lane order. local_counts: (k, F) int."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    payload = (list(mine), np.ascontiguousarray(local_counts))
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    frames = local_counts.shape[1]
    full = np.zeros((n_lanes, frames), local_counts.dtype)
    for idx, rows in gathered:
        for i, l in enumerate(idx):
            full[l] = rows[i]
    return full


def run_sharded(compute: Callable[[np.ndarray], np.ndarray], all_lanes, n_lanes: int, samples: int, src: int = 0):
    """Host-logic reference of the N-rank flow (used by the gloo tests): scatter -> compute on own lanes -> summary.
    `compute(lanes (k,S) float32) -> counts (k,F)`."""
    import torch.distributed as dist

    local, mine = scatter_lanes(all_lanes, n_lanes, samples, src)
    counts = compute(local.cpu().numpy()) if len(mine) else np.zeros((0, 0), np.uint32)
    frames = int(counts.shape[1]) if counts.size else 0
    summary = allgather_summary([len(mine) * frames, int(counts.astype(np.int64).sum()), dist.get_rank()])
    return counts, mine, summary
