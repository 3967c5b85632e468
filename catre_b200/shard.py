"""Batch sharding across the GPUs of one box.

Objects are independent, so the path shards with no data-path collective: rank r refines the contiguous
slice [lo, hi) of the B objects (the split the reference's InferenceSampler uses,
core/utils/my_distributed_sampler.py:190-193) with replicated weights, and ONE all-gather of the
per-object poses ([K+1, B/G, 15] fp32: R 9 + t 3 + s 3) at the end replaces the reference's pickled-object
``all_gather(self._predictions)`` (core/catre/engine/catre_custom_evaluator.py:202-203).
Works on any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_size(total: int, world: int) -> int:
    return (total + world - 1) // world


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank`; trailing ranks may be short or empty when world does not divide total."""
    per = shard_size(total, world)
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def pack_poses(poses: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """[K+1, b, 3, 4], [K+1, b, 3] -> [K+1, b, 15]"""
    k1, b = poses.shape[0], poses.shape[1]
    return torch.cat((poses.reshape(k1, b, 12), scales.reshape(k1, b, 3)), dim=2)


def unpack_poses(packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    k1, b = packed.shape[0], packed.shape[1]
    return packed[:, :, :12].reshape(k1, b, 3, 4), packed[:, :, 12:].reshape(k1, b, 3)


def gather_poses(poses: torch.Tensor, scales: torch.Tensor, total: int,
                 group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the local slice's poses into the full [K+1, total, ...] on every rank.
    The local slice is padded to ceil(total / world) so every rank contributes the same size."""
    world = dist.get_world_size(group)
    per = shard_size(total, world)
    local = pack_poses(poses, scales)
    k1, b = local.shape[0], local.shape[1]
    if b < per:
        local = torch.cat((local, local.new_zeros((k1, per - b, 15))), dim=1)
    local = local.contiguous()
    out = local.new_empty((world * k1, per, 15))  # concatenation along dim 0 (rank-major)
    dist.all_gather_into_tensor(out, local, group=group)
    full = out.reshape(world, k1, per, 15).permute(1, 0, 2, 3).reshape(k1, world * per, 15)[:, :total].contiguous()
    return unpack_poses(full)


def gather_packed(local: torch.Tensor, out: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """ONE collective and nothing else on the stream: ``local`` [per, 15] (the engine's packed final poses of this rank's
    contiguous slice, catre_pack_poses / catre_refine_host_packed) -> ``out`` [world * per, 15] in object order (rank-major =
    slice order).  Both buffers are preallocated by the caller; a short last slice is padded by the caller's buffer."""
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def gather_rows(rows: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather a ragged set of rows: rank r holds [n_r, D]; every rank gets [sum n_r, D] in rank order.
    One size exchange + one padded ``all_gather_into_tensor`` (the tensor replacement for the reference's
    pickled ``all_gather(self._predictions)``, core/catre/engine/catre_evaluator.py:172-177)."""
    if world_size() == 1:
        return rows
    world = dist.get_world_size(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    sizes = torch.empty(world, dtype=torch.int64, device=rows.device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    sizes = sizes.tolist()
    per = max(max(sizes), 1)
    d = rows.shape[1]
    local = rows.new_zeros((per, d))
    local[: rows.shape[0]] = rows
    out = rows.new_empty((world * per, d))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return torch.cat([out[r * per: r * per + sizes[r]] for r in range(world)], dim=0)


def gather_vocab(words, group: Optional[dist.ProcessGroup] = None):
    """Union (rank order, first occurrence first) of each rank's small list of strings (scene ids)."""
    if world_size() == 1:
        return list(dict.fromkeys(words))
    lists = [None] * dist.get_world_size(group)
    dist.all_gather_object(lists, list(dict.fromkeys(words)), group=group)
    return list(dict.fromkeys(w for l in lists for w in l))
