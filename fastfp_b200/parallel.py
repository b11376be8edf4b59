"""Multi-GPU plumbing: frequency sharding + one all-gather.

Every (frequency, draw) output is independent, so the packed pulsar arrays are replicated on
every rank, the frequency axis is cut into ``world_size`` contiguous shards, each rank sweeps its
shard for all pulsars, and the per-bin values are assembled with a single
``all_gather_into_tensor`` (NCCL over NVLink on GPUs; gloo in the CPU tests). There is no other
communication on the plain-Fp path. The noise-marginalised path shards the DRAW axis (``sharded_draws``) and, in its
two-dimensional form (``NMFP.calculate_nmfp_2d``), also the frequency axis of its draw-independent stage, whose tiles
are exchanged with one more all-gather (``tile_blocks``). One process per GPU (``torchrun``); ranks read
``RANK``/``LOCAL_RANK``/``WORLD_SIZE`` from the environment.
"""
from __future__ import annotations

from typing import Callable, Tuple


def shard_bounds(F: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Contiguous shard ``[lo, hi)`` of ``F`` frequencies for ``rank`` and the common padded
    shard length ``per`` (``ceil(F / world)``; the last shards may be short or empty)."""
    if F < 0 or world < 1 or not 0 <= rank < world:
        raise ValueError("need F >= 0 and 0 <= rank < world")
    per = -(-F // world) if F else 0
    lo = min(F, rank * per)
    hi = min(F, lo + per)
    return lo, hi, per


def tile_blocks(F: int, world: int) -> Tuple[int, int]:
    """Two-dimensional sharding of the noise-marginalised sweep: the draw-independent stage works in tiles of 32
    frequencies; returns ``(nt, per)`` = the number of tiles of ``F`` frequencies and the (even) number of tiles every
    rank computes, rank ``r`` owning tiles ``[r * per, (r + 1) * per)`` (those beyond ``nt`` are padding)."""
    if F < 1 or world < 1:
        raise ValueError("need F >= 1 and world >= 1")
    nt = -(-F // 32)
    per = -(-nt // world)
    per += per & 1
    return nt, per


def sharded_sweep(local_fn: Callable, freqs, group=None, lead_shape=()):
    """Run ``local_fn(freqs[lo:hi]) -> tensor(*lead_shape, hi-lo)`` on every rank and all-gather.

    ``freqs`` is a 1-D torch tensor (same on every rank). Returns the full
    ``(*lead_shape, F)`` result on every rank. With one process it is just ``local_fn(freqs)``.
    The frequency axis is gathered shard-major, so the device buffer is ``(world, *lead, per)``
    and is re-laid to ``(*lead, F)`` (a view for the plain-Fp case ``lead_shape == ()``)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local_fn(freqs)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    F = int(freqs.shape[0])
    lo, hi, per = shard_bounds(F, rank, world)
    local = local_fn(freqs[lo:hi])
    send = torch.zeros(*lead_shape, per, dtype=local.dtype, device=local.device)
    send[..., : hi - lo] = local
    recv = torch.empty(world, *lead_shape, per, dtype=local.dtype, device=local.device)
    # flat views: the gloo backend (CPU tests) only accepts the concatenated 1-D form
    dist.all_gather_into_tensor(recv.view(-1), send.contiguous().view(-1), group=group)
    if lead_shape == ():
        return recv.reshape(world * per)[:F]
    # (world, *lead, per) -> (*lead, world*per)
    nd = len(lead_shape)
    perm = list(range(1, nd + 1)) + [0, nd + 1]
    return recv.permute(*perm).reshape(*lead_shape, world * per)[..., :F]


def sharded_draws(local_fn: Callable, D: int, group=None):
    """Noise-marginalised sweeps shard the DRAW axis instead (every stage of that path -- the
    per-draw factorisation included -- then splits across ranks; only the small per-frequency
    stage is repeated). ``local_fn(lo, hi) -> tensor(hi-lo, F)`` runs on every rank for its
    contiguous draw range; the ``(D, F)`` result is assembled on every rank with one
    ``all_gather_into_tensor`` (draw-major output, so the gathered buffer needs no re-layout)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local_fn(0, D)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi, per = shard_bounds(D, rank, world)
    local = local_fn(lo, hi)
    F = int(local.shape[-1])
    send = local if hi - lo == per else torch.cat(
        [local, torch.zeros(per - (hi - lo), F, dtype=local.dtype, device=local.device)])
    recv = torch.empty(world * per, F, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(-1), send.contiguous().view(-1), group=group)
    return recv[:D]
