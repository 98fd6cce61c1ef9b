"""Multi-GPU plumbing for self-play: one process per GPU, games sharded statically, no collective
on the hot path.  Collectives (NCCL over NVLink; gloo in CPU tests) run once per iteration only:
the weight broadcast that replaces the per-game state_dict pickle of self_play.py:386-395, and the
sample gather that replaces the pickled results of :404-408."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_games: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of game ids owned by ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(int(n_games), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_weights(module: torch.nn.Module, src: int = 0) -> int:
    """Broadcast every parameter and buffer from ``src`` as ONE flat buffer per dtype
    (24,634,141 fp32 parameters = 98.5 MB for ChessNet).  Returns the bytes sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    groups: Dict[torch.dtype, list] = {}
    for t in list(module.parameters()) + list(module.buffers()):
        groups.setdefault(t.dtype, []).append(t)
    sent = 0
    for dtype, tensors in groups.items():
        flat = torch.cat([t.detach().reshape(-1) for t in tensors])
        dist.broadcast(flat, src=src)
        sent += flat.numel() * flat.element_size()
        off = 0
        with torch.no_grad():
            for t in tensors:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
    return sent


def gather_samples(samples: Dict[str, torch.Tensor], dst: int = 0) -> Optional[Dict[str, torch.Tensor]]:
    """Concatenate per-rank sample tensors (same keys, ragged first dimension) on ``dst`` in rank
    order.  One all_gather of counts + one padded all_gather per key."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return samples
    world, rank = dist.get_world_size(), dist.get_rank()
    first = next(iter(samples.values()))
    cnt = torch.tensor([first.shape[0]], dtype=torch.int64, device=first.device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c) for c in counts]
    mx = max(counts)
    out = {}
    for k, t in samples.items():
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        if rank == dst:
            out[k] = torch.cat([b[:c] for b, c in zip(bufs, counts)])
    return out if rank == dst else None


def distributed_self_play(evaluator, num_games: int, num_simulations: int, temperature: float = 1.0,
                          seed: int = 0, network: Optional[torch.nn.Module] = None, dst: int = 0,
                          opponent=None):
    """One self-play phase on all ranks (one process per GPU): broadcast ``network``'s weights
    from ``dst`` (if given), play this rank's block of game ids as one device batch, gather the
    training samples on ``dst``.  Replaces parallel_self_play's process pool + pickles
    (self_play.py:386-408); there is no collective inside the game loop.
    Returns (samples dict on ``dst`` / None elsewhere, this rank's BatchedSelfPlay)."""
    from .samples import training_tensors
    from .self_play import BatchedSelfPlay
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if network is not None:
        broadcast_weights(network, src=dst)
    lo, hi = shard_range(num_games, rank, world)
    sp = BatchedSelfPlay(evaluator, hi - lo, num_simulations, temperature, opponent_network=opponent,
                         seed=seed, first_game_id=lo)
    sp.play()
    mine = training_tensors(sp, red_only=opponent is not None)
    mine["game"] = mine["game"] + lo
    return gather_samples(mine, dst=dst), sp
