"""One full training iteration on N GPUs (cfg 5): weight broadcast -> sharded self-play ->
sample gather -> the value-loss update of ``Trainer.train_network`` on the trainer rank.

The reference's iteration is ``Trainer.collect_self_play_data`` + ``Trainer.train_network``
(trainer.py:147-362) on one process with a CPU worker pool; that code runs UNCHANGED on the
drop-in classes (tests/test_reference_consumers.py).  This module is the same iteration for one
process per GPU: games are sharded over the ranks with no collective inside self-play, the
samples stay on the device from the game loop to the optimizer (no Python tuples, no replay
deque), and the only collectives are the two per-iteration ones of ``dist.py``.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch
import torch.distributed as dist

from .dist import distributed_self_play
from .samples import training_batch

BATCH_SIZE = 64       # config.py:34
MAX_BATCHES = 50      # trainer.py:309


def train_on_samples(network: torch.nn.Module, optimizer: torch.optim.Optimizer,
                     samples: Dict[str, torch.Tensor], batch_size: int = BATCH_SIZE,
                     max_batches: int = MAX_BATCHES, generator: Optional[torch.Generator] = None) -> float:
    """``Trainer.train_network`` (trainer.py:298-362) on device-resident samples: per batch, 64
    samples drawn without replacement (``ReplayBuffer.sample``, :35-41), planes of
    ``encode_board(board, 1)`` (:316-319, the player flag is hard-wired to 1 there), MSE between
    the value head and the shaped reward (:328-333), gradient-norm clipping at 1.0 (:340), Adam
    step.  Returns the mean loss, like the reference."""
    n = int(samples["reward"].shape[0])
    num_batches = min(max_batches, n // batch_size)
    if num_batches == 0:
        return float("nan")
    network.train()
    dev = samples["board"].device
    total = torch.zeros((), dtype=torch.float32, device=dev)
    for _ in range(num_batches):
        idx = torch.randperm(n, device=dev, generator=generator)[:batch_size]
        states, target = training_batch(samples, idx)
        _, pred = network(states)
        loss = torch.nn.functional.mse_loss(pred, target)
        optimizer.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(network.parameters(), max_norm=1.0)
        optimizer.step()
        total += loss.detach()
    return float(total) / num_batches


def self_play_iteration(network: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer],
                        num_games: int, num_simulations: int, temperature: float = 1.0, seed: int = 0,
                        net_dtype: torch.dtype = torch.float32, trainer_rank: int = 0) -> Dict[str, float]:
    """All ranks call this.  Returns timings (seconds, host wall clock with device syncs at the
    phase boundaries) and counts; ``loss`` on the trainer rank only."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = next(network.parameters()).device
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    samples, sp = distributed_self_play(network, num_games, num_simulations, temperature, seed=seed,
                                        network=network, dst=trainer_rank, net_dtype=net_dtype)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    loss = None
    if rank == trainer_rank and optimizer is not None:
        loss = train_on_samples(network, optimizer, samples)
        torch.cuda.synchronize(dev)
    t2 = time.perf_counter()
    st = sp.stats()
    return {"self_play_s": t1 - t0, "train_s": t2 - t1, "seconds": t2 - t0, "plies": st["plies"],
            "sims": st["sims"], "samples": 0 if samples is None else int(samples["reward"].shape[0]),
            "loss": loss}
