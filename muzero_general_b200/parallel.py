"""One process per GPU: game sharding and the per-reporting-step counter exchange.

The reference scales by independent ``SelfPlay`` actors (``muzero.py:177-188``); here rank r owns the
global game ids ``r*B .. r*B+B-1`` (B = ``num_parallel_games``), searches them on its own GPU with no
data-path collective, and once per reporting step every rank contributes
``{games_finished, env_steps, simulations}`` to ONE all-gather (NCCL over NVLink on GPUs, gloo in the
CPU tests) so rank 0 can publish ``num_played_games / num_played_steps`` like ``replay_buffer.py:63-65``.
"""
from __future__ import annotations

import torch


def shard_game_ids(rank: int, world: int, games_per_rank: int):
    """Global ids of the games rank ``rank`` owns."""
    assert 0 <= rank < world
    return list(range(rank * games_per_rank, (rank + 1) * games_per_rank))


def gather_counters(dist, games: int, env_steps: int, simulations: int, device="cpu"):
    """ONE all-gather per reporting step; returns ([per-rank triples], totals)."""
    mine = torch.tensor([games, env_steps, simulations], dtype=torch.int64, device=device)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        rows = [mine]
    else:
        rows = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(rows, mine)
    table = [[int(v) for v in r.tolist()] for r in rows]
    totals = [sum(col) for col in zip(*table)]
    return table, totals


def max_over_ranks(dist, seconds: float, device="cpu") -> float:
    """A multi-GPU time is the slowest rank's time."""
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
