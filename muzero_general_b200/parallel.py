"""One process per GPU: game sharding, weight broadcast on refresh, the per-reporting-step counter exchange,
and the multi-GPU self-play entry point.

The reference scales by independent ``SelfPlay`` actors (``muzero.py:177-196``) that pull the newest weights from the
shared storage (``self_play.py:37``) and push games to the replay buffer (``self_play.py:52``).  Here rank r owns the
global game ids ``r*B + slot + k*world*B`` (B = games per rank; ``game_id_stride = world*B``), searches and plays them on its own GPU with no
data-path collective, and

* on every weight refresh rank 0's ``state_dict`` travels to the other ranks as ONE flat fp32 blob
  (``broadcast_weights``: ``ncclBroadcast`` over NVLink on GPUs, gloo in the CPU tests),
* once per reporting step every rank contributes ``{games_finished, env_steps, simulations}`` to ONE all-gather
  (``gather_counters``) so rank 0 can publish ``num_played_games / num_played_steps`` like ``replay_buffer.py:63-65``.

Run it as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        -m muzero_general_b200.parallel --game connect4 --games 8192 --reports 4 --moves-per-report 16

(one rank: ``python -m muzero_general_b200.parallel --game cartpole --games 4096``).  Rank 0 prints one JSON line per
reporting step and a final summary.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy
import torch


def shard_game_ids(rank: int, world: int, games_per_rank: int):
    """Global ids of the games rank ``rank`` owns first (its slots' later games add multiples of world * games_per_rank)."""
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


def broadcast_weights(dist, weights, keys_and_shapes, src: int = 0, device="cpu"):
    """Rank ``src``'s ``state_dict`` (``models.py:69-70``) to every rank as one flat fp32 blob - the replacement for each
    actor's ``shared_storage.get_info("weights")`` RPC (``self_play.py:37``).  ``keys_and_shapes`` is
    ``netspec.weights_spec(spec)``, known on every rank, so only the numbers travel (6 KB .. 2.9 MB).
    ``weights`` may be None on the other ranks.  Returns the ``{key: numpy array}`` dict on every rank."""
    float_keys = [(k, s) for k, s in keys_and_shapes if not k.endswith("num_batches_tracked")]
    total = sum(int(numpy.prod(s)) if len(s) else 1 for _, s in float_keys)
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else src
    if rank == src:
        parts = []
        for k, s in float_keys:
            v = weights[k]
            v = v.detach().cpu().numpy() if hasattr(v, "detach") else numpy.asarray(v)
            assert tuple(v.shape) == tuple(s), (k, v.shape, s)
            parts.append(numpy.ascontiguousarray(v, dtype=numpy.float32).ravel())
        blob = torch.from_numpy(numpy.concatenate(parts)).to(device)
    else:
        blob = torch.empty(total, dtype=torch.float32, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    flat = blob.cpu().numpy()
    out, off = {}, 0
    for k, s in keys_and_shapes:
        if k.endswith("num_batches_tracked"):
            out[k] = numpy.array(0, dtype=numpy.int64)          # carried by the state_dict, unused in eval mode
            continue
        n = int(numpy.prod(s)) if len(s) else 1
        out[k] = flat[off:off + n].reshape(s).copy()
        off += n
    return out


def history_digest(gh) -> str:
    """Content hash of a finished game (actions, visit counts, root values, observations)."""
    import hashlib
    h = hashlib.sha1()
    h.update(numpy.asarray([int(a) for a in gh.action_history], numpy.int64).tobytes())
    h.update(numpy.asarray(gh.child_visits, numpy.float64).tobytes())
    h.update(numpy.asarray(gh.root_values, numpy.float64).tobytes())
    h.update(numpy.asarray(gh.observation_history, numpy.float64).tobytes())
    return h.hexdigest()


def run_selfplay(game: str, total_games: int, reports: int, moves_per_report: int, num_simulations=None, seed=0,
                 temperature=1.0, refresh_every=1, weights_seed=0, emit=print, digests=None):
    """Rank-sharded self-play: ``total_games`` concurrent games split over the ranks of the current process group.
    ``digests``: optional dict filled with ``{global game id: history_digest}`` of the games this rank finished."""
    import torch.distributed as dist
    from . import self_play as sp
    from .games import load_game_module
    from .netspec import netspec_from_config, synthetic_weights, weights_spec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    have_own_gpu = torch.cuda.is_available() and torch.cuda.device_count() > local_rank
    if world > 1 and not dist.is_initialized():
        # NCCL needs one GPU per rank; ranks sharing a GPU (tests on a one-GPU box) talk over gloo
        backend = "nccl" if have_own_gpu and torch.cuda.device_count() >= world else "gloo"
        kw = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    use_dist = dist if world > 1 else None
    gpu = local_rank if have_own_gpu else 0
    if torch.cuda.is_available():
        torch.cuda.set_device(gpu)
    comm_device = torch.device("cuda", gpu) if (use_dist is not None and dist.get_backend() == "nccl") else "cpu"

    assert total_games % world == 0, "games must divide evenly over the ranks"
    B = total_games // world
    mod = load_game_module(game)
    cfg = mod.MuZeroConfig()
    cfg.num_parallel_games, cfg.rng_mode = B, "philox"
    if num_simulations:
        cfg.num_simulations = int(num_simulations)
    spec = netspec_from_config(cfg)
    keys = weights_spec(spec)
    # rank 0 plays the trainer: it owns the weights and publishes a (here: re-seeded synthetic) set on every refresh
    first = broadcast_weights(use_dist, synthetic_weights(spec, weights_seed) if rank == 0 else None, keys, device=comm_device)
    worker = sp.SelfPlay({"weights": first}, mod.Game, cfg, seed=seed, device=gpu, first_game_id=rank * B,
                         game_id_stride=world * B)
    totals_all = [0, 0, 0]
    lines = []
    t_start = time.perf_counter()
    for rep in range(reports):
        if rep and refresh_every and rep % refresh_every == 0:
            w = broadcast_weights(use_dist, synthetic_weights(spec, weights_seed) if rank == 0 else None, keys, device=comm_device)
            worker.model.set_weights(w)                       # self_play.py:37
        games0, steps0 = worker.played_games, worker.env_steps
        t0 = time.perf_counter()
        finished = worker.play_moves(moves_per_report, temperature, cfg.temperature_threshold)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if digests is not None:
            for gh in finished:
                digests[int(gh.game_id)] = history_digest(gh)
        dt = max_over_ranks(use_dist, time.perf_counter() - t0, device=comm_device)
        steps = worker.env_steps - steps0
        table, totals = gather_counters(use_dist, worker.played_games - games0, steps, steps * cfg.num_simulations,
                                        device=comm_device)
        for i in range(3):
            totals_all[i] += totals[i]
        line = {"report": rep, "world": world, "games_finished": totals[0], "env_steps": totals[1], "simulations": totals[2],
                "seconds": dt, "env_steps_per_s": totals[1] / dt, "per_rank": table, "path": worker.loop_path,
                "finished_on_this_rank": len(finished)}
        lines.append(line)
        if rank == 0 and emit:
            emit(json.dumps(line))
    wall = max_over_ranks(use_dist, time.perf_counter() - t_start, device=comm_device)
    summary = {"summary": True, "game": game, "world": world, "games_per_rank": B, "num_simulations": cfg.num_simulations,
               "num_played_games": totals_all[0], "num_played_steps": totals_all[1], "simulations": totals_all[2],
               "seconds": wall, "env_steps_per_s": totals_all[1] / wall, "numerics": worker.model.engine.numerics
               if hasattr(worker.model.engine, "numerics") else None}
    if rank == 0 and emit:
        emit(json.dumps(summary))
    return worker, lines, summary


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--game", default="cartpole")
    ap.add_argument("--games", type=int, default=4096, help="concurrent games over ALL ranks")
    ap.add_argument("--reports", type=int, default=4)
    ap.add_argument("--moves-per-report", type=int, default=16)
    ap.add_argument("--simulations", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--refresh-every", type=int, default=1, help="reports between two weight broadcasts (0 = never)")
    ap.add_argument("--dump-histories", default=None, help="directory: every rank writes content hashes of the games it finished (tests)")
    args = ap.parse_args(argv)
    import torch.distributed as dist
    digests = {} if args.dump_histories else None
    worker, lines, summary = run_selfplay(args.game, args.games, args.reports, args.moves_per_report, args.simulations,
                                          args.seed, args.temperature, args.refresh_every, digests=digests)
    if args.dump_histories:
        os.makedirs(args.dump_histories, exist_ok=True)
        rank = int(os.environ.get("RANK", "0"))
        with open(os.path.join(args.dump_histories, f"rank{rank}.json"), "w") as f:
            json.dump({"summary": summary, "lines": lines, "digests": {str(k): v for k, v in digests.items()}}, f)
    worker.close()
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
