"""world_size = 2 on the gloo backend (CPU): sharding, the counter all-gather, max-over-ranks timing, and
world-size invariance of the games themselves (host logic with the oracle-backed test double)."""
import os
import socket

import numpy
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from conftest import weights_for
    from fake_engine import FakeSearchEngine
    from muzero_general_b200 import parallel, self_play as sp
    from muzero_general_b200.games import load_game_module
    from muzero_general_b200.netspec import netspec_from_config
    sp.SearchEngine = FakeSearchEngine

    B = 2
    ids = parallel.shard_game_ids(rank, world, B)
    mod = load_game_module("tictactoe")
    cfg = mod.MuZeroConfig()
    cfg.num_simulations, cfg.num_parallel_games = 6, B
    spec = netspec_from_config(cfg)
    worker = sp.SelfPlay({"weights": weights_for("tictactoe", spec)}, mod.Game, cfg, seed=5, first_game_id=ids[0])
    games = worker.play_games(B, 1.0)[:B]
    steps = sum(len(g.action_history) - 1 for g in games)
    table, totals = parallel.gather_counters(dist, len(games), steps, steps * cfg.num_simulations)
    slow = parallel.max_over_ranks(dist, 1.0 + rank)
    torch.save(dict(ids=ids, table=table, totals=totals, slow=slow,
                    histories=[[int(a) for a in g.action_history] for g in games],
                    root_values=[list(g.root_values) for g in games]), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_games_and_exchange_counters(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"rank{k}.pt") for k in range(world)]
    assert r[0]["ids"] == [0, 1] and r[1]["ids"] == [2, 3]                      # disjoint, covering
    assert r[0]["table"] == r[1]["table"] and len(r[0]["table"]) == 2            # every rank sees the same table
    assert r[0]["totals"][0] == 4 and r[0]["totals"][2] == r[0]["totals"][1] * 6
    assert r[0]["slow"] == r[1]["slow"] == 2.0                                   # max over ranks

    # world-size invariance: the same four global games played by ONE rank with a batch of four
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import weights_for
    from fake_engine import FakeSearchEngine
    from muzero_general_b200 import self_play as sp
    from muzero_general_b200.games import load_game_module
    from muzero_general_b200.netspec import netspec_from_config
    old = sp.SearchEngine
    sp.SearchEngine = FakeSearchEngine
    try:
        mod = load_game_module("tictactoe")
        cfg = mod.MuZeroConfig()
        cfg.num_simulations, cfg.num_parallel_games = 6, 4
        spec = netspec_from_config(cfg)
        worker = sp.SelfPlay({"weights": weights_for("tictactoe", spec)}, mod.Game, cfg, seed=5)
        single = worker.play_games(8, 1.0)
    finally:
        sp.SearchEngine = old
    played = {tuple(int(a) for a in g.action_history): list(g.root_values) for g in single}
    for k in range(world):
        for hist, rv in zip(r[k]["histories"], r[k]["root_values"]):
            assert tuple(hist) in played and played[tuple(hist)] == rv


def _bcast_main(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from muzero_general_b200 import parallel
    from muzero_general_b200.games import load_game_module
    from muzero_general_b200.netspec import netspec_from_config, synthetic_weights, weights_spec
    spec = netspec_from_config(load_game_module("tictactoe").MuZeroConfig())
    mine = synthetic_weights(spec, 3) if rank == 0 else None          # only the "trainer" rank has the weights
    got = parallel.broadcast_weights(dist, mine, weights_spec(spec))
    torch.save({k: torch.from_numpy(numpy.asarray(v)) for k, v in got.items()}, os.path.join(out_dir, f"w{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_on_refresh(tmp_path):
    """Rank 0's state_dict reaches every rank bit for bit as one flat blob (the replacement for self_play.py:37)."""
    world, port = 2, _free_port()
    mp.spawn(_bcast_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from muzero_general_b200.games import load_game_module
    from muzero_general_b200.netspec import netspec_from_config, synthetic_weights
    want = synthetic_weights(netspec_from_config(load_game_module("tictactoe").MuZeroConfig()), 3)
    for k in range(world):
        got = torch.load(tmp_path / f"w{k}.pt")
        assert list(got) == list(want)
        for key, v in want.items():
            if not key.endswith("num_batches_tracked"):
                assert numpy.array_equal(got[key].numpy(), v), key
