"""Pins the C tree oracle (oracle/tree_oracle.c) against the Python oracle and the reference traces."""
import numpy
import pytest

from conftest import golden_json
from helpers import oracle_replay, random_teacher, teacher_from_cases
from oracle import build_c
from oracle import mcts as om


@pytest.mark.parametrize("name", ["cartpole_synth", "cartpole_pretrained", "tictactoe", "connect4", "breakout",
                                  "connect4_n200", "breakout_n50", "gomoku"])
def test_c_oracle_reproduces_reference_traces(name, game_configs):
    cfg = game_configs[name.split("_")[0]]
    A = len(cfg.action_space)
    for c in golden_json(f"mcts_{name}.json"):
        N = c["num_simulations"]
        t, legal, noise, first, to_play = teacher_from_cases([c], A, N)
        r = build_c.tree_search(1, N, A, len(cfg.players), cfg.discount, cfg.pb_c_base, cfg.pb_c_init,
                                cfg.root_exploration_fraction, legal, to_play, noise if c["add_noise"] else None, first,
                                cfg.seed, None, None, t)
        assert [int(r["visit_counts"][0, a]) for a in c["root_actions"]] == c["root_visits"]
        assert r["root_value"][0] == c["root_value"] and r["max_depth"][0] == c["max_tree_depth"]
        assert [[int(a) for a in r["actions"][0, s, :r["depth"][0, s]]] for s in range(N)] == [s["actions"] for s in c["sims"]]


@pytest.mark.parametrize("game,N,n", [("cartpole", 50, 40), ("tictactoe", 30, 24), ("connect4", 40, 12), ("gomoku", 40, 6)])
def test_c_oracle_matches_python_oracle(game, N, n, game_configs):
    cfg = game_configs[game]
    A, P = len(cfg.action_space), len(cfg.players)
    rs = numpy.random.RandomState(5)
    legal = (rs.uniform(size=(n, A)) < 0.7).astype(numpy.uint8)
    legal[numpy.arange(n), rs.randint(0, A, n)] = 1
    t = random_teacher(rs, n, N, A, reward_scale=1.0 if P == 1 else 0.0, legal=legal)
    if (A & (A - 1)) == 0:
        t["priors"][:4] = numpy.float32(1.0 / A); t["value"][:4] = 0; t["reward"][:4] = 0
    noise = rs.dirichlet([cfg.root_dirichlet_alpha] * A, size=n)
    to_play = rs.randint(0, P, n).astype(numpy.int32)
    gid = (77 + numpy.arange(n)).astype(numpy.int64)
    mv = rs.randint(0, 9, n).astype(numpy.int32)
    r = build_c.tree_search(n, N, A, P, cfg.discount, cfg.pb_c_base, cfg.pb_c_init, cfg.root_exploration_fraction,
                            legal, to_play, noise, None, cfg.seed, gid, mv, t)
    params = om.SearchParams.from_config(cfg, N)
    for i in range(n):
        acts = [a for a in range(A) if legal[i, a]]
        res, draws = oracle_replay(params, acts, int(to_play[i]),
                                   (t["root_value"][i], t["root_reward"][i], [t["root_priors"][i, a] for a in acts]),
                                   [(t["value"][i, s], t["reward"][i, s], t["priors"][i, s]) for s in range(N)],
                                   [noise[i, a] for a in acts], None, seed=cfg.seed, game=int(gid[i]), move=int(mv[i]))
        assert [int(r["visit_counts"][i, a]) for a in acts] == res.root_visits
        assert r["root_value"][i] == res.root_value and r["ties"][i] == draws.later_ties
        assert (r["range"][i, 0], r["range"][i, 1]) == (res.range_lo, res.range_hi)
        assert [[int(a) for a in r["actions"][i, s, :r["depth"][i, s]]] for s in range(N)] == [s.path_actions for s in res.sims]
