"""The reference-shaped SelfPlay / MCTS surface on the real device engine (C ABI underneath)."""
import numpy
import pytest

from conftest import golden_json, golden_npz, weights_for
from muzero_general_b200.games import load_game_module
from muzero_general_b200.netspec import netspec_from_config

pytestmark = pytest.mark.gpu


def _worker(name, seed, **over):
    from muzero_general_b200 import self_play as sp
    mod = load_game_module(name)
    cfg = mod.MuZeroConfig()
    for k, v in over.items():
        setattr(cfg, k, v)
    spec = netspec_from_config(cfg)
    return sp.SelfPlay({"weights": weights_for(name, spec)}, mod.Game, cfg, seed), cfg, sp


@pytest.mark.parametrize("name,mode", [("tictactoe", "off"), ("connect4", "off"), ("connect4", "x3")])
def test_play_game_reproduces_reference_games_on_device(name, mode, monkeypatch):
    """Whole games through SelfPlay.play_game on the GPU: the reference's action sequences and visit
    distributions exactly (fp32 CUDA-core networks, and the default fp32-grade tensor-core towers); root values within
    5e-3 (fp32 network differences of ~1e-6 per logit accumulate through 25-50 backed-up simulations)."""
    monkeypatch.setenv("MZ_TC_MODE", mode)
    for ref in golden_json("play.json")[name]:
        worker, cfg, sp = _worker(name, ref["seed"], num_simulations=ref["num_simulations"])
        gh = worker.play_game(ref["temperature"], cfg.temperature_threshold, False, "self", 0)
        assert [int(a) for a in gh.action_history] == ref["action_history"]
        assert [[float(x) for x in c] for c in gh.child_visits] == ref["child_visits"]
        assert [float(r) for r in gh.reward_history] == ref["reward_history"]
        assert [int(t) for t in gh.to_play_history] == ref["to_play_history"]
        numpy.testing.assert_allclose(gh.root_values, ref["root_values"], rtol=5e-3, atol=2e-3)
        worker.model.engine.close()


@pytest.mark.parametrize("name,tc", [("tictactoe", "off"), ("connect4", "off"), ("connect4", "x3"), ("connect4", "fp16")])
def test_mcts_run_node_graph_on_device(name, tc, monkeypatch):
    """MCTS(config).run returns a Node graph with the reference's shape; hidden states come back in NCHW
    order whatever the internal layout (P64C4 on the tensor-core path)."""
    monkeypatch.setenv("MZ_TC_MODE", tc)
    worker, cfg, sp = _worker(name, 0)
    c = golden_json(f"mcts_{name}.json")[0]
    cfg.num_simulations = c["num_simulations"]
    worker, cfg, sp = _worker(name, 0, num_simulations=c["num_simulations"])
    numpy.random.seed(c["seed"])
    obs = numpy.array(c["obs"]).reshape(c["obs_shape"])
    root, info = sp.MCTS(cfg).run(worker.model, obs, c["legal"], c["to_play"], True)
    assert list(root.children.keys()) == c["root_actions"]
    got = [root.children[a].visit_count for a in c["root_actions"]]
    if tc != "fp16":
        assert got == c["root_visits"]
    else:      # fp16 tensor-core towers: a visit may move between two nearly tied children
        assert 0.5 * sum(abs(x - y) for x, y in zip(got, c["root_visits"])) / c["num_simulations"] <= 0.05
    assert root.visit_count == c["num_simulations"] == sum(got)
    tol = 1e-4 if tc != "fp16" else 3e-2
    assert abs(root.value() - c["root_value"]) <= tol * max(1.0, abs(c["root_value"]))
    if tc != "fp16":
        assert info["max_tree_depth"] == c["max_tree_depth"]
    # root hidden state = the reference network's representation of the observation
    spec = netspec_from_config(cfg)
    from oracle.net import OracleNet
    h = OracleNet(spec, weights_for(name, spec)).initial_inference(obs[None].astype(numpy.float32))[3].numpy().ravel()
    numpy.testing.assert_allclose(root.hidden_state, h, rtol=2e-2 if tc == "fp16" else 2e-4, atol=2e-2 if tc == "fp16" else 2e-5)
    node = root
    for a in c["sims"][-1]["actions"][:-1]:
        node = node.children[a]
        assert node.expanded() and node.hidden_state is not None and node.hidden_state.shape == (spec.hidden_elems,)
    worker.model.engine.close()


def test_batched_self_play_on_device():
    """Lockstep batch on the device: every finished GameHistory is well formed; slot 0 equals the
    single-game run with the same seed (numpy draw order)."""
    w8, cfg, sp = _worker("tictactoe", 3, num_parallel_games=8, num_simulations=16)
    games = w8.play_games(8, 1.0)
    for g in games:
        T = len(g.action_history) - 1
        assert 5 <= T <= 9 and len(g.child_visits) == T == len(g.root_values)
        assert all(abs(sum(c) - 1) < 1e-12 for c in g.child_visits)
        assert g.observation_history[0].shape == (3, 3, 3)
    w1, _, _ = _worker("tictactoe", 3, num_parallel_games=1, num_simulations=16)
    solo = w1.play_games(1, 1.0)[0]
    same = [g for g in games if [int(a) for a in g.action_history] == [int(a) for a in solo.action_history]]
    assert same and same[0].child_visits == solo.child_visits
    w8.model.engine.close(); w1.model.engine.close()


def test_error_paths_on_device(game_configs):
    from muzero_general_b200 import _lib
    from muzero_general_b200.engine import SearchEngine
    cfg = game_configs["cartpole"]
    eng = SearchEngine(cfg, max_games=4, num_simulations=5)
    obs = numpy.zeros((4, 4), numpy.float32)
    with pytest.raises(_lib.MzError, match="weights not loaded"):
        eng.search(obs=obs)
    spec = netspec_from_config(cfg)
    w = dict(weights_for("cartpole", spec))
    bad = dict(w); bad.pop("prediction_value_network.module.2.bias")
    with pytest.raises(KeyError):
        eng.load_weights(bad)
    eng.load_weights(w)
    with pytest.raises(_lib.MzError, match="out of range"):
        eng.search(obs=numpy.zeros((5, 4), numpy.float32))
    with pytest.raises(ValueError):
        eng.search(obs=numpy.zeros((4, 5), numpy.float32))
    out = eng.search(obs=obs[:2])                      # fewer games than the capacity
    assert out.visit_counts.shape == (2, 2) and (out.visit_counts.sum(1) == 5).all()
    eng.close()


@pytest.mark.parametrize("name", ["tictactoe", "cartpole"])
def test_override_root_with_on_device(name, monkeypatch):
    """MCTS.run(..., override_root_with=node) (self_play.py:275-277): (a) the most visited child of a finished search
    becomes the root of a second search (subtree reuse); (b) a hand-expanded, unvisited node as diagnose_model.py:54-69
    builds it.  The device continues the imported tree and reproduces the reference's visit counts exactly."""
    monkeypatch.setenv("MZ_TC_MODE", "off")
    fx = golden_json("override_root.json")[name]
    first = fx["first"]
    worker, cfg, sp = _worker(name, 0, num_simulations=first["num_simulations"])
    A = len(cfg.action_space)
    obs = numpy.array(first["obs"]).reshape(first["obs_shape"])
    for case in fx["cases"]:
        numpy.random.seed(0)
        root, _ = sp.MCTS(cfg).run(worker.model, obs, first["legal"], first["to_play"], True)
        assert [root.children[a].visit_count for a in first["root_actions"]] == first["root_visits"]
        action = int(sp.SelfPlay.select_action(root, 0))
        assert action == case["action"]
        if case["kind"] == "subtree":
            node = root.children[action]
        else:
            r = worker.model.engine.recurrent_inference(root.hidden_state[None], [action])
            node = sp.Node(0)
            node.expand(cfg.action_space, case["to_play"], float(r["reward"][0]), r["policy_logits"], r["hidden"][0])
        assert node.visit_count == case["pre_visits"]
        root2, info2 = sp.MCTS(cfg).run(worker.model, None, cfg.action_space, case["to_play"], True, node)
        assert list(root2.children.keys()) == case["root_actions"]
        assert [root2.children[a].visit_count for a in case["root_actions"]] == case["root_visits"], case["kind"]
        assert root2.visit_count == case["root_visit_count"] == case["pre_visits"] + first["num_simulations"]
        assert info2["root_predicted_value"] is None and info2["max_tree_depth"] == case["max_tree_depth"]
        assert abs(root2.value() - case["root_value"]) <= 1e-4 * max(1.0, abs(case["root_value"]))
        numpy.testing.assert_allclose([root2.children[a].prior for a in case["root_actions"]], case["root_priors"], rtol=1e-5, atol=1e-7)
        numpy.testing.assert_allclose([root2.children[a].value_sum for a in case["root_actions"]], case["root_child_value_sums"],
                                      rtol=1e-3, atol=1e-3)
    worker.model.engine.close()


def test_batched_reanalyse_on_device(monkeypatch):
    """Reanalyse (replay_buffer.py:307-373) on the real engine: games played by continuous_self_play on the GPU are
    re-evaluated in ONE batched mz_initial_inference; the values equal the oracle network's per-game computation, the
    bulk priorities are attached, and the reference-shaped actor loop updates a buffer."""
    import torch
    monkeypatch.setenv("MZ_TC_MODE", "off")
    from muzero_general_b200 import reanalyse as ra
    from oracle.net import OracleNet, support_to_scalar
    worker, cfg, sp = _worker("tictactoe", 2, num_parallel_games=6, num_simulations=8, training_steps=6, ratio=None)
    spec = netspec_from_config(cfg)
    w = weights_for("tictactoe", spec)

    class Storage:
        def __init__(self):
            self.d = dict(weights=w, training_step=0, terminate=False, num_played_steps=0, num_played_games=0, num_reanalysed_games=0)
        def get_info(self, k):
            if k == "training_step":
                self.d[k] += 2
            return self.d[k]
        def set_info(self, k, v=None):
            self.d.update(k if isinstance(k, dict) else {k: v})

    class Buffer:
        def __init__(self):
            self.buffer, self.updated = {}, set()
        def save_game(self, gh, storage=None):
            self.buffer[len(self.buffer)] = gh
            if storage is not None:
                storage.set_info("num_played_games", len(self.buffer))
        def sample_game(self, force_uniform=False):
            i = int(numpy.random.randint(len(self.buffer)))
            return i, self.buffer[i], None
        def update_game_history(self, game_id, gh):
            self.updated.add(game_id); self.buffer[game_id] = gh

    st, buf = Storage(), Buffer()
    worker.continuous_self_play(st, buf)                       # a17 on the GPU engine (host loop, numpy draws)
    games = list(buf.buffer.values())
    assert len(games) >= 6
    for gh in games:                                            # bulk PER ingest (replay_buffer.py:39-51)
        gh.priorities, gh.game_priority = ra.initial_priorities(gh, cfg)
        assert gh.priorities.shape == (len(gh.root_values),) and gh.game_priority == gh.priorities.max()
    actor = ra.Reanalyse({"weights": w, "num_reanalysed_games": 0}, cfg, max_positions=32)
    actor.reanalyse_games(games)
    net = OracleNet(spec, w)
    for gh in games:
        T = len(gh.root_values)
        obs = numpy.array([gh.get_stacked_observations(i, cfg.stacked_observations, 9) for i in range(T)], dtype=numpy.float32)
        want = torch.squeeze(support_to_scalar(net.initial_inference(obs)[0], cfg.support_size)).numpy()
        assert gh.reanalysed_predicted_root_values.shape == want.shape
        numpy.testing.assert_allclose(gh.reanalysed_predicted_root_values, want, rtol=2e-4, atol=5e-4)
    st.d["training_step"] = 0
    actor.games_per_call = 4
    actor.reanalyse(buf, st)
    assert buf.updated and st.d["num_reanalysed_games"] > len(games)
    actor.close(); worker.model.engine.close()


def test_wide_action_space_game_through_the_host_loop():
    """games/gomoku.py (121 actions: four actions per lane in the tree kernels) through SelfPlay.play_moves on the host
    loop: every move is legal (one new stone per move on a free cell), visit counts sum to N over legal actions only."""
    worker, cfg, sp = _worker("gomoku", 0, num_parallel_games=4, num_simulations=20)
    assert worker.loop_path == "host"
    worker.play_moves(10, 1.0)
    env = worker._batched.env
    assert (numpy.abs(env.board).sum(1) == 10).all() and (env.board.sum(1) == 0).all()     # 5 stones each, all on distinct cells
    rec = worker._batched.records[-1]
    assert (rec["visits"].sum(1) == 20).all() and (rec["visits"][rec["legal"] == 0] == 0).all()
    # single-game MCTS.run returns a Node graph over the legal actions
    root, info = sp.MCTS(cfg).run(worker.model, numpy.zeros(cfg.observation_shape, numpy.float32), list(range(0, 121, 3)), 0, True)
    assert list(root.children) == list(range(0, 121, 3)) and root.visit_count == 20
    worker.model.engine.close()
