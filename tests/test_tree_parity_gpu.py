"""GPU parity of the tree kernels (bit-exact) and the FC network kernels (fp32 tolerance).

Protocol: (i) teacher forcing from reference traces, (ii) student forcing -
the device's own per-simulation outputs replayed through the oracle, (iii) network outputs
against the reference's, (iv) closed loop against the reference's visit counts.
Everything goes through the C ABI (muzero_general_b200.engine -> libmzb200.so)."""
import numpy
import pytest

from conftest import golden_json, golden_npz, weights_for
from helpers import oracle_replay, paths_from_trace, random_teacher, teacher_from_cases
from muzero_general_b200.netspec import netspec_from_config
from oracle import mcts as om

pytestmark = pytest.mark.gpu

SEARCH_FILES = ["cartpole_synth", "cartpole_pretrained", "tictactoe", "connect4", "breakout", "connect4_n200", "breakout_n50", "gomoku"]


def _engine(cfg, max_games, N):
    from muzero_general_b200.engine import SearchEngine
    return SearchEngine(cfg, max_games=max_games, num_simulations=N)


@pytest.mark.parametrize("stepwise", [False, True])
@pytest.mark.parametrize("name", SEARCH_FILES)
def test_teacher_forced_reference_traces(name, stepwise, game_configs):
    """Injected (priors, value, reward) from the reference's own runs -> identical trees."""
    cfg = game_configs[name.split("_")[0]]
    A = len(cfg.action_space)
    cases = golden_json(f"mcts_{name}.json")
    by_n = {}
    for c in cases:
        by_n.setdefault((c["num_simulations"], c["add_noise"]), []).append(c)
    for (N, add_noise), group in by_n.items():
        eng = _engine(cfg, len(group), N)
        t, legal, noise, first, to_play = teacher_from_cases(group, A, N)
        out = eng.search(legal_mask=legal, to_play=to_play, add_exploration_noise=add_noise, noise=noise,
                         first_index=first, teacher=t, trace=True, stepwise=stepwise, n_games=len(group))
        for i, c in enumerate(group):
            assert [int(out.visit_counts[i, a]) for a in c["root_actions"]] == c["root_visits"]
            assert out.root_value[i] == c["root_value"]                      # fp64 equality
            assert [out.root_priors[i, a] for a in c["root_actions"]] == c["root_priors"]
            assert out.max_tree_depth[i] == c["max_tree_depth"]
            assert out.tie_count[i] == 0
            assert paths_from_trace(out.trace, i, N) == [s["actions"] for s in c["sims"]]
        eng.close()


@pytest.mark.parametrize("stepwise", [False, True])
@pytest.mark.parametrize("game,N,n", [("cartpole", 50, 48), ("tictactoe", 50, 32), ("connect4", 64, 16), ("breakout", 30, 16), ("gomoku", 60, 24)])
def test_teacher_forced_synthetic_vs_oracle(game, N, n, stepwise, game_configs):
    """Random per-simulation tables, restricted legal sets, both player modes, Philox ties."""
    cfg = game_configs[game]
    A, P = len(cfg.action_space), len(cfg.players)
    rs = numpy.random.RandomState(1234)
    legal = (rs.uniform(size=(n, A)) < 0.7).astype(numpy.uint8)
    legal[numpy.arange(n), rs.randint(0, A, n)] = 1
    t = random_teacher(rs, n, N, A, reward_scale=1.0 if P == 1 else 0.0, legal=legal)
    # a few games with perfectly flat priors and zero values: forces exact ties at every level
    t["priors"][:4] = numpy.float32(1.0 / A) if (A & (A - 1)) == 0 else t["priors"][:4]
    t["value"][:4] = 0
    t["reward"][:4] = 0
    noise = rs.dirichlet([cfg.root_dirichlet_alpha] * A, size=n)
    to_play = rs.randint(0, P, n).astype(numpy.int32)
    game_id = (1000 + numpy.arange(n)).astype(numpy.int64)
    move = rs.randint(0, 40, n).astype(numpy.int32)
    eng = _engine(cfg, n, N)
    out = eng.search(legal_mask=legal, to_play=to_play, add_exploration_noise=True, noise=noise, game_id=game_id,
                     move_index=move, teacher=t, trace=True, stepwise=stepwise, n_games=n)
    params = om.SearchParams.from_config(cfg, N)
    total_ties = 0
    for i in range(n):
        acts = [a for a in range(A) if legal[i, a]]
        res, draws = oracle_replay(
            params, acts, int(to_play[i]),
            (t["root_value"][i], t["root_reward"][i], [t["root_priors"][i, a] for a in acts]),
            [(t["value"][i, s], t["reward"][i, s], t["priors"][i, s]) for s in range(N)],
            [noise[i, a] for a in acts], None, seed=cfg.seed, game=int(game_id[i]), move=int(move[i]))
        assert [int(out.visit_counts[i, a]) for a in acts] == res.root_visits, i
        assert out.root_value[i] == res.root_value
        assert out.max_tree_depth[i] == res.max_tree_depth
        assert out.tie_count[i] == draws.later_ties
        assert (out.value_range[i, 0], out.value_range[i, 1]) == (res.range_lo, res.range_hi)
        assert paths_from_trace(out.trace, i, N) == [s.path_actions for s in res.sims]
        total_ties += draws.later_ties
    if (A & (A - 1)) == 0:
        assert total_ties > 0        # the tie path was really exercised
    eng.close()


@pytest.mark.parametrize("name", ["cartpole", "cartpole_pretrained"])
def test_fc_network_matches_reference(name, game_configs):
    cfg = game_configs["cartpole"]
    spec = netspec_from_config(cfg)
    g = golden_npz(f"net_{name}.npz")
    eng = _engine(cfg, len(g["obs"]), 5)
    eng.load_weights(weights_for(name, spec))
    r0 = eng.initial_inference(g["obs"])
    tol = dict(rtol=2e-5, atol=2e-6)
    numpy.testing.assert_allclose(r0["value_logits"], g["init_value"], **tol)
    numpy.testing.assert_allclose(r0["policy_logits"], g["init_policy"], **tol)
    numpy.testing.assert_allclose(r0["hidden"], g["init_hidden"].reshape(len(g["obs"]), -1), **tol)
    numpy.testing.assert_allclose(r0["value"], g["init_value_scalar"], **tol)
    assert (r0["reward"] == 0).all() and numpy.signbit(r0["reward"]).all()     # -0.0 like the reference
    r1 = eng.recurrent_inference(g["init_hidden"], g["action"])
    numpy.testing.assert_allclose(r1["value_logits"], g["rec_value"], **tol)
    numpy.testing.assert_allclose(r1["reward_logits"], g["rec_reward"], **tol)
    numpy.testing.assert_allclose(r1["policy_logits"], g["rec_policy"], **tol)
    numpy.testing.assert_allclose(r1["hidden"], g["rec_hidden"].reshape(len(g["obs"]), -1), **tol)
    numpy.testing.assert_allclose(r1["value"], g["rec_value_scalar"], **tol)
    numpy.testing.assert_allclose(r1["reward"], g["rec_reward_scalar"], **tol)
    eng.close()


@pytest.mark.parametrize("stepwise", [False, True])
def test_fc_student_forced(stepwise, game_configs):
    """The device's own network outputs, replayed through the oracle tree, give the same search;
    and those outputs agree with the oracle network on the states the device visited."""
    import torch
    from oracle.net import OracleNet, support_to_scalar
    cfg = game_configs["cartpole"]
    spec = netspec_from_config(cfg)
    N, n, A = 50, 64, 2
    rs = numpy.random.RandomState(7)
    obs = rs.uniform(-0.05, 0.05, size=(n, 1, 1, 4)).astype(numpy.float32)
    noise = rs.dirichlet([cfg.root_dirichlet_alpha] * A, size=n)
    first = rs.randint(0, A, n).astype(numpy.int32)
    w = weights_for("cartpole_pretrained", spec)
    eng = _engine(cfg, n, N)
    eng.load_weights(w)
    out = eng.search(obs=obs, add_exploration_noise=True, noise=noise, first_index=first, trace=True,
                     keep_tree=True, stepwise=stepwise)
    params = om.SearchParams.from_config(cfg, N)
    net = OracleNet(spec, w)
    for i in range(n):
        tr = out.trace
        res, draws = oracle_replay(
            params, [0, 1], 0, (out.root_predicted_value[i], tr["root_reward"][i], list(tr["root_priors_raw"][i])),
            [(tr["value"][i, s], tr["reward"][i, s], tr["priors"][i, s]) for s in range(N)],
            list(noise[i]), int(first[i]), seed=cfg.seed, game=i)
        assert [int(v) for v in out.visit_counts[i]] == res.root_visits
        assert out.root_value[i] == res.root_value
        assert paths_from_trace(tr, i, N) == [s.path_actions for s in res.sims]
        if i < 4:
            # network parity along the device's own trajectory: hidden states from the exported tree
            tree = eng.export_tree(i, with_hidden=True)
            assert tree["n_expansions"] == N + 1 and tree["root_visit"] == N
            v, r, pol, h = net.initial_inference(obs[i:i + 1])
            numpy.testing.assert_allclose(tree["hidden"][0], h.numpy()[0], rtol=2e-5, atol=2e-6)
            numpy.testing.assert_allclose(out.root_predicted_value[i], support_to_scalar(v, 10).item(), rtol=2e-5)
            # expansion e = s+1 was produced from (parent expansion, action) of simulation s
            exp_of_slot = tree["child_expansion"]
            parent_of = {int(exp_of_slot[s]): (s // A, s % A) for s in range(len(exp_of_slot)) if exp_of_slot[s] >= 0}
            for e in (1, 2, N // 2, N):
                pe, a = parent_of[e]
                v, r, pol, h = net.recurrent_inference(torch.from_numpy(tree["hidden"][pe:pe + 1]), torch.tensor([[a]]))
                numpy.testing.assert_allclose(tree["hidden"][e], h.numpy()[0], rtol=1e-4, atol=1e-5)
                numpy.testing.assert_allclose(tr["value"][i, e - 1], support_to_scalar(v, 10).item(), rtol=1e-4, atol=1e-5)
                numpy.testing.assert_allclose(tr["reward"][i, e - 1], support_to_scalar(r, 10).item(), rtol=1e-4, atol=1e-5)
                numpy.testing.assert_allclose(tr["priors"][i, e - 1], torch.softmax(pol[0], 0).numpy(), rtol=1e-4, atol=1e-6)
    eng.close()


@pytest.mark.parametrize("name", ["cartpole_pretrained", "cartpole_synth"])
def test_fc_closed_loop_matches_reference_counts(name, game_configs):
    """Whole MCTS.run on the device (own networks, reference's noise and first pick) reproduces the
    reference's visit counts; root value within fp32 network tolerance."""
    cfg = game_configs["cartpole"]
    spec = netspec_from_config(cfg)
    for c in golden_json(f"mcts_{name}.json"):
        eng = _engine(cfg, 1, c["num_simulations"])
        eng.load_weights(weights_for(name, spec))
        obs = numpy.array(c["obs"], numpy.float32).reshape(1, *c["obs_shape"])
        noise = numpy.array([c["noise"]]) if c["noise"] else None
        out = eng.search(obs=obs, add_exploration_noise=c["add_noise"], noise=noise,
                         first_index=numpy.array([c["first_index"]], numpy.int32))
        assert [int(v) for v in out.visit_counts[0]] == c["root_visits"]
        assert abs(out.root_value[0] - c["root_value"]) <= 1e-4 * max(1.0, abs(c["root_value"]))
        assert abs(out.root_predicted_value[0] - c["root_predicted_value"]) <= 1e-4 * max(1.0, abs(c["root_predicted_value"]))
        assert out.max_tree_depth[0] == c["max_tree_depth"]
        eng.close()


def test_full_size_invariants(game_configs):
    """BASELINE config 2 size (4096 games, N=50): size-independent properties of every tree."""
    cfg = game_configs["cartpole"]
    spec = netspec_from_config(cfg)
    n, N, A = 4096, 50, 2
    rs = numpy.random.RandomState(3)
    obs = rs.uniform(-0.05, 0.05, size=(n, 4)).astype(numpy.float32)
    noise = rs.dirichlet([0.25] * A, size=n)
    eng = _engine(cfg, n, N)
    eng.load_weights(weights_for("cartpole", spec))
    a = eng.search(obs=obs, add_exploration_noise=True, noise=noise)
    b = eng.search(obs=obs, add_exploration_noise=True, noise=noise, stepwise=True)
    assert (a.visit_counts.sum(1) == N).all()                     # appendix A.9
    assert (a.max_tree_depth >= 1).all() and (a.max_tree_depth <= N).all()
    assert numpy.isfinite(a.root_value).all()
    assert (a.value_range[:, 0] <= a.value_range[:, 1]).all()
    # the fused kernel and the step-wise pipeline are the same function
    assert (a.visit_counts == b.visit_counts).all() and (a.root_value == b.root_value).all()
    # determinism and batch-composition independence
    c = eng.search(obs=obs[::-1].copy(), add_exploration_noise=True, noise=noise[::-1].copy(),
                   game_id=numpy.arange(n)[::-1].astype(numpy.int64))
    assert (c.visit_counts[::-1] == a.visit_counts).all() and (c.root_value[::-1] == a.root_value).all()
    eng.close()


@pytest.mark.parametrize("stepwise", [False, True])
@pytest.mark.parametrize("game,n,N", [("cartpole", 4096, 50), ("tictactoe", 2048, 50), ("connect4", 1024, 200), ("gomoku", 256, 120)])
def test_teacher_forced_full_size_vs_c_oracle(game, n, N, stepwise, game_configs):
    """BASELINE-sized batches, injected outputs: every game's visit counts, root value, depth,
    tie count, value range and every selected path equal the C oracle's, bit for bit."""
    from oracle import build_c
    cfg = game_configs[game]
    A, P = len(cfg.action_space), len(cfg.players)
    rs = numpy.random.RandomState(99)
    legal = (rs.uniform(size=(n, A)) < 0.8).astype(numpy.uint8)
    legal[numpy.arange(n), rs.randint(0, A, n)] = 1
    t = random_teacher(rs, n, N, A, reward_scale=1.0 if P == 1 else 0.0, legal=legal)
    noise = rs.dirichlet([cfg.root_dirichlet_alpha] * A, size=n)
    to_play = rs.randint(0, P, n).astype(numpy.int32)
    gid = rs.randint(0, 1 << 40, n).astype(numpy.int64)
    mv = rs.randint(0, 400, n).astype(numpy.int32)
    D = 64 if A <= 32 else 160              # wide, flat action spaces search deep
    ref = build_c.tree_search(n, N, A, P, cfg.discount, cfg.pb_c_base, cfg.pb_c_init, cfg.root_exploration_fraction,
                              legal, to_play, noise, None, cfg.seed, gid, mv, t, D=D)
    eng = _engine(cfg, n, N)
    out = eng.search(legal_mask=legal, to_play=to_play, add_exploration_noise=True, noise=noise, game_id=gid,
                     move_index=mv, teacher=t, trace=True, trace_depth=D, stepwise=stepwise, n_games=n)
    assert (out.visit_counts == ref["visit_counts"]).all()
    assert (out.root_value == ref["root_value"]).all()
    assert (out.max_tree_depth == ref["max_depth"]).all()
    assert (out.tie_count == ref["ties"]).all()
    assert (out.value_range == ref["range"]).all()
    assert (out.trace["depth"] == ref["depth"]).all()
    assert ref["max_depth"].max() <= D
    mask = numpy.arange(D)[None, None, :] < ref["depth"][:, :, None]
    assert (numpy.where(mask, out.trace["actions"], 0) == numpy.where(mask, ref["actions"], 0)).all()
    eng.close()


def test_device_drawn_dirichlet_noise(game_configs):
    """noise=NULL: the root noise is drawn on the device (Philox + Marsaglia-Tsang gamma).  It is a
    Dirichlet(alpha) sample over the legal actions, keyed by (seed, game id, move), and replaying the
    search through the oracle with the exported noise reproduces the tree exactly."""
    cfg = game_configs["connect4"]
    A, P, N, n = 7, 2, 30, 2048
    rs = numpy.random.RandomState(21)
    legal = (rs.uniform(size=(n, A)) < 0.8).astype(numpy.uint8)
    legal[numpy.arange(n), rs.randint(0, A, n)] = 1
    t = random_teacher(rs, n, N, A, reward_scale=0.0, legal=legal)
    to_play = rs.randint(0, P, n).astype(numpy.int32)
    gid = numpy.arange(n).astype(numpy.int64) + 10_000
    eng = _engine(cfg, n, N)
    a = eng.search(legal_mask=legal, to_play=to_play, add_exploration_noise=True, game_id=gid, teacher=t, trace=True, n_games=n)
    b = eng.search(legal_mask=legal, to_play=to_play, add_exploration_noise=True, game_id=gid, teacher=t, trace=True,
                   stepwise=True, n_games=n)
    nz = a.trace["noise"]
    assert (nz == b.trace["noise"]).all() and (a.visit_counts == b.visit_counts).all()     # deterministic, both paths
    assert ((nz > 0) == (legal > 0)).all()
    numpy.testing.assert_allclose(nz.sum(1), 1.0, rtol=1e-12)
    # Dirichlet(alpha=0.3) over 7 actions: mean 1/7, var = (1/7)(6/7)/(7*0.3+1); all-legal batch for the statistics
    d = eng.search(to_play=to_play, add_exploration_noise=True, game_id=gid, teacher=t, trace=True, n_games=n)
    numpy.testing.assert_allclose(d.trace["noise"].mean(0), 1 / 7, atol=0.015)
    numpy.testing.assert_allclose(d.trace["noise"].var(0), (1 / 7) * (6 / 7) / (A * cfg.root_dirichlet_alpha + 1), rtol=0.2)
    c = eng.search(legal_mask=legal, to_play=to_play, add_exploration_noise=True, game_id=gid + 1, teacher=t, trace=True, n_games=n)
    assert (c.trace["noise"] != nz).any()
    params = om.SearchParams.from_config(cfg, N)
    for i in range(0, n, 64):
        acts = [k for k in range(A) if legal[i, k]]
        res, _ = oracle_replay(params, acts, int(to_play[i]),
                               (t["root_value"][i], t["root_reward"][i], [t["root_priors"][i, k] for k in acts]),
                               [(t["value"][i, s], t["reward"][i, s], t["priors"][i, s]) for s in range(N)],
                               [nz[i, k] for k in acts], None, seed=cfg.seed, game=int(gid[i]))
        assert [int(a.visit_counts[i, k]) for k in acts] == res.root_visits and a.root_value[i] == res.root_value
    eng.close()


def test_fixed_shape_network_path_is_bit_identical_to_the_generic_one(game_configs, monkeypatch):
    """The fused FC kernel evaluates CartPole's networks through fully unrolled fixed-shape code (fc_net.cuh::
    fc_recurrent_fixed); MZ_FC_GENERIC=1 walks the layer descriptors instead.  Same operations in the same order: visit
    counts, root values, value ranges and every hidden state of the exported trees are equal bit for bit."""
    cfg = game_configs["cartpole"]
    spec = netspec_from_config(cfg)
    n, N, A = 1500, 50, 2
    rs = numpy.random.RandomState(17)
    obs = rs.uniform(-0.05, 0.05, size=(n, 4)).astype(numpy.float32)
    noise = rs.dirichlet([0.25] * A, size=n)
    outs, trees = [], []
    for generic in ("1", "0"):
        monkeypatch.setenv("MZ_FC_GENERIC", generic)
        for wname in ("cartpole", "cartpole_pretrained"):
            eng = _engine(cfg, n, N)
            eng.load_weights(weights_for(wname, spec))
            outs.append(eng.search(obs=obs, add_exploration_noise=True, noise=noise, keep_tree=True))
            trees.append([eng.export_tree(i, with_hidden=True) for i in (0, 1, n // 3, n - 1)])
            eng.close()
    for a, b in ((outs[0], outs[2]), (outs[1], outs[3])):
        assert numpy.array_equal(a.visit_counts, b.visit_counts)
        assert numpy.array_equal(a.root_value, b.root_value)
        assert numpy.array_equal(a.value_range, b.value_range)
        assert numpy.array_equal(a.root_predicted_value, b.root_predicted_value)
    for ta, tb in list(zip(trees[0], trees[2])) + list(zip(trees[1], trees[3])):
        for k in ta:
            assert numpy.array_equal(numpy.asarray(ta[k]), numpy.asarray(tb[k])), k
