"""Pins the oracle (oracle/net.py, oracle/mcts.py) against outputs of the reference itself.

The fixtures were produced by oracle/gen_golden.py from the unmodified reference
(self_play.py / models.py).  Everything compared here is CPU-only."""
import math

import numpy
import pytest
import torch

from conftest import golden_json, golden_npz, weights_for
from muzero_general_b200.netspec import netspec_from_config
from oracle import mcts as om
from oracle.net import OracleNet, support_to_scalar

torch.set_num_threads(1)


def _net(name, cfgs):
    base = name.split("_")[0]
    spec = netspec_from_config(cfgs[base])
    return spec, OracleNet(spec, weights_for(name, spec))


@pytest.mark.parametrize("name", ["cartpole", "cartpole_pretrained", "tictactoe", "connect4", "breakout", "gomoku", "connect4_b64",
                                  "connect4_stress_large", "connect4_stress_overflow", "connect4_stress_tiny"])
def test_network_matches_reference_outputs(name, game_configs):
    spec, net = _net(name, game_configs)
    g = golden_npz(f"net_{name}.npz")
    v0, r0, p0, h0 = net.initial_inference(g["obs"])
    v1, r1, p1, h1 = net.recurrent_inference(h0, g["action"])
    v2, r2, p2, h2 = net.recurrent_inference(h1, (g["action"] + 1) % spec.action_space)
    tol = dict(rtol=1e-5, atol=1e-6)     # same ATen calls; allows for a different CPU ISA
    if "_stress_" in name:               # logits of 1e4..1e7: the absolute term scales with the data
        tol = dict(rtol=1e-5, atol=1e-6 * max(1.0, float(numpy.abs(g["rec_value"]).max())))
    for got, key in ((v0, "init_value"), (p0, "init_policy"), (h0, "init_hidden"),
                     (v1, "rec_value"), (r1, "rec_reward"), (p1, "rec_policy"), (h1, "rec_hidden"),
                     (v2, "rec2_value"), (r2, "rec2_reward"), (p2, "rec2_policy"), (h2, "rec2_hidden")):
        numpy.testing.assert_allclose(got.numpy(), g[key], err_msg=key, **tol)
    S = spec.support_size
    numpy.testing.assert_allclose(support_to_scalar(v1, S).numpy()[:, 0], g["rec_value_scalar"], **tol)
    numpy.testing.assert_allclose(support_to_scalar(r1, S).numpy()[:, 0], g["rec_reward_scalar"], **tol)
    assert (support_to_scalar(r0, S).numpy() == 0).all()


def test_large_configuration_network_matches_reference_outputs(game_configs):
    """games/atari.py: 131 stacked planes of 96x96, DownSample stem, 16 blocks x 256 channels, 601-bin heads."""
    spec, net = _net("atari", game_configs)
    g = golden_npz("net_atari.npz")
    obs = numpy.random.RandomState(int(g["obs_seed"])).random_sample((2, spec.in_channels, 96, 96)).astype(numpy.float32)
    v0, r0, p0, h0 = net.initial_inference(obs)
    v1, r1, p1, h1 = net.recurrent_inference(h0, g["action"])
    tol = dict(rtol=2e-5, atol=2e-6)
    for got, key in ((v0, "init_value"), (p0, "init_policy"), (h0, "init_hidden"), (v1, "rec_value"), (r1, "rec_reward"),
                     (p1, "rec_policy"), (h1, "rec_hidden")):
        numpy.testing.assert_allclose(got.numpy(), g[key], err_msg=key, **tol)
    numpy.testing.assert_allclose(support_to_scalar(v1, spec.support_size).numpy()[:, 0], g["rec_value_scalar"], rtol=1e-4, atol=1e-4)


def test_support_to_scalar_kat():
    k = golden_json("kat.json")["support_to_scalar"]
    out = support_to_scalar(torch.tensor(k["logits"], dtype=torch.float32), 10)[:, 0]
    numpy.testing.assert_allclose(out.numpy(), k["out"], rtol=1e-6)
    assert abs(k["out"][0] - 4.885034561157227) < 1e-6
    centre = torch.log(torch.zeros(1, 21).scatter(1, torch.tensor([[10]]), 1.0))
    c = support_to_scalar(centre, 10).item()
    assert c == 0 and (math.copysign(1, c) < 0) == k["centre_sign_negative"]


def test_ucb_score_kat():
    for c in golden_json("kat.json")["ucb_score"]:
        p = om.SearchParams(50, [0, 1], list(range(c["players"])), c["discount"], c["pb_c_base"],
                            c["pb_c_init"], 0.25, 0.25)
        t = om.Tree()
        t.visit[0] = c["parent_visits"]
        t.expand(0, [0], 0, 0, [c["prior"]], None)
        t.visit[1], t.vsum[1], t.reward[1] = c["visits"], c["value_sum"], c["reward"]
        rng = om.RunningRange()
        if c["lo"] is not None:
            rng.update(c["lo"]); rng.update(c["hi"])
        assert om.TreeSearch(p)._score(t, 0, 1, rng) == c["score"]
    # the value quoted in SURVEY.md 8c
    assert golden_json("kat.json")["ucb_score"][0]["score"] == 0.8299265960706028


def test_select_action_kat():
    for c in golden_json("kat.json")["select_action"]:
        temp = float("inf") if c["temperature"] == "inf" else c["temperature"]
        draws = om.LegacyNumpyDraws(numpy.random.RandomState(c["seed"]))
        assert om.select_action(c["actions"], c["counts"], temp, draws) == c["action"]


def test_stacked_observations_and_statistics_kat():
    k = golden_json("kat.json")
    so = k["stacked_observations"]
    obs = [numpy.array(o, dtype="int32") for o in so["observations"]]
    for c in so["cases"]:
        got = om.stacked_observation(obs, so["actions"], c["index"], c["stacked"], so["A"])
        assert list(got.shape) == c["shape"] and str(got.dtype) == c["dtype"]
        assert got.ravel().tolist() == c["data"]
    st = k["search_statistics"]
    assert om.child_visit_policy(list(range(9)), [0, 4, 8], [6, 18, 1]) == st["child_visits"][0]


SEARCH_FILES = ["cartpole_synth", "cartpole_pretrained", "tictactoe", "connect4", "breakout", "connect4_n200", "breakout_n50", "gomoku"]


@pytest.mark.parametrize("name", SEARCH_FILES)
def test_search_reproduces_reference_bit_for_bit(name, game_configs):
    """Same weights, same legacy numpy seed -> identical tree (fp64 equality, not tolerance)."""
    spec, net = _net(name, game_configs)
    cfg = game_configs[name.split("_")[0]]
    for case in golden_json(f"mcts_{name}.json"):
        params = om.SearchParams.from_config(cfg, case["num_simulations"])
        obs = numpy.array(case["obs"]).reshape(case["obs_shape"])
        draws = om.LegacyNumpyDraws(numpy.random.RandomState(case["seed"]))
        res = om.TreeSearch(params).run(om.ModelEvaluator(net, spec.support_size), obs, case["legal"],
                                        case["to_play"], case["add_noise"], draws)
        assert res.root_actions == case["root_actions"]
        assert res.root_visits == case["root_visits"]
        assert res.root_value == case["root_value"]
        assert res.root_priors == case["root_priors"]
        assert res.max_tree_depth == case["max_tree_depth"]
        assert res.root_predicted_value == case["root_predicted_value"]
        assert [s.path_actions for s in res.sims] == [s["actions"] for s in case["sims"]]
        assert [s.value for s in res.sims] == [s["value"] for s in case["sims"]]
        assert [s.reward for s in res.sims] == [s["reward"] for s in case["sims"]]
        assert [s.priors for s in res.sims] == [s["priors"] for s in case["sims"]]


@pytest.mark.parametrize("name", SEARCH_FILES)
def test_search_teacher_forced_from_trace(name, game_configs):
    """Replaying the recorded per-simulation outputs (no network) gives the same tree:
    this is the protocol the device tree kernels are tested with."""
    cfg = game_configs[name.split("_")[0]]
    for case in golden_json(f"mcts_{name}.json"):
        params = om.SearchParams.from_config(cfg, case["num_simulations"])
        ev = om.TableEvaluator((case["root_predicted_value"], case["root_reward"], case["root_priors_raw"]),
                               [(s["value"], s["reward"], s["priors"]) for s in case["sims"]])
        draws = om.InjectedDraws(case["noise"], case["first_index"])
        res = om.TreeSearch(params).run(ev, None, case["legal"], case["to_play"], case["add_noise"], draws)
        assert draws.later_ties == case["later_ties"] == 0
        assert res.root_visits == case["root_visits"] and res.root_value == case["root_value"]
        assert [s.path_actions for s in res.sims] == [s["actions"] for s in case["sims"]]


def test_published_survey_vectors():
    """The numbers SURVEY.md 8c quotes for the shipped CartPole checkpoint."""
    cases = golden_json("mcts_cartpole_pretrained.json")
    assert cases[0]["root_visits"] == [2, 23] and cases[0]["root_value"] == 103.03345453874076
    assert cases[0]["root_priors"] == [0.38765144048778205, 0.612348559512218]
    assert cases[0]["max_tree_depth"] == 6 and cases[0]["root_predicted_value"] == 103.25457763671875
    assert cases[1]["root_visits"] == [7, 43] and cases[1]["root_value"] == 103.43967241245127
    assert cases[2]["root_visits"] == [7, 43] and cases[2]["root_value"] == 103.40166359742176
