"""Bulk callers either side of the self-play path (muzero_general_b200/reanalyse.py) on the CPU: PER priorities against
the unmodified reference ReplayBuffer, batched Reanalyse against the oracle network."""
import copy

import numpy
import pytest
import torch

from conftest import weights_for
from fake_engine import FakeSearchEngine
from muzero_general_b200 import reanalyse as ra
from muzero_general_b200 import self_play as sp
from muzero_general_b200.games import load_game_module
from muzero_general_b200.netspec import netspec_from_config
from oracle.refload import reference_available

torch.set_num_threads(1)


def _random_history(rs, cfg, T, players):
    gh = sp.GameHistory()
    A = len(cfg.action_space)
    gh.action_history = [0] + [int(a) for a in rs.randint(0, A, T)]
    gh.observation_history = [rs.random_sample(cfg.observation_shape).astype(numpy.float32) for _ in range(T + 1)]
    gh.reward_history = [0] + [float(r) for r in rs.choice([0.0, 1.0, -1.0, 0.5], T)]
    gh.to_play_history = [int(i % players) for i in range(T + 1)] if players > 1 else [0] * (T + 1)
    cv = rs.random_sample((T, A)); gh.child_visits = (cv / cv.sum(1, keepdims=True)).tolist()
    gh.root_values = [float(v) for v in rs.standard_normal(T)]
    return gh


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (not present on the GPU box)")
@pytest.mark.parametrize("name,td,discount,alpha,reanalysed", [("tictactoe", 20, 1, 0.5, False), ("cartpole", 50, 0.997, 0.5, False),
                                                               ("cartpole", 7, 0.9, 1.0, True), ("connect4", 3, 1, 0.7, True)])
def test_bulk_priorities_equal_the_reference_save_game(name, td, discount, alpha, reanalysed):
    """initial_priorities == what the unmodified ReplayBuffer.save_game computes, bit for bit (float32 priorities, game
    priority), for one- and two-player games, with and without reanalysed values, short and long td horizons."""
    from oracle.refload import load_reference, load_reference_game
    _, _, ref_rb, _ = load_reference()
    ref_cfg = load_reference_game(name).MuZeroConfig()
    ref_cfg.td_steps, ref_cfg.discount, ref_cfg.PER_alpha, ref_cfg.PER = td, discount, alpha, True
    ck = {"num_played_games": 0, "num_played_steps": 0}
    rs = numpy.random.RandomState(4)
    for T in (1, 2, 9, 42, 130):
        gh = _random_history(rs, ref_cfg, T, len(ref_cfg.players))
        if reanalysed:
            gh.reanalysed_predicted_root_values = rs.standard_normal(T).astype(numpy.float32)
        mine, top = ra.initial_priorities(copy.deepcopy(gh), ref_cfg)
        buf = ref_rb.ReplayBuffer(copy.deepcopy(ck), {}, ref_cfg)
        theirs = copy.deepcopy(gh)
        buf.save_game(theirs)
        assert mine.dtype == theirs.priorities.dtype == numpy.float32
        assert numpy.array_equal(mine, theirs.priorities), (name, T)
        assert top == theirs.game_priority
        # and save_games attaches them so that the reference keeps them as they are
        buf2 = ref_rb.ReplayBuffer(copy.deepcopy(ck), {}, ref_cfg)
        g2 = copy.deepcopy(gh)
        ra.save_games(buf2, [g2], ref_cfg)
        assert numpy.array_equal(buf2.buffer[0].priorities, theirs.priorities) and buf2.num_played_steps == T


def test_batched_reanalyse_matches_per_game_inference(monkeypatch):
    """One batched call over many games == the reference's per-game computation (replay_buffer.py:345-366) done with
    the oracle network; the actor loop updates the buffer and the counter."""
    from oracle.net import OracleNet, support_to_scalar
    monkeypatch.setattr(ra, "SearchEngine", FakeSearchEngine)
    mod = load_game_module("tictactoe")
    cfg = mod.MuZeroConfig()
    cfg.training_steps = 3
    spec = netspec_from_config(cfg)
    w = weights_for("tictactoe", spec)
    rs = numpy.random.RandomState(2)
    games = [_random_history(rs, cfg, T, 2) for T in (1, 5, 9, 3)]
    for g in games:
        g.observation_history = [rs.randint(0, 2, cfg.observation_shape).astype(numpy.int32) for _ in g.observation_history]
    actor = ra.Reanalyse({"weights": w, "num_reanalysed_games": 0}, cfg, max_positions=7)      # forces several chunks
    actor.reanalyse_games(games)
    net = OracleNet(spec, w)
    for g in games:
        T = len(g.root_values)
        obs = numpy.array([g.get_stacked_observations(i, cfg.stacked_observations, 9) for i in range(T)], dtype=numpy.float32)
        want = torch.squeeze(support_to_scalar(net.initial_inference(obs)[0], cfg.support_size)).numpy()
        assert g.reanalysed_predicted_root_values.shape == want.shape and g.reanalysed_predicted_root_values.dtype == numpy.float32
        numpy.testing.assert_allclose(g.reanalysed_predicted_root_values, want, rtol=1e-6, atol=1e-7)
    assert actor.num_reanalysed_games == 4

    class Storage:
        def __init__(self):
            self.d = dict(weights=w, training_step=0, terminate=False, num_played_games=4, num_reanalysed_games=0)
        def get_info(self, k):
            if k == "training_step":
                self.d[k] += 1
            return self.d[k]
        def set_info(self, k, v=None):
            self.d.update(k if isinstance(k, dict) else {k: v})

    class Buffer:
        def __init__(self):
            self.buffer = {i: copy.deepcopy(g) for i, g in enumerate(games)}
            self.updated = set()
        def sample_game(self, force_uniform=False):
            i = int(rs.randint(len(self.buffer)))
            return i, self.buffer[i], None
        def update_game_history(self, game_id, gh):
            self.updated.add(game_id); self.buffer[game_id] = gh

    st, buf = Storage(), Buffer()
    actor.games_per_call = 3
    actor.reanalyse(buf, st)
    assert buf.updated and st.d["num_reanalysed_games"] == actor.num_reanalysed_games > 4
