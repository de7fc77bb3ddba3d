"""Test double for muzero_general_b200.engine.SearchEngine built on the ORACLE.

Lets the host-side logic (SelfPlay, BatchedSelfPlay, MCTS/Node views, GameHistory assembly)
be exercised on a machine without a GPU.  Test infrastructure only - never importable from
the product package."""
import numpy

from muzero_general_b200.engine import SearchOutput
from muzero_general_b200.netspec import netspec_from_config
from oracle import mcts as om
from oracle import philox
from oracle.net import OracleNet


class FakeSearchEngine:
    def __init__(self, config, max_games=1, device=0, seed=None, num_simulations=None, extra_expansions=0):
        self.config = config
        self.spec = netspec_from_config(config)
        self.A = self.spec.action_space
        self.N = int(config.num_simulations if num_simulations is None else num_simulations)
        self.max_games = max_games
        self.seed = int(config.seed if seed is None else seed)
        self.hidden_elems = self.spec.hidden_elems
        self.obs_elems = self.spec.obs_elems
        self.net = None
        self._last = None
        self.launch_count = 0
        self.last_search_ms = 0.0

    def close(self):
        pass

    def load_weights(self, state_dict):
        sd = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else numpy.asarray(v)) for k, v in state_dict.items()}
        self.net = OracleNet(self.spec, sd)

    def search(self, obs=None, legal_mask=None, to_play=None, add_exploration_noise=False, noise=None,
               first_index=None, game_id=None, move_index=None, keep_tree=False, **_):
        n, A, N = obs.shape[0], self.A, self.N
        params = om.SearchParams.from_config(self.config, N)
        shape = (self.spec.in_channels,) + tuple(self.spec.obs_shape[1:])
        out = SearchOutput(numpy.zeros((n, A), numpy.int32), numpy.zeros(n), numpy.zeros(n, numpy.float32),
                           numpy.zeros(n, numpy.int32), numpy.zeros(n, numpy.int32), numpy.zeros((n, A)),
                           numpy.zeros((n, 2)))
        self._last = []
        for g in range(n):
            acts = [a for a in range(A) if legal_mask is None or legal_mask[g, a]]
            gid = int(game_id[g]) if game_id is not None else g
            mv = int(move_index[g]) if move_index is not None else 0
            draws = om.InjectedDraws(
                [noise[g, a] for a in acts] if add_exploration_noise else None,
                None if first_index is None else int(first_index[g]),
                tie_fn=lambda k, ctx, gid=gid, mv=mv: philox.tie_index(self.seed, gid, mv, ctx[0], ctx[1], k))
            res = om.TreeSearch(params).run(om.ModelEvaluator(self.net, self.spec.support_size),
                                            numpy.asarray(obs[g]).reshape(shape), acts,
                                            int(to_play[g]) if to_play is not None else 0, add_exploration_noise, draws)
            for a, v, p in zip(res.root_actions, res.root_visits, res.root_priors):
                out.visit_counts[g, a] = v
                out.root_priors[g, a] = p
            out.root_value[g] = res.root_value
            out.root_predicted_value[g] = res.root_predicted_value
            out.max_tree_depth[g] = res.max_tree_depth
            out.tie_count[g] = draws.later_ties
            out.value_range[g] = (res.range_lo, res.range_hi)
            self._last.append(res)
        return out

    def initial_inference(self, obs):
        from oracle.net import support_to_scalar
        obs = numpy.asarray(obs, dtype=numpy.float32)
        shape = (obs.shape[0], self.spec.in_channels) + tuple(self.spec.obs_shape[1:])
        v, r, p, h = self.net.initial_inference(obs.reshape(shape))
        return dict(value_logits=v.numpy(), reward_logits=r.numpy(), policy_logits=p.numpy(),
                    hidden=h.numpy().reshape(obs.shape[0], -1), value=support_to_scalar(v, self.spec.support_size).numpy()[:, 0],
                    reward=support_to_scalar(r, self.spec.support_size).numpy()[:, 0])

    def export_tree(self, game, with_hidden=False):
        """Oracle tree -> the SoA layout of mz_export_tree (expansion e owns slots [e*A, e*A+A))."""
        res, A, N = self._last[game], self.A, self.N
        t = res.tree
        S = (N + 1) * A
        out = dict(child_visit=numpy.zeros(S, numpy.int32), child_value_sum=numpy.zeros(S), child_reward=numpy.zeros(S, numpy.float32),
                   child_prior=numpy.zeros(S), child_expansion=numpy.full(S, -1, numpy.int32))
        if with_hidden:
            out["hidden"] = numpy.zeros((N + 1, self.hidden_elems), numpy.float32)
            out["hidden"][0] = t.state[0].numpy().ravel()
        for e, (base, acts) in enumerate(zip(t.base, t.acts)):
            for k, a in enumerate(acts):
                s_o, s_d = base + k, e * A + a
                out["child_visit"][s_d] = t.visit[s_o]
                out["child_value_sum"][s_d] = t.vsum[s_o]
                out["child_prior"][s_d] = t.prior[s_o]
                if t.block[s_o] >= 0:
                    out["child_expansion"][s_d] = t.block[s_o]
                    out["child_reward"][s_d] = t.reward[s_o]
                    if with_hidden:
                        out["hidden"][t.block[s_o]] = t.state[s_o].numpy().ravel()
        out["n_expansions"] = len(t.base)
        out["root_visit"] = t.visit[0]
        out["root_value_sum"] = float(t.vsum[0])
        return out
