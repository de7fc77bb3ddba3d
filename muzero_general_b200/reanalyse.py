"""The callers either side of the self-play path, in bulk.

* ``Reanalyse`` - same constructor and ``reanalyse(replay_buffer, shared_storage)`` loop as the reference actor
  (``replay_buffer.py:307-373``), but the fresh root values of MANY games come from ONE batched
  ``mz_initial_inference`` per call (the representation + prediction kernels of the search path, with
  ``support_to_scalar`` fused behind the value head) instead of one game per RPC on the CPU.
* ``initial_priorities`` / ``save_games`` - the prioritised-replay priorities ``ReplayBuffer.save_game`` computes one
  position at a time in Python (``replay_buffer.py:33-51`` calling ``compute_target_value``, ``:230-262``), evaluated
  for a whole game with array arithmetic in the reference's operation order (bit-identical float32 priorities), and
  attached before the game is handed over, so the unmodified ``save_game`` skips its loop
  (``if game_history.priorities is not None``).

Nothing here imports torch; the engine does the arithmetic on the GPU.
"""
from __future__ import annotations

import time

import numpy

from .engine import SearchEngine


def _call(obj, method, *args, **kw):
    fn = getattr(obj, method)
    if hasattr(fn, "remote"):
        import ray
        return ray.get(fn.remote(*args, **kw))
    return fn(*args, **kw)


def _fire(obj, method, *args):
    fn = getattr(obj, method)
    return fn.remote(*args) if hasattr(fn, "remote") else fn(*args)


class Reanalyse:
    """Updates games of the replay buffer with fresh value estimates (MuZero paper, appendix Reanalyse)."""

    def __init__(self, initial_checkpoint, config, device=0, max_positions=None, games_per_call=None):
        self.config = config
        numpy.random.seed(config.seed)                     # replay_buffer.py:318
        self.max_positions = int(max_positions or getattr(config, "reanalyse_max_positions", 4096))
        self.games_per_call = int(games_per_call or getattr(config, "reanalyse_games_per_call", 64))
        # inference only: num_simulations = 0 keeps the node / hidden-state pools at one entry per position
        self.engine = SearchEngine(config, max_games=self.max_positions, device=device, num_simulations=0)
        self.engine.load_weights(initial_checkpoint["weights"])
        self.num_reanalysed_games = initial_checkpoint.get("num_reanalysed_games", 0)

    def set_weights(self, weights):
        self.engine.load_weights(weights)

    def close(self):
        self.engine.close()

    # ------------------------------------------------------------------ the batched core
    def fresh_root_values(self, game_histories):
        """``models.support_to_scalar(model.initial_inference(observations)[0])`` (replay_buffer.py:345-366) for every
        position of every game, batched over games; returns one float32 array per game (``torch.squeeze`` shape:
        ``[T]``, or 0-d for a one-position game)."""
        cfg = self.config
        A = len(cfg.action_space)
        obs, counts = [], []
        for gh in game_histories:
            T = len(gh.root_values)
            counts.append(T)
            for i in range(T):
                obs.append(numpy.asarray(gh.get_stacked_observations(i, cfg.stacked_observations, A), dtype=numpy.float32))
        if not obs:
            return [numpy.zeros(0, numpy.float32) for _ in game_histories]
        obs = numpy.stack(obs).reshape(len(obs), -1)
        values = numpy.empty(len(obs), numpy.float32)
        for lo in range(0, len(obs), self.max_positions):
            hi = min(len(obs), lo + self.max_positions)
            values[lo:hi] = self.engine.initial_inference(obs[lo:hi])["value"]
        out, off = [], 0
        for T in counts:
            v = values[off:off + T].copy()
            out.append(v.reshape(()) if T == 1 else v)
            off += T
        return out

    def reanalyse_games(self, game_histories):
        """Set ``reanalysed_predicted_root_values`` on every history (one batched inference); returns the histories."""
        if self.config.use_last_model_value:
            for gh, v in zip(game_histories, self.fresh_root_values(game_histories)):
                gh.reanalysed_predicted_root_values = v
        self.num_reanalysed_games += len(game_histories)
        return game_histories

    # ------------------------------------------------------------------ the reference's actor loop
    def reanalyse(self, replay_buffer, shared_storage):
        cfg = self.config
        while _call(shared_storage, "get_info", "num_played_games") < 1:
            time.sleep(0.1)
        while (_call(shared_storage, "get_info", "training_step") < cfg.training_steps
               and not _call(shared_storage, "get_info", "terminate")):
            self.set_weights(_call(shared_storage, "get_info", "weights"))
            sampled = [_call(replay_buffer, "sample_game", force_uniform=True) for _ in range(self.games_per_call)]
            games = {}
            for game_id, game_history, _ in sampled:       # the same game may be drawn twice: analyse it once
                games.setdefault(game_id, game_history)
            self.reanalyse_games(list(games.values()))
            for game_id, game_history in games.items():
                _fire(replay_buffer, "update_game_history", game_id, game_history)
            _fire(shared_storage, "set_info", "num_reanalysed_games", self.num_reanalysed_games)


# ----------------------------------------------------------------------------------------------------------------
# bulk ingest: PER priorities of whole games
# ----------------------------------------------------------------------------------------------------------------
def _weak_scalar_dtype(dtype):
    """dtype of ``dtype.type(1) * 1.0``: float32 under NumPy >= 2 (NEP 50, Python floats are weak), float64 under the
    value-based casting of the NumPy 1.21 the reference pins - whichever the installed NumPy does, the reference's
    scalar arithmetic on ``reanalysed_predicted_root_values`` (a float32 array) does the same."""
    return (numpy.dtype(dtype).type(1) * 1.0).dtype


def target_values(game_history, config):
    """``ReplayBuffer.compute_target_value`` (replay_buffer.py:230-262) for every position of a game at once.

    The reference starts from ``last_step_value * discount**td_steps`` (or the int 0 when the bootstrap index is past the
    end) and adds the signed rewards ``reward * discount**i`` for i = 0, 1, ... in that order; the same additions happen
    here in the same order and in the same floating-point type, element-wise over all positions, so every value is
    bit-identical to the scalar loop.  Returns a list of numpy scalars (float64, or float32 where the reference's own
    arithmetic stays in float32 because the bootstrap value comes from a float32 array)."""
    T = len(game_history.root_values)
    td, discount = int(config.td_steps), config.discount
    reanalysed = game_history.reanalysed_predicted_root_values is not None
    src = game_history.reanalysed_predicted_root_values if reanalysed else game_history.root_values
    if reanalysed:
        roots = numpy.asarray(src).reshape(-1)
        boot_dtype = _weak_scalar_dtype(roots.dtype)
    else:
        roots = numpy.asarray([0.0 if r is None else r for r in src], dtype=numpy.float64)
        boot_dtype = numpy.dtype(numpy.float64)
    to_play = numpy.asarray(game_history.to_play_history, dtype=numpy.int64)
    rewards = numpy.asarray(game_history.reward_history, dtype=numpy.float64)
    idx = numpy.arange(T)
    boot = idx + td
    has = boot < T
    acc = numpy.zeros(T, boot_dtype)                     # positions WITH a bootstrap value: its dtype rules
    plain = numpy.zeros(T, numpy.float64)                # positions without: Python-float arithmetic
    if has.any():
        b = boot[has]
        last = numpy.where(to_play[b] == to_play[idx[has]], roots[b], -roots[b]).astype(roots.dtype)
        acc[has] = (last * discount ** td).astype(boot_dtype)
    n_hist = len(rewards)
    for i in range(td):
        j = idx + 1 + i                                  # reward_history[index + 1 + i], while inside [index+1, bootstrap]
        ok = j < n_hist
        if not ok.any():
            break
        jj = numpy.where(ok, j, 0)
        same = to_play[idx] == to_play[numpy.minimum(idx + i, len(to_play) - 1)]
        term = numpy.where(same, rewards[jj], -rewards[jj]) * discount ** i
        plain = numpy.where(ok, plain + term, plain)
        acc = numpy.where(ok, acc + term.astype(boot_dtype), acc)       # a weak Python float joins in acc's own type
    return [acc[i] if has[i] else plain[i] for i in range(T)]


def initial_priorities(game_history, config):
    """The ``priorities`` array and ``game_priority`` that ``save_game`` would compute (replay_buffer.py:39-51)."""
    tv = target_values(game_history, config)
    alpha = config.PER_alpha
    pri = [numpy.abs(root_value - tv[i]) ** alpha for i, root_value in enumerate(game_history.root_values)]
    pri = numpy.array(pri, dtype="float32")
    return pri, numpy.max(pri)


def save_games(replay_buffer, game_histories, config, shared_storage=None):
    """Hand a batch of finished games to an (unmodified) ``ReplayBuffer``: priorities are attached first, so its
    per-position Python loop is skipped; every other effect of ``save_game`` (eviction, counters) is the reference's."""
    for gh in game_histories:
        if config.PER and gh.priorities is None:
            gh.priorities, gh.game_priority = initial_priorities(gh, config)
        _fire(replay_buffer, "save_game", gh, shared_storage)
