import sys, os, numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_general_b200 import self_play as sp
from muzero_general_b200.games import load_game_module
from muzero_general_b200.netspec import netspec_from_config, synthetic_weights
mod = load_game_module("tictactoe")
cfg = mod.MuZeroConfig()
cfg.num_parallel_games, cfg.rng_mode, cfg.num_simulations = 8192, "philox", 50
spec = netspec_from_config(cfg)
w = sp.SelfPlay({"weights": synthetic_weights(spec, 0)}, mod.Game, cfg, seed=0)
w.play_moves(1, 1.0)
dl = w._device_loop
loop = dl.loop
prev = loop.stats.env_steps
for i in range(12):
    k = [1, 2, 4, 8, 16, 16, 16, 16, 16, 16, 16, 16][i]
    st = loop.moves(k, 1.0)
    buf, idx = loop.drain()
    print(f"call {i}: k={k} env_steps +{st.env_steps - prev} ({(st.env_steps - prev) / k:.0f}/move) staged_games={st.staged_games} staged_bytes={st.staged_bytes} parked={st.parked_slots} cap={st.staging_capacity} dev_ms={st.device_ms:.1f} drained={len(idx)}")
    prev = st.env_steps
