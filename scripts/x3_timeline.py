#!/usr/bin/env python
"""Per-layer timeline of conv_tower_x3_kernel (CTA 0) on the Connect4 BASELINE shape.

Run as  MZ_NO_GRAPH=1 MZ_X3_TIMELINE=<file> python scripts/x3_timeline.py  : the library appends, for every tower launch,
clock64 stamps per (layer, tile): MMA issue start / issued / epilogue sees the accumulator / tile rewritten.  This script
runs a 6-simulation search of 1024 games and prints the median per-layer durations of the dynamics-tower launches."""
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muzero_general_b200.engine import SearchEngine
from muzero_general_b200.games import load_game_module
from muzero_general_b200.netspec import netspec_from_config, synthetic_weights

path = os.environ["MZ_X3_TIMELINE"]
cfg = load_game_module("connect4").MuZeroConfig()
spec = netspec_from_config(cfg)
B = 1024
eng = SearchEngine(cfg, max_games=B, device=0, num_simulations=6, seed=0)
eng.load_weights(synthetic_weights(spec, 0))
rs = numpy.random.RandomState(0)
obs = rs.random_sample((B, eng.obs_elems)).astype(numpy.float32)
for _ in range(2):
    eng.search(obs=obs, add_exploration_noise=False, game_id=numpy.arange(B, dtype=numpy.int64))
eng.close()

launches, cur = [], None
for line in open(path):
    if line.startswith("launch"):
        cur = {"hdr": line.split(), "rows": []}
        launches.append(cur)
    else:
        cur["rows"].append([int(v) for v in line.split()])
full = [l for l in launches if l["hdr"][1] == "boards=512"]
by_layers = {}
for l in full:
    by_layers.setdefault(len(l["rows"]) // 2, []).append(numpy.array(l["rows"], dtype=numpy.int64))
for nl, arr in sorted(by_layers.items()):
    a = numpy.median(numpy.stack(arr[len(arr) // 2:]), axis=0)          # warm launches
    st = numpy.median(numpy.array([[int(v.split("=")[1]) for v in l["hdr"][3:6]] for l in full if len(l["rows"]) // 2 == nl][len(arr) // 2:]), axis=0)
    print(f"### {nl}-layer towers ({len(arr)} launches of 512 boards), cycles from the first MMA issue of CTA 0")
    print(f"kernel entry {int(st[0])}, setup done {int(st[1])}, outputs stored {int(st[2])}\n")
    print("| layer | tile | MMA issue starts | issued | epilogue starts | tile rewritten | issue | issue -> accumulator | epilogue |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in a:
        l, k, t0, t1, t2, t3 = [int(v) for v in r]
        print(f"| {l} | {k} | {t0} | {t1} | {t2} | {t3} | {t1 - t0} | {t2 - t0} | {t3 - t2} |")
