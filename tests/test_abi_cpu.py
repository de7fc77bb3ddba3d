"""CPU-side checks of the boundary: the library loads, exports every declared symbol, the
Python structs match the header, and there is no CPU fallback."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from muzero_general_b200 import _lib
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "mzb200.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+)?[a-z0-9_]+\s*\*?\s*(mz_[a-z0-9_]+)\s*\(", header, re.M))
    assert declared, "no prototypes parsed"
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mz_abi_version() == 1


def test_struct_sizes_match_header_layout():
    from muzero_general_b200 import _lib
    # MzNetDesc: 7 + 5*(1+8) + 5 + 3*(1+8) + 1 int32
    assert ctypes.sizeof(_lib.MzNetDesc) == 4 * (7 + 5 * 9 + 5 + 3 * 9 + 1)
    assert ctypes.sizeof(_lib.MzSearchDesc) == 16 + 5 * 8 + 8 + 24
    assert ctypes.sizeof(_lib.MzSearchIO) == 8 + 3 * 8 + 8 + 4 * 8 + 7 * 8 + 2 * 8


def test_no_cpu_fallback(game_configs):
    """Without a GPU mz_create must fail loudly rather than compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from muzero_general_b200 import _lib
    from muzero_general_b200.engine import SearchEngine
    with pytest.raises(_lib.MzError) as e:
        SearchEngine(game_configs["cartpole"], max_games=4)
    assert "no CPU fallback" in str(e.value)


def test_more_than_two_players_rejected(game_configs):
    """self_play.py:429-430 raises NotImplementedError; so does the boundary."""
    import copy
    cfg = copy.deepcopy(game_configs["cartpole"])
    cfg.players = [0, 1, 2]
    from muzero_general_b200.engine import SearchEngine
    with pytest.raises(NotImplementedError):
        SearchEngine(cfg, max_games=1)


def test_bench_conv_flops_agree_with_survey_table():
    """bench.py derives the tensor roofline's algorithmic FLOPs from the network shape; the 3x3 convolutions must account
    for (almost) all of the per-inference FLOPs quoted in SURVEY.md section 8 (the rest are the 1x1 convs and FC heads)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from muzero_general_b200.games import load_game_module
    from muzero_general_b200.netspec import netspec_from_config
    for game, N in (("tictactoe", 50), ("connect4", 200), ("breakout", 50)):
        ns = netspec_from_config(load_game_module(game).MuZeroConfig())
        f0, f1 = bench.NET_FLOPS[game]
        conv = bench.conv3x3_flops(ns, N)
        total = f0 + N * f1
        assert 0.85 * total < conv <= total, (game, conv, total)
