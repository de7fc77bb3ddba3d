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
