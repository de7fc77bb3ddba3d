"""ORACLE support: import the UNMODIFIED reference from /root/reference in this container.

The reference imports ``ray`` at module top and decorates its actors with ``@ray.remote``
(``self_play.py:5,11``); ``ray`` and ``gym`` are not installed here, so two stub modules
are placed in ``sys.modules`` first (``remote`` = identity decorator, ``get`` = identity).
Nothing is copied from the reference; it is only executed, to produce the golden
fixtures under ``tests/golden/`` and to cross-check the restatement in ``oracle/``.

This only works where /root/reference exists (NOT on the GPU box): GPU tests, smoke() and
bench.py never import this file.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MZ_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "self_play.py"))


def _stub_modules():
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")

        def remote(*args, **kwargs):
            if len(args) == 1 and callable(args[0]) and not kwargs:
                return args[0]
            return lambda cls: cls

        ray.remote = remote
        ray.get = lambda x: x
        sys.modules["ray"] = ray
    if "gym" not in sys.modules:
        sys.modules["gym"] = types.ModuleType("gym")


def load_reference():
    """Return (self_play, models, replay_buffer, trainer) modules of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _stub_modules()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mods = [importlib.import_module(n) for n in ("self_play", "models", "replay_buffer", "trainer")]
    return tuple(mods)


def load_reference_game(name):
    load_reference()
    return importlib.import_module(f"games.{name}")
