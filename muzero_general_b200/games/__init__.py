"""Game plug-ins (``MuZeroConfig`` + ``Game`` per module, loaded by name like muzero.py:44-47)."""
import importlib


def load_game_module(name):
    return importlib.import_module(f"{__name__}.{name}")
