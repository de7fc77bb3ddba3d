"""ORACLE support: compile oracle/tree_oracle.c with gcc into oracle/_ref/libtree_oracle.so."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tree_oracle.c")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libtree_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", LIB, SRC, "-lm"])
    return LIB


def load():
    lib = ctypes.CDLL(build())
    lib.mz_oracle_tree_search.restype = ctypes.c_int
    return lib


def tree_search(n, N, A, P, discount, pb_c_base, pb_c_init, frac, legal, to_play, noise, first_index, seed,
                game_id, move_index, teacher, D=None):
    """numpy wrapper; returns dict(visit_counts, root_value, max_depth, ties, range, depth, actions)."""
    import numpy
    lib = load()
    D = D or max(N, 1)
    c = lambda a, dt: None if a is None else numpy.ascontiguousarray(a, dtype=dt)
    legal, to_play, noise = c(legal, numpy.uint8), c(to_play, numpy.int32), c(noise, numpy.float64)
    first_index, game_id, move_index = c(first_index, numpy.int32), c(game_id, numpy.int64), c(move_index, numpy.int32)
    t = {k: c(v, numpy.float32) for k, v in teacher.items()}
    out = dict(visit_counts=numpy.zeros((n, A), numpy.int32), root_value=numpy.zeros(n), max_depth=numpy.zeros(n, numpy.int32),
               ties=numpy.zeros(n, numpy.int32), range=numpy.zeros((n, 2)), depth=numpy.zeros((n, N), numpy.int32),
               actions=numpy.zeros((n, N, D), numpy.uint8))
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.mz_oracle_tree_search(
        ctypes.c_int(n), ctypes.c_int(N), ctypes.c_int(A), ctypes.c_int(P), ctypes.c_double(discount),
        ctypes.c_double(pb_c_base), ctypes.c_double(pb_c_init), ctypes.c_double(frac),
        p(legal), p(to_play), p(noise), p(first_index), ctypes.c_uint64(seed), p(game_id), p(move_index),
        p(t["root_reward"]), p(t["root_priors"]), p(t["value"]), p(t["reward"]), p(t["priors"]),
        p(out["visit_counts"]), p(out["root_value"]), p(out["max_depth"]), p(out["ties"]), p(out["range"]),
        p(out["depth"]), p(out["actions"]), ctypes.c_int(D))
    assert rc == 0
    return out


if __name__ == "__main__":
    print(build(force=True))
