"""ORACLE (test infrastructure). Philox4x32-10 counter-based generator in plain Python.

Restates the published algorithm (Salmon et al., "Parallel random numbers: as easy as
1, 2, 3", SC'11; constants M0=0xD2511F53, M1=0xCD9E8D57, W0=0x9E3779B9, W1=0xBB67AE85)
that ``muzero_general_b200/csrc/philox.cuh`` implements on the device, so that tie-breaks
decided on the GPU can be replayed on the CPU.  Known-answer vectors from the Random123
distribution are checked in ``tests/test_philox.py``.
"""
M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = (int(x) & MASK for x in counter)
    k0, k1 = (int(x) & MASK for x in key)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


# Stream tags (key word 1 is seed_hi ^ tag); must match csrc/philox.cuh
TAG_TIE = 0x7169E001
TAG_NOISE = 0x7169E002
TAG_ACTION = 0x7169E003


def tie_index(seed, game, move, sim, depth, n_tied):
    """Index in [0, n_tied) used by the device for an exact UCB tie (mulhi of word 0)."""
    w = philox4x32_10((game & MASK, move, sim, depth), (seed & MASK, ((seed >> 32) & MASK) ^ TAG_TIE))
    return (w[0] * n_tied) >> 32
