"""Kernel-level check of the relabelling store (trailing SWAP ops of a pass folded into the tile's
write-back): passes made only of swaps, through the C ABI, against NumPy axis swaps.  Shared by
the CPU test (emulated kernels) and the GPU test."""
import itertools

import numpy as np

from qiskit_aakash_b200 import capi, schedule


def swap_pass(tile_digits, swaps):
    P = np.zeros(1, dtype=capi.PASS_DTYPE)
    K = len(tile_digits)
    P[0]["n_tile_digits"] = K
    P[0]["tile_digit"][:K] = tile_digits
    for k, (a, b) in enumerate(swaps):
        op = P[0]["ops"][k]
        op["kind"] = capi.OP_SWAP
        op["a"], op["b"] = a, b
        op["fd"][:K - 2] = schedule.lane_order(K, a, b)
    P[0]["n_ops"] = len(swaps)
    return P


def expected(vec, n, tile_digits, swaps):
    arr = vec.reshape([4] * n)                      # axis n-1-p <-> digit position p
    for a, b in swaps:
        arr = np.swapaxes(arr, n - 1 - tile_digits[a], n - 1 - tile_digits[b])
    return np.ascontiguousarray(arr).reshape(-1)


def check_relabelling_store(make_engine, n=8, tile_digits=(0, 1, 3, 4, 6, 7)):
    """Every single transposition and a sample of double / triple ones; returns folded-swap count."""
    rng = np.random.default_rng(8)
    vec = rng.normal(size=4 ** n)
    pairs = list(itertools.combinations(range(6), 2))
    combos = [[p] for p in pairs]
    combos += [[(0, x), (1, y)] for x in range(2, 6) for y in range(2, 6) if x != y]
    combos += [[(1, 2), (0, 2)], [(0, 3), (0, 4), (1, 3)], [(2, 5), (3, 4)], [(0, 1), (1, 5)], [(0, 5), (0, 5)]]
    e = make_engine(n)
    folded0 = e.stats().get("folded_swaps", 0)
    for swaps in combos:
        e.upload(vec)
        e.ctx.apply_passes(e.sptr, e.n_bits, swap_pass(list(tile_digits), swaps))
        got = np.empty(4 ** n)
        e.ctx.download(e.sptr, got)
        want = expected(vec, n, list(tile_digits), swaps)
        assert np.array_equal(got, want), swaps
    return e.stats().get("folded_swaps", 0) - folded0, sum(len(s) for s in combos)
