"""Randomised validity check of the SHARDED compile (``ShardedPauliEngine.compile``), TEST TOOL: random circuits
(n up to 18, 2 / 4 / 8 ranks, random caps / strategies / parking option, compiled in 1-3 chunks like an engine that
meets readouts) checked by the symbolic replay of tests/test_sharded_compile_symbolic.py -- no state allocated.
1 500 cases were clean when this was committed.

    python tests/harness/fuzz_sharded_compile.py FIRST_SEED END_SEED
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import test_sharded_compile_symbolic as S  # noqa: E402
from qiskit_aakash_b200 import capi, circuits as C  # noqa: E402


def one(seed):
    rng = np.random.default_rng(seed)
    world = int(rng.choice([2, 4, 8]))
    n = int(rng.integers({2: 3, 4: 4, 8: 6}[world], 19))
    if rng.random() < 0.5:
        circ = cases._rand_circuit(n, int(rng.integers(5, 300)), seed, two_qubit_frac=float(rng.uniform(0.2, 0.9)))
    else:
        circ = C.random_layered(n, int(rng.integers(1, 25)), seed, readout=False)
    pairs = [tuple(i.qubits) for i in circ.instructions if i.name == "cx"]
    if not pairs:
        return 0
    e = S._bare_engine(n, world)
    e.max_ops_per_pass = int(rng.integers(3, 17))
    e.park_in_last_pass = bool(rng.integers(2))
    e.strategy = int(rng.integers(2))
    queue = [("2q", capi.OP_CX, a, b, None, None, [float(k + 1)]) for k, (a, b) in enumerate(pairs)]
    pos0 = list(e.pos)
    try:
        cuts = sorted(set(int(x) for x in rng.integers(0, len(pairs) + 1, size=int(rng.integers(0, 3)))))
        steps, prev = [], 0
        for c in cuts + [len(pairs)]:
            e.queue = queue[prev:c]
            prev = c
            steps += e.compile(final=(c == len(pairs)))
        S._replay_steps(steps, pairs, pos0, e.pos, n, e.n_loc, e.m)
    except Exception as ex:  # noqa: BLE001
        print("seed", seed, "n", n, "world", world, type(ex).__name__, str(ex)[:150])
        return 1
    return 0


def run(first, end):
    return sum(one(seed) for seed in range(first, end))


if __name__ == "__main__":
    n_bad = run(int(sys.argv[1]), int(sys.argv[2]))
    print("bad", n_bad)
    sys.exit(1 if n_bad else 0)
