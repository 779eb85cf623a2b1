"""Randomised validity check of the native gate-fusion scheduler (``dmb_schedule``), TEST TOOL: random op
streams (n up to 20 qubits, brick-wall or arbitrary pairs), both strategies, random caps, one-shot or streamed
with a look-ahead tail -- every schedule must be a valid reordering by symbolic replay
(tests/test_native_schedule.py::_replay).  5 000 cases were clean when this was committed.

    python tests/harness/fuzz_schedule.py FIRST_SEED END_SEED
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_native_schedule as T  # noqa: E402
from qiskit_aakash_b200 import capi, schedule  # noqa: E402


def one(lib, seed):
    """0 if the schedule of this seed's stream replays correctly, else 1 (and a line on stdout)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 21))
    max_ops = int(rng.integers(1, 17))
    brick = bool(rng.integers(2))
    qops = T._rand_stream(rng, n, int(rng.integers(1, 300)), brick=brick)
    for k, op in enumerate(qops):
        op.kind, op.coef = (capi.OP_MATS if op.db is None else capi.OP_CX), [float(k + 1)]     # tag = position + 1
    nd = max(n, 2)
    pos0 = [int(x) for x in rng.permutation(n)]
    strat = int(rng.integers(2))
    try:
        pos = list(pos0)
        if rng.random() < 0.5:
            passes = schedule.relabel_passes(lib, qops, pos, nd, max_ops=max_ops, native=True, strategy=strat)
            T._replay(passes, qops, pos0, pos, nd, max_ops)
        else:                                  # streamed: cut anywhere, keep a tail queued as look-ahead
            queue, parts = [], []
            cut, tail = int(rng.integers(1, 60)), int(rng.integers(1, 40))
            for i in range(0, len(qops), cut):
                queue += qops[i:i + cut]
                p, queue = schedule.relabel_passes(lib, queue, pos, nd, max_ops=max_ops, min_tail=tail, native=True,
                                                   strategy=strat)
                if len(p):
                    parts.append(p)
            p = schedule.relabel_passes(lib, queue, pos, nd, max_ops=max_ops, native=True, strategy=strat)
            if len(p):
                parts.append(p)
            allp = np.concatenate(parts) if parts else np.zeros(0, dtype=capi.PASS_DTYPE)
            T._replay(allp, qops, pos0, pos, nd, max_ops)
    except Exception as e:  # noqa: BLE001
        print("seed", seed, "n", n, "max_ops", max_ops, "strategy", strat, type(e).__name__, str(e)[:200])
        return 1
    return 0


def run(first, end):
    lib = capi.load_library()
    return sum(one(lib, seed) for seed in range(first, end))


if __name__ == "__main__":
    n_bad = run(int(sys.argv[1]), int(sys.argv[2]))
    print("bad", n_bad)
    sys.exit(1 if n_bad else 0)
