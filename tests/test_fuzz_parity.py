"""Randomised parity (CPU): random instruction lists with every measurement mode, resets,
barriers and random option sets through the backend on the emulated kernels vs the oracle
(tests/harness/fuzz_emu.py is the long-running form of the same hunt; 1700 seeds were clean when this
slice was committed).  A case where the reference itself crashes must raise here too."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))

import fuzz_emu  # noqa: E402


@pytest.mark.parametrize("seed", range(40))
def test_random_small_circuit_matches_oracle(seed):
    status, msg = fuzz_emu.one(seed, 7)
    assert status in ("ok", "both-raise"), msg


@pytest.mark.parametrize("seed", range(9000, 9004))
def test_random_long_circuit_matches_oracle(seed):
    status, msg = fuzz_emu.one(seed, 8, 8, 200)
    assert status in ("ok", "both-raise"), msg


@pytest.mark.parametrize("world,mode,seed", [(2, "pull", 1), (4, "pull", 1), (8, "pull", 9), (4, "pull", 123), (8, "pull", 67),
                                             (2, "push", 5), (4, "nccl", 6), (8, "push", 7), (8, "nccl", 8)])
def test_random_circuit_on_the_sharded_engine_matches_oracle(world, mode, seed):
    """Thread cluster + emulated kernels (tests/harness/fuzz_sharded.py; 3000+ clean runs).  The pull seeds are
    the ones that exposed the scratch-shard race: a rank that finished its fused pull used the
    scratch shard as N-basis readout workspace while a slower peer was still pulling from it
    (fixed by ShardedPauliEngine._own_scratch)."""
    import fuzz_sharded
    status, msg = fuzz_sharded.one(seed, world, mode)
    assert status in ("ok", "both-raise"), msg


def test_scratch_race_minimal_case():
    """cx on a global qubit (one fused pull) followed at once by an N-basis partial measurement."""
    import copy
    import numpy as np
    import test_fused_exchange_cpu as tfx
    from oracle import dm_oracle
    from qiskit_aakash_b200 import circuits as C
    n = 5
    circ = C.Circuit(n)
    circ.cx(0, 1)
    vec = np.array([0.67113843, -0.43624435, -0.56956367])
    circ.measure([1, 4], [1, 4], basis="N", add_param=vec)
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), {})
    for _ in range(3):
        for res, _x, _p in tfx._run_world(2, n, circ, {}, True, "pull"):
            for k in ("partial_probability", "coeffmatrix"):
                assert np.max(np.abs(fuzz_emu.as_arr(res["data"][k]) - fuzz_emu.as_arr(ref["data"][k]))) <= 1e-10
