"""Randomised parity (CPU): random instruction lists with every measurement mode, resets,
barriers and random option sets through the backend on the emulated kernels vs the oracle
(tools/fuzz_emu.py is the long-running form of the same hunt; 1700 seeds were clean when this
slice was committed).  A case where the reference itself crashes must raise here too."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import fuzz_emu  # noqa: E402


@pytest.mark.parametrize("seed", range(40))
def test_random_small_circuit_matches_oracle(seed):
    status, msg = fuzz_emu.one(seed, 7)
    assert status in ("ok", "both-raise"), msg


@pytest.mark.parametrize("seed", range(9000, 9004))
def test_random_long_circuit_matches_oracle(seed):
    status, msg = fuzz_emu.one(seed, 8, 8, 200)
    assert status in ("ok", "both-raise"), msg
