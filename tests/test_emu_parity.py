"""GPU-less checks of everything above the silicon: the backend (options, merge, partition,
measurement dispatch, lazy op queue, pass scheduler, readouts) driving the *same per-thread
kernel bodies* as the CUDA library, compiled for the CPU (tests/emu), against the golden
fixtures generated from the unmodified reference.  The `-m gpu` tests repeat this through
the real library on a B200."""
import numpy as np
import pytest

import cases
from emu_backend import emu_backend
from golden_check import check_against_golden
from qiskit_aakash_b200 import assemble, circuits as C


def run_case(name, backend=None):
    case = cases.get(name)
    cases.write_files(case, ".")
    circ = C.Circuit(case["n"], name)
    circ.instructions = case["instrs"]
    backend = backend or emu_backend()
    res = backend.run(assemble(circ), backend_options=case["options"]).result()
    assert res["success"] and res["status"] == "COMPLETED"
    return res["results"][0]


@pytest.mark.parametrize("name", list(cases.CASES))
def test_backend_on_emulated_kernels_matches_golden(name, golden, case_dir):
    check_against_golden(golden, name, run_case(name))


@pytest.mark.parametrize("name", ["layered_n8_d6_noisy", "layered_n9_d4_memnoise", "layered_n10_d3_noisy", "rand_n7_fullnoise",
                                  "qft8_binary", "grover4_noisy"])
def test_fused_remap_swaps_match_golden(name, golden, case_dir, monkeypatch):
    """dmb_op.post_swap (remap swap folded into the previous op's store), opt-in via
    schedule.FUSE_SWAPS_DEFAULT: same numbers as explicit swap ops."""
    from qiskit_aakash_b200 import schedule
    monkeypatch.setattr(schedule, "FUSE_SWAPS_DEFAULT", True)
    check_against_golden(golden, name, run_case(name))
