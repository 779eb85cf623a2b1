"""Generate ``tests/golden/golden_n12.npz``: FULL-circuit results of the reference's own, unmodified
``dm_simulator.py`` / ``basicaertools.py`` (loaded by path through ``oracle/ref_harness.py``) at n = 12,
the size BASELINE.json's north_star quotes the 1e-10 parity bar at.

    python tests/golden/make_golden_n12.py          # ~10 CPU-minutes, needs /root/reference

Cases (``tests/cases_n12.py``): Grover-12 (BASELINE configs[1]: 7 search + 5 ancilla qubits, noise.ipynb
option set, partial Z readout; 315 instructions, 176 levels) and the layered U3+CX circuit of configs[2]
at n = 12 (depth 6, per-gate noise, ensemble-Z readout).  The 4^12 coefficient vector (128 MiB) is kept
as a strided sample ``vec[::STRIDE]`` plus sum / sum of squares / a signed checksum, the probability
dictionaries in full.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases_n12  # noqa: E402
from oracle import ref_harness  # noqa: E402


def main():
    out = {}
    for name in cases_n12.CASES:
        case = cases_n12.get(name)
        t0 = time.time()
        res = ref_harness.run_reference(case["n"], case["instrs"], case["options"], name=name)
        cases_n12.pack(name, res, out)
        print("%-24s n=%d levels=%d keys=%s  %.0f s" % (name, case["n"], res["number_of_clock_cycles"],
                                                       sorted(res["data"].keys()), time.time() - t0), flush=True)
    path = os.path.join(HERE, "golden_n12.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
