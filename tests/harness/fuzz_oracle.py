"""Randomised pinning of the ORACLE against the LIVE reference simulator (TEST TOOL; needs
/root/reference, build container only): tests/harness/fuzz_emu.py's random circuits and option sets are
run through the unmodified ``DmSimulatorPy.run_experiment`` (oracle/ref_harness.py) and through
oracle/dm_oracle.py; every result key must agree to 1e-13 and both must raise on the same inputs.

    python tests/harness/fuzz_oracle.py [--seeds 300] [--start 0] [--max-n 6]
"""
import argparse
import contextlib
import copy
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=300)
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--max-n", type=int, default=6)
    a = ap.parse_args()
    from oracle import dm_oracle, ref_harness
    import fuzz_emu
    counts = {}
    for seed in range(a.start, a.start + a.seeds):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(1, a.max_n + 1))
        circ = fuzz_emu.random_circuit(rng, n, 60)
        opts = fuzz_emu.random_options(rng)
        fuzz_emu.random_init(rng, n, opts)
        if n <= 5 and rng.random() < 0.5:
            opts["compute_densitymatrix"] = True
        ref = got = e_ref = e_got = None
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                ref = ref_harness.run_reference(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
            except Exception as e:  # noqa: BLE001
                e_ref = e
            try:
                got = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
            except Exception as e:  # noqa: BLE001
                e_got = e
        st, msg = "ok", ""
        if e_ref or e_got:
            st = "both-raise" if (e_ref and e_got) else "FAIL"
            msg = "n=%d reference: %r / oracle: %r" % (n, e_ref, e_got)
        elif ref["number_of_clock_cycles"] != got["number_of_clock_cycles"] or set(ref["data"]) != set(got["data"]):
            st, msg = "FAIL", "levels / keys"
        else:
            for k, v in ref["data"].items():
                x, y = fuzz_emu.as_arr(v), fuzz_emu.as_arr(got["data"][k])
                if isinstance(v, dict) and list(v.keys()) != list(got["data"][k].keys()):
                    st, msg = "FAIL", "dict keys of %s" % k
                elif x.shape != y.shape or (x.size and float(np.max(np.abs(x - y))) > 1e-13):
                    st, msg = "FAIL", "n=%d %s differs by %.3e" % (n, k, float(np.max(np.abs(x - y))) if x.shape == y.shape else -1)
        counts[st] = counts.get(st, 0) + 1
        if st == "FAIL":
            print("seed %d: %s" % (seed, msg), flush=True)
    print("summary:", counts)
    return 1 if counts.get("FAIL") else 0


if __name__ == "__main__":
    sys.exit(main())
