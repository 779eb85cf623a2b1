"""Randomised parity hunt for the SHARDED engine (TEST TOOL, CPU only): the ranks are threads of
one process (tests/test_fused_exchange_cpu.py's ThreadCluster), kernels are the emulated ones, the
circuits / option sets are tests/harness/fuzz_emu.py's (every measurement mode, resets, barriers).

    python tests/harness/fuzz_sharded.py [--seeds 100] [--start 0] [--worlds 2,4,8] [--modes pull,push,nccl]
"""
import argparse
import copy
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))

import fuzz_emu  # noqa: E402
from oracle import dm_oracle  # noqa: E402
import test_fused_exchange_cpu as tfx  # noqa: E402


def one(seed, world, mode, max_ops=80):
    rng = np.random.default_rng(seed)
    lo = {2: 3, 4: 4, 8: 6}[world]           # at least one local slot pair besides the global digits
    n = int(rng.integers(lo, 9))
    circ = fuzz_emu.random_circuit(rng, n, max_ops)
    opts = fuzz_emu.random_options(rng)
    fuzz_emu.random_init(rng, n, opts)
    if n <= 5 and rng.random() < 0.5:
        opts["compute_densitymatrix"] = True
    ref_exc = got_exc = None
    try:
        ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    except Exception as e:  # noqa: BLE001
        ref_exc = e
    try:
        outs = tfx._run_world(world, n, circ, opts, mode != "nccl", mode)
    except Exception as e:  # noqa: BLE001
        got_exc = e
    if ref_exc or got_exc:
        if ref_exc and got_exc:
            return "both-raise", ""
        return "FAIL", "n=%d one side raised: oracle=%r backend=%r" % (n, ref_exc, got_exc)
    for rank, (got, _x, _p) in enumerate(outs):
        if got["number_of_clock_cycles"] != ref["number_of_clock_cycles"]:
            return "FAIL", "levels"
        if set(got["data"]) != set(ref["data"]):
            return "FAIL", "n=%d rank %d keys %s vs %s" % (n, rank, sorted(got["data"]), sorted(ref["data"]))
        for k, v in ref["data"].items():
            a, b = fuzz_emu.as_arr(v), fuzz_emu.as_arr(got["data"][k])
            if a.shape != b.shape:
                return "FAIL", "shape of %s" % k
            if a.size and float(np.max(np.abs(a - b))) > 1e-10:
                return "FAIL", "n=%d rank %d %s max|d| = %.3e" % (n, rank, k, float(np.max(np.abs(a - b))))
    return "ok", ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--worlds", default="2,4,8")
    ap.add_argument("--modes", default="pull,push,nccl")
    a = ap.parse_args()
    counts = {}
    for seed in range(a.start, a.start + a.seeds):
        for world in [int(x) for x in a.worlds.split(",")]:
            for mode in a.modes.split(","):
                try:
                    st, msg = one(seed, world, mode)
                except Exception:  # noqa: BLE001
                    st, msg = "FAIL", traceback.format_exc()
                counts[st] = counts.get(st, 0) + 1
                if st == "FAIL":
                    print("seed %d world %d mode %s: %s" % (seed, world, mode, msg), flush=True)
    print("summary:", counts)
    return 1 if counts.get("FAIL") else 0


if __name__ == "__main__":
    sys.exit(main())
