"""Randomised parity hunt (TEST TOOL, CPU only): random instruction lists with every measurement
mode, resets, barriers and random option sets are run through the backend bound to the emulated
kernels (tests/emu) and through the oracle; any difference > 1e-10, any key mismatch and any
exception raised by only one side is reported with the seed that reproduces it.

    python tests/harness/fuzz_emu.py [--seeds 200] [--start 0] [--min-n 1] [--max-n 8] [--max-ops 60]
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

from qiskit_aakash_b200 import circuits as C  # noqa: E402
from qiskit_aakash_b200.dm_simulator import assemble  # noqa: E402
from oracle import dm_oracle  # noqa: E402
from emu_backend import emu_backend  # noqa: E402

PI = np.pi


def random_options(rng):
    o = {}
    if rng.random() < 0.5:
        o["rotation_error"] = {k: [float(rng.uniform(0.9, 1.0)), float(rng.uniform(-0.05, 0.05))] for k in ("rx", "ry", "rz")}
    if rng.random() < 0.5:
        o["tsp_model_error"] = [float(rng.uniform(0.9, 1.0)), float(rng.uniform(-0.05, 0.05)) if rng.random() < 0.6 else 0.0]
    if rng.random() < 0.4:
        o["thermal_factor"] = float(rng.uniform(0, 1))
        o["decoherence_factor"] = float(rng.uniform(0.8, 1))
        o["decay_factor"] = float(rng.uniform(0.8, 1))
    if rng.random() < 0.4:
        o["depolarization_factor"] = float(rng.uniform(0.8, 1))
    if rng.random() < 0.2:
        o["merge"] = False
    if rng.random() < 0.15:
        o["chop_threshold"] = 1e-6
    return o


def random_init(rng, n, o):
    r = rng.random()
    if r < 0.12:
        o["custom_densitymatrix"] = "max_mixed"
    elif r < 0.24:
        o["custom_densitymatrix"] = "uniform_superpos"
    elif r < 0.36:
        o["custom_densitymatrix"] = "thermal_state"
        o.setdefault("thermal_factor", float(rng.uniform(0, 1)))
    elif r < 0.5:
        o["custom_densitymatrix"] = "binary_string"
        o["initial_densitymatrix"] = "".join(rng.choice(["0", "1"], size=n))


def random_circuit(rng, n, max_ops=60):
    c = C.Circuit(n, "fuzz")
    n_ops = int(rng.integers(1, max_ops))
    for _ in range(n_ops):
        r = rng.random()
        if n > 1 and r < 0.3:
            a, b = rng.choice(n, size=2, replace=False)
            c.cx(int(a), int(b))
        elif r < 0.75:
            q = int(rng.integers(n))
            ang = rng.uniform(-2 * PI, 2 * PI, 3)
            k = int(rng.integers(4))
            if k == 0:
                c.u1(ang[0], q)
            elif k == 1:
                c.u2(ang[0], ang[1], q)
            elif k == 2:
                c.u3(ang[0], ang[1], ang[2], q)
            else:
                c.iden(q)
        elif r < 0.79:
            c.barrier()
        elif r < 0.83:
            c.reset(int(rng.integers(n)))
        elif r < 0.90:
            k = int(rng.integers(1, n + 1))
            qs = sorted(int(x) for x in rng.choice(n, size=k, replace=False))
            basis = str(rng.choice(["X", "Y", "Z", "N"]))
            if rng.random() < 0.3 and k >= 2:          # mixed bases in one layer (skip quirk)
                for q in qs:
                    b = str(rng.choice(["X", "Y", "Z"]))
                    c.measure(q, q, basis=b)
            elif basis == "N":
                c.measure(qs, qs, basis="N", add_param=np.asarray(rng.uniform(-1, 1, 3)))
            else:
                c.measure(qs, qs, basis=basis)
        elif r < 0.93 and n >= 2:
            a, b = sorted(int(x) for x in rng.choice(n, size=2, replace=False))
            if n <= 10:
                c.measure(0, 0, basis="Bell", add_param="%d%d" % (a, b))
        elif r < 0.96:
            s = "".join(rng.choice(list("IXYZ"), size=n))
            c.measure(0, 0, basis="Expect", add_param=s)
        else:
            b = str(rng.choice(["X", "Y", "Z", "N"]))
            if b == "N":
                vec = np.asarray(rng.uniform(-1, 1, 3))
                c.barrier()
                for q in range(n):     # the form the reference's dispatcher expects (dm_simulator.py:1130-1132)
                    c.instructions.append(C.instr("measure", [q], ["Ensemble", ["N", vec.copy()]], memory=[q]))
                c.barrier()
            else:
                c.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param=b)
    return c


def as_arr(v):
    if isinstance(v, dict):
        return np.array(list(v.values()), dtype=complex)
    return np.asarray(v, dtype=complex)


KNOBS = ("DMB_MAX_OPS_PER_PASS", "DMB_SCHED_STRATEGY", "DMB_RESERVE_LOW", "DMB_DRAIN_TAIL", "DMB_DRAIN_CHUNK",
         "DMB_RELABEL", "DMB_FOLD_SWAPS", "DMB_DRAIN_THRESHOLD")


def random_knobs(rng):
    """Scheduler / streaming knobs the engine reads from the environment (engine.PauliEngine.__init__)."""
    for k in KNOBS:
        os.environ.pop(k, None)
    env = {}
    if rng.random() < 0.7:
        env["DMB_MAX_OPS_PER_PASS"] = int(rng.integers(1, 17))
    if rng.random() < 0.5:
        env["DMB_SCHED_STRATEGY"] = int(rng.integers(0, 2))
    if rng.random() < 0.3:
        env["DMB_RESERVE_LOW"] = int(rng.integers(0, 3))
    if rng.random() < 0.7:
        env["DMB_DRAIN_TAIL"] = int(rng.integers(1, 12))
        env["DMB_DRAIN_CHUNK"] = int(rng.integers(1, 8))
    if rng.random() < 0.2:
        env["DMB_RELABEL"] = 0
        env["DMB_DRAIN_THRESHOLD"] = int(rng.integers(1, 10))
    if rng.random() < 0.3:
        env["DMB_FOLD_SWAPS"] = 0
    for k, v in env.items():
        os.environ[k] = str(v)
    return env


def one(seed, max_n, min_n=1, max_ops=60, knobs=False, backend=None):
    """``backend``: factory of the backend under test (default: emulated kernels; the -m gpu tests pass the real one)."""
    rng = np.random.default_rng(seed)
    if knobs:
        env = random_knobs(np.random.default_rng(seed + 77777))
    n = int(rng.integers(min_n, max_n + 1))
    circ = random_circuit(rng, n, max_ops)
    opts = random_options(rng)
    random_init(rng, n, opts)
    if n <= 5 and rng.random() < 0.5:
        opts["compute_densitymatrix"] = True
    ref_exc = got_exc = None
    try:
        ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    except Exception as e:  # noqa: BLE001
        ref_exc = e
    try:
        got = (backend or emu_backend)().run(assemble(circ), backend_options=copy.deepcopy(opts)).result()["results"][0]
    except Exception as e:  # noqa: BLE001
        got_exc = e
    if ref_exc or got_exc:
        if ref_exc and got_exc:
            return "both-raise", "%s | %s" % (type(ref_exc).__name__, type(got_exc).__name__)
        return "FAIL", "one side raised: oracle=%r backend=%r" % (ref_exc, got_exc)
    if got["number_of_clock_cycles"] != ref["number_of_clock_cycles"]:
        return "FAIL", "levels %s vs %s" % (got["number_of_clock_cycles"], ref["number_of_clock_cycles"])
    if set(got["data"]) != set(ref["data"]):
        return "FAIL", "keys %s vs %s" % (sorted(got["data"]), sorted(ref["data"]))
    worst = 0.0
    for k, v in ref["data"].items():
        a, b = as_arr(v), as_arr(got["data"][k])
        if a.shape != b.shape:
            return "FAIL", "shape of %s" % k
        if isinstance(v, dict) and list(v.keys()) != list(got["data"][k].keys()):
            return "FAIL", "dict keys of %s" % k
        if a.size:
            worst = max(worst, float(np.max(np.abs(a - b))))
    if worst > 1e-10:
        return "FAIL", "n=%d max|d| = %.3e %s" % (n, worst, env if knobs else "")
    return "ok", "%.1e" % worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=200)
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--max-n", type=int, default=8)
    ap.add_argument("--min-n", type=int, default=1)
    ap.add_argument("--max-ops", type=int, default=60)
    ap.add_argument("--knobs", action="store_true", help="randomise the scheduler / streaming knobs too")
    a = ap.parse_args()
    counts = {}
    for seed in range(a.start, a.start + a.seeds):
        try:
            st, msg = one(seed, a.max_n, a.min_n, a.max_ops, a.knobs)
        except Exception:  # noqa: BLE001
            st, msg = "FAIL", traceback.format_exc()
        counts[st] = counts.get(st, 0) + 1
        if st != "ok":
            print("seed %d: %s %s" % (seed, st, msg), flush=True)
    print("summary:", counts)
    return 1 if counts.get("FAIL") else 0


if __name__ == "__main__":
    sys.exit(main())
