"""Randomised check of the circuit facade against the LIVE reference front-end (TEST TOOL; needs
/root/reference, build container only): the same random program is built through
qiskit_aakash_b200.frontend and through the reference's QuantumCircuit (oracle/frontend_harness.py),
and the instruction lists both hand to the simulator must be identical record by record
(names, order, qubit / clbit indices, parameter bits).

    python tests/harness/fuzz_frontend.py [--seeds 300] [--start 0]
"""
import argparse
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ONE = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "iden"]
ONE_P = [("rx", 1), ("ry", 1), ("rz", 1), ("u1", 1), ("u2", 2), ("u3", 3)]
TWO = ["cx", "cy", "cz", "ch", "swap"]
TWO_P = [("crz", 1), ("cu1", 1), ("cu3", 3), ("rzz", 1)]
THREE = ["ccx", "cswap"]


def program(seed):
    """A register layout and a list of (method, args) records, independent of either API."""
    rng = np.random.default_rng(seed)
    n_regs = int(rng.integers(1, 4))
    sizes = [int(rng.integers(1, 6)) for _ in range(n_regs)]
    while sum(sizes) < 3:
        sizes[0] += 1
    n = sum(sizes)
    names = list(rng.permutation(["q", "anc", "b", "zz", "a10"]))[:n_regs]   # names drive the str(qargs) ordering
    steps = []
    for _ in range(int(rng.integers(1, 70))):
        r = rng.random()
        qs = [int(x) for x in rng.choice(n, 3, replace=False)]
        if r < 0.25:
            steps.append((ONE[int(rng.integers(len(ONE)))], [], qs[:1]))
        elif r < 0.45:
            name, k = ONE_P[int(rng.integers(len(ONE_P)))]
            steps.append((name, [float(x) for x in rng.uniform(-7, 7, k)], qs[:1]))
        elif r < 0.6:
            steps.append((TWO[int(rng.integers(len(TWO)))], [], qs[:2]))
        elif r < 0.72:
            name, k = TWO_P[int(rng.integers(len(TWO_P)))]
            steps.append((name, [float(x) for x in rng.uniform(-7, 7, k)], qs[:2]))
        elif r < 0.78:
            steps.append((THREE[int(rng.integers(2))], [], qs))
        elif r < 0.82:
            steps.append(("barrier", [], sorted(set(qs[:int(rng.integers(0, 4))]))))
        elif r < 0.86:
            steps.append(("reset", [], qs[:1]))
        elif r < 0.9:
            steps.append(("reg_h", [], [int(rng.integers(n_regs))]))                 # whole-register broadcast
        elif r < 0.96:
            b = ["X", "Y", "Z", None][int(rng.integers(4))]
            steps.append(("measure", [b], qs[:1]))
        else:
            steps.append(("measure_n", [[float(x) for x in rng.uniform(-1, 1, 3)]], qs[:1]))
    fin = rng.random()
    if fin < 0.3:
        steps.append(("ensemble", ["XYZ"[int(rng.integers(3))]], []))
    elif fin < 0.45:
        steps.append(("expect", ["".join(rng.choice(list("IXYZ"), size=n))], []))
    elif fin < 0.55:
        a, b = sorted(int(x) for x in rng.choice(n, 2, replace=False))
        steps.append(("bell", ["%d%d" % (a, b)], []))
    return sizes, names, steps


def build(api, sizes, names, steps):
    qregs = [api.QuantumRegister(s, nm) for s, nm in zip(sizes, names)]
    n = sum(sizes)
    creg = api.ClassicalRegister(n, "c")
    qc = api.QuantumCircuit(*qregs, creg)
    bits = [r[i] for r in qregs for i in range(len(r))]
    for name, params, qs in steps:
        if name == "barrier":
            qc.barrier(*[bits[q] for q in qs])
        elif name == "reset":
            qc.reset(bits[qs[0]])
        elif name == "reg_h":
            qc.h(qregs[qs[0]])
        elif name == "measure":
            if params[0] is None:
                qc.measure(bits[qs[0]], creg[qs[0]])
            else:
                qc.measure(bits[qs[0]], creg[qs[0]], basis=params[0])
        elif name == "measure_n":
            qc.measure(bits[qs[0]], creg[qs[0]], basis="N", add_param=np.array(params[0]))
        elif name == "ensemble":
            for r in qregs:
                pass
            qc.measure(qregs[0], [creg[i] for i in range(len(qregs[0]))], basis="Ensemble", add_param=params[0])
        elif name == "expect":
            qc.measure(bits[0], creg[0], basis="Expect", add_param=params[0])
        elif name == "bell":
            qc.measure(bits[0], creg[0], basis="Bell", add_param=params[0])
        else:
            getattr(qc, name)(*params, *[bits[q] for q in qs])
    return qc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=300)
    ap.add_argument("--start", type=int, default=0)
    a = ap.parse_args()
    warnings.simplefilter("ignore")
    from oracle import frontend_harness
    ref_api = frontend_harness.load()
    from qiskit_aakash_b200 import frontend
    import test_frontend as tf
    my_api = SimpleNamespace(QuantumCircuit=frontend.QuantumCircuit, QuantumRegister=frontend.QuantumRegister,
                             ClassicalRegister=frontend.ClassicalRegister, pi=frontend.pi)
    counts = {}
    for seed in range(a.start, a.start + a.seeds):
        sizes, names, steps = program(seed)
        ref = err_ref = got = err_got = None
        try:
            ref = ref_api.lower(build(ref_api, sizes, names, steps))
        except Exception as e:  # noqa: BLE001
            err_ref = e
        try:
            n_q, n_c, instrs = build(my_api, sizes, names, steps).lowered()
            got = tf.as_records(instrs)
        except Exception as e:  # noqa: BLE001
            err_got = e
        if err_ref or err_got:
            st = "both-raise" if (err_ref and err_got) else "FAIL"
            msg = "reference: %r / facade: %r" % (err_ref, err_got)
        elif (n_q, n_c) != (ref["n_qubits"], ref["memory_slots"]):
            st, msg = "FAIL", "sizes"
        elif got != ref["instructions"]:
            k = next((i for i, (x, y) in enumerate(zip(got, ref["instructions"])) if x != y), min(len(got), len(ref["instructions"])))
            st, msg = "FAIL", "instruction %d of %d/%d: %r != %r" % (k, len(got), len(ref["instructions"]),
                                                                     got[k] if k < len(got) else None,
                                                                     ref["instructions"][k] if k < len(ref["instructions"]) else None)
        else:
            st, msg = "ok", ""
        counts[st] = counts.get(st, 0) + 1
        if st != "ok":
            print("seed %d: %s %s" % (seed, st, msg), flush=True)
    print("summary:", counts)
    return 1 if counts.get("FAIL") else 0


if __name__ == "__main__":
    sys.exit(main())
