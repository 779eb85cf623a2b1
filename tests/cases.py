"""Named parity cases shared by the golden-fixture generator, the oracle tests and the
GPU parity tests.  Every case is ``name -> dict(n=..., instrs=[...], options={...})`` and is
fully determined by its name (seeded), so the fixture produced in the build container by
the live reference can be re-derived anywhere.

Coverage follows the reference's behaviours listed in SURVEY.md section 8a: all initial
states (a4), u1/u2/u3 merging (a12-a14), both CX orientations at all distances (a9), the
three noise families (a7 rotation_error, a9 tsp_model_error, a11 memory noise), every
measurement mode and its quirks (a18-a23), reset (a24), barriers/partition (a15-a17).
"""
from __future__ import annotations

import math

import numpy as np

from qiskit_aakash_b200 import circuits as C

PI = math.pi


def _rand_circuit(n, n_ops, seed, two_qubit_frac=0.35, with_u1u2=True):
    """Unstructured random circuit: arbitrary qubit pairs (both orientations), u1/u2/u3/id."""
    rng = np.random.default_rng(seed)
    c = C.Circuit(n, "rnd")
    for _ in range(n_ops):
        if n > 1 and rng.random() < two_qubit_frac:
            a, b = rng.choice(n, size=2, replace=False)
            c.cx(int(a), int(b))
        else:
            q = int(rng.integers(n))
            kind = int(rng.integers(4)) if with_u1u2 else 2
            ang = rng.uniform(-2 * PI, 2 * PI, 3)
            if kind == 0:
                c.u1(ang[0], q)
            elif kind == 1:
                c.u2(ang[0], ang[1], q)
            elif kind == 2:
                c.u3(ang[0], ang[1], ang[2], q)
            else:
                c.iden(q)
                c.u3(ang[0], ang[1], ang[2], q)
    return c


FULL_NOISE = {
    "rotation_error": {"rx": [0.99, 0.01], "ry": [0.98, -0.02], "rz": [0.97, 0.03]},
    "tsp_model_error": [0.985, 0.04],
    "thermal_factor": 0.2, "decoherence_factor": 0.93, "decay_factor": 0.96,
    "depolarization_factor": 0.97,
}


def _cases():
    cs = {}

    def add(name, circ, options=None, files=None):
        cs[name] = dict(n=circ.n_qubits, instrs=circ.instructions, options=dict(options or {}),
                        files=dict(files or {}))

    # ---- recorded examples of the reference (README.md:44-62 and user-guide notebooks) ----
    c = C.Circuit(2); c.x(1); c.cx(0, 1)
    add("readme_x_cx", c)
    add("init_max_mixed_1", C.Circuit(1), {"custom_densitymatrix": "max_mixed"})
    add("init_binary_01", C.Circuit(2), {"custom_densitymatrix": "binary_string",
                                         "initial_densitymatrix": "01"})
    c = C.Circuit(1); c.s(0); c.measure(0, 0, basis="X")
    add("nb_measure_x", c)
    c = C.Circuit(1); c.s(0); c.measure(0, 0, basis="N", add_param=np.array([1, 2, 3]))
    add("nb_measure_n", c)
    c = C.Circuit(3); c.measure(0, 0, basis="Bell", add_param="01")
    add("nb_bell_000", c)
    c = C.ghz(3); c.measure([0, 1, 2], [0, 1, 2], basis="Ensemble", add_param="X")
    add("nb_ghz_ensemble_x", c)
    c = C.ghz(3); c.measure(0, 0, basis="Expect", add_param="ZIZ")
    add("nb_ghz_expect_ziz", c)
    # noise.ipynb cell 5 (insertion order; SURVEY.md section 8c records the oracle values)
    def noise_nb():
        c = C.Circuit(3)
        c.u1(3.6, 0); c.cx(0, 1); c.u1(2.6, 2); c.cx(1, 0); c.s(2); c.y(2)
        c.measure([0, 1, 2], [0, 1, 2], basis="Ensemble", add_param="Z")
        return c
    add("nb_noise_clean", noise_nb())
    add("nb_noise_noisy", noise_nb(), {"thermal_factor": 0., "decoherence_factor": .9,
                                       "depolarization_factor": 0.99,
                                       "bell_depolarization_factor": 0.99, "decay_factor": 0.99,
                                       "rotation_error": {"rx": [1., 0.], "ry": [1., 0.], "rz": [1., 0.]},
                                       "tsp_model_error": [1., 0.]})
    # partition.ipynb cell 6 in the transpiled order shown by its recorded output
    c = C.Circuit(3)
    c.u1(3.6, 0); c.cx(0, 1); c.u1(2.6, 2)
    c.measure(1, 1, basis="X"); c.measure(0, 0, basis="Y")
    c.cx(1, 0); c.s(2); c.y(2)
    c.measure(1, 1, basis="Bell", add_param="12"); c.measure(0, 0)
    c.measure([0, 1, 2], [0, 1, 2], basis="Ensemble", add_param="X")
    add("nb_partition_mixed_measures", c)

    # ---- initial states (a4) ----
    for mode in ("max_mixed", "uniform_superpos", "thermal_state"):
        add("init_%s_rand3" % mode, _rand_circuit(3, 12, 11),
            {"custom_densitymatrix": mode, "thermal_factor": 0.3})
    add("init_binary_rand4", _rand_circuit(4, 14, 12),
        {"custom_densitymatrix": "binary_string", "initial_densitymatrix": "0110"})
    rng = np.random.default_rng(5)
    v = rng.normal(size=64) * 0.01; v[0] = 2.0 ** -3
    # a raw vector is only reachable through the 'stored_density_matrix.npy' file (a4)
    add("init_stored_rand3", _rand_circuit(3, 10, 13),
        {"custom_densitymatrix": "stored_density_matrix", "initial_densitymatrix": True},
        files={"stored_density_matrix.npy": v})
    # compare / fidelity against 'stored_coefficients.npy' (a27), evaluated at the ensemble measure
    w = rng.normal(size=64) * 0.05
    c = _rand_circuit(3, 10, 14); c.measure([0, 1, 2], [0, 1, 2], basis="Ensemble", add_param="Z")
    add("compare_fidelity_rand3", c, {"compare": True}, files={"stored_coefficients.npy": w})

    # ---- gates and noise, unstructured circuits, n = 1..7 ----
    for n, n_ops, seed in ((1, 9, 21), (2, 16, 22), (3, 25, 23), (4, 30, 24), (5, 40, 25),
                           (6, 40, 26), (7, 50, 27)):
        add("rand_n%d_clean" % n, _rand_circuit(n, n_ops, seed))
        add("rand_n%d_fullnoise" % n, _rand_circuit(n, n_ops, seed + 100), FULL_NOISE)
    add("rand_n4_nomerge", _rand_circuit(4, 30, 31), dict(FULL_NOISE, merge=False))
    add("rand_n3_chop", _rand_circuit(3, 20, 32), {"chop_threshold": 1e-3})
    add("rand_n5_nomatrix", _rand_circuit(5, 30, 33), {"compute_densitymatrix": False})
    # near-inverse u3 pairs: exercises the ill-conditioned arccos of U3_merge (a12)
    c = C.Circuit(2)
    c.u3(0.3, 0.2, 0.1, 0); c.u3(-0.3 + 1e-7, -0.1, -0.2, 0); c.cx(0, 1); c.h(1); c.h(1)
    add("merge_near_identity", c, {"rotation_error": {"rz": [0.999, 0.001]}})

    # ---- measurement modes (a18-a23) ----
    base = _rand_circuit(4, 20, 41)
    for b in ("X", "Y", "Z"):
        c = C.Circuit(4); c.instructions = list(_rand_circuit(4, 20, 41).instructions)
        c.measure(2, 2, basis=b)
        add("measure_single_%s" % b, c, {"depolarization_factor": 0.9})
        c = C.Circuit(4); c.instructions = list(_rand_circuit(4, 20, 41).instructions)
        c.measure([0, 1, 2, 3], [0, 1, 2, 3], basis="Ensemble", add_param=b)
        add("measure_ensemble_%s" % b, c, {"depolarization_factor": 0.9})
        c = C.Circuit(4); c.instructions = list(_rand_circuit(4, 20, 41).instructions)
        c.measure(3, 3, basis=b); c.measure(1, 1, basis=b)
        c.u3(0.4, 0.5, 0.6, 1)
        add("measure_partial_%s" % b, c, dict(FULL_NOISE))
    c = C.Circuit(4); c.instructions = list(base.instructions)
    c.measure(0, 0)      # no basis given -> 'Z', params attribute absent
    add("measure_default_z", c)
    c = C.Circuit(4); c.instructions = list(_rand_circuit(4, 20, 41).instructions)
    c.measure(0, 0, basis="N", add_param=[0.0, 0.6, 0.8]); c.measure(2, 2, basis="N", add_param=[0.0, 0.6, 0.8])
    add("measure_partial_N", c, {"depolarization_factor": 0.95})
    # Ensemble readout along a direction: the only form the reference's front-end lets through and its
    # dispatcher accepts is an object ndarray ['N', direction] (circuit/instruction.py:142-143,
    # dm_simulator.py:1130-1132)
    c = C.Circuit(4); c.instructions = list(_rand_circuit(4, 20, 41).instructions)
    c.barrier()
    for q in range(4):
        prm = np.empty(2, dtype=object)
        prm[0], prm[1] = "N", np.array([1.0, -2.0, 0.5])
        c.instructions.append(C.instr("measure", [q], ["Ensemble", prm], memory=[q]))
    c.barrier()
    c.u3(0.3, 0.1, 0.2, 1)
    add("measure_ensemble_N", c, dict(FULL_NOISE))
    # mixed bases in one level: every second measure is skipped (dm_simulator.py:1100)
    c = C.Circuit(4); c.instructions = list(_rand_circuit(4, 20, 42).instructions)
    c.measure(0, 0, basis="X"); c.measure(1, 1, basis="Y"); c.measure(2, 2, basis="Z"); c.measure(3, 3, basis="X")
    add("measure_mixed_skip_quirk", c, {"depolarization_factor": 0.9})
    c = C.Circuit(4); c.instructions = list(_rand_circuit(4, 20, 43).instructions)
    c.measure(0, 0, basis="Expect", add_param="XIZY")
    c.u3(0.1, 0.2, 0.3, 0)
    add("measure_expect_xizy", c, dict(FULL_NOISE))
    for pair in ("01", "13", "20"):
        c = C.Circuit(4); c.instructions = list(_rand_circuit(4, 24, 44).instructions)
        c.measure(0, 0, basis="Bell", add_param=pair)
        c.u3(0.1, 0.2, 0.3, 2)
        add("measure_bell_%s" % pair, c, {"decoherence_factor": 0.95, "bell_depolarization_factor": 0.5})
    # two measure layers with gates between; second layer coalesces (a15)
    c = C.Circuit(3); c.instructions = list(_rand_circuit(3, 12, 45).instructions)
    c.measure(0, 0, basis="Z"); c.u3(0.3, 0.1, 0.2, 1); c.measure(1, 1, basis="Z"); c.measure(2, 2, basis="Z")
    add("measure_two_layers", c, {"decay_factor": 0.97, "thermal_factor": 0.1})

    # ---- reset and barriers (a24, a15-a17) ----
    c = C.Circuit(3); c.instructions = list(_rand_circuit(3, 15, 51).instructions)
    c.reset(1); c.u3(0.5, 0.4, 0.3, 1); c.reset(0); c.cx(1, 2)
    add("reset_coalesce", c, {"decay_factor": 0.98, "thermal_factor": 0.4})
    c = C.Circuit(2); c.u3(0.1, 0.2, 0.3, 0); c.barrier(); c.u3(0.4, 0.5, 0.6, 1); c.barrier()
    c.u3(0.7, 0.8, 0.9, 0); c.u3(1.0, 1.1, 1.2, 1)
    add("barrier_levels", c, {"decoherence_factor": 0.9})

    # ---- BASELINE.json configs at oracle-friendly sizes ----
    add("qft5", C.qft(5))
    add("qft8_binary", C.qft(8), {"custom_densitymatrix": "binary_string",
                                  "initial_densitymatrix": "01101001",
                                  "compute_densitymatrix": False})
    add("qft8", C.qft(8), {"compute_densitymatrix": False})
    g = C.grover(3, "101", 1)
    add("grover3_noisy", g, C.grover_options())
    g = C.grover(4, "1011", 2)
    add("grover4_noisy", g, dict(C.grover_options(), compute_densitymatrix=False))
    add("layered_n6_d10_noisy", C.random_layered(6, 10, 600), C.noisy_options())
    add("layered_n8_d6_noisy", C.random_layered(8, 6, 800),
        dict(C.noisy_options(), compute_densitymatrix=False))
    add("layered_n9_d4_memnoise", C.random_layered(9, 4, 900),
        dict(C.noisy_options(), compute_densitymatrix=False, **C.grover_options()))
    add("layered_n10_d3_noisy", C.random_layered(10, 3, 1000),
        dict(C.noisy_options(), compute_densitymatrix=False))
    # circuits lowered by the reference's REAL front-end (tests/golden/frontend_golden.json, written by
    # tests/golden/make_frontend_golden.py): golden results for these pin the whole reference stack,
    # QuantumCircuit -> transpile -> assemble -> DmSimulatorPy
    for name, opts in FRONTEND_CASES.items():
        n, instrs = frontend_instructions(name)
        circ = C.Circuit(n, "fe_" + name)
        circ.instructions = instrs
        add("fe_" + name, circ, opts)
    return cs


FRONTEND_CASES = {
    "readme_example": {},
    "teleport_style": FULL_NOISE,
    "phase_estimation_style": {"decoherence_factor": 0.95, "decay_factor": 0.97, "thermal_factor": 0.3},
    "qft_register": dict(FULL_NOISE, compute_densitymatrix=False),
    "every_gate": FULL_NOISE,
    "two_registers": {"rotation_error": {"rx": [1.0, 0.0], "ry": [0.99, 0.02], "rz": [0.98, -0.01]}},
    "broadcast_and_barriers": {"tsp_model_error": [0.97, 0.05]},
    "resets": FULL_NOISE,
    "measurement_modes": FULL_NOISE,
    "mid_circuit_measures": FULL_NOISE,
    "symbolic_parameters": {},
    "random_mixed": dict(FULL_NOISE, compute_densitymatrix=False),
}


def frontend_instructions(name):
    """(n_qubits, instruction namespaces) recorded from the reference front-end for a builder of
    tests/frontend_cases.py."""
    import json
    import os
    global _FE_GOLDEN
    if _FE_GOLDEN is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_golden.json")) as fh:
            _FE_GOLDEN = json.load(fh)
    gold = _FE_GOLDEN[name]
    out = []
    for rec in gold["instructions"]:
        params = None
        if "params" in rec:
            params = [np.array([float.fromhex(x) for x in p["array"]]) if "array" in p else
                      p["symbol"] if "symbol" in p else float.fromhex(p["float"]) for p in rec["params"]]
        out.append(C.instr(rec["name"], rec.get("qubits", []), params, rec.get("memory")))
    return gold["n_qubits"], out


_FE_GOLDEN = None
CASES = _cases()

#: cases whose full coefficient vector is too large to commit; the fixture keeps a strided
#: sample plus moments instead (see tests/golden/make_golden.py)
SAMPLE_STRIDE = {8: 7, 9: 29, 10: 113}


def get(name):
    import copy
    d = CASES[name]
    return dict(n=d["n"], instrs=copy.deepcopy(d["instrs"]), options=copy.deepcopy(d["options"]),
                files=d["files"])


def write_files(case, directory):
    """Materialise the .npy side files a case expects in the current working directory."""
    import os
    for fname, arr in case["files"].items():
        np.save(os.path.join(directory, fname), arr)
