"""Full-circuit parity cases at n = 12 (the size BASELINE.json quotes the 1e-10 bar at).  The fixture
``tests/golden/golden_n12.npz`` holds what the UNMODIFIED reference returns for them
(``tests/golden/make_golden_n12.py``); the oracle (CPU suite) and the CUDA backend (``-m gpu``) are both
compared with it through ``check``."""
from __future__ import annotations

import copy
import os

import numpy as np

from qiskit_aakash_b200 import circuits as C

STRIDE = 1021                    # prime: the sample walks through every digit pattern
TOL = 1e-10
TRACE_TOL = 1e-12


def _cases():
    cs = {}
    g = C.grover(7, "1011001", 1)                       # BASELINE configs[1]
    assert g.n_qubits == 12
    cs["grover12_noisy"] = dict(n=12, instrs=g.instructions,
                                options=dict(C.grover_options(), compute_densitymatrix=False))
    lay = C.random_layered(12, 6, 1200)                  # BASELINE configs[2] shape at n = 12
    cs["layered_n12_d6_noisy"] = dict(n=12, instrs=lay.instructions,
                                      options=dict(C.noisy_options(), compute_densitymatrix=False))
    return cs


CASES = _cases()


def get(name):
    d = CASES[name]
    return dict(n=d["n"], instrs=copy.deepcopy(d["instrs"]), options=copy.deepcopy(d["options"]))


def signed_checksum(vec):
    """sum_i s_i * vec_i with a fixed +-1 pattern (catches permuted coefficients that sum / sumsq miss)."""
    i = np.arange(vec.size, dtype=np.uint64)
    s = 1.0 - 2.0 * (((i * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(63)).astype(np.float64))
    return float(np.dot(s, vec))


def pack(name, result, out):
    out[name + "/levels"] = np.array(result["number_of_clock_cycles"])
    for key, val in result["data"].items():
        if key == "coeffmatrix":
            vec = np.asarray(val, dtype=float).reshape(-1)
            out[name + "/coeff"] = vec[::STRIDE].copy()
            out[name + "/coeff_sum"] = np.array(vec.sum())
            out[name + "/coeff_sumsq"] = np.array(np.dot(vec, vec))
            out[name + "/coeff_signed"] = np.array(signed_checksum(vec))
        elif isinstance(val, dict):
            out[name + "/" + key] = np.array(list(val.values()), dtype=float)
            out[name + "/" + key + "/keys"] = np.array(list(val.keys()))
        else:
            out[name + "/" + key] = np.asarray(val, dtype=float)


def load_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_n12.npz"))


def check(golden, name, result):
    """max |delta| of a result dict against the reference's recorded output; asserts the 1e-10 / 1e-12 bars."""
    n = CASES[name]["n"]
    assert int(golden[name + "/levels"]) == result["number_of_clock_cycles"], "level count"
    data = result["data"]
    vec = np.asarray(data["coeffmatrix"], dtype=float).reshape(-1)
    assert vec.size == 4 ** n
    worst = float(np.max(np.abs(vec[::STRIDE] - golden[name + "/coeff"])))
    assert abs(vec.sum() - float(golden[name + "/coeff_sum"])) <= 1e-9
    assert abs(np.dot(vec, vec) - float(golden[name + "/coeff_sumsq"])) <= 1e-9
    assert abs(signed_checksum(vec) - float(golden[name + "/coeff_signed"])) <= 1e-9
    assert abs(vec[0] * 2 ** n - 1.0) <= TRACE_TOL, "trace"
    seen = 0
    for key, val in data.items():
        if key == "coeffmatrix":
            continue
        g = golden[name + "/" + key]
        if isinstance(val, dict):
            assert list(val.keys()) == [str(x) for x in golden[name + "/" + key + "/keys"]], key
            arr = np.array(list(val.values()), dtype=float)
        else:
            arr = np.asarray(val, dtype=float)
        worst = max(worst, float(np.max(np.abs(arr - g))))
        seen += 1
    assert seen >= 1, "no probability readout in the result"
    assert worst <= TOL, "max |delta| = %g" % worst
    return worst
