"""Partial trace / reduced state readout (a23's reduced two-qubit matrix generalised
to any qubit subset) against a NumPy partial trace of the oracle's density matrix."""
import copy
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import dm_oracle  # noqa: E402
from qiskit_aakash_b200 import assemble, circuits as C  # noqa: E402

PAULI = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])]


def numpy_partial_trace(rho, n, keep):
    """Tr over the qubits not in ``keep`` (qubit 0 = most significant bit of row and column, the
    reference's convention ``dm_simulator.py:1230-1253``); ``keep[0]`` = most significant kept bit."""
    t = rho.reshape([2] * (2 * n))
    rest = [q for q in range(n) if q not in keep]
    t = np.transpose(t, list(keep) + rest + [n + q for q in keep] + [n + q for q in rest])
    k, r = len(keep), len(rest)
    t = t.reshape(2 ** k, 2 ** r, 2 ** k, 2 ** r)
    return np.einsum("arbr->ab", t)


def pauli_vector(rho, k):
    """Pauli coefficients r[P] = Tr(rho P) / 2^k, first qubit = most significant base-4 digit."""
    out = np.zeros(4 ** k)
    for j in range(4 ** k):
        P = np.array([[1.0]])
        for i in range(k):
            P = np.kron(P, PAULI[(j >> (2 * (k - 1 - i))) & 3])
        out[j] = np.real(np.trace(rho @ P)) / 2 ** k
    return out


def check_reduced(backend_factory, n, seed, keep, opts):
    circ = cases._rand_circuit(n, 30, seed)
    circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Z")
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), dict(copy.deepcopy(opts), compute_densitymatrix=True))
    want = numpy_partial_trace(np.asarray(ref["data"]["densitymatrix"]), n, keep)
    c2 = C.Circuit(n)
    c2.instructions = copy.deepcopy(circ.instructions)
    res = backend_factory().run(assemble(c2), backend_options=dict(copy.deepcopy(opts), reduced_state=list(keep),
                                                                    compute_densitymatrix=False)).result()
    data = res["results"][0]["data"]
    assert np.max(np.abs(data["reduced_densitymatrix"] - want)) <= 1e-10
    assert np.max(np.abs(data["reduced_coeffmatrix"] - pauli_vector(want, len(keep)))) <= 1e-10
    assert abs(data["reduced_coeffmatrix"][0] * 2 ** len(keep) - 1) <= 1e-12
    assert np.max(np.abs(data["coeffmatrix"] - ref["data"]["coeffmatrix"])) <= 1e-10      # state untouched
    return data


@pytest.mark.parametrize("n,seed,keep", [(5, 1, [3, 1]), (5, 2, [0]), (6, 3, [5, 0, 2]), (7, 4, [1, 2, 3, 4]),
                                         (4, 5, [0, 1, 2, 3]), (1, 6, [0]), (8, 7, [7, 6])])
def test_reduced_state_matches_numpy_partial_trace_emulated(n, seed, keep):
    from emu_backend import emu_backend
    check_reduced(emu_backend, n, seed, keep, cases.FULL_NOISE if seed % 2 else {})


def test_reduced_state_of_a_bell_pair_is_maximally_mixed():
    from emu_backend import emu_backend
    c = C.ghz(3)
    res = emu_backend().run(assemble(c), backend_options={"reduced_state": [1]}).result()
    assert np.allclose(res["results"][0]["data"]["reduced_densitymatrix"], np.eye(2) / 2, atol=1e-14)
    res = emu_backend().run(assemble(c), backend_options={}).result()            # the option is not sticky
    assert "reduced_densitymatrix" not in res["results"][0]["data"]


def test_reduced_state_rejects_bad_qubit_lists():
    from emu_backend import emu_backend
    from qiskit_aakash_b200.exceptions import BasicAerError
    for bad in ([0, 0], [5], []):
        with pytest.raises(BasicAerError):
            emu_backend().run(assemble(C.ghz(3)), backend_options={"reduced_state": bad}).result()


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,keep", [(6, 11, [4, 1]), (9, 12, [8, 0, 3]), (10, 13, [2, 3, 4, 5, 6])])
def test_reduced_state_matches_numpy_partial_trace_cuda(n, seed, keep):
    from qiskit_aakash_b200 import DmSimulatorB200
    check_reduced(DmSimulatorB200, n, seed, keep, cases.FULL_NOISE)
