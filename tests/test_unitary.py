"""``unitary`` instructions: rejected like the reference by default, unrolled
to u3 / cx (Euler angles, three-CNOT KAK) with ``reference_quirks=False`` -- checked against dense
matrix algebra (rho -> U rho U^dagger with the a25 convention: qubit 0 is the MSB of the index)."""
import numpy as np
import pytest
from scipy.stats import unitary_group

from qiskit_aakash_b200 import QuantumCircuit, execute, unitary as U
from qiskit_aakash_b200.exceptions import BasicAerError

SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
CX_Q0_CONTROL = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)   # little-endian


def _embed(mat, qubits, n):
    """Full 2^n x 2^n operator of ``mat`` on ``qubits`` (Qiskit order: qubits[0] = LSB of mat's index) in the
    simulator's convention (qubit 0 = most significant bit of the row / column index)."""
    k = len(qubits)
    t = np.asarray(mat, dtype=complex).reshape([2] * (2 * k))      # out bits (MSB..LSB), in bits (MSB..LSB)
    full = np.eye(2 ** n, dtype=complex).reshape([2] * (2 * n))
    # contract the in-indices of mat with the row axes of `full` that belong to the gate's qubits
    row_axes = [qubits[k - 1 - j] for j in range(k)]               # mat's MSB is qubits[-1]
    out = np.tensordot(t, full, axes=(list(range(k, 2 * k)), row_axes))
    # tensordot puts mat's out-indices first; move them back to the qubit positions
    rest = [a for a in range(n) if a not in row_axes]
    order = row_axes + rest
    perm = [order.index(a) for a in range(n)] + list(range(n, 2 * n))
    return np.transpose(out, perm).reshape(2 ** n, 2 ** n)


def _gate_list_matrix(gates, n):
    m = np.eye(2 ** n, dtype=complex)
    for g in gates:
        if g.name == "cx":
            m = _embed(CX_Q0_CONTROL, g.qubits, n) @ m
        else:
            m = _embed(U.u3_matrix(*g.params), g.qubits, n) @ m
    return m


def test_embed_helper_conventions():
    # cx(control 0, target 1) on |q0 q1> = |10> -> |11> (qubit 0 is the MSB)
    m = _embed(CX_Q0_CONTROL, [0, 1], 2)
    v = np.zeros(4); v[0b10] = 1
    assert np.argmax(np.abs(m @ v)) == 0b11
    assert np.allclose(_embed(SWAP, [0, 2], 3), _embed(SWAP, [2, 0], 3))


@pytest.mark.parametrize("seed", range(40))
def test_two_qubit_decomposition_rebuilds_the_matrix(seed):
    mat = unitary_group.rvs(4, random_state=seed)
    for qubits in ([0, 1], [1, 0], [2, 0]):
        gates = U.decompose_unitary(mat, qubits)
        assert sum(g.name == "cx" for g in gates) == 3
        n = 3
        assert U.equal_up_to_phase(_gate_list_matrix(gates, n), _embed(mat, qubits, n), 1e-8)


def test_special_two_qubit_matrices():
    rng = np.random.default_rng(5)
    specials = [np.eye(4), np.diag([1, 1, 1, -1]), SWAP, CX_Q0_CONTROL, np.diag(np.exp(1j * rng.uniform(0, 6, 4))),
                np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]]),          # iSWAP
                np.kron(unitary_group.rvs(2, random_state=1), unitary_group.rvs(2, random_state=2)),
                np.exp(0.3j) * np.eye(4)]
    for mat in specials:
        gates = U.decompose_unitary(mat, [0, 1])
        assert U.equal_up_to_phase(_gate_list_matrix(gates, 2), _embed(mat, [0, 1], 2), 1e-8)
    # products of one-qubit unitaries need no CNOT at all
    assert not any(g.name == "cx" for g in U.decompose_unitary(specials[-2], [0, 1]))
    assert U.decompose_unitary(np.eye(4), [0, 1]) == []


@pytest.mark.parametrize("seed", range(20))
def test_one_qubit_euler_angles(seed):
    mat = unitary_group.rvs(2, random_state=seed) if seed > 4 else \
        [np.eye(2), [[0, 1], [1, 0]], np.diag([1, 1j]), np.array([[1, 1], [1, -1]]) / np.sqrt(2), [[0, -1j], [1j, 0]]][seed]
    (g,) = U.decompose_unitary(np.asarray(mat, dtype=complex), [3])
    assert g.name == "u3" and g.qubits == [3]
    assert U.equal_up_to_phase(U.u3_matrix(*g.params), np.asarray(mat, dtype=complex))


def test_rejects_what_it_cannot_decompose():
    with pytest.raises(BasicAerError):
        U.decompose_unitary(np.ones((2, 2)), [0])
    with pytest.raises(BasicAerError):
        U.decompose_unitary(np.eye(8), [0, 1, 2])
    with pytest.raises(BasicAerError):
        U.decompose_unitary(np.eye(4), [1, 1])
    with pytest.raises(BasicAerError):
        U.decompose_unitary(np.eye(4), [1])


def _circuit(seed, n=4):
    rng = np.random.default_rng(seed)
    qc = QuantumCircuit(n, n)
    full = np.eye(2 ** n, dtype=complex)
    for k in range(8):
        if k % 2 == 0:
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            mat = unitary_group.rvs(4, random_state=seed * 100 + k)
            qc.unitary(mat, [qc.qubits[a], qc.qubits[b]])
            full = _embed(mat, [a, b], n) @ full
        else:
            a = int(rng.integers(n))
            mat = unitary_group.rvs(2, random_state=seed * 100 + k)
            qc.unitary(mat, [a])
            full = _embed(mat, [a], n) @ full
        qc.h(int(rng.integers(n)))
        full = _embed(np.array([[1, 1], [1, -1]]) / np.sqrt(2), [qc.data[-1][1][0].index], n) @ full
    return qc, full


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """The emulated kernels in the CPU suite, the real library on a B200 under ``-m gpu``."""
    if request.param == "cuda":
        import __graft_entry__ as g
        g.build()
        from qiskit_aakash_b200 import DmSimulatorB200
        return DmSimulatorB200
    from emu_backend import emu_backend
    return emu_backend


def test_default_rejects_unitary_like_the_reference(backend):
    qc, _ = _circuit(1)
    with pytest.raises(BasicAerError, match="unrecognized instruction: unitary"):
        execute(qc, backend()).result()


@pytest.mark.parametrize("seed,n", [(0, 4), (1, 4), (2, 4), (3, 4), (4, 7), (5, 8)])
def test_unitary_gates_without_quirks_match_dense_evolution(seed, n, backend):
    qc, full = _circuit(seed, n)
    res = execute(qc, backend(), reference_quirks=False, compute_densitymatrix=True).result()["results"][0]
    rho0 = np.zeros((2 ** n, 2 ** n), dtype=complex)
    rho0[0, 0] = 1
    rho = full @ rho0 @ full.conj().T
    assert np.max(np.abs(res["data"]["densitymatrix"] - rho)) <= 1e-10
    assert abs(res["data"]["coeffmatrix"][0] * 2 ** n - 1) <= 1e-12


def test_unitary_goes_through_the_noise_model(backend):
    """Unrolled gates are ordinary u3 / cx: the TSP error is charged three times per two-qubit unitary."""
    qc = QuantumCircuit(2, 2)
    qc.unitary(unitary_group.rvs(4, random_state=9), [0, 1])
    be = backend()
    clean = execute(qc, be, reference_quirks=False).result()["results"][0]["data"]["coeffmatrix"]
    noisy = execute(qc, be, reference_quirks=False, tsp_model_error=[0.9, 0.0]).result()["results"][0]["data"]["coeffmatrix"]
    assert np.dot(noisy, noisy) < np.dot(clean, clean) - 1e-3      # purity drops
