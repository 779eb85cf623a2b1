"""``unitary`` instructions -> the simulator's own basis (u3 / cx).

The reference advertises ``unitary`` in ``basis_gates`` (``dm_simulator.py:85``), so its
transpiler hands ``UnitaryGate`` instructions to the backend untouched -- where
``single_gate_merge`` rejects them (``basicaertools.py:286-287``).  With
``reference_quirks=True`` (default) this package raises the same error.  With
``reference_quirks=False`` the backend does what the reference's own
``UnitaryGate._define`` (``qiskit/extensions/unitary.py:99-109``) would have done had the gate
been unrolled: a one-qubit matrix becomes one ``u3`` (Euler angles), a two-qubit matrix becomes
three ``cx`` plus ``u3`` gates (KAK decomposition), and those then go through the ordinary merge /
partition / noise model like any other gate.

Everything here is host-side float64/complex128 NumPy on 2x2 and 4x4 matrices; the decomposition
is verified at the end (``decompose_unitary`` rebuilds the matrix from the emitted gates and
refuses to return a list that is off by more than 1e-9 up to a global phase).

Conventions (Qiskit's): a matrix given for qubits ``[q0, q1]`` acts on ``|q1 q0>``, i.e. ``q0`` is
the least significant bit of the matrix index.
"""
from __future__ import annotations

import cmath
import math
from types import SimpleNamespace as NS

import numpy as np

from .exceptions import BasicAerError

_I2 = np.eye(2, dtype=complex)
# magic basis (columns): Bell-like states in which SU(2) x SU(2) is SO(4) and XX, YY, ZZ are diagonal
_MAGIC = np.array([[1, 0, 0, 1j], [0, 1j, 1, 0], [0, 1j, -1, 0], [1, 0, 0, -1j]], dtype=complex) / math.sqrt(2)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.diag([1, -1]).astype(complex)


def u3_matrix(theta, phi, lam):
    """``U3Gate.to_matrix`` (``extensions/standard/u3.py``)."""
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -cmath.exp(1j * lam) * s],
                     [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]], dtype=complex)


def _cx_matrix(control_is_first):
    """CNOT on kron(first, second) with index = 2 * first + second."""
    m = np.zeros((4, 4), dtype=complex)
    for f in range(2):
        for s in range(2):
            g, t = (f, s ^ f) if control_is_first else (f ^ s, s)
            m[2 * g + t, 2 * f + s] = 1
    return m


def is_unitary(mat, tol=1e-9):
    mat = np.asarray(mat)
    return mat.ndim == 2 and mat.shape[0] == mat.shape[1] and \
        np.allclose(mat.conj().T @ mat, np.eye(mat.shape[0]), atol=tol)


def equal_up_to_phase(a, b, tol=1e-9):
    k = int(np.argmax(np.abs(b)))
    if abs(b.flat[k]) < 1e-12:
        return False
    ph = a.flat[k] / b.flat[k]
    return abs(abs(ph) - 1) <= tol and np.allclose(a, ph * b, atol=tol)


def euler_u3(mat):
    """(theta, phi, lam) with ``mat == e^{i alpha} u3(theta, phi, lam)``."""
    u = np.asarray(mat, dtype=complex)
    if u.shape != (2, 2) or not is_unitary(u):
        raise BasicAerError("unitary: not a 2x2 unitary matrix")
    u = u / cmath.sqrt(np.linalg.det(u))                         # SU(2)
    c, s = abs(u[0, 0]), abs(u[1, 0])
    theta = 2 * math.atan2(s, c)
    # SU(2): u11 = e^{i(phi+lam)/2} cos, u10 = e^{i(phi-lam)/2} sin
    plus = 2 * cmath.phase(u[1, 1]) if c > 1e-12 else 0.0
    minus = 2 * cmath.phase(u[1, 0]) if s > 1e-12 else 0.0
    return theta, (plus + minus) / 2, (plus - minus) / 2


def _split_tensor(k):
    """kron(a, b) == k (up to the numerical noise of k) for k in SU(2) x SU(2): nearest Kronecker
    product from the dominant singular pair of the reshuffled matrix."""
    r = k.reshape(2, 2, 2, 2).transpose(0, 2, 1, 3).reshape(4, 4)      # [(a row, a col), (b row, b col)]
    u, s, vh = np.linalg.svd(r)
    if s[1] > 1e-7:
        raise BasicAerError("unitary: internal error, local factor is not a tensor product")
    a = (u[:, 0] * math.sqrt(s[0])).reshape(2, 2)
    b = (vh[0, :] * math.sqrt(s[0])).reshape(2, 2)
    return a, b


def _as_tensor_product(u4):
    """(a, b) with kron(a, b) == u4 when u4 is a product of one-qubit unitaries, else None."""
    r = np.asarray(u4, dtype=complex).reshape(2, 2, 2, 2).transpose(0, 2, 1, 3).reshape(4, 4)
    u, s, vh = np.linalg.svd(r)
    if s[1] > 1e-10:
        return None
    return (u[:, 0] * math.sqrt(s[0])).reshape(2, 2), (vh[0, :] * math.sqrt(s[0])).reshape(2, 2)


def _kak(u4):
    """u4 (4x4 unitary, kron(first, second) ordering) = phase * kron(a1, b1) . N(x, y, z) . kron(a0, b0)
    with N = exp(i (x XX + y YY + z ZZ)).  Returns (a1, b1, (x, y, z), a0, b0)."""
    u = np.asarray(u4, dtype=complex)
    u = u / np.linalg.det(u) ** 0.25                              # SU(4) (any fourth root: a global phase)
    up = _MAGIC.conj().T @ u @ _MAGIC
    m2 = up.T @ up                                                # symmetric unitary = O0^T D^2 O0
    # its real and imaginary parts are commuting real symmetric matrices: the eigenvectors of a generic
    # combination diagonalise both
    p = None
    for t in (0.6180339887, 0.3141592653, 0.7071067811, 0.1234567891, 0.8765432109, 0.4567891234):
        _, cand = np.linalg.eigh(t * m2.real + (1 - t) * m2.imag)
        d = cand.T @ m2 @ cand
        if np.allclose(d, np.diag(np.diag(d)), atol=1e-10):
            p = cand
            break
    if p is None:
        raise BasicAerError("unitary: KAK diagonalisation failed")
    if np.linalg.det(p) < 0:
        p[:, 0] = -p[:, 0]
    d2 = np.diag(p.T @ m2 @ p)
    theta = np.angle(d2) / 2                                      # D = exp(i theta), any branch
    dmat = np.exp(1j * theta)
    if (np.prod(dmat)).real < 0:                                  # det D = -1: take the other root once
        theta[0] += math.pi
        dmat = np.exp(1j * theta)
    o1 = up @ p @ np.diag(1 / dmat)                               # real, det +1
    if not np.allclose(o1.imag, 0, atol=1e-8):
        raise BasicAerError("unitary: KAK factor is not real")
    k1 = _MAGIC @ o1.real @ _MAGIC.conj().T
    k0 = _MAGIC @ p.T @ _MAGIC.conj().T
    # D = exp(i (x XX + y YY + z ZZ)) in the magic basis: read the three angles off the phases
    ex = np.real(np.diag(_MAGIC.conj().T @ np.kron(_X, _X) @ _MAGIC))
    ey = np.real(np.diag(_MAGIC.conj().T @ np.kron(_Y, _Y) @ _MAGIC))
    ez = np.real(np.diag(_MAGIC.conj().T @ np.kron(_Z, _Z) @ _MAGIC))
    theta = theta - np.round(theta.sum() / (2 * math.pi)) * np.array([2 * math.pi, 0, 0, 0])   # sum = 0 exactly
    x, y, z = (float(ex @ theta) / 4, float(ey @ theta) / 4, float(ez @ theta) / 4)
    a1, b1 = _split_tensor(k1)
    a0, b0 = _split_tensor(k0)
    return a1, b1, (x, y, z), a0, b0


def _rz(t):
    return np.diag([cmath.exp(-0.5j * t), cmath.exp(0.5j * t)])


def _ry(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=complex)


def two_qubit_gates(u4):
    """Gate list ``[(name, local qubits, matrix-or-None)]`` equal to ``u4`` up to a global phase; local
    qubit 1 is the FIRST tensor factor (most significant index bit), local qubit 0 the second.

    The canonical part uses three CNOTs (control/target alternate), found by matching the 3-CNOT
    template of Vatan & Williams (quant-ph/0308006, fig. 6) against exp(i(x XX + y YY + z ZZ)):

        N = Rz_1(pi/2) . CX(0->1) . Ry_0(-2y - pi/2) . CX(1->0) . [Rz_1(-2z - pi/2) Ry_0(2x + pi/2)] . CX(0->1) . Rz_0(-pi/2)
    """
    a1, b1, (x, y, z), a0, b0 = _kak(u4)
    h = math.pi / 2
    seq = [("1q", (1,), a0), ("1q", (0,), b0),
           ("1q", (0,), _rz(-h)),
           ("cx", (0, 1), None),
           ("1q", (1,), _rz(-2 * z - h)), ("1q", (0,), _ry(2 * x + h)),
           ("cx", (1, 0), None),
           ("1q", (0,), _ry(-2 * y - h)),
           ("cx", (0, 1), None),
           ("1q", (1,), _rz(h)),
           ("1q", (1,), a1), ("1q", (0,), b1)]
    # fold runs of one-qubit matrices between the CNOTs into one matrix per qubit
    out, pend = [], {0: None, 1: None}

    def flush():
        for q in (0, 1):
            if pend[q] is not None:
                out.append(("1q", (q,), pend[q]))
                pend[q] = None

    for name, qs, mat in seq:
        if name == "1q":
            q = qs[0]
            pend[q] = mat if pend[q] is None else mat @ pend[q]
        else:
            flush()
            out.append((name, qs, None))
    flush()
    return out


def _matrix_of(gates):
    """4x4 matrix of a local gate list (verification)."""
    m = np.eye(4, dtype=complex)
    for name, qs, mat in gates:
        if name == "cx":
            g = _cx_matrix(control_is_first=(qs[0] == 1))
        else:
            g = np.kron(mat, _I2) if qs[0] == 1 else np.kron(_I2, mat)
        m = g @ m
    return m


def decompose_unitary(matrix, qubits):
    """``unitary`` on ``qubits`` -> list of ``u3`` / ``cx`` instruction namespaces."""
    mat = np.asarray(matrix, dtype=complex)
    qubits = [int(q) for q in qubits]
    dim = 2 ** len(qubits)
    if mat.shape != (dim, dim):
        raise BasicAerError("unitary: a %dx%d matrix does not act on %d qubit(s)" % (mat.shape + (len(qubits),)))
    if not is_unitary(mat):
        raise BasicAerError("unitary: input matrix is not unitary")
    if len(qubits) == 1:
        th, ph, la = euler_u3(mat)
        if not equal_up_to_phase(u3_matrix(th, ph, la), mat):
            raise BasicAerError("unitary: Euler decomposition failed verification")
        return [NS(name="u3", qubits=[qubits[0]], params=[th, ph, la])]
    if len(qubits) != 2:
        raise BasicAerError("unitary: only one- and two-qubit matrices can be decomposed "
                            "(like UnitaryGate._define, extensions/unitary.py:99-109)")
    if len(set(qubits)) != 2:
        raise BasicAerError("unitary: duplicate qubit arguments")
    local = _as_tensor_product(mat)
    gates = [("1q", (1,), local[0]), ("1q", (0,), local[1])] if local else two_qubit_gates(mat)
    if not equal_up_to_phase(_matrix_of(gates), mat):
        raise BasicAerError("unitary: KAK decomposition failed verification")
    out = []
    for name, qs, m in gates:
        if name == "cx":
            out.append(NS(name="cx", qubits=[qubits[qs[0]], qubits[qs[1]]]))
        else:
            if equal_up_to_phase(m, _I2):
                continue
            th, ph, la = euler_u3(m)
            out.append(NS(name="u3", qubits=[qubits[qs[0]]], params=[th, ph, la]))
    return out


def expand_unitaries(instructions):
    """Replace every ``unitary`` instruction by its u3 / cx decomposition (order kept)."""
    if not any(getattr(i, "name", None) == "unitary" for i in instructions):
        return instructions
    out = []
    for ins in instructions:
        if ins.name != "unitary":
            out.append(ins)
            continue
        params = getattr(ins, "params", None) or []
        if len(params) != 1:
            raise BasicAerError("unitary: expected one matrix parameter")
        out.extend(decompose_unitary(params[0], ins.qubits))
    return out
