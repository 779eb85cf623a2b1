"""Circuits written against the reference's user API (QuantumCircuit / QuantumRegister /
ClassicalRegister, ``qc.measure(q, c, basis=..., add_param=...)``).  Each builder takes an ``api``
namespace, so the same source drives the real reference front-end (``oracle/frontend_harness.py``,
when generating ``tests/golden/frontend_golden.json``) and ``qiskit_aakash_b200.frontend``.
"""
import numpy as np


def readme_example(api):
    qc = api.QuantumCircuit(2)
    qc.x(1)
    qc.cx(0, 1)
    return qc


def teleport_style(api):
    q = api.QuantumRegister(3, "q")
    c = api.ClassicalRegister(3, "c")
    qc = api.QuantumCircuit(q, c)
    qc.h(q[1])
    qc.cx(q[1], q[2])
    qc.h(q[0])
    qc.cx(q[0], q[1])
    qc.h(q[0])
    qc.measure(q[0], c[0])
    qc.measure(q[1], c[1])
    qc.z(q[2])
    qc.x(q[2])
    qc.measure(q[2], c[2])
    return qc


def phase_estimation_style(api):
    n = 5
    q = api.QuantumRegister(n, "q")
    c = api.ClassicalRegister(3, "c")
    qc = api.QuantumCircuit(q, c)
    for j in range(3):
        qc.h(q[j])
    qc.cx(q[2], q[3])
    qc.cx(q[2], q[4])
    qc.h(q[1])
    qc.cu1(2 * np.pi / 4, q[1], q[0])
    qc.h(q[0])
    qc.cu1(2 * np.pi / 8, q[1], q[2])
    qc.cu1(2 * np.pi / 4, q[0], q[2])
    qc.h(q[2])
    for j in range(3):
        qc.measure(q[j], c[j])
    return qc


def qft_register(api, n=6):
    q = api.QuantumRegister(n, "q")
    c = api.ClassicalRegister(n, "c")
    qc = api.QuantumCircuit(q, c)
    qc.x(q[0])
    qc.x(q[3])
    for w in range(n - 1):
        qc.h(q[w])
        for r in range(w + 1):
            qc.cu1(2 * np.pi / 2 ** (w + 2 - r), q[w + 1], q[r])
    qc.h(q[n - 1])
    qc.measure(q, c, basis="Ensemble", add_param="Z")
    return qc


def every_gate(api):
    q = api.QuantumRegister(4, "q")
    c = api.ClassicalRegister(4, "c")
    qc = api.QuantumCircuit(q, c)
    qc.h(q[0]); qc.x(q[1]); qc.y(q[2]); qc.z(q[3])
    qc.s(q[0]); qc.sdg(q[1]); qc.t(q[2]); qc.tdg(q[3])
    qc.rx(0.3, q[0]); qc.ry(0.4, q[1]); qc.rz(0.5, q[2])
    qc.u1(0.6, q[3]); qc.u2(0.7, 0.8, q[0]); qc.u3(0.9, 1.0, 1.1, q[1])
    qc.iden(q[2]); qc.i(q[3]); qc.u0(1, q[0])
    qc.cx(q[0], q[1]); qc.cy(q[1], q[2]); qc.cz(q[2], q[3]); qc.ch(q[3], q[0])
    qc.crz(1.2, q[0], q[2]); qc.cu1(1.3, q[1], q[3]); qc.cu3(1.4, 1.5, 1.6, q[2], q[0])
    qc.swap(q[0], q[3]); qc.ccx(q[0], q[1], q[2]); qc.cswap(q[3], q[1], q[0]); qc.rzz(1.7, q[2], q[3])
    qc.measure(q, c, basis="Ensemble", add_param="Y")
    return qc


def two_registers(api):
    a = api.QuantumRegister(2, "a")
    b = api.QuantumRegister(3, "b")
    m = api.ClassicalRegister(2, "m")
    k = api.ClassicalRegister(3, "k")
    qc = api.QuantumCircuit(b, a, k, m)              # b first: wire order differs from name order
    qc.h(a)
    qc.cx(a[0], b)                                    # one control, register of targets
    qc.cx(a, b[0:2])                                  # pairwise
    qc.t(b[2])
    qc.cz(b[1], a[1])
    qc.ry(0.25, [b[0], a[1]])
    qc.measure(a, m)
    qc.measure(b, k)
    return qc


def twelve_qubits_string_order(api):
    """str(qargs) keys order q[10], q[11] before q[2]."""
    n = 12
    qc = api.QuantumCircuit(n, n)
    rng = np.random.default_rng(12)
    for _ in range(60):
        r = rng.integers(4)
        a, b = (int(x) for x in rng.choice(n, 2, replace=False))
        if r == 0:
            qc.u3(float(rng.uniform(0, 6)), float(rng.uniform(0, 6)), float(rng.uniform(0, 6)), a)
        elif r == 1:
            qc.cx(a, b)
        elif r == 2:
            qc.h(a)
        else:
            qc.cu1(float(rng.uniform(0, 3)), a, b)
    qc.measure(range(n), range(n), basis="Ensemble", add_param="X")
    return qc


def broadcast_and_barriers(api):
    q = api.QuantumRegister(4, "q")
    c = api.ClassicalRegister(4, "c")
    qc = api.QuantumCircuit(q, c)
    qc.h(q)
    qc.barrier()
    qc.cx(q[0], [q[1], q[2]])
    qc.barrier(q[0], q[2])
    qc.x([0, 3])
    qc.barrier(q)
    qc.ccx([q[0], q[1]], [q[1], q[2]], [q[2], q[3]])
    qc.u1(0.1, slice(1, 3))
    qc.measure(q, c)
    return qc


def resets(api):
    q = api.QuantumRegister(3, "q")
    c = api.ClassicalRegister(3, "c")
    qc = api.QuantumCircuit(q, c)
    qc.reset(q[0])                # leading resets are dropped (RemoveResetInZeroState)
    qc.reset(q[0])
    qc.h(q[1])
    qc.reset(q[1])                # kept
    qc.reset(q[1])
    qc.cx(q[1], q[2])
    qc.reset(q)                   # q[0]'s is still leading
    qc.h(q[0])
    qc.measure(q, c, basis="Ensemble")
    return qc


def measurement_modes(api):
    q = api.QuantumRegister(3, "q")
    c = api.ClassicalRegister(3, "c")
    qc = api.QuantumCircuit(q, c)
    qc.h(q[0]); qc.cx(q[0], q[1]); qc.cx(q[1], q[2])
    qc.s(q[2])
    qc.measure(q[2], c[2], basis="X")
    qc.u3(0.3, 0.2, 0.1, q[1])
    qc.measure(q[1], c[1], basis="N", add_param=np.array([1.0, 2.0, 3.0]))
    qc.h(q[2])
    qc.measure(q[0], c[0], basis="Bell", add_param="01")
    qc.t(q[0])
    qc.measure(q[0], c[0], basis="Expect", add_param="ZIZ")
    qc.h(q[1])
    qc.measure(q, c, basis="Ensemble", add_param="X")
    return qc


def mid_circuit_measures(api):
    """Plain measures between gates: their position relative to gates on OTHER qubits is decided
    by the front-end's topological order (and changes the clock-cycle partition)."""
    q = api.QuantumRegister(4, "q")
    c = api.ClassicalRegister(2, "c")
    qc = api.QuantumCircuit(q, c)
    qc.h(q[3]); qc.h(q[1])
    qc.measure(q[3], c[0])
    qc.cx(q[1], q[0])
    qc.measure(q[1], c[0])        # same classical bit: ordered after the first measure
    qc.t(q[2]); qc.h(q[2]); qc.cx(q[2], q[3])
    qc.measure(q[0], c[1], basis="Y")
    qc.rx(0.4, q[0])
    qc.measure([q[2], q[3]], [c[0], c[1]], basis="Z")
    return qc


def symbolic_parameters(api):
    pi = api.pi
    q = api.QuantumRegister(3, "q")
    qc = api.QuantumCircuit(q)
    qc.u1(pi / 3, q[0])
    qc.rx(pi / 2, q[1])
    qc.cu1(pi / 8, q[0], q[1])
    qc.cu3(pi / 3, pi / 5, pi / 7, q[1], q[2])
    qc.crz(-pi / 9, q[2], q[0])
    qc.u3(2 * pi / 3, 1, 2, q[2])            # ints
    qc.u2(0.5, pi, q[0])
    qc.rzz(pi / 11, q[0], q[2])
    qc.cu1(1, q[2], q[1])
    return qc


def random_mixed(api, n=7, gates=120, seed=71):
    rng = np.random.default_rng(seed)
    q = api.QuantumRegister(n, "qr")
    c = api.ClassicalRegister(n, "cr")
    qc = api.QuantumCircuit(q, c)
    one = ["h", "x", "y", "z", "s", "sdg", "t", "tdg"]
    for _ in range(gates):
        r = rng.integers(10)
        a, b, d = (int(x) for x in rng.choice(n, 3, replace=False))
        if r < 3:
            getattr(qc, one[int(rng.integers(len(one)))])(q[a])
        elif r == 3:
            qc.u3(float(rng.uniform(0, 6)), float(rng.uniform(0, 6)), float(rng.uniform(0, 6)), q[a])
        elif r == 4:
            qc.cx(q[a], q[b])
        elif r == 5:
            qc.ccx(q[a], q[b], q[d])
        elif r == 6:
            qc.cu3(float(rng.uniform(0, 3)), float(rng.uniform(0, 3)), float(rng.uniform(0, 3)), q[a], q[b])
        elif r == 7:
            qc.swap(q[a], q[b])
        elif r == 8:
            qc.measure(q[a], c[b], basis="XYZ"[int(rng.integers(3))])
        else:
            qc.reset(q[a])
    qc.measure(q, c, basis="Ensemble", add_param="Z")
    return qc


def random_mixed_13(api):
    return random_mixed(api, n=13, gates=150, seed=1313)


CASES = {f.__name__: f for f in (readme_example, teleport_style, phase_estimation_style, qft_register, every_gate,
                                 two_registers, twelve_qubits_string_order, broadcast_and_barriers, resets,
                                 measurement_modes, mid_circuit_measures, symbolic_parameters, random_mixed,
                                 random_mixed_13)}
