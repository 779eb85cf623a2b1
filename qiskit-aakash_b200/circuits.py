"""Circuit generators for the BASELINE.json configs, emitted as qobj-style instruction lists.

The reference's front end (``QuantumCircuit`` -> transpile -> assemble) is out of scope
(SURVEY.md section 8b); what the dm_simulator backend actually receives is a list of
duck-typed instruction objects (``name``, ``qubits``, ``params``, ``memory``) in the basis
``u1/u2/u3/cx`` plus ``measure/reset/barrier``.  ``Circuit`` below lowers the standard gates
with the reference's own decomposition rules so the same list can be fed to this
package's backend and to the oracle:

    h   -> u2(0, pi)                       qiskit/extensions/standard/h.py:37-48
    x   -> u3(pi, 0, pi)                   .../x.py:35
    y   -> u3(pi, pi/2, pi/2)              .../y.py
    z/s/sdg/t/tdg -> u1(pi | +-pi/2 | +-pi/4)
    cu1 -> u1 a; cx a,b; u1 b; cx a,b; u1 b   .../cu1.py:33-52
    ccx -> 15-gate Toffoli                 .../ccx.py:38-68
    measure(q, c, basis, add_param): 'Ensemble'/'Expect'/'Bell' are wrapped in barriers
                                           qiskit/circuit/measure.py:29-59,92-95
"""
from __future__ import annotations

import math
from types import SimpleNamespace as NS

import numpy as np

PI = math.pi


def instr(name, qubits, params=None, memory=None):
    ns = NS(name=name, qubits=list(qubits))
    if params is not None:
        ns.params = list(params)
    if memory is not None:
        ns.memory = list(memory)
    return ns


class Circuit:
    """Minimal gate-list builder (not a QuantumCircuit replacement)."""

    def __init__(self, n_qubits, name="circuit"):
        self.n_qubits = int(n_qubits)
        self.name = name
        self.instructions = []

    # -- basis gates ---------------------------------------------------------------
    def u1(self, lam, q):
        self.instructions.append(instr("u1", [q], [float(lam)]))
        return self

    def u2(self, phi, lam, q):
        self.instructions.append(instr("u2", [q], [float(phi), float(lam)]))
        return self

    def u3(self, theta, phi, lam, q):
        self.instructions.append(instr("u3", [q], [float(theta), float(phi), float(lam)]))
        return self

    def cx(self, c, t):
        self.instructions.append(instr("cx", [c, t]))
        return self

    def iden(self, q):
        self.instructions.append(instr("id", [q]))
        return self

    def barrier(self):
        self.instructions.append(instr("barrier", list(range(self.n_qubits))))
        return self

    def reset(self, q):
        self.instructions.append(instr("reset", [q]))
        return self

    # -- lowered standard gates ----------------------------------------------------
    def h(self, q):
        return self.u2(0.0, PI, q)

    def x(self, q):
        return self.u3(PI, 0.0, PI, q)

    def y(self, q):
        return self.u3(PI, PI / 2, PI / 2, q)

    def z(self, q):
        return self.u1(PI, q)

    def s(self, q):
        return self.u1(PI / 2, q)

    def sdg(self, q):
        return self.u1(-PI / 2, q)

    def t(self, q):
        return self.u1(PI / 4, q)

    def tdg(self, q):
        return self.u1(-PI / 4, q)

    def rx(self, theta, q):
        return self.u3(theta, -PI / 2, PI / 2, q)

    def ry(self, theta, q):
        return self.u3(theta, 0.0, 0.0, q)

    def rz(self, phi, q):
        return self.u1(phi, q)

    def cz(self, a, b):
        return self.h(b).cx(a, b).h(b)

    def swap(self, a, b):
        return self.cx(a, b).cx(b, a).cx(a, b)

    def cu1(self, lam, a, b):
        return self.u1(lam / 2, a).cx(a, b).u1(-lam / 2, b).cx(a, b).u1(lam / 2, b)

    def ccx(self, a, b, c):
        self.h(c).cx(b, c).tdg(c).cx(a, c).t(c).cx(b, c).tdg(c).cx(a, c)
        self.t(b).t(c).h(c).cx(a, b).t(a).tdg(b).cx(a, b)
        return self

    # -- measurement (circuit/measure.py) ---------------------------------------------
    def measure(self, q, c, basis=None, add_param=None):
        qs = list(q) if isinstance(q, (list, tuple, range)) else [q]
        cs = list(c) if isinstance(c, (list, tuple, range)) else [c]
        special = basis in ("Ensemble", "Expect", "Bell")
        if basis == "Ensemble" and add_param is None:
            add_param = "Z"
        if basis in ("N", "Bell", "Expect") and add_param is None:
            raise ValueError("basis %r needs add_param" % basis)
        if special:
            self.barrier()
        for qq, cc in zip(qs, cs):
            if basis is None:
                prm = []
            elif add_param is None:
                prm = [basis[0]]
            elif isinstance(add_param, str):
                prm = [basis, add_param]
            else:
                prm = [basis, np.array(add_param, dtype=float)]
            self.instructions.append(instr("measure", [qq], prm if prm else None, memory=[cc]))
        if special:
            self.barrier()
        return self

    def measure_all_ensemble(self, add_param="Z"):
        n = self.n_qubits
        return self.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param=add_param)

    def count_ops(self):
        out = {}
        for i in self.instructions:
            out[i.name] = out.get(i.name, 0) + 1
        return out


# --------------------------------------------------------------------------------------
# BASELINE.json configs
# --------------------------------------------------------------------------------------

def qft(n, readout=True):
    """QFT as in ``circuits/qft.py:36-40`` (the loop once): 148 basis gates at n=8."""
    c = Circuit(n, "qft%d" % n)
    for wire in range(n - 1):
        c.h(wire)
        for rot in range(wire + 1):
            c.cu1(2 * PI / 2 ** (wire + 2 - rot), wire + 1, rot)
    c.h(n - 1)
    if readout:
        c.measure_all_ensemble("Z")
    return c


def random_layered(n, depth, seed, readout=True):
    """Config 3/5: layer l = u3 on every qubit (angles ~U[0,2pi), drawn in qubit order),
    then CX on pairs (q, q+1) for q = l mod 2 (mod 2), orientation from rng.integers(2)."""
    rng = np.random.default_rng(seed)
    c = Circuit(n, "rand_u3cx_n%d_d%d" % (n, depth))
    for layer in range(depth):
        for q in range(n):
            th, ph, lam = rng.uniform(0.0, 2 * PI, 3)
            c.u3(th, ph, lam, q)
        for q in range(layer % 2, n - 1, 2):
            if rng.integers(2):
                c.cx(q + 1, q)
            else:
                c.cx(q, q + 1)
    if readout:
        c.measure_all_ensemble("Z")
    return c


def noisy_options(r=0.999, depol=0.99):
    """Per-gate 'depolarizing' knobs for config 3/5."""
    return {"rotation_error": {"rx": [r, 0.0], "ry": [r, 0.0], "rz": [r, 0.0]},
            "tsp_model_error": [r, 0.0], "depolarization_factor": depol}


def grover_options():
    """``features/noise.ipynb`` cell 5 option set (config 2)."""
    return {"thermal_factor": 0.0, "decoherence_factor": 0.9, "decay_factor": 0.99,
            "depolarization_factor": 0.99}


def _mcx_vchain(c, controls, target, anc):
    """Multi-controlled X with a V-chain of Toffolis (len(anc) >= len(controls)-2)."""
    k = len(controls)
    if k == 1:
        c.cx(controls[0], target)
        return
    if k == 2:
        c.ccx(controls[0], controls[1], target)
        return
    c.ccx(controls[0], controls[1], anc[0])
    for i in range(2, k - 1):
        c.ccx(controls[i], anc[i - 2], anc[i - 1])
    c.ccx(controls[k - 1], anc[k - 3], target)
    for i in range(k - 2, 1, -1):
        c.ccx(controls[i], anc[i - 2], anc[i - 1])
    c.ccx(controls[0], controls[1], anc[0])


def grover(n_search, marked=None, iterations=1, n_total=None, readout=True):
    """Grover search on ``n_search`` qubits with ``n_search-2`` V-chain ancillas
    (config 2 uses 7 + 5 = 12 qubits).  ``marked`` is a bitstring over the search
    register (char i = qubit i); readout is a partial Z measurement of that register."""
    n_anc = max(n_search - 2, 0)
    n = n_total or (n_search + n_anc)
    if marked is None:
        marked = "1" * n_search
    c = Circuit(n, "grover%d" % n)
    srch = list(range(n_search))
    anc = list(range(n_search, n_search + n_anc))
    for q in srch:
        c.h(q)
    for _ in range(iterations):
        # oracle: phase-flip |marked>
        for q in srch:
            if marked[q] == "0":
                c.x(q)
        c.h(srch[-1])
        _mcx_vchain(c, srch[:-1], srch[-1], anc)
        c.h(srch[-1])
        for q in srch:
            if marked[q] == "0":
                c.x(q)
        # diffusion
        for q in srch:
            c.h(q)
            c.x(q)
        c.h(srch[-1])
        _mcx_vchain(c, srch[:-1], srch[-1], anc)
        c.h(srch[-1])
        for q in srch:
            c.x(q)
            c.h(q)
    if readout:
        c.measure(srch, srch, basis="Z")
    return c


def ghz(n):
    c = Circuit(n, "ghz%d" % n)
    c.h(0)
    for q in range(n - 1):
        c.cx(q, q + 1)
    return c
