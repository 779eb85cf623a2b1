"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the reference dm_simulator path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; it is the *checker*, never the
product.  The product path (``qiskit-aakash_b200``) never imports it and fails loudly
when its CUDA library is missing.

Parity status: **pinned**.  ``tests/test_oracle_pinned.py`` checks this restatement
against (i) every recorded output the reference holds for the path (README example and
the user-guide notebooks), (ii) fixtures under ``tests/golden/``
produced by running the reference's *own* unmodified ``dm_simulator.py`` /
``basicaertools.py`` through ``oracle/ref_harness.py`` (generator script:
``tests/golden/make_golden.py``) and (iii), when ``/root/reference`` is present, the live
reference on freshly generated random circuits.

Each function cites the reference lines it follows (paths relative to
``/root/reference/qiskit/providers/basicaer/``).  The arithmetic expressions are kept in
the reference's order so results agree to the last bit wherever NumPy evaluates them the
same way; what is *not* reproduced is the reference's habit of copying the whole state
before every update (``basicaertools.py:119,344``; ``dm_simulator.py:415``) -- only the
planes that are read after being overwritten are copied, so this port is a somewhat
faster CPU baseline than the reference itself (stated wherever it is timed).

State layout (``dm_simulator.py:61-68,376-377``): real float64 vector of ``4**n`` Pauli
coefficients, C-order tensor of shape ``n*[4]``, qubit ``q`` = axis ``q`` (qubit 0 is
the most significant base-4 digit), digit values 0..3 = I, X, Y, Z.
"""
from __future__ import annotations

import copy
import itertools
import time
from types import SimpleNamespace as NS

import numpy as np


class OracleError(Exception):
    """Stands in for BasicAerError / QiskitError (``exceptions.py:22-32``)."""


# --------------------------------------------------------------------------------------
# instructions
# --------------------------------------------------------------------------------------

def as_instruction(ins):
    """Accept dicts or attribute objects; return a fresh SimpleNamespace (value equality,
    like the reference's ``BaseModel(SimpleNamespace)``, ``validation/base.py:306``)."""
    d = dict(ins) if isinstance(ins, dict) else dict(vars(ins))
    ns = NS(name=d["name"], qubits=list(d.get("qubits", [])))
    if d.get("params") is not None:
        ns.params = [copy.deepcopy(p) for p in d["params"]]
    for key in ("memory", "register"):
        if key in d:
            setattr(ns, key, list(d[key]))
    if d.get("conditional") is not None:
        ns.conditional = copy.deepcopy(d["conditional"])
    return ns


def conditional_skips(op):
    """``dm_simulator.py:1020-1034``: an instruction with an int ``conditional`` runs only when that bit of the
    classical register is set, one with a mask / val ``conditional`` only when the masked classical memory equals
    val.  Neither is ever written on this path (measurements do not record outcomes, and the ``bfunc`` that would
    set a register bit cannot pass the partitioner, ``basicaertools.py:549``), so both stay 0: an int conditional
    always skips, a mask conditional skips unless val == 0."""
    cond = getattr(op, "conditional", None)
    if isinstance(cond, int):
        return True
    if cond is not None:
        if int(cond.mask, 16) > 0 and int(cond.val, 16) != 0:
            return True
    return False


# --------------------------------------------------------------------------------------
# a12-a14: single-qubit gate merging   (basicaertools.py:129-307)
# --------------------------------------------------------------------------------------

def u3_merge(xi, theta1, theta2):
    """Ry(theta1) Rz(xi) Ry(theta2) -> (beta, alpha, gamma) of Rz(alpha) Ry(beta) Rz(gamma).
    ``basicaertools.py:162-178`` -- same formulae, same (ill-conditioned) arccos."""
    sxi, cxi = np.sin(xi * 0.5), np.cos(xi * 0.5)
    sp, cp = np.sin((theta1 + theta2) * 0.5), np.cos((theta1 + theta2) * 0.5)
    sm, cm = np.sin((theta1 - theta2) * 0.5), np.cos((theta1 - theta2) * 0.5)
    apg2 = np.arctan2(sxi * cm, cxi * cp)
    amg2 = np.arctan2(-sxi * sm, cxi * sp)
    cb2 = np.sqrt((cxi * cp) ** 2 + (sxi * cm) ** 2)
    return 2 * np.arccos(cb2), apg2 + amg2, apg2 - amg2


def merge_pair(g1, g2):
    """``mergeU`` (``basicaertools.py:181-228``): g1 acts first.  Entries are [inst, index]."""
    keep = copy.deepcopy(g1 if g1[1] < g2[1] else g2)
    a, b = g1[0], g2[0]
    if a.name == "u1" and b.name == "u1":
        keep[0].params[0] = a.params[0] + b.params[0]
    elif a.name == "u1" or b.name == "u1":
        if keep[0].name == "u1":
            keep[0].name = "u3"
            keep[0].params.extend([0, 0])
        if a.name == "u1" and b.name == "u3":
            keep[0].params[0] = b.params[0]
            keep[0].params[1] = b.params[1]
            keep[0].params[2] = b.params[2] + a.params[0]
        elif a.name == "u3" and b.name == "u1":
            keep[0].params[0] = a.params[0]
            keep[0].params[1] = a.params[1] + b.params[0]
            keep[0].params[2] = a.params[2]
    elif a.name == "u3" and b.name == "u3":
        beta, alpha, gamma = u3_merge(float(b.params[2] + a.params[1]),
                                      float(b.params[0]), float(a.params[0]))
        keep[0].params[0] = beta
        keep[0].params[1] = b.params[1] + alpha
        keep[0].params[2] = a.params[2] + gamma
    else:
        raise OracleError("unrecognized instructions in merge: %s, %s" % (a.name, b.name))
    return keep


def merge_run(run):
    """``merge_gates`` (``basicaertools.py:231-248``): left fold of merge_pair."""
    acc = run[0]
    for nxt in run[1:]:
        acc = merge_pair(acc, nxt)
    return acc[0]


def single_gate_merge(instrs, n, merge=True):
    """``single_gate_merge`` (``basicaertools.py:251-307``).  Any cx / measure / bfunc /
    reset / barrier flushes the pending runs of *all* qubits, in qubit order."""
    out = []
    if not merge:
        for op in instrs:
            if op.name == "CX":
                op.name = "cx"
            elif op.name == "U":
                op.name = "u3"
            elif op.name == "u2":
                op.name = "u3"
                op.params.insert(0, np.pi / 2)
            if op.name not in ("id", "u0"):
                out.append(op)
        return out
    pending = [[] for _ in range(n)]
    for pos, op in enumerate(instrs):
        if op.name in ("CX", "cx", "measure", "bfunc", "reset", "barrier"):
            for q in range(n):
                if pending[q]:
                    out.append(merge_run(pending[q]))
                    pending[q] = []
            if op.name == "CX":
                op.name = "cx"
            out.append(op)
        elif op.name in ("U", "u1", "u2", "u3"):
            if op.name == "U":
                op.name = "u3"
            elif op.name == "u2":
                op.name = "u3"
                op.params.insert(0, np.pi / 2)
            pending[op.qubits[0]].append([op, pos])
        elif op.name in ("id", "u0"):
            continue
        else:
            raise OracleError("Encountered unrecognized instruction: %s" % op)
    for q in range(n):
        if pending[q]:
            out.append(merge_run(pending[q]))
    return out


# --------------------------------------------------------------------------------------
# a15-a17: level partition   (basicaertools.py:536-748)
# --------------------------------------------------------------------------------------

def _qubit_stacks(instrs, n):
    """``qubit_stack`` (``basicaertools.py:536-591``)."""
    stacks = [[] for _ in range(n)]
    for ins in instrs:
        if ins.name not in ("measure", "reset"):
            for q in ins.qubits:
                stacks[q].append(ins)
            continue
        dummy_name = "dummy_" + ins.name
        own = stacks[ins.qubits[0]]
        if own and own[-1].name == dummy_name:
            own[-1] = ins
            continue
        own.append(ins)
        dummy = copy.deepcopy(ins)
        dummy.name = dummy_name
        dummy.qubits[0] = -1
        for q in range(n):
            if q not in ins.qubits:
                stacks[q].append(dummy)
    return stacks, max(len(s) for s in stacks)


def _partition_segment(instrs, n):
    """``partition_helper`` (``basicaertools.py:594-712``).  ``instrs`` is consumed."""
    stacks, depth = _qubit_stacks(instrs, n)
    level, seq = 0, [[] for _ in range(depth)]
    while instrs:
        taken = []
        if level == len(seq):
            seq.append([])
        for q in range(n):
            if not stacks[q]:
                continue
            g = stacks[q][0]
            if g.name in ("dummy_measure", "dummy_reset"):
                continue
            if g.name in ("u3", "u1"):
                if q in taken:
                    continue
                seq[level].append(g)
                taken.append(q)
                instrs.remove(g)
                stacks[q].pop(0)
            elif g.name in ("CX", "cx"):
                other = list(set(g.qubits).difference({q}))[0]
                head = stacks[other][0]
                if q in taken or other in taken:
                    continue
                if g == head:
                    taken += [q, other]
                    seq[level].append(g)
                    instrs.remove(g)
                    stacks[q].pop(0)
                    stacks[other].pop(0)
            elif g.name in ("measure", "reset"):
                kinds = (g.name, "dummy_" + g.name)
                if all((not stacks[x]) or stacks[x][0].name in kinds for x in range(n)):
                    if seq[level]:
                        taken = []
                        level += 1
                        if level == len(seq):
                            seq.append([])
                    for x in range(n):
                        if not stacks[x]:
                            continue
                        if stacks[x][0].name == g.name:
                            taken.append(x)
                            seq[level].append(stacks[x][0])
                            instrs.remove(stacks[x][0])
                        stacks[x].pop(0)
                    break
            if not instrs:
                break
        level += 1
    return seq, level


def partition(instrs, n):
    """``partition`` (``basicaertools.py:714-748``): barriers split segments."""
    segments, cur = [], []
    for ins in instrs:
        if ins.name != "barrier":
            cur.append(ins)
        else:
            segments.append(cur)
            cur = []
    if cur:
        segments.append(cur)
    parts, levels = [], 0
    for seg in segments:
        if not seg:
            continue
        # basicaertools.py:740-742 has a special branch for segments that start with a
        # 'Bell'/'Expect'/'Ensemble' measure, but it is dead in the real flow: params[0]
        # arrives as a sympy.Symbol and ``Symbol('Bell') in ['Bell', ...]`` is False (with
        # plain-str params the reference takes it and then crashes at dm_simulator.py:1006;
        # SURVEY.md a17).  Our params are plain str, so the branch is omitted, not restated.
        seq, lv = _partition_segment(seg, n)
        parts.append(seq)
        levels += lv
    return list(itertools.chain(*parts)), levels


# --------------------------------------------------------------------------------------
# a6-a11, a19-a26: state updates
# --------------------------------------------------------------------------------------

DEFAULT_OPTIONS = {
    "initial_densitymatrix": None, "chop_threshold": 1e-15, "thermal_factor": 1.,
    "decoherence_factor": 1., "depolarization_factor": 1., "bell_depolarization_factor": 1.,
    "decay_factor": 1., "tsp_model_error": [1., 0.],
}


class OracleSim:
    """Restatement of ``DmSimulatorPy`` (``dm_simulator.py:61-1313``) without the provider
    plumbing.  One object = one run; options as in ``_set_options`` (``:177-271``)."""

    def __init__(self, options=None):
        o = dict(options or {})
        self.rotation_error = {"rx": [1., 0.], "ry": [1., 0.], "rz": [1., 0.]}
        if "rotation_error" in o:
            re = o["rotation_error"]
            if type(re) != dict or not all(x in ["rx", "ry", "rz"] for x in re):
                raise OracleError("Incorrect Rotation Error parameters")
            self.rotation_error.update(re)
        self.tsp = DEFAULT_OPTIONS["tsp_model_error"]
        if "tsp_model_error" in o:
            t = o["tsp_model_error"]
            if type(t) != list or len(t) != 2 or t[0] > 1 or t[1] > 1:
                raise OracleError("Incorrect transition model error parameter")
            self.tsp = t
        self.p = o.get("thermal_factor", 1.)
        self.f = o.get("decoherence_factor", 1.)
        self.g = o.get("decay_factor", 1.)
        self.depol = o.get("depolarization_factor", 1.)
        self.bell_depol = 1.0  # dm_simulator.py:240 stores the option under the wrong name
        self.chop = o.get("chop_threshold", 1e-15)
        self.compute_dm = o.get("compute_densitymatrix", True)
        self.merge = o.get("merge", True)
        self.custom = o.get("custom_densitymatrix", None)
        self.initial = None
        if "initial_densitymatrix" in o:
            self.initial = np.array(o["initial_densitymatrix"], dtype=float)
        if self.custom in ("binary_string", "stored_density_matrix"):
            self.initial = o["initial_densitymatrix"]
        self.store_local = bool(o.get("store_densitymatrix", False))
        self.stored, self.compare = None, False
        if "compare" in o:                                   # :265-271
            try:
                self.stored = np.load("stored_coefficients.npy")
                self.compare = bool(o["compare"])
            except FileNotFoundError:
                print("Stored Coefficient File does not exist")
        self.fidelity = None
        self.n = 0
        self.dm = None

    # ---- a4/a5 -------------------------------------------------------------------
    def initialize(self, n):
        """``_initialize_densitymatrix`` / ``_validate_initial_densitymatrix`` (``:284-365``)."""
        self.n = n
        if self.initial is None:
            if self.custom is None:
                v = [1, 0, 0, 1]
            elif self.custom == "max_mixed":
                v = [1, 0, 0, 0]
            elif self.custom == "uniform_superpos":
                v = [1, 1, 0, 0]
            elif self.custom == "thermal_state":
                v = [1, 0, 0, 2 * self.p - 1]
            else:
                raise OracleError("_custom_densitymatrix value is invalid")
            dm = np.array(v, dtype=float)
            for _ in range(n - 1):
                dm = np.kron(v, dm)
            dm = dm * 0.5 ** n
        elif self.custom == "binary_string":
            s = self.initial
            if len(s) != n:
                raise OracleError("Wrong input binary string length")
            dm = np.array([1, 0, 0, 1 if s[0] == "0" else -1], dtype=float)
            for ch in s[1:]:
                dm = np.kron([1, 0, 0, 1 if ch == "0" else -1], dm)   # last char = qubit 0
            dm = dm * 0.5 ** n
        elif self.custom == "stored_density_matrix":
            dm = np.load("stored_density_matrix.npy")        # :337-343
            if len(dm) != 4 ** n:
                raise OracleError("Wrong input stored density matrix")
        else:
            # a bare ``initial_densitymatrix`` vector always ends here (:344-345)
            raise OracleError("_custom_densitymatrix value is invalid")
        if self.initial is not None:
            if dm.size != 4 ** n:
                raise OracleError("initial densitymatrix is incorrect length")
            if dm[0] != 2.0 ** (-n):
                raise OracleError("Trace of initial densitymatrix is not one")
        self.dm = np.ascontiguousarray(dm, dtype=float).reshape(-1)

    def _v3(self, q):
        return self.dm.reshape(4 ** q, 4, 4 ** (self.n - q - 1))

    # ---- a6-a8 -------------------------------------------------------------------
    def rotate(self, axis, angle, q):
        """``rot_gate_dm_matrix`` (``basicaertools.py:93-126``)."""
        r, delta = self.rotation_error[axis]
        c = r * np.cos(angle + delta)
        s = r * np.sin(angle + delta)
        k0, k1 = {"rz": (1, 2), "ry": (3, 1), "rx": (2, 3)}[axis]
        v = self._v3(q)
        t1 = v[:, k0, :].copy()
        t2 = v[:, k1, :].copy()
        v[:, k0, :] = c * t1 - s * t2
        v[:, k1, :] = c * t2 + s * t1

    def single_gate(self, name, params, q):
        """``single_gate_dm_matrix`` + ``_add_unitary_single`` (``basicaertools.py:69-90``,
        ``dm_simulator.py:367-384``): u3 = rz(lam), ry(theta), rz(phi) in that order."""
        prm = list(map(float, params))
        if name in ("U", "u3"):
            seq = [("rz", prm[2]), ("ry", prm[0]), ("rz", prm[1])]
        elif name == "u1":
            seq = [("rz", prm[0])]
        else:
            raise OracleError("Gate is not among the valid types: %s" % name)
        for axis, ang in seq:
            self.rotate(axis, ang, q)

    # ---- a9/a10 ------------------------------------------------------------------
    def cx(self, ctrl, tgt):
        """``cx_gate_dm_matrix`` (``basicaertools.py:310-392``).  ``blk[c][t]`` below is the
        plane with control digit c and target digit t, whichever qubit is the outer axis."""
        n = self.n
        cav, e1 = self.tsp
        c2av = 4 * cav - 3
        c = cav * np.cos(e1)
        s = cav * np.sin(e1)
        c2 = 0.5 * (1 + c2av * np.cos(2 * e1))
        s2 = 0.5 * (1 - c2av * np.cos(2 * e1))
        cs = c2av * np.sin(e1) * np.cos(e1)
        if ctrl == tgt or ctrl >= n or tgt >= n:
            raise OracleError("Qubit Labels out of bound in CX Gate")
        lo, hi = min(ctrl, tgt), max(ctrl, tgt)
        v = self.dm.reshape(4 ** lo, 4, 4 ** (hi - lo - 1), 4, 4 ** (n - hi - 1))
        old = v.copy()

        def o(ci, ti):
            return old[:, ci, :, ti, :] if ctrl < tgt else old[:, ti, :, ci, :]

        def put(ci, ti, val):
            if ctrl < tgt:
                v[:, ci, :, ti, :] = val
            else:
                v[:, ti, :, ci, :] = val

        I, X, Y, Z = 0, 1, 2, 3
        # {IY, ZY, IZ, ZZ}   (:346-353 / :371-378)
        put(I, Y, s2 * o(I, Y) + c2 * o(Z, Y) - cs * (o(I, Z) - o(Z, Z)))
        put(Z, Y, c2 * o(I, Y) + s2 * o(Z, Y) + cs * (o(I, Z) - o(Z, Z)))
        put(I, Z, s2 * o(I, Z) + c2 * o(Z, Z) + cs * (o(I, Y) - o(Z, Y)))
        put(Z, Z, c2 * o(I, Z) + s2 * o(Z, Z) - cs * (o(I, Y) - o(Z, Y)))
        # control X row (:355-358 / :380-383)
        put(X, I, c * o(X, X) - s * o(Y, I))
        put(X, X, c * o(X, I) - s * o(Y, X))
        put(X, Y, -s * o(Y, Y) + c * o(Y, Z))
        put(X, Z, -c * o(Y, Y) - s * o(Y, Z))
        # control Y row (:360-363 / :385-388)
        put(Y, I, s * o(X, I) + c * o(Y, X))
        put(Y, X, s * o(X, X) + c * o(Y, I))
        put(Y, Y, s * o(X, Y) - c * o(X, Z))
        put(Y, Z, c * o(X, Y) + s * o(X, Z))

    # ---- a11 -----------------------------------------------------------------------
    def memory_noise(self):
        """``_add_decoherence_and_amp_decay`` (``dm_simulator.py:397-425``)."""
        off = np.sqrt(self.g) * self.f
        dd = (1 - self.g) * (2 * self.p - 1)
        for q in range(self.n):
            v = self._v3(q)
            v[:, 1, :] = off * v[:, 1, :]
            v[:, 2, :] = off * v[:, 2, :]
            v[:, 3, :] = self.g * v[:, 3, :] + dd * v[:, 0, :]

    # ---- a19 -----------------------------------------------------------------------
    def measure_axis(self, q, basis, err):
        """``_add_qasm_measure_X/Y/Z`` (``dm_simulator.py:574-664``)."""
        keep = {"X": 1, "Y": 2, "Z": 3}[basis]
        v = self._v3(q)
        for comp in (1, 2, 3):
            if comp == keep:
                v[:, comp, :] *= err
            else:
                v[:, comp, :] = 0

    def measure_n(self, q, nvec, err):
        """``_add_qasm_measure_N`` (``dm_simulator.py:666-704``)."""
        v = self._v3(q)
        t = nvec[0] * v[:, 1, :] + nvec[1] * v[:, 2, :] + nvec[2] * v[:, 3, :]
        t *= err
        v[:, 1, :] = t * nvec[0]
        v[:, 2, :] = t * nvec[1]
        v[:, 3, :] = t * nvec[2]

    @staticmethod
    def unit(nvec):
        """``_unit_vector_normalisation`` (``dm_simulator.py:706-719``)."""
        nvec = np.array(nvec, dtype=float)
        norm = np.linalg.norm(nvec)
        return nvec if norm == 1 else nvec / norm

    # ---- a20 -----------------------------------------------------------------------
    def ensemble(self, basis, add_param, err):
        """``_add_ensemble_measure`` (``dm_simulator.py:427-481``) evaluated as the scaled
        Walsh-Hadamard transform it is (the reference builds the dense 2^n x 2^n kron
        matrix, ``:451-455``; same numbers up to summation order)."""
        n = self.n
        t = self.dm.reshape(n * [4])
        if basis != "N":
            b = {"X": 1, "Y": 2, "Z": 3}[basis]
            m = t
            for q in range(n):
                m = np.take(m, [0, b], axis=q)
            w = np.array([1.0, err])
        else:
            nv = np.asarray(add_param, dtype=float) * err
            m = t
            for q in range(n):
                m = np.moveaxis(m, q, -1)
                m = np.stack([m[..., 0], m[..., 1] * nv[0] + m[..., 2] * nv[1] + m[..., 3] * nv[2]], axis=-1)
                m = np.moveaxis(m, -1, q)
            w = np.array([1.0, 1.0])
        m = np.array(m, dtype=float)
        for q in range(n):
            m = np.moveaxis(m, q, -1)
            a0, a1 = m[..., 0] * w[0], m[..., 1] * w[1]
            m = np.stack([a0 + a1, a0 - a1], axis=-1)
            m = np.moveaxis(m, -1, q)
        probs = m.reshape(2 ** n)
        keys = ["".join(s) for s in itertools.product("01", repeat=n)]
        if self.store_local:
            np.save("stored_coefficients", self.dm.reshape(4 ** n))        # :476-477, :1271-1275
        if self.compare:
            self.fidelity = np.dot(self.stored, self.dm) * 2 ** n          # :479-480, :1277-1282
        return dict(zip(keys, probs))

    # ---- a21 -----------------------------------------------------------------------
    def partial(self, qubits, basis, add_param, err):
        """``_add_partial_measure`` (``dm_simulator.py:492-538``)."""
        n = self.n
        ens = np.array(list(self.ensemble(basis, add_param, err).values()))
        axes = tuple(set(range(n)) - set(qubits))
        probs = np.reshape(np.sum(ens.reshape(n * [2]), axis=axes), 2 ** len(qubits))
        keys = ["".join(s) for s in itertools.product("01", repeat=len(qubits))]
        for q in qubits:
            if basis == "N" and add_param is not None:
                self.measure_n(q, add_param, self.depol)
            elif basis == "N":
                self.measure_n(q, np.array([0.0, 0.0, 1.0]), self.depol)
            else:
                self.measure_axis(q, basis, self.depol)
        return dict(zip(keys, probs))

    # ---- a22 -----------------------------------------------------------------------
    def expectation(self, letters, err):
        """``_pauli_string_expectation`` (``dm_simulator.py:540-572``)."""
        for q in range(self.n):
            if letters[q] in "XYZ":
                self.measure_axis(q, letters[q], err)
        idx = tuple("IXYZ".index(ch) for ch in letters)
        return self.dm.reshape(self.n * [4])[idx] * 2 ** self.n

    # ---- a23 -----------------------------------------------------------------------
    def bell(self, qa, qb, err):
        """``_add_bell_basis_measure`` (``dm_simulator.py:721-777``).  NB the reshape counts
        axes from the other end: the digits touched are qubits n-1-max and n-1-min."""
        n = self.n
        q1, q2 = min(qa, qb), max(qa, qb)
        v = self.dm.reshape(4 ** (n - q2 - 1), 4, 4 ** (q2 - q1 - 1), 4, 4 ** q1)
        red = np.zeros((4, 4))
        for i in range(4):
            for j in range(4):
                red[i, j] = v[0, i, 0, j, 0] * 2 ** (n - 2)
        for i in range(4):
            for j in range(4):
                if i != j:
                    v[:, i, :, j, :] = 0
        for i in (1, 2, 3):
            v[:, i, :, i, :] *= err
        k = [v[0, i, 0, i, 0] * 2 ** n for i in range(4)]
        probs = [0.25 * (k[0] + k[1] - k[2] + k[3]), 0.25 * (k[0] - k[1] + k[2] + k[3]),
                 0.25 * (k[0] + k[1] + k[2] - k[3]), 0.25 * (k[0] - k[1] - k[2] - k[3])]
        return dict(zip(["Bell_1", "Bell_2", "Bell_3", "Bell_4"], probs)), red

    # ---- a24 -----------------------------------------------------------------------
    def reset(self, q):
        """``_add_qasm_reset`` (``dm_simulator.py:810-823``)."""
        v = self._v3(q)
        v[:, 1, :] = 0
        v[:, 2, :] = 0
        v[:, 3, :] = v[:, 0, :].copy()

    # ---- a25/a26 -------------------------------------------------------------------
    def to_matrix(self):
        """``_compute_densitymatrix`` (``dm_simulator.py:1198-1255``): per digit
        [I,X,Y,Z] -> [I+Z, X-iY, X+iY, I-Z], digit mu=(rowbit,colbit), qubit 0 = MSB."""
        n = self.n
        t = self.dm.astype(complex)
        for q in range(n):
            v = t.reshape(4 ** q, 4, 4 ** (n - q - 1))
            w = np.zeros_like(v)
            w[:, 0, :] = v[:, 0, :] + v[:, 3, :]
            w[:, 1, :] = v[:, 1, :] + complex(0, -1) * v[:, 2, :]
            w[:, 2, :] = v[:, 1, :] + complex(0, 1) * v[:, 2, :]
            w[:, 3, :] = v[:, 0, :] + (-1) * v[:, 3, :]
            t = w
        t = t.reshape([2, 2] * n)                     # (r0,c0,r1,c1,...)
        perm = list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2))
        return np.ascontiguousarray(np.transpose(t, perm)).reshape(2 ** n, 2 ** n)

    def final_state(self):
        """``_get_densitymatrix`` (``dm_simulator.py:1257-1269``): matrix first, then chop."""
        mat = self.to_matrix() if self.compute_dm else None
        vec = self.dm.reshape(4 ** self.n)
        vec[abs(vec) < self.chop] = 0.0
        return vec, mat

    # ---- a18 -----------------------------------------------------------------------
    def run_experiment(self, n, instrs, name="circuit"):
        """``run_experiment`` (``dm_simulator.py:950-1196``) including its measurement
        dispatch quirks (skipped second measure after list.remove while iterating, Bell
        axis reversal, noise after measure levels)."""
        t0 = time.time()
        self.initialize(n)
        ops = single_gate_merge([as_instruction(i) for i in instrs], n, self.merge)
        levels_list, levels = partition(ops, n)
        t1 = time.time()
        data = {}
        part_measure = None
        for clock in range(levels):
            level = levels_list[clock]
            for op in level:                                             # :1006-1018
                if op.name == "measure":
                    basis = str(getattr(level[0], "params", ["Z"])[0])
                    prm = getattr(op, "params", ["Z"])
                    if str(prm[0]) in ["X", "Y", "Z", "N"]:
                        if str(prm[0]) == basis:
                            part_measure = True
                        else:
                            part_measure = False
                            break
                    else:
                        part_measure = False
                        continue
            it = iter(_LiveIter(level))
            for op in it:
                if conditional_skips(op):
                    continue
                if op.name in ("u1", "u3"):
                    self.single_gate(op.name, op.params, op.qubits[0])
                elif op.name == "cx":
                    self.cx(op.qubits[0], op.qubits[1])
                elif op.name == "reset":
                    self.reset(op.qubits[0])
                elif op.name == "barrier":
                    pass
                elif op.name == "measure":
                    if not hasattr(op, "params"):
                        op.params = ["Z"]
                    prm = op.params
                    q = op.qubits[0]
                    kind = str(prm[0])
                    single = False
                    if kind == "Ensemble":
                        mode = "ens"
                    elif kind == "Expect":
                        mode = "exp"
                    elif len(level) == 1 or part_measure is not True:
                        mode, single = "single", True
                    else:
                        mode = "part"
                    if len(prm) == 1:
                        prm.append(None)
                    if single:
                        if kind in ("X", "Y"):
                            self.measure_axis(q, kind, self.depol)
                        elif kind == "N":
                            prm[1] = self.unit(prm[1])
                            self.measure_n(q, prm[1], self.depol)
                        elif kind == "Bell":
                            s = str(prm[1])
                            probs, red = self.bell(int(s[0]), int(s[1]), self.bell_depol)
                            data["bell_probabilities" + s[0] + s[1]] = probs
                            data["reduced_bell_densitymatrix" + s[0] + s[1]] = red
                        else:
                            self.measure_axis(q, "Z", self.depol)
                        level.remove(op)                                 # :1100 (quirk)
                    elif mode == "part":
                        qs = [x.qubits[0] for x in level]
                        add = prm[1] if kind == "N" else None
                        data["partial_probability"] = self.partial(qs, kind, add, self.depol)
                        break
                    elif mode == "exp":
                        data["Pauli_string_expectation"] = self.expectation(str(prm[1]), self.depol)
                        break
                    else:
                        if len(str(prm[1])) == 1:
                            add, basis = None, str(prm[1])
                        elif prm[1][0] == "N" or str(prm[1][0]) == "N":
                            add, basis = self.unit(prm[1][1]), "N"
                        data["ensemble_probability"] = self.ensemble(basis, add, self.depol)
                        break
                elif op.name == "bfunc":
                    pass   # classical registers are never written (:1138-1166 is inert)
                else:
                    raise OracleError('encountered unrecognized operation "%s"' % op.name)
            self.memory_noise()                                          # :1173-1177
        vec, mat = self.final_state()
        data["coeffmatrix"] = vec
        if self.compute_dm:
            data["densitymatrix"] = mat
        if self.fidelity is not None:
            data["fidelity"] = self.fidelity
        t2 = time.time()
        return {"name": name, "number_of_clock_cycles": levels, "data": data, "status": "DONE",
                "success": True, "processing_time_taken": t1 - t0, "running_time_taken": t2 - t1,
                "header": {"name": name}}


class _LiveIter:
    """Index-based iteration over a list that may shrink underneath us -- the behaviour of
    Python's list iterator that ``dm_simulator.py:1100`` relies on (and trips over)."""

    def __init__(self, lst):
        self.lst = lst

    def __iter__(self):
        i = 0
        while i < len(self.lst):
            yield self.lst[i]
            i += 1


def run_oracle(n_qubits, instrs, options=None, name="circuit"):
    """Convenience wrapper mirroring ``oracle.ref_harness.run_reference``."""
    return OracleSim(options).run_experiment(n_qubits, instrs, name)
