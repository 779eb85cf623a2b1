"""Host-side instruction passes: single-qubit gate merging and level ("clock cycle") partitioning.

These two passes are part of the *numerical* contract of the reference, not just of its
performance: rotation errors are charged per merged gate and memory noise per level
("hard parts"), so both are reproduced with the reference's exact
semantics, including the Euler-angle formulae of ``U3_merge`` in float64:

* ``merge_single_qubit_gates``  <->  ``single_gate_merge`` / ``merge_gates`` / ``mergeU`` /
  ``U3_merge``  (``qiskit/providers/basicaer/basicaertools.py:129-307``)
* ``partition_levels``          <->  ``partition`` / ``partition_helper`` / ``qubit_stack``
  (``basicaertools.py:536-748``)

Instructions come in as the duck-typed qobj objects the reference backend reads
(``name``, ``qubits``, optional ``params`` / ``memory`` / ``register``); internally they are
turned into ``Op`` records so the caller's objects are never mutated (the reference edits
them in place).
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np

from .exceptions import BasicAerError

_ROT = ("U", "u1", "u2", "u3")
_FLUSHERS = ("CX", "cx", "measure", "bfunc", "reset", "barrier")


class Op:
    """One instruction after renaming: name in {u1,u3,cx,measure,reset,barrier,bfunc}."""
    __slots__ = ("name", "qubits", "params", "memory", "register", "order", "conditional")

    def __init__(self, name, qubits, params=None, memory=None, register=None, order=0, conditional=None):
        self.name = name
        self.qubits = list(qubits)
        self.params = params
        self.memory = memory
        self.register = register
        self.order = order
        self.conditional = conditional      # carried through untouched (dm_simulator.py:1020-1034)

    def __repr__(self):
        return "Op(%s%s%s)" % (self.name, self.qubits, "" if self.params is None else self.params)


def _to_op(ins, order):
    if isinstance(ins, dict):
        get = ins.get
        name, params, mem, reg = ins["name"], get("params"), get("memory"), get("register")
        qubits, cond = get("qubits"), get("conditional")
    else:
        name, params = ins.name, getattr(ins, "params", None)
        mem, reg = getattr(ins, "memory", None), getattr(ins, "register", None)
        qubits, cond = getattr(ins, "qubits", None), getattr(ins, "conditional", None)
    return Op(name, qubits or [], list(params) if params is not None else None, list(mem) if mem is not None else None,
              list(reg) if reg is not None else None, order, cond)


def zyz_from_yzy(xi, theta1, theta2):
    """Ry(theta1) Rz(xi) Ry(theta2)  ->  (beta, alpha, gamma) with Rz(alpha) Ry(beta) Rz(gamma).

    Same float64 expressions as ``U3_merge`` (``basicaertools.py:162-178``); the arccos is
    ill-conditioned near the identity and parity at 1e-10 depends on keeping it."""
    half = 0.5
    s_xi, c_xi = np.sin(xi * half), np.cos(xi * half)
    s_sum, c_sum = np.sin((theta1 + theta2) * half), np.cos((theta1 + theta2) * half)
    s_dif, c_dif = np.sin((theta1 - theta2) * half), np.cos((theta1 - theta2) * half)
    half_sum = np.arctan2(s_xi * c_dif, c_xi * c_sum)
    half_dif = np.arctan2(-s_xi * s_dif, c_xi * s_sum)
    beta = 2 * np.arccos(np.sqrt((c_xi * c_sum) ** 2 + (s_xi * c_dif) ** 2))
    return beta, half_sum + half_dif, half_sum - half_dif


def _fuse(first, second):
    """Compose two rotations on one qubit (``first`` acts first) -> (name, params).
    Rules of ``mergeU`` (``basicaertools.py:198-224``)."""
    (n1, p1), (n2, p2) = first, second
    if n1 == "u1" and n2 == "u1":
        return "u1", [p1[0] + p2[0]]
    if n1 == "u1" and n2 == "u3":
        return "u3", [p2[0], p2[1], p2[2] + p1[0]]
    if n1 == "u3" and n2 == "u1":
        return "u3", [p1[0], p1[1] + p2[0], p1[2]]
    if n1 == "u3" and n2 == "u3":
        beta, alpha, gamma = zyz_from_yzy(float(p2[2] + p1[1]), float(p2[0]), float(p1[0]))
        return "u3", [beta, p2[1] + alpha, p1[2] + gamma]
    raise BasicAerError("Encountered unrecognized instructions: %s, %s" % (n1, n2))


def _rename(op):
    """U -> u3, CX -> cx, u2(phi,lam) -> u3(pi/2,phi,lam)  (``basicaertools.py:273-282``)."""
    if op.name == "CX":
        op.name = "cx"
    elif op.name == "U":
        op.name = "u3"
    elif op.name == "u2":
        op.name = "u3"
        op.params = [np.pi / 2] + list(op.params)
    return op


def merge_single_qubit_gates(instructions, n_qubits, merge=True):
    """Return the merged instruction list as ``Op`` records.

    Consecutive rotations on a qubit are folded left to right; ANY cx / measure / bfunc /
    reset / barrier first emits the pending rotation of every qubit (ascending qubit
    number) and then itself; ``id`` / ``u0`` vanish; anything else is an error."""
    out = []
    if not merge:
        for k, ins in enumerate(instructions):
            op = _rename(_to_op(ins, k))
            if op.name not in ("id", "u0"):
                out.append(op)
        return out

    # per qubit: (name, params, order of first gate, conditional of the first gate -- mergeU keeps a copy of the
    # earlier instruction, basicaertools.py:193-196, and with it that instruction's `conditional`)
    pending = [None] * n_qubits

    def flush():
        for q in range(n_qubits):
            if pending[q] is not None:
                name, params, order, cond = pending[q]
                out.append(Op(name, [q], params, order=order, conditional=cond))
                pending[q] = None

    for k, ins in enumerate(instructions):
        op = _to_op(ins, k)
        if op.name in _FLUSHERS:
            flush()
            out.append(_rename(op))
        elif op.name in _ROT:
            _rename(op)
            q = op.qubits[0]
            if pending[q] is None:
                pending[q] = (op.name, list(op.params), k, op.conditional)
            else:
                name, params, order, cond = pending[q]
                fused_name, fused_params = _fuse((name, params), (op.name, op.params))
                pending[q] = (fused_name, fused_params, order, cond)
        elif op.name in ("id", "u0"):
            continue
        else:
            raise BasicAerError("Encountered unrecognized instruction: %s" % op.name)
    flush()
    return out


# --------------------------------------------------------------------------------------
# level partition
# --------------------------------------------------------------------------------------

class _Marker:
    """Placeholder a measure/reset pushes on every other qubit's queue (``dummy_*``)."""
    __slots__ = ("name",)

    def __init__(self, kind):
        self.name = "dummy_" + kind


def _per_qubit_queues(ops, n_qubits):
    queues = [deque() for _ in range(n_qubits)]
    for op in ops:
        if op.name in ("measure", "reset"):
            q = op.qubits[0]
            own = queues[q]
            if own and own[-1].name == "dummy_" + op.name:
                own[-1] = op                  # joins the layer opened by an earlier measure/reset
                continue
            own.append(op)
            marker = _Marker(op.name)
            for other in range(n_qubits):
                if other != q:
                    queues[other].append(marker)
        else:
            for q in op.qubits:
                queues[q].append(op)
    return queues


def _partition_segment(ops, n_qubits):
    """Greedy ASAP levelling of one barrier-free segment (``partition_helper``)."""
    queues = _per_qubit_queues(ops, n_qubits)
    depth = max(len(q) for q in queues) if queues else 0
    levels = [[] for _ in range(depth)]
    remaining = len(ops)
    level = 0
    while remaining:
        busy = set()
        if level == len(levels):
            levels.append([])
        for q in range(n_qubits):
            if not queues[q]:
                continue
            head = queues[q][0]
            kind = head.name
            if kind.startswith("dummy_"):
                continue
            if kind in ("u3", "u1"):
                if q in busy:
                    continue
                levels[level].append(head)
                busy.add(q)
                queues[q].popleft()
                remaining -= 1
            elif kind in ("cx", "CX"):
                other = head.qubits[0] if head.qubits[1] == q else head.qubits[1]
                if q in busy or other in busy:
                    continue
                if queues[other][0] is head:
                    busy.update((q, other))
                    levels[level].append(head)
                    queues[q].popleft()
                    queues[other].popleft()
                    remaining -= 1
            elif kind in ("measure", "reset"):
                allowed = (kind, "dummy_" + kind)
                if all((not qu) or qu[0].name in allowed for qu in queues):
                    if levels[level]:
                        busy = set()
                        level += 1
                        if level == len(levels):
                            levels.append([])
                    for x in range(n_qubits):
                        if not queues[x]:
                            continue
                        if queues[x][0].name == kind:
                            busy.add(x)
                            levels[level].append(queues[x][0])
                            remaining -= 1
                        queues[x].popleft()
                    break
            else:
                # the reference would spin forever here (no branch of partition_helper matches)
                raise BasicAerError('cannot schedule instruction "%s"' % kind)
            if not remaining:
                break
        level += 1
    return levels, level


def partition_levels(ops, n_qubits):
    """Split at barriers (dropped), level each segment, concatenate.

    Returns ``(levels, n_levels)`` exactly like the reference: ``levels`` may be longer
    than ``n_levels`` (a segment can leave empty trailing levels behind, which the
    reference then walks instead of the real tail -- reproduced, not repaired)."""
    segments, cur = [], []
    for op in ops:
        if op.name == "barrier":
            segments.append(cur)
            cur = []
        else:
            cur.append(op)
    if cur:
        segments.append(cur)
    all_levels, total = [], 0
    for seg in segments:
        if not seg:
            continue
        lv, cnt = _partition_segment(seg, n_qubits)
        all_levels.extend(lv)
        total += cnt
    return all_levels, total
