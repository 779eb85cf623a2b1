"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference front-end.

Pins ``qiskit-aakash_b200/frontend.py`` (the QuantumCircuit / execute facade of SURVEY.md
section 8(f)1): circuit construction, argument broadcasting, gate decomposition
(``Unroller`` over ``extensions/standard/*._define``), ``RemoveResetInZeroState`` and the
lexicographic topological order of ``dag_to_circuit`` are all run through the reference's own
files under ``/root/reference/qiskit`` and the resulting instruction list (what ``assemble``
hands to ``DmSimulatorPy.run``) is recorded as a golden fixture by
``tests/golden/make_frontend_golden.py``.

``import qiskit`` does not work in this image (marshmallow, ply, matplotlib and the compiled
cython passes are missing).  The pieces used here do not need them
functionally, so:

  * ``qiskit``, ``qiskit.transpiler`` and ``qiskit.transpiler.passes`` are created as bare
    packages pointing at the reference directories (their ``__init__`` files, which pull in
    everything, are not executed);
  * ``qiskit.validation`` and ``qiskit.qobj*`` (marshmallow models) are replaced by
    ``SimpleNamespace`` stand-ins -- ``Instruction.assemble`` only needs an attribute bag;
  * ``ply``, ``matplotlib`` and ``sympy.printing.ccode`` (renamed in current SymPy) are stubbed,
    and the ``numpy.float``-style aliases the old code mentions are restored for this process.

Must run in its own process (it claims the ``qiskit`` module name; ``oracle/ref_harness.py``
claims it differently).  Nothing in the product package, ``bench.py`` or the ``-m gpu`` tests
may import this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("DMB_REFERENCE_ROOT", "/root/reference")
BASIS_GATES = ["u1", "u2", "u3", "cx", "id", "unitary"]       # dm_simulator.py:85

_api = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "qiskit", "circuit", "quantumcircuit.py"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _package(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, key):
        return _Inert()


class _Model(SimpleNamespace):
    def as_dict(self):
        return dict(self.__dict__)

    to_dict = as_dict


def load():
    """Import the reference front-end; returns a namespace with QuantumCircuit, QuantumRegister,
    ClassicalRegister, pi and ``lower(circuit) -> [instruction dict]``."""
    global _api
    if _api is not None:
        return _api
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "qiskit" in sys.modules:
        raise RuntimeError("a module named qiskit is already imported in this process")
    import numpy
    for alias, typ in (("float", float), ("int", int), ("complex", complex), ("bool", bool)):
        if alias not in numpy.__dict__:
            setattr(numpy, alias, typ)
    import sympy.printing.c as _c
    _stub("sympy.printing.ccode", ccode=_c.ccode)
    ply = _stub("ply")
    ply.lex = _stub("ply.lex", lex=_Inert())
    ply.yacc = _stub("ply.yacc", yacc=_Inert())
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        _stub(name, Axes3D=object)
    qdir = os.path.join(REFERENCE_ROOT, "qiskit")
    _package("qiskit", qdir)
    _stub("qiskit.validation", BaseModel=_Model, BaseSchema=object, bind_schema=lambda schema: (lambda cls: cls),
          ModelTypeValidator=object)
    _stub("qiskit.qobj")
    _stub("qiskit.qobj.models")
    _stub("qiskit.qobj.models.qasm", QasmQobjInstruction=_Model)
    _stub("qiskit.qobj.models.pulse", PulseQobjInstruction=_Model)
    _package("qiskit.transpiler", os.path.join(qdir, "transpiler"))
    _package("qiskit.transpiler.passes", os.path.join(qdir, "transpiler", "passes"))
    for mod in ("qiskit.exceptions", "qiskit.circuit", "qiskit.extensions.standard", "qiskit.dagcircuit",
                "qiskit.converters", "qiskit.transpiler.exceptions", "qiskit.transpiler.basepasses"):
        importlib.import_module(mod)
    circuit = sys.modules["qiskit.circuit"]
    converters = sys.modules["qiskit.converters"]
    unroller = importlib.import_module("qiskit.transpiler.passes.unroller")
    rrz = importlib.import_module("qiskit.transpiler.passes.remove_reset_in_zero_state")
    qasm = importlib.import_module("qiskit.qasm")

    def lower(qc):
        """transpile (default_pass_manager_simulator, preset_passmanagers/default.py:95-112) +
        the per-instruction part of assemble_circuits (assembler/assemble_circuits.py:82-104)."""
        dag = converters.circuit_to_dag(qc)
        dag = unroller.Unroller(list(BASIS_GATES)).run(dag)
        depth = None
        while True:                                   # [RemoveResetInZeroState, Depth, FixedPoint('depth')]
            dag = rrz.RemoveResetInZeroState().run(dag)
            d = dag.depth()
            if d == depth:
                break
            depth = d
        out = converters.dag_to_circuit(dag)
        qubit_labels = [[r.name, j] for r in out.qregs for j in range(r.size)]
        clbit_labels = [[r.name, j] for r in out.cregs for j in range(r.size)]
        instrs = []
        for inst, qargs, cargs in out.data:
            a = inst.assemble()
            rec = {"name": a.name}
            if qargs:
                rec["qubits"] = [qubit_labels.index([q.register.name, q.index]) for q in qargs]
            if cargs:
                rec["memory"] = [clbit_labels.index([c.register.name, c.index]) for c in cargs]
            if hasattr(a, "params"):
                rec["params"] = [_plain(p) for p in a.params]
            instrs.append(rec)
        return {"n_qubits": len(qubit_labels), "memory_slots": len(clbit_labels), "name": out.name,
                "instructions": instrs}

    _api = SimpleNamespace(QuantumCircuit=circuit.QuantumCircuit, QuantumRegister=circuit.QuantumRegister,
                           ClassicalRegister=circuit.ClassicalRegister, pi=qasm.pi, lower=lower, kind="reference")
    return _api


def _plain(p):
    """JSON-able view of an assembled parameter: floats as hex (exact), symbols as str."""
    import numpy as np
    import sympy
    if isinstance(p, np.ndarray):
        return {"array": [float(x).hex() for x in p.reshape(-1)]}
    if isinstance(p, sympy.Symbol):
        return {"symbol": str(p)}
    if isinstance(p, (sympy.Basic, float, int)):
        return {"float": float(p).hex()}
    return {"repr": repr(p)}
