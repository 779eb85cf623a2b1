"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference dm_simulator.

This module exists to pin ``oracle/dm_oracle.py`` (the CPU restatement) and to
generate the golden fixtures under ``tests/golden/``.  It imports the reference's
own files **by path** from ``/root/reference`` (read-only; present in the build
container, absent on the GPU box) or, where that tree does not exist, from the copy of
exactly these three files that ``oracle/stage_ref.py`` (run by ``__graft_entry__.build()`` in
the build container) stages under the git-ignored ``oracle/_ref/`` so that it travels to the
GPU box with the snapshot -- ``bench.py``'s CPU legs then time the real reference there:

    <root>/qiskit/providers/basicaer/exceptions.py
    <root>/qiskit/providers/basicaer/basicaertools.py
    <root>/qiskit/providers/basicaer/dm_simulator.py

``import qiskit`` itself cannot work here (marshmallow / ply / matplotlib are not
installed), therefore the handful of modules the two hot-path
files import are replaced by inert stubs before loading them.  Nothing in the
product package, in ``bench.py`` or in the ``-m gpu`` tests may import this file.
"""
from __future__ import annotations

import copy
import importlib.util
import os
import sys
import types
from types import SimpleNamespace as NS

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root():
    env = os.environ.get("DMB_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "qiskit", "providers", "basicaer", "dm_simulator.py")):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _pick_root()
_BASICAER = os.path.join(REFERENCE_ROOT, "qiskit", "providers", "basicaer")

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(_BASICAER, "dm_simulator.py"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load():
    """Return (dm_simulator module, basicaertools module) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "qiskit" in sys.modules and not getattr(sys.modules["qiskit"], "_dmb_stub", False):
        raise RuntimeError("a real qiskit is already imported; refusing to stub over it")

    class QiskitError(Exception):
        def __init__(self, *message):
            super().__init__(" ".join(str(m) for m in message))
            self.message = " ".join(str(m) for m in message)

    class BaseBackend:
        def __init__(self, configuration, provider=None):
            self._configuration = configuration
            self._provider = provider

        def name(self):
            return self._configuration.backend_name

        def configuration(self):
            return self._configuration

    class QasmBackendConfiguration(NS):
        @classmethod
        def from_dict(cls, d):
            return cls(**d)

    import psutil

    plt = _stub("matplotlib.pyplot")
    cm = _stub("matplotlib.cm")
    _stub("matplotlib", pyplot=plt, cm=cm)
    m3d = _stub("mpl_toolkits.mplot3d", Axes3D=object)
    _stub("mpl_toolkits", mplot3d=m3d)

    q = _stub("qiskit", _dmb_stub=True)
    q.__path__ = []
    _stub("qiskit.exceptions", QiskitError=QiskitError)
    _stub("qiskit.util", local_hardware_info=lambda: {
        "memory": psutil.virtual_memory().total / (1024 ** 3), "cpus": os.cpu_count()})
    prov = _stub("qiskit.providers", BaseBackend=BaseBackend)
    prov.__path__ = []
    _stub("qiskit.providers.models", QasmBackendConfiguration=QasmBackendConfiguration)
    _stub("qiskit.result", Result=object)
    pkg = _stub("qiskit.providers.basicaer")
    pkg.__path__ = [_BASICAER]
    _stub("qiskit.providers.basicaer.basicaerjob", BasicAerJob=object)

    mods = {}
    for short in ("exceptions", "basicaertools", "dm_simulator"):
        full = "qiskit.providers.basicaer." + short
        spec = importlib.util.spec_from_file_location(full, os.path.join(_BASICAER, short + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        mods[short] = mod
    _loaded = (mods["dm_simulator"], mods["basicaertools"])
    return _loaded


def to_reference_instructions(instrs):
    """Our plain instruction dicts/namespaces -> the duck-typed objects the reference reads.

    String measure params must arrive as ``sympy.Symbol`` (the assembler does that,
    ``circuit/instruction.py:140-141``); numeric ones stay floats.
    """
    import numpy as np
    import sympy

    out = []
    for ins in instrs:
        d = dict(ins) if isinstance(ins, dict) else dict(vars(ins))
        ns = NS(name=d["name"], qubits=list(d.get("qubits", [])))
        if "params" in d and d["params"] is not None:
            ps = []
            for p in d["params"]:
                if isinstance(p, str):
                    ps.append(sympy.Symbol(p))
                elif isinstance(p, (list, tuple, np.ndarray)) and len(p) == 2 and isinstance(p[0], str):
                    # Ensemble + 'N': the only form the real front-end lets through is an object
                    # ndarray ['N', direction] (circuit/instruction.py:142-143 accepts ndarrays,
                    # rejects lists); dm_simulator.py:1130 then compares its element 0 with 'N'
                    arr = np.empty(2, dtype=object)
                    arr[0], arr[1] = str(p[0]), np.array(p[1], dtype=float)
                    ps.append(arr)
                elif isinstance(p, (list, tuple, np.ndarray)):
                    ps.append(np.array(p, dtype=float))
                else:
                    ps.append(float(p))
            ns.params = ps
        if "memory" in d:
            ns.memory = list(d["memory"])
        if "register" in d:
            ns.register = list(d["register"])
        if d.get("conditional") is not None:
            ns.conditional = copy.deepcopy(d["conditional"])
        out.append(ns)
    return out


def run_reference(n_qubits, instrs, options=None, name="circuit"):
    """Run one experiment through the reference's ``DmSimulatorPy.run_experiment``.

    A fresh simulator object and fresh class-level default options are used for every
    call because ``_set_options`` mutates the shared ``DEFAULT_OPTIONS['rotation_error']``
    dict in place (``dm_simulator.py:182,212-213``).
    """
    dm, _ = load()
    cls = dm.DmSimulatorPy
    cls.DEFAULT_OPTIONS = dict(cls.DEFAULT_OPTIONS)
    cls.DEFAULT_OPTIONS["rotation_error"] = {"rx": [1., 0.], "ry": [1., 0.], "rz": [1., 0.]}
    sim = cls()
    sim._set_options(None, copy.deepcopy(options) if options else {})
    exp = NS(config=NS(n_qubits=n_qubits, memory_slots=n_qubits),
             instructions=to_reference_instructions(instrs),
             header=NS(name=name, as_dict=lambda: {"name": name}))
    return sim.run_experiment(exp)
