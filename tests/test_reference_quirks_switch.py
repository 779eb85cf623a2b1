"""``reference_quirks=False``: the reference's accidental behaviours switched off.
Expected values come from the oracle driven in a way that avoids each quirk."""
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


def _run(be, circ, opts):
    c2 = C.Circuit(circ.n_qubits)
    c2.instructions = copy.deepcopy(circ.instructions)
    return be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]


def _prep(n, seed):
    circ = cases._rand_circuit(n, 25, seed)
    return circ


_KIND = "emu"


@pytest.fixture(autouse=True, params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def _which_backend(request):
    """Every test of this module runs on the emulated kernels in the CPU suite and on the real library under
    ``-m gpu``."""
    global _KIND
    _KIND = request.param
    if _KIND == "cuda":
        import __graft_entry__ as g
        g.build()
    yield
    _KIND = "emu"


def _backend():
    if _KIND == "cuda":
        from qiskit_aakash_b200 import DmSimulatorB200
        return DmSimulatorB200()
    from emu_backend import emu_backend
    return emu_backend()


def test_mixed_basis_level_measures_every_qubit():
    n = 4
    base = _prep(n, 21)
    mixed = C.Circuit(n)
    mixed.instructions = copy.deepcopy(base.instructions)
    mixed.barrier()
    for q, b in ((0, "X"), (1, "Z"), (2, "Y"), (3, "Z")):
        mixed.measure(q, q, basis=b)
    separate = C.Circuit(n)                       # one measure per level: the reference skips nothing here
    separate.instructions = copy.deepcopy(base.instructions)
    for q, b in ((0, "X"), (1, "Z"), (2, "Y"), (3, "Z")):
        separate.barrier()
        separate.measure(q, q, basis=b)
    opts = {"depolarization_factor": 0.9, "compute_densitymatrix": False}
    want = dm_oracle.run_oracle(n, copy.deepcopy(separate.instructions), copy.deepcopy(opts))["data"]["coeffmatrix"]
    quirky = dm_oracle.run_oracle(n, copy.deepcopy(mixed.instructions), copy.deepcopy(opts))["data"]["coeffmatrix"]
    assert np.max(np.abs(want - quirky)) > 1e-3                     # the quirk is visible on this circuit
    got = _run(_backend(), mixed, dict(opts, reference_quirks=False))["data"]["coeffmatrix"]
    assert np.max(np.abs(got - want)) <= 1e-12
    got_q = _run(_backend(), mixed, opts)["data"]["coeffmatrix"]     # default: the reference, quirk included
    assert np.max(np.abs(got_q - quirky)) <= 1e-12


def test_bell_measure_acts_on_the_named_qubits_and_depolarizes():
    n, a, b, f = 4, 0, 2, 0.8
    base = _prep(n, 22)
    named = C.Circuit(n)
    named.instructions = copy.deepcopy(base.instructions)
    named.measure(0, 0, basis="Bell", add_param="%d%d" % (a, b))
    mirrored = C.Circuit(n)                        # what the reference needs to be told to hit qubits a, b
    mirrored.instructions = copy.deepcopy(base.instructions)
    mirrored.measure(0, 0, basis="Bell", add_param="%d%d" % (n - 1 - b, n - 1 - a))
    opts = {"compute_densitymatrix": False}
    ref = dm_oracle.run_oracle(n, copy.deepcopy(mirrored.instructions), copy.deepcopy(opts))["data"]
    got = _run(_backend(), named, dict(opts, reference_quirks=False))["data"]
    key_ref, key_got = "%d%d" % (n - 1 - b, n - 1 - a), "%d%d" % (a, b)
    assert np.max(np.abs(got["coeffmatrix"] - ref["coeffmatrix"])) <= 1e-12
    for k, v in ref["bell_probabilities" + key_ref].items():
        assert abs(got["bell_probabilities" + key_got][k] - v) <= 1e-12
    assert np.max(np.abs(got["reduced_bell_densitymatrix" + key_got] - ref["reduced_bell_densitymatrix" + key_ref])) <= 1e-12
    # bell_depolarization_factor is honoured: the i == j != 0 terms of the pair are scaled by f
    dep = _run(_backend(), named, dict(opts, reference_quirks=False, bell_depolarization_factor=f))["data"]
    t = ref["coeffmatrix"].reshape([4] * n).copy()
    idx = [slice(None)] * n
    for i in (1, 2, 3):
        idx[a], idx[b] = i, i
        t[tuple(idx)] *= f
    assert np.max(np.abs(dep["coeffmatrix"] - t.reshape(-1))) <= 1e-12
    same = _run(_backend(), named, dict(opts, bell_depolarization_factor=f))["data"]      # reference: ignored
    ref_named = dm_oracle.run_oracle(n, copy.deepcopy(named.instructions), dict(opts, bell_depolarization_factor=f))["data"]
    assert np.max(np.abs(same["coeffmatrix"] - ref_named["coeffmatrix"])) <= 1e-12


def test_options_do_not_stick_between_runs():
    n = 3
    circ = C.ghz(n)
    be = _backend()
    first = _run(be, circ, {"custom_densitymatrix": "max_mixed", "compute_densitymatrix": False,
                            "rotation_error": {"rz": [0.9, 0.0]}})
    assert "densitymatrix" not in first["data"]
    sticky = _run(be, circ, {})                                      # the reference keeps all three settings
    assert "densitymatrix" not in sticky["data"]
    assert np.max(np.abs(sticky["data"]["coeffmatrix"] - first["data"]["coeffmatrix"])) <= 1e-15
    clean = _run(be, circ, {"reference_quirks": False})
    want = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), {})["data"]
    assert np.max(np.abs(clean["data"]["coeffmatrix"] - want["coeffmatrix"])) <= 1e-12
    assert np.max(np.abs(clean["data"]["densitymatrix"] - want["densitymatrix"])) <= 1e-12
