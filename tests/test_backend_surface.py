"""The reference-facing surface of the backend: result-dict schema,
option handling incl. the reference's sticky/leaky quirks, error behaviour, provider lookup.
Runs on the CPU emulation of the kernels (host logic is what is under test)."""
import copy

import numpy as np
import pytest

from emu_backend import emu_backend
from oracle import dm_oracle
from qiskit_aakash_b200 import BasicAer, BasicAerError, Circuit, assemble, circuits as C
from qiskit_aakash_b200.dm_simulator import DmSimulatorB200


def _run(be, circ, **opts):
    return be.run(assemble(circ), backend_options=opts).result()


def test_result_dict_schema():
    """dm_simulator.py:938-946 and :1189-1196."""
    res = _run(emu_backend(), C.qft(3))
    assert set(res) == {"backend_name", "backend_version", "qobj_id", "job_id", "results", "status", "success",
                        "time_taken", "header"}
    assert res["backend_name"] == "dm_simulator" and res["status"] == "COMPLETED" and res["success"] is True
    r = res["results"][0]
    assert set(r) == {"name", "number_of_clock_cycles", "data", "status", "success", "processing_time_taken",
                      "running_time_taken", "header"}
    assert r["status"] == "DONE" and r["name"] == "qft3"
    assert list(r["data"]) == ["ensemble_probability", "coeffmatrix", "densitymatrix"]
    assert r["data"]["coeffmatrix"].shape == (64,) and r["data"]["densitymatrix"].shape == (8, 8)
    assert list(r["data"]["ensemble_probability"])[:3] == ["000", "001", "010"]


def test_several_experiments_in_one_qobj():
    be = emu_backend()
    res = be.run(assemble([C.ghz(2), C.ghz(3)]), backend_options={}).result()
    assert [r["name"] for r in res["results"]] == ["ghz2", "ghz3"]
    assert res["results"][1]["data"]["coeffmatrix"].size == 64


def test_rotation_error_leaks_into_later_runs_like_the_reference():
    """_set_options mutates the default dict in place (dm_simulator.py:182,212-213)."""
    be = emu_backend()
    c = Circuit(1); c.h(0)
    clean = _run(be, c)["results"][0]["data"]["coeffmatrix"].copy()
    _run(be, c, rotation_error={"rz": [0.9, 0.0]})
    leaked = _run(be, c)["results"][0]["data"]["coeffmatrix"]
    ref = dm_oracle.run_oracle(1, copy.deepcopy(c.instructions), {"rotation_error": {"rz": [0.9, 0.0]}})
    assert not np.allclose(leaked, clean)
    assert np.allclose(leaked, ref["data"]["coeffmatrix"], atol=1e-15)
    fresh = _run(emu_backend(), c)["results"][0]["data"]["coeffmatrix"]
    assert np.allclose(fresh, clean, atol=1e-15)


def test_custom_densitymatrix_and_compute_flag_are_sticky():
    """dm_simulator.py:200-205 (never reset) and :247-248."""
    be = emu_backend()
    c = Circuit(2)
    d1 = _run(be, c, custom_densitymatrix="max_mixed", compute_densitymatrix=False)["results"][0]["data"]
    assert "densitymatrix" not in d1 and np.allclose(d1["coeffmatrix"], [0.25] + [0] * 15)
    d2 = _run(be, c)["results"][0]["data"]
    assert "densitymatrix" not in d2 and np.allclose(d2["coeffmatrix"], [0.25] + [0] * 15)


def test_bell_depolarization_factor_is_ineffective():
    """dm_simulator.py:240 stores it under an attribute nobody reads."""
    c = Circuit(3); c.h(0); c.cx(0, 1); c.measure(0, 0, basis="Bell", add_param="12")
    a = _run(emu_backend(), c)["results"][0]["data"]
    b = _run(emu_backend(), c, bell_depolarization_factor=0.1)["results"][0]["data"]
    assert a["bell_probabilities12"] == b["bell_probabilities12"]
    assert np.array_equal(a["coeffmatrix"], b["coeffmatrix"])


def test_bad_options_raise_basicaererror():
    be = emu_backend()
    with pytest.raises(BasicAerError):
        _run(be, Circuit(1), rotation_error=[1.0, 0.0])
    with pytest.raises(BasicAerError):
        _run(be, Circuit(1), rotation_error={"rq": [1.0, 0.0]})
    with pytest.raises(BasicAerError):
        _run(be, Circuit(1), tsp_model_error=[1.5, 0.0])
    with pytest.raises(BasicAerError):
        _run(emu_backend(), Circuit(2), custom_densitymatrix="binary_string", initial_densitymatrix="011")
    with pytest.raises(BasicAerError):
        _run(emu_backend(), Circuit(1), custom_densitymatrix="no_such_state")
    with pytest.raises(BasicAerError):      # a bare vector is rejected exactly like the reference (:344-345)
        _run(emu_backend(), Circuit(1), initial_densitymatrix=[0.5, 0, 0, 0.5])


def test_unknown_instruction_and_too_many_qubits():
    be = emu_backend()
    c = Circuit(1)
    c.instructions.append(C.instr("unitary", [0], [np.eye(2)]))
    with pytest.raises(Exception):
        _run(be, c)
    big = Circuit(DmSimulatorB200.MAX_QUBITS_MEMORY + 1)
    with pytest.raises(BasicAerError):
        _run(be, big)


def test_job_object_and_provider():
    be = emu_backend()
    job = be.run(assemble(C.ghz(2)), backend_options={})
    assert job.status() == "DONE" and isinstance(job.job_id(), str) and job.backend() is be
    assert BasicAer.get_backend("dm_simulator") is BasicAer.get_backend("dm_simulator_py")
    assert BasicAer.get_backend("dm_simulator").name() == "dm_simulator"
    assert BasicAer.get_backend().configuration().basis_gates == ["u1", "u2", "u3", "cx", "id", "unitary"]
    with pytest.raises(BasicAerError):
        BasicAer.get_backend("ibmq_qasm_simulator")
    other = BasicAer.get_backend("qasm_simulator")        # out of scope: placeholder that refuses to run
    assert other.name() == "qasm_simulator"
    with pytest.raises(BasicAerError):
        other.run(assemble(C.ghz(2)))


def test_show_final_state_flag_and_symbol_like_params():
    """params may arrive as sympy.Symbol (the reference's assembler): only str() is used."""
    class Sym:
        def __init__(self, s):
            self.s = s

        def __str__(self):
            return self.s

    c = C.ghz(3)
    c.barrier()
    c.instructions.append(C.instr("measure", [0], [Sym("Ensemble"), Sym("X")], memory=[0]))
    c.barrier()
    be = emu_backend()
    be.SHOW_FINAL_STATE = False
    d = _run(be, c)["results"][0]["data"]
    assert list(d) == ["ensemble_probability"]
    assert abs(d["ensemble_probability"]["000"] - 0.25) < 1e-15


def test_store_and_compare_round_trip(case_dir):
    """store_densitymatrix writes stored_coefficients.npy at the ensemble measure; compare reads it
    back and reports fidelity = dot(a, b) * 2^n = purity for the same state (dm_simulator.py:476-480)."""
    c = C.ghz(3); c.measure([0, 1, 2], [0, 1, 2], basis="Ensemble", add_param="Z")
    _run(emu_backend(), c, store_densitymatrix=True)
    stored = np.load("stored_coefficients.npy")
    assert stored.shape == (64,) and abs(stored[0] * 8 - 1) < 1e-15
    d = _run(emu_backend(), c, compare=True)["results"][0]["data"]
    assert abs(d["fidelity"] - 1.0) < 1e-12


def test_fidelity_of_a_compare_job_does_not_leak_into_the_next_job(tmp_path, monkeypatch):
    """The reference runs each job in a forked worker, so the fidelity computed by a 'compare' job is gone when
    the next job starts (found by tests/harness/fuzz_sessions.py against the live reference)."""
    from emu_backend import emu_backend
    from qiskit_aakash_b200 import assemble, circuits as C
    monkeypatch.chdir(tmp_path)
    be = emu_backend()
    c = C.ghz(3)
    c.measure([0, 1, 2], [0, 1, 2], basis="Ensemble", add_param="Z")
    be.run(assemble(c), backend_options={"store_densitymatrix": True}).result()
    second = be.run(assemble(c), backend_options={"compare": True, "store_densitymatrix": False}).result()["results"][0]["data"]
    assert abs(second["fidelity"] - 1.0) <= 1e-12
    plain = C.Circuit(3)
    plain.h(0)                                   # no ensemble readout -> no comparison in this job
    third = be.run(assemble(plain), backend_options={}).result()["results"][0]["data"]
    assert "fidelity" not in third


def _conditional_program():
    """Raw-qobj instructions carrying ``conditional`` (dm_simulator.py:1020-1034): an int (register bit, never set
    on this path -> skipped), a mask / val object with val 0 (runs) and with val 1 (skipped); one conditional gate
    is the FIRST of a merged run (the merged gate inherits it, basicaertools.py:193-196), one the second (lost)."""
    from types import SimpleNamespace as NS
    c = Circuit(3)
    c.u3(0.3, 0.2, 0.1, 0)
    c.u3(0.5, 0.4, 0.3, 1); c.instructions[-1].conditional = 0
    c.cx(0, 1)
    c.cx(1, 2); c.instructions[-1].conditional = 1
    c.u3(0.7, 0.1, 0.2, 2); c.instructions[-1].conditional = NS(mask="0x1", val="0x0")
    c.u3(0.2, 0.9, 0.4, 2)
    c.cx(2, 0)
    c.u3(1.1, 0.3, 0.6, 0)
    c.u3(0.4, 0.8, 0.5, 0); c.instructions[-1].conditional = 2
    c.cx(0, 2)
    c.u3(0.9, 0.2, 0.7, 1); c.instructions[-1].conditional = NS(mask="0x3", val="0x1")
    c.measure([0, 1, 2], [0, 1, 2], basis="Ensemble", add_param="Z")
    return c


def test_conditional_instructions_are_skipped_like_the_reference(case_dir):
    from oracle import ref_harness
    circ = _conditional_program()
    opts = {"rotation_error": {"rz": [0.99, 0.01]}, "decay_factor": 0.98}
    got = _run(emu_backend(), circ, **copy.deepcopy(opts))["results"][0]
    ora = dm_oracle.run_oracle(3, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    assert np.max(np.abs(got["data"]["coeffmatrix"] - ora["data"]["coeffmatrix"])) <= 1e-13
    plain = Circuit(3)
    plain.instructions = [i for i in copy.deepcopy(circ.instructions)]
    for i in plain.instructions:
        if hasattr(i, "conditional"):
            del i.conditional
    unconditional = _run(emu_backend(), plain, **copy.deepcopy(opts))["results"][0]
    assert np.max(np.abs(got["data"]["coeffmatrix"] - unconditional["data"]["coeffmatrix"])) > 1e-3
    if ref_harness.available():
        ref = ref_harness.run_reference(3, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
        assert ref["number_of_clock_cycles"] == got["number_of_clock_cycles"]
        assert np.max(np.abs(got["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"])) <= 1e-13
        assert np.max(np.abs(ora["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"])) <= 1e-13


@pytest.mark.parametrize("letters", ["ZZX", "Z"])
def test_pauli_string_of_the_wrong_length_raises_like_the_reference(letters, case_dir):
    """dm_simulator.py:553-569 indexes the string per qubit and the n-axis array with one index per letter."""
    from oracle import ref_harness
    c = C.ghz(2)
    c.measure(0, 0, basis="Expect", add_param=letters)
    with pytest.raises(IndexError):
        _run(emu_backend(), c)
    with pytest.raises(IndexError):
        dm_oracle.run_oracle(2, copy.deepcopy(c.instructions), {})
    if ref_harness.available():
        with pytest.raises(IndexError):
            ref_harness.run_reference(2, copy.deepcopy(c.instructions), {})
