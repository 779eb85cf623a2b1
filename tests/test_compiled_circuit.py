"""``DmSimulatorB200.compile`` / ``CompiledCircuit.run`` (small-n latency): a job is recorded once
as the sequence of its device-facing steps and replayed without any host-side lowering.  The replay must return
what ``run`` returns -- bit for bit, same keys in the same order -- and what the reference returned (golden
fixtures); jobs that read single coefficients or files are refused (``compile`` returns None)."""
import copy

import numpy as np
import pytest

import cases
from golden_check import check_against_golden
from qiskit_aakash_b200 import assemble, circuits as C

REPLAYABLE = ["readme_x_cx", "qft5", "qft8", "qft8_binary", "grover3_noisy", "grover4_noisy", "layered_n6_d10_noisy",
              "layered_n8_d6_noisy", "rand_n2_clean", "rand_n5_nomatrix", "rand_n7_fullnoise", "rand_n3_chop",
              "measure_partial_Z", "measure_single_X", "measure_ensemble_Y", "measure_mixed_skip_quirk",
              "reset_coalesce", "barrier_levels", "init_binary_rand4", "init_thermal_state_rand3"]
REFUSED = ["measure_expect_xizy", "measure_bell_01", "measure_ensemble_N", "compare_fidelity_rand3", "init_stored_rand3",
           "rand_n1_clean"]


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "cuda":
        import __graft_entry__ as g
        g.build()
        from qiskit_aakash_b200 import DmSimulatorB200
        return DmSimulatorB200
    from emu_backend import emu_backend
    return emu_backend


def _qobj(case):
    c = C.Circuit(case["n"], "c")
    c.instructions = case["instrs"]
    return assemble(c)


def _same(a, b):
    if isinstance(a, dict):
        return list(a) == list(b) and np.array_equal(np.array(list(a.values())), np.array(list(b.values())))
    return np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("name", REPLAYABLE)
def test_replay_returns_what_run_returns(name, backend, golden, case_dir):
    case = cases.get(name)
    cases.write_files(case, ".")
    compiled = backend().compile(_qobj(case), backend_options=copy.deepcopy(case["options"]))
    assert compiled is not None
    direct = backend().run(_qobj(cases.get(name)), backend_options=copy.deepcopy(case["options"])).result()
    assert compiled.first_result["results"][0]["data"].keys() == direct["results"][0]["data"].keys()
    for _ in range(3):
        res = compiled.run()
        assert set(res) == set(direct) and res["success"] and res["status"] == "COMPLETED"
        r, d = res["results"][0], direct["results"][0]
        assert set(r) == set(d) and r["number_of_clock_cycles"] == d["number_of_clock_cycles"]
        assert list(r["data"]) == list(d["data"])
        for key in d["data"]:
            assert _same(r["data"][key], d["data"][key]), key
        check_against_golden(golden, name, r)
    assert all(e[0] in ("init", "passes", "marginal", "chop", "to_matrix", "download") for e in compiled._tape)


@pytest.mark.parametrize("name", REFUSED)
def test_jobs_that_cannot_be_replayed_are_refused(name, backend, case_dir):
    case = cases.get(name)
    cases.write_files(case, ".")
    assert backend().compile(_qobj(case), backend_options=copy.deepcopy(case["options"])) is None


def test_replays_do_not_share_buffers(backend, case_dir):
    case = cases.get("qft5")
    compiled = backend().compile(_qobj(case), backend_options={})
    a = compiled.run()["results"][0]["data"]
    b = compiled.run()["results"][0]["data"]
    a["coeffmatrix"][:] = 7.0
    assert not np.shares_memory(a["coeffmatrix"], b["coeffmatrix"]) and abs(b["coeffmatrix"][0] * 32 - 1) < 1e-12
