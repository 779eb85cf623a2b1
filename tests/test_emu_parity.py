"""GPU-less checks of everything above the silicon: the backend (options, merge, partition,
measurement dispatch, lazy op queue, pass scheduler, readouts) driving the *same per-thread
kernel bodies* as the CUDA library, compiled for the CPU (tests/emu), against the golden
fixtures generated from the unmodified reference.  The `-m gpu` tests repeat this through
the real library on a B200."""
import numpy as np
import pytest

import cases
from emu_backend import emu_backend
from golden_check import check_against_golden
from qiskit_aakash_b200 import assemble, circuits as C


def run_case(name, backend=None):
    case = cases.get(name)
    cases.write_files(case, ".")
    circ = C.Circuit(case["n"], name)
    circ.instructions = case["instrs"]
    backend = backend or emu_backend()
    res = backend.run(assemble(circ), backend_options=case["options"]).result()
    assert res["success"] and res["status"] == "COMPLETED"
    return res["results"][0]


@pytest.mark.parametrize("name", list(cases.CASES))
def test_backend_on_emulated_kernels_matches_golden(name, golden, case_dir):
    check_against_golden(golden, name, run_case(name))


def test_mid_circuit_readouts_vs_oracle_on_emulated_kernels():
    """Same scenario as the GPU test of that name, n = 7, against the oracle."""
    import copy
    from oracle import dm_oracle
    n = 7
    circ = cases._rand_circuit(n, 40, 8011)
    circ.measure(2, 2, basis="X")
    circ.u3(0.3, 0.2, 0.1, 2); circ.cx(2, n - 1)
    circ.measure([0, 3, n - 2], [0, 3, n - 2], basis="Z")
    circ.reset(1); circ.u3(1.0, 0.5, 0.25, 1); circ.cx(1, n - 3)
    circ.measure(0, 0, basis="Bell", add_param="0%d" % (n - 1))
    circ.cx(n - 1, 0); circ.u3(0.7, 0.1, 0.9, n - 1)
    circ.measure(0, 0, basis="Expect", add_param=("ZXIY" * n)[:n])
    circ.cx(0, 1)
    circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Y")
    opts = dict(cases.FULL_NOISE)
    got = emu_backend().run(assemble(circ), backend_options=copy.deepcopy(opts)).result()["results"][0]
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    assert got["number_of_clock_cycles"] == ref["number_of_clock_cycles"] and set(got["data"]) == set(ref["data"])
    for k, v in ref["data"].items():
        a = np.array(list(v.values())) if isinstance(v, dict) else np.asarray(v)
        w = got["data"][k]
        b = np.array(list(w.values())) if isinstance(w, dict) else np.asarray(w)
        assert np.max(np.abs(a - b)) <= 1e-10, k


@pytest.mark.parametrize("tail,chunk", [(3, 2), (8, 4)])
@pytest.mark.parametrize("name", ["layered_n8_d6_noisy", "layered_n10_d3_noisy", "rand_n7_fullnoise", "qft8_binary"])
def test_streaming_drain_matches_golden(name, tail, chunk, golden, case_dir, monkeypatch):
    """The engine launches passes while later levels are still being lowered, keeping the newest
    ``drain_tail`` ops queued as look-ahead (engine.PauliEngine.drain).  Forced here with tiny
    tails so that small circuits drain many times."""
    monkeypatch.setenv("DMB_DRAIN_TAIL", str(tail))
    monkeypatch.setenv("DMB_DRAIN_CHUNK", str(chunk))
    engines = []
    from emu_backend import emu_engine
    from qiskit_aakash_b200.dm_simulator import DmSimulatorB200

    def factory(n):
        e = emu_engine(n)
        engines.append(e)
        return e

    check_against_golden(golden, name, run_case(name, DmSimulatorB200(_engine_factory=factory)))
    assert engines[0].drain_tail == tail
    assert engines[0].passes_run >= 2


def test_trailing_swaps_are_folded_into_the_store(golden, case_dir, monkeypatch):
    """The relabelling store (dmb_make_lean_pass(fold_swaps)): trailing SWAP ops of a pass become
    a digit permutation of the write-back.  Same numbers as running them as shared-memory ops
    (bit for bit: a swap moves data, it does no arithmetic), and the counter shows it happened."""
    import json
    import subprocess
    import sys
    results = {}
    for fold in ("1", "0"):
        code = ("import os, sys, json; os.environ['DMB_FOLD_SWAPS']=%r; sys.path[:0]=%r; import numpy as np; "
                "import cases, emu_backend; from qiskit_aakash_b200 import assemble, circuits as C; "
                "from qiskit_aakash_b200.dm_simulator import DmSimulatorB200; "
                "case=cases.get('layered_n10_d3_noisy'); c=C.Circuit(case['n']); c.instructions=case['instrs']; "
                "es=[]; f=lambda n: (es.append(emu_backend.emu_engine(n)) or es[-1]); "
                "r=DmSimulatorB200(_engine_factory=f).run(assemble(c), backend_options=case['options']).result(); "
                "v=r['results'][0]['data']['coeffmatrix']; "
                "print(json.dumps({'folded': es[0].stats()['folded_swaps'], 'sum': float(v.sum()).hex(), "
                "'dot': float(v @ v).hex(), 'sample': [float(x).hex() for x in v[::4099]]}))"
                % (fold, sys.path))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        results[fold] = json.loads(out.stdout.strip().splitlines()[-1])
    assert results["1"]["folded"] > 0 and results["0"]["folded"] == 0
    for key in ("sum", "dot", "sample"):
        assert results["1"][key] == results["0"][key]


@pytest.mark.parametrize("name", ["qft8", "grover4_noisy", "layered_n10_d3_noisy"])
def test_chained_ops_share_a_round_trip_and_change_nothing(name):
    """dmb_chain_ops: two consecutive CNOT-kind ops on the same ordered digit pair (the two CNOTs of a controlled-phase
    gate, ...) keep their 16-blocks in registers -- one shared-memory round trip for both; a partner that is separated
    by ops on other digits is hoisted next to it.  Same arithmetic on the same values (hoisting reorders commuting ops,
    which may move the last bit), and the counter shows it happened where the circuit has such pairs."""
    import json
    import subprocess
    import sys
    if name not in cases.CASES:
        pytest.skip("no such case")
    results = {}
    for chain in ("1", "0"):
        code = ("import os, sys, json; os.environ['DMB_CHAIN_OPS']=%r; sys.path[:0]=%r; import numpy as np; "
                "import cases, emu_backend; from qiskit_aakash_b200 import assemble, circuits as C; "
                "from qiskit_aakash_b200.dm_simulator import DmSimulatorB200; "
                "case=cases.get(%r); c=C.Circuit(case['n']); c.instructions=case['instrs']; "
                "es=[]; f=lambda n: (es.append(emu_backend.emu_engine(n)) or es[-1]); "
                "r=DmSimulatorB200(_engine_factory=f).run(assemble(c), backend_options=case['options']).result(); "
                "v=r['results'][0]['data']['coeffmatrix']; st=es[0].stats(); "
                "print(json.dumps({'chained': st['chained_ops'], 'ops': st['fused_ops'], "
                "'all': [float(x) for x in v[::(1 if v.size < 70000 else 257)]]}))"
                % (chain, sys.path, name))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        results[chain] = json.loads(out.stdout.strip().splitlines()[-1])
    assert results["0"]["chained"] == 0 and results["0"]["ops"] == results["1"]["ops"]
    if name == "qft8":                                   # every cu1 is two CNOTs on one pair
        assert results["1"]["chained"] >= 20
    a, b = np.array(results["1"]["all"]), np.array(results["0"]["all"])
    assert np.max(np.abs(a - b)) <= 1e-15


@pytest.mark.parametrize("variant", [0, 1])
def test_kernel_variants_match_golden(variant, golden, case_dir):
    """The shipped tile kernel (variant 0: one thread plays virtual threads 2u and 2u + 1, ops that leave tile digit 0
    free move both 16-blocks with 128-bit accesses -- dmb_lean_op_pair) and the generic register-staged A/B baseline
    (variant 1) run the same per-thread bodies here as on the GPU."""
    import ctypes
    from emu_backend import emu_engine, emu_lib
    lib = emu_lib()
    hook = lib.dmb_emu_paired_ops
    hook.restype = ctypes.c_long
    before = hook()
    ctx = emu_engine(6).ctx
    ctx.set_tile_variant(variant)
    try:
        for name in ("layered_n8_d6_noisy", "layered_n9_d4_memnoise", "rand_n7_fullnoise", "qft8_binary", "rand_n6_clean"):
            if name in cases.CASES:
                check_against_golden(golden, name, run_case(name))
    finally:
        ctx.set_tile_variant(0)
    assert (hook() > before) == (variant == 0)
