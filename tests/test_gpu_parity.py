"""Parity tests proper (``-m gpu``): the CUDA library on a B200, called through the C ABI,
against (1) the golden fixtures generated from the unmodified reference, (2) the oracle on
seeded circuits at sizes it finishes in seconds, (3) size-independent properties at the
BASELINE sizes (n = 14: trace, normalisation, schedule independence, U.U^-1 round trip,
GHZ known answer).  Tolerances are BASELINE.json's: max |delta| <= 1e-10 on Pauli
coefficients and probabilities (float64), trace preserved to 1e-12."""
import copy

import numpy as np
import pytest

import cases
from golden_check import TOL, TRACE_TOL, check_against_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import __graft_entry__ as g
    g.build()
    from qiskit_aakash_b200 import DmSimulatorB200
    return lambda: DmSimulatorB200()


def _run(backend, n, instrs, options, name="c"):
    from qiskit_aakash_b200 import assemble, circuits as C
    circ = C.Circuit(n, name)
    circ.instructions = instrs
    res = backend().run(assemble(circ), backend_options=options).result()
    assert res["success"]
    return res["results"][0]


def test_cuda_library_is_the_one_loaded(backend):
    """The product path must be the in-tree CUDA library (no silent fallback)."""
    from qiskit_aakash_b200 import capi, engine
    e = engine.PauliEngine(3)
    assert e.lib._name == capi.DEFAULT_LIB
    e.init_product([[1, 0, 0, 1]] * 3, 0.125)
    assert np.allclose(e.download()[[0, 3, 63]], 0.125)
    assert e.stats()["other_launches"] >= 1


@pytest.mark.parametrize("name", list(cases.CASES))
def test_cuda_backend_matches_golden(name, golden, backend, case_dir):
    case = cases.get(name)
    cases.write_files(case, ".")
    check_against_golden(golden, name, _run(backend, case["n"], case["instrs"], case["options"], name))


@pytest.mark.parametrize("name", ["qft8", "qft8_binary", "rand_n7_fullnoise", "rand_n6_fullnoise", "rand_n4_nomerge",
                                  "layered_n8_d6_noisy"])
def test_small_states_run_their_pass_list_in_one_launch(name, golden, backend, case_dir):
    """(f)4: at most 4^8 coefficients -> all passes of a flush are ONE launch (k_small_passes6: the tiles' CTAs are a
    thread-block cluster that meets at the cluster barrier between passes; k_small_passes<K> below one tile), and
    the results are the reference's (golden fixtures; QFT-8 is BASELINE configs[0])."""
    case = cases.get(name)
    cases.write_files(case, ".")
    be = backend()
    res = be.run(__import__("qiskit_aakash_b200").assemble(_circ(case)), backend_options=case["options"]).result()
    check_against_golden(golden, name, res["results"][0])
    st = be.last_engine_stats
    if st["passes"] > st["tile_pass_launches"]:        # some flush had more than one pass: it went out as one launch
        assert st["small_plan_launches"] >= 1
    else:                                              # every flush of this circuit was a single pass
        assert st["small_plan_launches"] == 0


def _circ(case):
    from qiskit_aakash_b200 import circuits as C
    c = C.Circuit(case["n"], "c")
    c.instructions = case["instrs"]
    return c


def _compare_with_oracle(backend, n, instrs, options):
    from oracle import dm_oracle
    got = _run(backend, n, copy.deepcopy(instrs), copy.deepcopy(options))
    ref = dm_oracle.run_oracle(n, copy.deepcopy(instrs), copy.deepcopy(options))
    assert got["number_of_clock_cycles"] == ref["number_of_clock_cycles"]
    assert set(got["data"]) == set(ref["data"])
    worst = 0.0
    for k, v in ref["data"].items():
        a = np.array(list(v.values())) if isinstance(v, dict) else np.asarray(v)
        w = got["data"][k]
        b = np.array(list(w.values())) if isinstance(w, dict) else np.asarray(w)
        worst = max(worst, float(np.max(np.abs(a - b))))
    assert worst <= TOL, worst
    assert abs(got["data"]["coeffmatrix"][0] * 2 ** n - 1) <= TRACE_TOL
    return worst


@pytest.mark.parametrize("n,seed", [(6, 1), (7, 2), (8, 3), (9, 4), (10, 5), (11, 6)])
def test_random_circuits_vs_oracle(backend, n, seed):
    circ = cases._rand_circuit(n, 60, 7000 + seed)
    circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="XYZ"[seed % 3])
    opts = dict(cases.FULL_NOISE, compute_densitymatrix=(n <= 8))
    _compare_with_oracle(backend, n, circ.instructions, opts)


@pytest.mark.parametrize("n,seed", [(8, 11), (9, 12), (10, 13)])
def test_mid_circuit_measure_reset_bell_expect_vs_oracle(backend, n, seed):
    """Every state-changing readout in the middle of a noisy circuit, then more gates (the lazy
    queue is flushed by each readout; layouts are relabelled in between)."""
    circ = cases._rand_circuit(n, 40, 8000 + seed)
    circ.measure(2, 2, basis="X")                       # single projective measurement
    circ.u3(0.3, 0.2, 0.1, 2); circ.cx(2, n - 1)
    circ.measure([0, 3, n - 2], [0, 3, n - 2], basis="Z")    # partial measurement of three qubits
    circ.reset(1); circ.u3(1.0, 0.5, 0.25, 1); circ.cx(1, n - 3)
    circ.measure(0, 0, basis="Bell", add_param="0%d" % (n - 1))
    circ.cx(n - 1, 0); circ.u3(0.7, 0.1, 0.9, n - 1)
    circ.measure(0, 0, basis="Expect", add_param=("ZXIY" * n)[:n])
    circ.cx(0, 1)
    circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Y")
    opts = dict(cases.FULL_NOISE, compute_densitymatrix=(n <= 8))
    _compare_with_oracle(backend, n, circ.instructions, opts)


def test_baseline_configs_at_oracle_sizes(backend):
    from qiskit_aakash_b200 import circuits as C
    _compare_with_oracle(backend, 10, C.qft(10).instructions, {"compute_densitymatrix": False})
    _compare_with_oracle(backend, 11, C.random_layered(11, 6, 1100).instructions,
                         dict(C.noisy_options(), compute_densitymatrix=False))
    g = C.grover(5, "10110", 1)                       # 5 search + 3 ancilla qubits
    _compare_with_oracle(backend, g.n_qubits, g.instructions, dict(C.grover_options(), compute_densitymatrix=False))


@pytest.mark.parametrize("n,noisy", [(10, False), (11, False), (9, True)])
def test_chained_cnots_of_controlled_phase_gates_on_cuda(backend, n, noisy):
    """QFT above one tile: the two CNOTs of every cu1 act on one ordered digit pair and share a shared-memory round trip
    in k_tile_pass6 (DMB_CHAIN, dmb_stats.chained_ops); ideal CNOTs and, with the noise options, TSP CNOTs with full
    maps.  Against the oracle."""
    from oracle import dm_oracle
    from qiskit_aakash_b200 import assemble, circuits as C
    circ = C.qft(n)
    opts = dict(C.noisy_options() if noisy else {}, compute_densitymatrix=False)
    if noisy:
        opts.update(C.grover_options())
    be = backend()
    c2 = C.Circuit(n)
    c2.instructions = copy.deepcopy(circ.instructions)
    got = be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]
    st = be.last_engine_stats
    assert st["chained_ops"] >= n * (n - 1) // 4, st          # at least half of the n (n - 1) / 2 controlled phases
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    assert np.max(np.abs(got["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"])) <= TOL
    p = np.array(list(got["data"]["ensemble_probability"].values()))
    q = np.array(list(ref["data"]["ensemble_probability"].values()))
    assert np.max(np.abs(p - q)) <= TOL and abs(got["data"]["coeffmatrix"][0] * 2 ** n - 1) <= TRACE_TOL


@pytest.mark.parametrize("name", ["grover12_noisy", "layered_n12_d6_noisy"])
def test_full_circuits_at_n12_match_the_reference(backend, name):
    """north_star: results within 1e-10 of the reference at n <= 12.  FULL circuits at n = 12 -- Grover-12
    (BASELINE configs[1]: 315 instructions, 176 noisy levels, partial Z readout) and the layered U3+CX circuit of
    configs[2] -- against what the UNMODIFIED reference returned for them (tests/golden/golden_n12.npz, written by
    tests/golden/make_golden_n12.py: strided sample of the 4^12 coefficients, sum / sum of squares / signed
    checksum of all of them, every probability)."""
    import cases_n12
    case = cases_n12.get(name)
    res = _run(backend, case["n"], case["instrs"], case["options"], name)
    worst = cases_n12.check(cases_n12.load_golden(), name, res)
    assert worst <= TOL


# ---- every op kind on every digit pair, engine level, against NumPy on a random state ------

@pytest.mark.parametrize("n", [2, 3, 5, 6, 8])
def test_each_op_kind_on_every_digit_pair(backend, n):
    from oracle import dm_oracle
    from qiskit_aakash_b200 import engine
    rng = np.random.default_rng(n)
    vec = rng.normal(size=4 ** n)
    vec[0] = 0.5 ** n
    pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
    if len(pairs) > 14:
        pairs = [pairs[i] for i in rng.choice(len(pairs), size=14, replace=False)]
    for tsp in ([1.0, 0.0], [0.97, 0.05]):
        e = engine.PauliEngine(n, max_ops_per_pass=1)
        e.upload(vec)
        o = dm_oracle.OracleSim({"tsp_model_error": tsp, "rotation_error": {"rz": [0.99, 0.01], "ry": [0.98, 0.02]},
                                 "decay_factor": 0.9, "thermal_factor": 0.3, "decoherence_factor": 0.95})
        o.n, o.dm = n, vec.copy()
        for (a, b) in pairs:
            ang = rng.uniform(0, 6, 3)
            e.apply_1q(a, engine.gate_matrix("u3", ang, o.rotation_error))
            o.single_gate("u3", ang, a)
            e.apply_cx(a, b, tsp)
            o.cx(a, b)
            e.apply_1q_all(engine.memory_noise_matrix(o.f, o.p, o.g))
            o.memory_noise()
        got = e.download()
        assert np.max(np.abs(got - o.dm)) <= 1e-12 * max(1.0, np.max(np.abs(vec)))


# ---- size-independent properties at the BASELINE size (n = 14, 2 GiB state) ------------------

def test_n14_ghz_known_answer(backend):
    from qiskit_aakash_b200 import circuits as C
    c = C.ghz(14)
    c.measure(list(range(14)), list(range(14)), basis="Ensemble", add_param="Z")
    res = _run(backend, 14, c.instructions, {"compute_densitymatrix": False})
    p = res["data"]["ensemble_probability"]
    assert abs(p["0" * 14] - 0.5) <= TOL and abs(p["1" * 14] - 0.5) <= TOL
    assert abs(sum(p.values()) - 1) <= TOL
    vec = res["data"]["coeffmatrix"]
    assert abs(vec[0] * 2 ** 14 - 1) <= TRACE_TOL
    assert abs(np.dot(vec, vec) * 2 ** 14 - 1) <= 1e-9          # purity of a pure state


def test_n14_unitary_round_trip(backend):
    """U then U^-1 (noise free) must return to |0..0>: coefficients 2^-n on the I/Z strings."""
    from qiskit_aakash_b200 import circuits as C
    fwd = C.random_layered(14, 6, 42, readout=False)
    c = C.Circuit(14)
    c.instructions = list(fwd.instructions)
    for ins in reversed(fwd.instructions):
        if ins.name == "cx":
            c.cx(ins.qubits[0], ins.qubits[1])
        else:
            th, ph, lam = ins.params
            c.barrier()                                   # keep U and U^-1 from being merged away
            c.u3(-th, -lam, -ph, ins.qubits[0])
    c.measure(list(range(14)), list(range(14)), basis="Ensemble", add_param="Z")
    res = _run(backend, 14, c.instructions, {"compute_densitymatrix": False})
    p = res["data"]["ensemble_probability"]
    assert abs(p["0" * 14] - 1.0) <= 1e-9
    vec = res["data"]["coeffmatrix"].reshape([4] * 14)
    assert abs(vec[(3,) * 14] * 2 ** 14 - 1) <= 1e-9 and abs(vec[(1,) + (0,) * 13]) <= 1e-12


def test_n14_noisy_invariants_and_schedule_independence(backend):
    """Config 3 shape (n = 14, per-gate noise): trace, normalisation, and the same numbers
    from the fused schedule and from the one-op-per-pass schedule."""
    from qiskit_aakash_b200 import DmSimulatorB200, assemble, circuits as C, engine
    circ = C.random_layered(14, 8, 1400)
    opts = dict(C.noisy_options(), compute_densitymatrix=False, **C.grover_options())
    outs = []
    for max_ops, reserve, relabel in ((10, 2, True), (1, 1, False)):
        be = DmSimulatorB200(_engine_factory=lambda n: engine.PauliEngine(n, max_ops_per_pass=max_ops, reserve_low=reserve,
                                                                          relabel=relabel))
        c2 = C.Circuit(14); c2.instructions = copy.deepcopy(circ.instructions)
        r = be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]
        outs.append((r["data"]["coeffmatrix"].copy(), np.array(list(r["data"]["ensemble_probability"].values()))))
        del r
    (v1, p1), (v2, p2) = outs
    assert abs(v1[0] * 2 ** 14 - 1) <= TRACE_TOL
    assert abs(p1.sum() - 1) <= 1e-10 and p1.min() >= -1e-12
    assert np.max(np.abs(v1 - v2)) <= 1e-12 and np.max(np.abs(p1 - p2)) <= 1e-12


def test_n16_round_trip_at_the_single_gpu_limit():
    """The largest register one B200 advertises (n = 16: 2^32 coefficients, 34 GB -- byte offsets beyond 32 bits in every
    kernel): GHZ known answer and U . U^-1 = 1 through the public backend, probabilities only (``SHOW_FINAL_STATE =
    False``, a reference flag, ``dm_simulator.py:134``: the 34 GB vector stays on the device), plus single coefficients
    read through ``dmb_read_coeffs``."""
    import torch
    from qiskit_aakash_b200 import DmSimulatorB200, assemble, circuits as C, engine
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs 34 GB of free device memory")
    n = 16
    engines = []

    def factory(nq):
        engines.append(engine.PauliEngine(nq))
        return engines[-1]

    be = DmSimulatorB200(_engine_factory=factory)
    be.SHOW_FINAL_STATE = False
    ghz = C.ghz(n)
    ghz.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Z")
    res = be.run(assemble(ghz), backend_options={"compute_densitymatrix": False}).result()["results"][0]
    p = res["data"]["ensemble_probability"]
    assert abs(p["0" * n] - 0.5) <= TOL and abs(p["1" * n] - 0.5) <= TOL and abs(sum(p.values()) - 1) <= TOL
    e = engines[-1]
    e.flush()
    pos = e.pos                                             # digit position of qubit q in the current layout

    def index(digits):                                      # Pauli string (digit of qubit q) -> vector index
        return sum(int(d) << (2 * pos[q]) for q, d in enumerate(digits))

    got = e.ctx.read_coeffs(e.sptr, [index([0] * n), index([3] * n), index([1] * n), index([3, 3] + [0] * (n - 2)),
                                     index([1] + [0] * (n - 1))])
    # GHZ: <I..I> = <Z..Z (even n)> = <X..X> = <Z Z I..I> = 1, <X I..I> = 0; coefficients carry 2^-n
    assert np.max(np.abs(got * 2 ** n - np.array([1, 1, 1, 1, 0]))) <= 1e-12
    del e
    engines.clear()

    fwd = C.random_layered(n, 3, 1600, readout=False)
    c = C.Circuit(n)
    c.instructions = list(fwd.instructions)
    for ins in reversed(fwd.instructions):
        if ins.name == "cx":
            c.cx(ins.qubits[0], ins.qubits[1])
        else:
            th, ph, lam = ins.params
            c.barrier()
            c.u3(-th, -lam, -ph, ins.qubits[0])
    c.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Z")
    res = be.run(assemble(c), backend_options={"compute_densitymatrix": False}).result()["results"][0]
    p = res["data"]["ensemble_probability"]
    assert abs(p["0" * n] - 1.0) <= 1e-9 and abs(sum(p.values()) - 1) <= 1e-9
    assert "coeffmatrix" not in res["data"]


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


@pytest.mark.parametrize("name", list(cases.FRONTEND_CASES))
def test_facade_on_cuda_matches_full_reference_stack(name, golden, backend, case_dir):
    """User-level source (tests/frontend_cases.py) through frontend.QuantumCircuit + execute on the
    GPU vs the reference's real front-end + real simulator (golden ``fe_<name>``)."""
    import frontend_cases
    from types import SimpleNamespace
    from qiskit_aakash_b200 import execute, frontend
    api = SimpleNamespace(QuantumCircuit=frontend.QuantumCircuit, QuantumRegister=frontend.QuantumRegister,
                          ClassicalRegister=frontend.ClassicalRegister, pi=frontend.pi)
    qc = frontend_cases.CASES[name](api)
    res = execute(qc, backend(), **copy.deepcopy(cases.FRONTEND_CASES[name])).result()
    check_against_golden(golden, "fe_" + name, res["results"][0])


def test_relabelling_store_is_an_exact_digit_permutation_on_cuda(backend):
    """Swap-only passes through libdmb200.so (k_tile_pass6<..., PERM128 / SPLIT64>) == NumPy axis
    swaps, bit for bit; all swaps folded into the write-back."""
    import store_cases
    from qiskit_aakash_b200 import engine
    folded, total = store_cases.check_relabelling_store(lambda n: engine.PauliEngine(n))
    assert folded == total
    folded, total = store_cases.check_relabelling_store(lambda n: engine.PauliEngine(n), n=9,
                                                        tile_digits=(0, 1, 2, 5, 7, 8))
    assert folded == total


def test_random_programs_vs_oracle_on_cuda(backend):
    """tests/harness/fuzz_emu.py's generator (every measurement mode, resets, barriers, random option sets and
    initial states) on the real kernels: 160 programs with n <= 8, 12 long ones at n = 8..10, and 40
    with randomised scheduler / streaming knobs."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "harness"))
    import fuzz_emu
    bad = []
    plan = [(s, 8, 1, 60, False) for s in range(100000, 100160)] + \
           [(s, 10, 8, 250, False) for s in range(101000, 101012)] + \
           [(s, 9, 2, 150, True) for s in range(102000, 102040)]
    try:
        for seed, max_n, min_n, max_ops, knobs in plan:
            status, msg = fuzz_emu.one(seed, max_n, min_n, max_ops, knobs, backend=backend)
            if status == "FAIL":
                bad.append((seed, msg))
    finally:
        for k in fuzz_emu.KNOBS:
            os.environ.pop(k, None)
    assert not bad, bad[:5]


def _extra_variants():
    import os
    return [int(x) for x in os.environ.get("DMB_TEST_TILE_VARIANT", "").split(",") if x]


@pytest.mark.parametrize("variant", [1] + _extra_variants())
def test_tile_variant_parity(variant, backend, golden, case_dir, monkeypatch):
    """Parity of the non-default tile-kernel variants that ship in the library (1 = the generic register-staged
    A/B baseline; DMB_TEST_TILE_VARIANT adds experimental ones while they exist): the golden cases that
    reach the K = 6 kernel, 60 random programs with n >= 6, and the n = 14 round trip."""
    import os
    import sys
    monkeypatch.setenv("DMB_TILE_VARIANT", str(variant))
    try:
        for name in cases.CASES:
            case = cases.get(name)
            if case["n"] >= 6:
                cases.write_files(case, ".")
                check_against_golden(golden, name, _run(backend, case["n"], case["instrs"], case["options"], name))
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "harness"))
        import fuzz_emu
        bad = [(s, m) for s in range(103000, 103060) for st, m in [fuzz_emu.one(s, 10, 6, 200, False, backend=backend)]
               if st == "FAIL"]
        assert not bad, bad[:5]
        test_n14_unitary_round_trip(backend)
    finally:
        from qiskit_aakash_b200 import engine
        for ctx in engine._CONTEXTS.values():          # contexts are shared: put the default kernel back
            ctx.set_tile_variant(0)
