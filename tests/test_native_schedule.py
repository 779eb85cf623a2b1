"""The native gate-fusion scheduler (``dmb_schedule``, csrc/dm_schedule.h) against its executable
specification ``schedule.build_passes_relabel``: byte-identical ``dmb_pass`` arrays, layouts,
leftover ops and leftover moves -- on the real CUDA library (host-only entry point, loads without a
GPU) and on the emulation library."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from qiskit_aakash_b200 import capi, circuits as C, schedule  # noqa: E402


def _libs():
    from emu_backend import emu_lib
    return [("cuda", capi.load_library()), ("emu", emu_lib())]


def _rand_map(rng):
    m = np.zeros((4, 4))
    m[0, 0] = 1.0
    m[1:, :] = rng.standard_normal((3, 4))
    return m


def _rand_stream(rng, n_qubits, n_ops, local=None, brick=False):
    """Random two-qubit op stream on qubit ids (``local``: the qubits ops may touch)."""
    qs = list(range(n_qubits)) if local is None else list(local)
    ops = []
    for i in range(n_ops):
        if brick:
            a = qs[(2 * (i % (len(qs) // 2)) + (i // (len(qs) // 2)) % 2) % (len(qs) - 1)]
            b = qs[qs.index(a) + 1]
            if rng.integers(2):
                a, b = b, a
        else:
            a, b = (int(x) for x in rng.choice(qs, size=2, replace=False))
        kind = int(rng.choice([capi.OP_CX, capi.OP_CX_TSP, capi.OP_DIAG2, capi.OP_MATS]))
        pa = _rand_map(rng) if rng.random() < 0.7 else None
        pb = _rand_map(rng) if rng.random() < 0.7 else None
        coef = None
        if kind == capi.OP_CX_TSP:
            coef = rng.standard_normal(5)
        elif kind == capi.OP_DIAG2:
            coef = rng.standard_normal(16)
        if kind == capi.OP_MATS and rng.random() < 0.3:
            ops.append(schedule.DevOp(kind, a, None, pa if pa is not None else _rand_map(rng), None))
        else:
            ops.append(schedule.DevOp(kind, a, b, pa, pb, coef))
    return ops


@pytest.mark.parametrize("seed", range(12))
def test_native_scheduler_is_byte_identical_to_the_python_specification(seed):
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(2, 17))
    nd = max(n, 2)
    ops = _rand_stream(rng, n, int(rng.integers(1, 260)), brick=(seed % 3 == 0 and n >= 4))
    max_ops = int(rng.choice([4, 8, 10, 12, 16]))
    pos0 = [int(x) for x in rng.permutation(n)]
    want_pos = list(pos0)
    want = schedule.build_passes_relabel(list(ops), want_pos, nd, max_ops=max_ops)
    for name, lib in _libs():
        pos = list(pos0)
        got = schedule.relabel_passes(lib, list(ops), pos, nd, max_ops=max_ops, native=True)
        assert pos == want_pos, name
        assert got.dtype == want.dtype and len(got) == len(want), name
        assert got.tobytes() == want.tobytes(), name


@pytest.mark.parametrize("seed", range(6))
def test_native_scheduler_streaming_tail_matches(seed):
    """min_tail: same passes, same leftover ops (by identity), same layout."""
    rng = np.random.default_rng(9100 + seed)
    n = int(rng.integers(4, 15))
    ops = _rand_stream(rng, n, 300, brick=bool(seed % 2))
    tail = int(rng.choice([16, 64, 100]))
    for name, lib in _libs():
        pos_a, pos_b = [n - 1 - q for q in range(n)], [n - 1 - q for q in range(n)]
        qa, qb = [], []
        for i in range(0, len(ops), 50):
            qa += ops[i:i + 50]
            qb += ops[i:i + 50]
            pa, qa = schedule.build_passes_relabel(qa, pos_a, n, max_ops=10, min_tail=tail)
            pb, qb = schedule.relabel_passes(lib, qb, pos_b, n, max_ops=10, min_tail=tail, native=True)
            assert pa.tobytes() == pb.tobytes() and pos_a == pos_b, name
            assert len(qa) == len(qb) and all(x is y for x, y in zip(qa, qb)), name
        fa = schedule.build_passes_relabel(qa, pos_a, n, max_ops=10)
        fb = schedule.relabel_passes(lib, qb, pos_b, n, max_ops=10, native=True)
        assert fa.tobytes() == fb.tobytes() and pos_a == pos_b, name


@pytest.mark.parametrize("seed", range(8))
def test_native_scheduler_final_moves_match(seed):
    """Sharded use: ops on local qubits only, global qubits at positions >= n_digits, evictees
    parked in the top local slots by trailing swaps where the last tile has room."""
    rng = np.random.default_rng(9200 + seed)
    n = int(rng.integers(7, 17))
    m = int(rng.integers(1, 3))
    n_loc = n - m
    pos0 = [int(x) for x in rng.permutation(n)]
    local = [q for q in range(n) if pos0[q] < n_loc]
    ops = _rand_stream(rng, n, int(rng.integers(1, 120)), local=local)
    victims = [int(x) for x in rng.choice(local, size=m, replace=False)]
    moves = [(v, n_loc - m + i) for i, v in enumerate(victims)]
    if seed == 7:
        moves = []
    want_pos = list(pos0)
    want, want_left = schedule.build_passes_relabel(list(ops), want_pos, n_loc, max_ops=10,
                                                    final_moves=list(moves))
    for name, lib in _libs():
        pos = list(pos0)
        got, left = schedule.relabel_passes(lib, list(ops), pos, n_loc, max_ops=10, final_moves=list(moves), native=True)
        assert got.tobytes() == want.tobytes() and pos == want_pos, name
        assert [tuple(x) for x in left] == [tuple(x) for x in want_left], name


def test_native_scheduler_on_the_benchmark_circuit_shape():
    """BASELINE configs[2] shape (n = 14 brick-wall): identical schedule, two orders of magnitude faster."""
    import time
    n, depth = 14, 40
    circ = C.random_layered(n, depth, 1400)
    rng = np.random.default_rng(1)
    ops = [schedule.DevOp(capi.OP_CX_TSP, ins.qubits[0], ins.qubits[1], _rand_map(rng), _rand_map(rng),
                          rng.standard_normal(5))
           for ins in circ.instructions if ins.name == "cx"]
    pos_a, pos_b = [n - 1 - q for q in range(n)], [n - 1 - q for q in range(n)]
    t0 = time.perf_counter()
    want = schedule.build_passes_relabel(list(ops), pos_a, n, max_ops=10)
    t1 = time.perf_counter()
    got = schedule.relabel_passes(capi.load_library(), list(ops), pos_b, n, max_ops=10, native=True)
    t2 = time.perf_counter()
    assert got.tobytes() == want.tobytes() and pos_a == pos_b
    assert (t2 - t1) < (t1 - t0)


def test_native_scheduler_reports_errors():
    lib = capi.load_library()
    op = schedule.DevOp(capi.OP_CX, 0, 1)
    with pytest.raises(capi.DmbError):
        schedule.relabel_passes(lib, [op], [0, 0], 2, native=True)           # pos not injective
    with pytest.raises(capi.DmbError):
        schedule.relabel_passes(lib, [op], [0, 5], 3, native=True)           # op on a non-local qubit
    with pytest.raises(ValueError):
        bad = np.eye(4)
        bad[0, 1] = 0.5
        schedule.relabel_passes(lib, [schedule.DevOp(capi.OP_CX, 0, 1, bad)], [0, 1], 2, native=True)


# ---- tile-search strategy (native only): validity by symbolic replay, and schedule quality -------

def _replay(passes, qops, pos0, pos_end, nd, max_ops=16):
    """Every op lands on the digit positions its qubits occupy at that moment (SWAPs move qubits),
    per-qubit program order is kept, every op runs exactly once, the returned layout is right."""
    sim = {p: q for q, p in enumerate(pos0)}
    last, seen = {}, []
    for p in passes:
        K = int(p["n_tile_digits"])
        tile = [int(x) for x in p["tile_digit"][:K]]
        assert tile[0] == 0 and tile[1] == 1 and tile == sorted(set(tile)) and K == min(6, nd)
        assert 1 <= p["n_ops"] <= max_ops
        for o in p["ops"][:p["n_ops"]]:
            da, db = tile[o["a"]], tile[o["b"]]
            assert sorted([int(o["a"]), int(o["b"])] + [int(x) for x in o["fd"][:K - 2]]) == list(range(K))
            if o["kind"] == capi.OP_SWAP and o["flags"] == 0 and o["coef"][0] == 0:
                sim[da], sim[db] = sim.get(db), sim.get(da)
                continue
            tag = int(o["coef"][0]) - 1
            src = qops[tag]
            assert sim[da] == src.da and (src.db is None or sim[db] == src.db)
            for q in src.digits():
                assert last.get(q, -1) < tag, "ops sharing a qubit were reordered"
                last[q] = tag
            seen.append(tag)
    assert sorted(seen) == list(range(len(qops)))
    for q, p in enumerate(pos_end):
        assert sim[p] == q


@pytest.mark.parametrize("n_qubits,max_ops,brick", [(2, 16, False), (3, 4, False), (5, 6, False), (7, 10, True),
                                                    (14, 16, True), (14, 16, False), (16, 12, False), (11, 3, True)])
def test_tile_search_schedule_is_a_valid_reordering(n_qubits, max_ops, brick):
    rng = np.random.default_rng(n_qubits * 100 + max_ops)
    qops = _rand_stream(rng, n_qubits, 220, brick=brick)
    for k, op in enumerate(qops):
        op.kind, op.coef = (capi.OP_MATS if op.db is None else capi.OP_CX), [float(k + 1)]     # tag = position + 1
    nd = max(n_qubits, 2)
    pos0 = [int(x) for x in rng.permutation(n_qubits)]
    for name, lib in _libs():
        pos = list(pos0)
        passes = schedule.relabel_passes(lib, qops, pos, nd, max_ops=max_ops, native=True,
                                         strategy=capi.SCHED_TILE_SEARCH)
        _replay(passes, qops, pos0, pos, nd, max_ops)
    # streaming: cut anywhere, keep a tail -> still valid, every op exactly once
    pos, queue, parts = list(pos0), [], []
    for i in range(0, len(qops), 37):
        queue += qops[i:i + 37]
        p, queue = schedule.relabel_passes(capi.load_library(), queue, pos, nd, max_ops=max_ops, min_tail=20,
                                           native=True, strategy=capi.SCHED_TILE_SEARCH)
        parts.append(p)
    parts.append(schedule.relabel_passes(capi.load_library(), queue, pos, nd, max_ops=max_ops, native=True,
                                         strategy=capi.SCHED_TILE_SEARCH))
    _replay(np.concatenate(parts), qops, pos0, pos, nd, max_ops)


def test_tile_search_needs_fewer_passes_on_the_benchmark_shape():
    """BASELINE configs[2] (n = 14 brick-wall, depth 200): the tile search finds time-skewed windows
    -- ~130 passes of ~10 CNOTs where program-order list scheduling needs ~200 of 6.5."""
    n = 14
    circ = C.random_layered(n, 200, 1400)
    qops = [schedule.DevOp(capi.OP_CX, i.qubits[0], i.qubits[1]) for i in circ.instructions if i.name == "cx"]
    lib = capi.load_library()
    counts = {}
    for strategy, cap in ((capi.SCHED_PROGRAM_ORDER, 10), (capi.SCHED_PROGRAM_ORDER, 16), (capi.SCHED_TILE_SEARCH, 16)):
        pos = [n - 1 - q for q in range(n)]
        counts[strategy, cap] = len(schedule.relabel_passes(lib, qops, pos, n, max_ops=cap, native=True, strategy=strategy))
    assert counts[capi.SCHED_TILE_SEARCH, 16] <= 140
    assert counts[capi.SCHED_TILE_SEARCH, 16] < 0.7 * min(counts[capi.SCHED_PROGRAM_ORDER, 10],
                                                            counts[capi.SCHED_PROGRAM_ORDER, 16])


def test_sharded_final_moves_with_tile_search():
    """final_moves under the tile-search strategy: leftover moves are exactly those not at their target."""
    rng = np.random.default_rng(77)
    n, m = 12, 2
    n_loc = n - m
    pos0 = [int(x) for x in rng.permutation(n)]
    local = [q for q in range(n) if pos0[q] < n_loc]
    qops = _rand_stream(rng, n, 60, local=local)
    for k, op in enumerate(qops):
        op.kind, op.coef = (capi.OP_MATS if op.db is None else capi.OP_CX), [float(k + 1)]
    victims = [int(x) for x in rng.choice(local, size=m, replace=False)]
    moves = [(v, n_loc - m + i) for i, v in enumerate(victims)]
    pos = list(pos0)
    passes, left = schedule.relabel_passes(capi.load_library(), qops, pos, n_loc, max_ops=16, final_moves=moves,
                                           native=True, strategy=capi.SCHED_TILE_SEARCH)
    _replay(passes, qops, pos0, pos, n_loc)
    assert [tuple(x) for x in left] == [(q, t) for q, t in moves if pos[q] != t]


def test_random_streams_caps_and_cut_points_give_valid_schedules():
    """Slice of tests/harness/fuzz_schedule.py (5 000 cases clean): n up to 20, both strategies, streamed or one-shot."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "harness"))
    import fuzz_schedule
    assert fuzz_schedule.run(9000, 9150) == 0
