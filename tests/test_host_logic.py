"""CPU-only tests of the product's host logic and of the C-ABI library's loadability."""
import ctypes
import itertools
import math
import os
import re

import numpy as np
import pytest

from qiskit_aakash_b200 import capi, circuits as C, engine, hostpass, schedule
from qiskit_aakash_b200.exceptions import BasicAerError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI: the CUDA library loads here (no GPU needed) and exports what the header declares ----

def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    header = open(os.path.join(ROOT, "include", "dmb200.h")).read()
    declared = set(re.findall(r"\b(dmb_[a-z0-9_]+)\s*\(", header))
    declared -= {"dmb_ctx"}
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(capi.DEFAULT_LIB)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib2 = capi.load_library()
    assert lib2.dmb_abi_version() == capi.ABI_VERSION == 5
    assert lib2.dmb_sizeof_op() == capi.OP_DTYPE.itemsize == 336
    assert lib2.dmb_sizeof_pass() == capi.PASS_DTYPE.itemsize == 32 + 16 * 336
    assert lib2.dmb_sizeof_qop() == capi.QOP_DTYPE.itemsize == 336


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.DmbError):
        engine.PauliEngine(3)
    from qiskit_aakash_b200 import BasicAer, execute
    job = execute(C.ghz(3), BasicAer.get_backend("dm_simulator"))
    with pytest.raises(capi.DmbError):
        job.result()


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(capi.DmbError):
        capi.load_library(str(tmp_path / "nope.so"))


# ---- merge / partition known answers -------------------------------

def _sig(levels, n):
    return [[(g.name, tuple(g.qubits)) for g in lv] for lv in levels[:n]]


def _levels(build, n, merge=True):
    c = C.Circuit(n)
    build(c)
    ops = hostpass.merge_single_qubit_gates(c.instructions, n, merge)
    return hostpass.partition_levels(ops, n)


def test_u3_merge_kats():
    b, a, g = hostpass.zyz_from_yzy(0.7, 0.4, 1.1)
    assert np.allclose([b, a, g], [1.4178514323015705, 0.6198495499898623, 0.25664125730160314], atol=1e-15)
    b, a, g = hostpass.zyz_from_yzy(math.pi, math.pi / 2, math.pi / 2)
    assert abs(b) < 1e-7 and abs(a - math.pi / 2) < 1e-12 and abs(g - math.pi / 2) < 1e-12


def test_merge_kats_do_not_mutate_input():
    c = C.Circuit(1)
    c.u1(.5, 0); c.u1(.25, 0); c.u3(.1, .2, .3, 0); c.u1(.125, 0); c.u2(.3, .4, 0)
    before = [(i.name, list(i.params)) for i in c.instructions]
    out = hostpass.merge_single_qubit_gates(c.instructions, 1)
    assert [(i.name, list(i.params)) for i in c.instructions] == before
    assert len(out) == 1 and out[0].name == "u3"
    assert np.allclose(out[0].params, [1.6455912993793047, 0.3664375536449294, 1.7774866699020393], atol=1e-15)
    c = C.Circuit(1); c.u3(.1, .2, .3, 0); c.iden(0); c.u3(.1, .2, .3, 0)
    out = hostpass.merge_single_qubit_gates(c.instructions, 1)
    assert np.allclose(out[0].params, [0.1937626418121086, 0.45120320789875784, 0.5512032078987579], atol=1e-15)
    c = C.Circuit(3); c.u1(3.6, 0); c.cx(0, 1); c.cx(1, 0); c.u1(2.6, 2); c.s(2); c.y(2)
    out = hostpass.merge_single_qubit_gates(c.instructions, 3)
    q2 = [g for g in out if g.qubits == [2]]
    assert len(q2) == 1 and [round(x, 6) for x in q2[0].params] == [3.141593, 1.570796, 5.741593]
    # insertion order: u1(2.6) is flushed by cx(1,0), s.y -> u3(pi, pi/2, pi)
    c = C.Circuit(3); c.u1(3.6, 0); c.cx(0, 1); c.u1(2.6, 2); c.cx(1, 0); c.s(2); c.y(2)
    out = hostpass.merge_single_qubit_gates(c.instructions, 3)
    q2 = [g for g in out if g.qubits == [2]]
    assert [g.name for g in q2] == ["u1", "u3"]
    assert np.allclose(q2[1].params, [math.pi, math.pi / 2, math.pi], atol=1e-12)


def test_merge_rejects_unknown_gate():
    c = C.Circuit(1)
    c.instructions.append(C.instr("unitary", [0], [np.eye(2)]))
    with pytest.raises(BasicAerError):
        hostpass.merge_single_qubit_gates(c.instructions, 1)


def test_partition_kats():
    u = lambda c, q: c.u3(.1, .2, .3, q)
    lv, cnt = _levels(lambda c: (u(c, 0), u(c, 1), c.cx(0, 1), u(c, 0), c.cx(1, 2), u(c, 2)), 3)
    assert cnt == 4 and _sig(lv, cnt) == [[("u3", (0,)), ("u3", (1,))], [("cx", (0, 1))],
                                         [("u3", (0,)), ("cx", (1, 2))], [("u3", (2,))]]
    lv, cnt = _levels(lambda c: (u(c, 0), c.measure(0, 0), u(c, 1), c.measure(1, 1)), 2)
    assert cnt == 4 and [len(x) for x in lv[:cnt]] == [1, 1, 1, 1]
    lv, cnt = _levels(lambda c: (u(c, 0), c.cx(0, 1), c.measure(0, 0), c.measure(1, 1), c.measure(2, 2), u(c, 2)), 3)
    assert cnt == 4 and [len(x) for x in lv[:cnt]] == [1, 1, 3, 1]
    lv, cnt = _levels(lambda c: (u(c, 0), c.reset(0), u(c, 0), c.reset(1)), 2)
    assert cnt == 3 and [[g.name for g in x] for x in lv[:cnt]] == [["u3"], ["reset", "reset"], ["u3"]]
    lv, cnt = _levels(lambda c: (c.cx(0, 1), c.cx(1, 2), c.cx(0, 2), u(c, 1)), 3)
    assert cnt == 3 and _sig(lv, cnt)[2] == [("cx", (0, 2)), ("u3", (1,))]
    lv, cnt = _levels(lambda c: (u(c, 0), c.barrier(), u(c, 1), c.barrier(), u(c, 0), u(c, 1)), 2)
    assert cnt == 3
    assert _levels(lambda c: c.instructions.extend(C.qft(8).instructions), 8)[1] == 54
    for n in (8, 9, 10):
        assert _levels(lambda c: c.instructions.extend(C.random_layered(n, 4, 1, readout=False).instructions), n)[1] == 8


@pytest.mark.parametrize("seed", range(8))
def test_hostpass_agrees_with_oracle_restatement(seed):
    """Two independent restatements of merge+partition (oracle/, product) on random circuits."""
    import cases
    from oracle import dm_oracle as O
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 7))
    circ = cases._rand_circuit(n, int(rng.integers(5, 60)), 300 + seed)
    if seed % 2:
        qs = sorted(rng.choice(n, size=max(1, n // 2), replace=False).tolist())
        circ.measure(qs, qs, basis="Z"); circ.u3(.1, .2, .3, 0); circ.reset(n - 1); circ.measure(0, 0, basis="X")
    a_ops = O.single_gate_merge([O.as_instruction(i) for i in circ.instructions], n)
    a_lv, a_cnt = O.partition(a_ops, n)
    b_ops = hostpass.merge_single_qubit_gates(circ.instructions, n)
    b_lv, b_cnt = hostpass.partition_levels(b_ops, n)
    assert a_cnt == b_cnt and len(a_lv) == len(b_lv)
    for x, y in zip(a_lv, b_lv):
        assert [(g.name, tuple(g.qubits)) for g in x] == [(g.name, tuple(g.qubits)) for g in y]
        for g, h in zip(x, y):
            if g.name in ("u1", "u3"):
                assert np.array_equal(np.array(g.params, float), np.array(h.params, float))


# ---- shared-memory swizzle: every op's access pattern is bank-conflict free -------------------

def _swz(l):
    d2, d3, d4, d5 = (l >> 4) & 3, (l >> 6) & 3, (l >> 8) & 3, (l >> 10) & 3
    hi = d2 ^ d3 ^ d4 ^ d5
    odd = d3 ^ d5
    return l ^ (hi << 2) ^ (((odd ^ (odd >> 1)) & 1) << 1)


def test_swizzle_is_a_bijection_keeping_pairs():
    img = [_swz(l) for l in range(4096)]
    assert sorted(img) == list(range(4096))
    assert all((_swz(l) ^ _swz(l ^ 1)) == 1 for l in range(0, 4096, 2))
    assert all(_swz(l) >> 4 == l >> 4 for l in range(4096))


@pytest.mark.parametrize("a,b", [(a, b) for a in range(6) for b in range(6) if a != b])
def test_lane_order_is_bank_conflict_free(a, b):
    """Model of the LSU: a 64-bit warp access is served per half-warp (16 lanes must hit 16
    distinct 8-byte slots of the 128-byte row space), a 128-bit access per quarter-warp
    (8 lanes, 8 distinct 16-byte chunks).  Checked for every element (i, j) of the block."""
    K = 6
    fd = schedule.lane_order(K, a, b)
    assert sorted(fd + [a, b]) == list(range(K))
    mode_b = 0 in (a, b)
    for warp in range(8):
        for i, j in itertools.product(range(4), repeat=2):
            addrs = []
            for lane in range(32):
                t = warp * 32 + lane
                bl = 0
                for m in range(K - 2):
                    bl |= ((t >> (2 * m)) & 3) << (2 * fd[m])
                addrs.append(_swz(bl | (i << (2 * a)) | (j << (2 * b))))
            if mode_b:
                for quarter in range(4):
                    chunks = {(x >> 1) & 7 for x in addrs[8 * quarter: 8 * quarter + 8]}
                    assert len(chunks) == 8
            else:
                for half in range(2):
                    slots = {x & 15 for x in addrs[16 * half: 16 * half + 16]}
                    assert len(slots) == 16


@pytest.mark.parametrize("a,b", [(a, b) for a in range(1, 6) for b in range(1, 6) if a != b])
def test_paired_kernel_mode_a_is_bank_conflict_free(a, b):
    """Tile kernel, paired op body (dmb_lean_op_pair): real lane u plays virtual threads 2u / 2u + 1 and
    moves both blocks with one 128-bit access per (i, j).  lane_order puts tile digit 0 first, so the pair
    (2u, 2u + 1) is one 16-byte chunk, and a quarter-warp of real lanes must hit 8 distinct chunks."""
    K = 6
    fd = schedule.lane_order(K, a, b)
    assert fd[0] == 0                                 # the DMB_PAIRABLE condition (dm_device.h)
    for warp in range(4):                             # 128 real threads
        for i, j in itertools.product(range(4), repeat=2):
            chunk_of_lane = []
            for lane in range(32):
                u = warp * 32 + lane
                addrs = []
                for t in (2 * u, 2 * u + 1):
                    bl = 0
                    for m in range(K - 2):
                        bl |= ((t >> (2 * m)) & 3) << (2 * fd[m])
                    addrs.append(_swz(bl | (i << (2 * a)) | (j << (2 * b))))
                assert addrs[1] == addrs[0] + 1 and addrs[0] % 2 == 0      # the two halves of one aligned pair
                chunk_of_lane.append((addrs[0] >> 1) & 7)
            for quarter in range(4):
                assert len(set(chunk_of_lane[8 * quarter: 8 * quarter + 8])) == 8


# ---- pass scheduler ------------------------------------------------------------------------

def _random_devops(rng, n_digits, count):
    ops = []
    for k in range(count):
        if rng.random() < 0.2:
            ops.append(schedule.DevOp(capi.OP_MATS, int(rng.integers(n_digits)), None, np.eye(4), None))
        else:
            a, b = rng.choice(n_digits, size=2, replace=False)
            ops.append(schedule.DevOp(capi.OP_CX, int(a), int(b), np.eye(4), None))
        ops[-1].coef = [float(k)]           # tag = program position
    return ops


@pytest.mark.parametrize("n_digits,max_ops", [(2, 16), (3, 4), (6, 16), (9, 1), (14, 16), (14, 5)])
def test_scheduler_preserves_program_order_per_digit(n_digits, max_ops):
    rng = np.random.default_rng(n_digits * 100 + max_ops)
    ops = _random_devops(rng, n_digits, 200)
    passes = schedule.build_passes(ops, n_digits, max_ops=max_ops)
    seen = []
    last_on_digit = {}
    for p in passes:
        K = int(p["n_tile_digits"])
        tile = [int(x) for x in p["tile_digit"][:K]]
        assert K == min(6, n_digits) and tile[0] == 0 and tile == sorted(set(tile))
        assert 1 <= p["n_ops"] <= max_ops
        for o in p["ops"][:p["n_ops"]]:
            tag = int(o["coef"][0])
            src = ops[tag]
            assert tile[o["a"]] == src.da
            if src.db is not None:
                assert tile[o["b"]] == src.db
            assert sorted([int(o["a"]), int(o["b"])] + [int(x) for x in o["fd"][:K - 2]]) == list(range(K))
            for d in src.digits():
                assert last_on_digit.get(d, -1) < tag, "ops sharing a digit were reordered"
                last_on_digit[d] = tag
            seen.append(tag)
    assert sorted(seen) == list(range(len(ops)))


@pytest.mark.parametrize("n_qubits,max_ops", [(2, 16), (3, 4), (5, 6), (7, 10), (14, 10), (14, 4)])
def test_relabel_scheduler_is_a_valid_reordering(n_qubits, max_ops):
    """Replay the relabelled schedule symbolically: every op lands on the digit positions its
    qubits occupy at that moment (SWAPs move qubits), per-qubit program order is kept, and the
    returned layout matches the replay."""
    rng = np.random.default_rng(n_qubits * 10 + max_ops)
    qops = []
    for k in range(150):
        if n_qubits == 1 or rng.random() < 0.15:
            qops.append(schedule.DevOp(capi.OP_MATS, int(rng.integers(n_qubits)), None, np.eye(4), None, [float(k)]))
        else:
            a, b = rng.choice(n_qubits, size=2, replace=False)
            qops.append(schedule.DevOp(capi.OP_CX, int(a), int(b), np.eye(4), None, [float(k)]))
    nd = max(n_qubits, 2)
    pos = [n_qubits - 1 - q for q in range(n_qubits)]
    sim = {p: q for q, p in enumerate(pos)}                  # digit position -> qubit
    passes = schedule.build_passes_relabel(qops, pos, nd, max_ops=max_ops)
    last, seen = {}, []
    for p in passes:
        K = int(p["n_tile_digits"])
        tile = [int(x) for x in p["tile_digit"][:K]]
        assert tile[0] == 0 and tile == sorted(set(tile)) and K == min(6, nd) and p["n_ops"] <= 16
        if nd >= 2:
            assert tile[1] == 1
        for o in p["ops"][:p["n_ops"]]:
            da, db = tile[o["a"]], tile[o["b"]]
            if o["kind"] == capi.OP_SWAP:
                sim[da], sim[db] = sim.get(db), sim.get(da)
                continue
            tag = int(o["coef"][0])
            src = qops[tag]
            assert sim[da] == src.da
            if src.db is not None:
                assert sim[db] == src.db
            for q in src.digits():
                assert last.get(q, -1) < tag
                last[q] = tag
            seen.append(tag)
    assert sorted(seen) == list(range(len(qops)))
    for q in range(n_qubits):
        assert sim[pos[q]] == q


def test_scheduler_fuses_brick_layers():
    """Config 3 shape: the fused schedule needs far fewer HBM round trips than gates."""
    from emu_backend import NumpyAllocator, emu_lib
    circ = C.random_layered(14, 20, 1400, readout=False)
    n_cx = circ.count_ops()["cx"]
    e = engine.PauliEngine.__new__(engine.PauliEngine)     # planning only: no buffers
    e.n, e.nd = 14, 14
    e.pos = [13 - q for q in range(14)]
    e.pending, e.queue, e.max_ops_per_pass, e.reserve_low, e.drain_threshold, e.relabel = [None] * 14, [], 16, 2, 0, False
    for ins in circ.instructions:
        if ins.name == "u3":
            e.apply_1q(ins.qubits[0], engine.gate_matrix("u3", ins.params, {"rz": [1, 0], "ry": [1, 0]}))
        else:
            e.apply_cx(ins.qubits[0], ins.qubits[1])
    passes = e.plan()
    assert int(passes["n_ops"].sum()) >= n_cx
    assert len(passes) <= n_cx // 2, (len(passes), n_cx)
    assert all(p["tile_digit"][1] == 1 for p in passes)


# ---- single-qubit map builders ---------------------------------------------------------------

def test_gate_matrix_matches_sequential_rotations():
    err = {"rx": [1., 0.], "ry": [0.98, -0.02], "rz": [0.97, 0.03]}
    m = engine.gate_matrix("u3", [0.3, 1.1, -2.0], err)
    v = np.array([0.5, 0.1, -0.2, 0.3])
    w = v.copy()
    for axis, ang in (("rz", -2.0), ("ry", 0.3), ("rz", 1.1)):
        r, d = err[axis]
        c, s = r * np.cos(ang + d), r * np.sin(ang + d)
        k0, k1 = {"rz": (1, 2), "ry": (3, 1)}[axis]
        w[k0], w[k1] = c * w[k0] - s * w[k1], c * w[k1] + s * w[k0]
    assert np.allclose(m @ v, w, atol=1e-15)
    assert np.array_equal(m[0], [1, 0, 0, 0])


def test_streaming_schedule_equals_one_shot_schedule():
    """build_passes_relabel(min_tail=...) called repeatedly on a growing op stream produces the
    same passes as one call on the whole stream once the tail covers the look-ahead."""
    from qiskit_aakash_b200 import schedule
    rng = np.random.default_rng(3)
    n = 10
    qops = []
    for layer in range(40):
        for q in range(layer % 2, n - 1, 2):
            a, b = (q, q + 1) if rng.integers(2) else (q + 1, q)
            qops.append(schedule.DevOp(capi.OP_CX, a, b, None, None, None))
    pos0 = [n - 1 - q for q in range(n)]
    pos_a = list(pos0)
    whole = schedule.build_passes_relabel(list(qops), pos_a, n, max_ops=10)
    pos_b = list(pos0)
    queue, parts = [], []
    for i, op in enumerate(qops):
        queue.append(op)
        if len(queue) >= 64 + 16:
            p, queue = schedule.build_passes_relabel(queue, pos_b, n, max_ops=10, min_tail=64)
            parts.append(p)
    parts.append(schedule.build_passes_relabel(queue, pos_b, n, max_ops=10))
    streamed = np.concatenate(parts)
    assert len(parts) > 3 and pos_a == pos_b
    assert streamed.tobytes() == whole.tobytes()


def test_split_store_thread_order_is_bank_conflict_free():
    """Relabelling store with tile digit 0 swapped away (two 64-bit shared-memory reads per pair):
    the library picks a thread order whose half-warps read 16 distinct 8-byte bank slots, and whose
    warps still cover store-index bits 1..3 (whole 128-byte lines in global memory)."""
    import ctypes
    import itertools
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
    import build_emu
    lib = ctypes.CDLL(build_emu.build())
    lib.dmb_emu_split_order.restype = ctypes.c_int
    checked = 0
    for perm in itertools.permutations(range(6)):
        if perm[0] == 0:
            continue                                   # digit 0 stays: 128-bit path, nothing to choose
        moved = sum(1 for k in range(6) if perm[k] != k)
        if moved > 4:
            continue                                   # the scheduler emits at most two transpositions
        p = (ctypes.c_int32 * 6)(*perm)
        tb = (ctypes.c_int32 * 8)()
        worst = lib.dmb_emu_split_order(p, tb)
        assert worst == 1, (perm, list(tb), worst)
        assert {1, 2, 3} <= set(tb[:5]) and len(set(tb)) == 8 and all(1 <= b <= 11 for b in tb)
        checked += 1
    assert checked > 50


def test_relabelling_store_is_an_exact_digit_permutation():
    """Swap-only passes on the emulated kernels == NumPy axis swaps, bit for bit, and every swap
    was realised by the write-back (none ran as a shared-memory op)."""
    from emu_backend import emu_engine
    import store_cases
    folded, total = store_cases.check_relabelling_store(emu_engine)
    assert folded == total


def test_chained_op_order_keeps_every_dependency():
    """dmb_chain_ops hoists the second CNOT of a same-pair couple next to the first one.  Random K = 6 passes: the order
    the library runs is a permutation of the ops in front of the folded swaps, any two ops that share a tile digit keep
    their relative order (ops on disjoint digits commute), and exactly the adjacent same-pair, same-kind couples with
    at least one map are chained."""
    import ctypes
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
    import build_emu
    from qiskit_aakash_b200 import capi, schedule
    lib = ctypes.CDLL(build_emu.build())
    lib.dmb_emu_pass_chained.restype = ctypes.c_int
    rng = np.random.default_rng(7)
    total_chained = 0
    for trial in range(300):
        P = np.zeros(1, dtype=capi.PASS_DTYPE)
        n_ops = int(rng.integers(1, capi.MAX_OPS + 1))
        P[0]["n_tile_digits"] = 6
        P[0]["n_ops"] = n_ops
        P[0]["tile_digit"][:6] = [0, 1, 2, 3, 4, 5]
        pairs = []
        for k in range(n_ops):
            if pairs and rng.random() < 0.35:
                a, b = pairs[int(rng.integers(len(pairs)))]          # revisit an earlier pair (same orientation)
            else:
                a, b = (int(x) for x in rng.choice(6, 2, replace=False))
            pairs.append((a, b))
            op = P[0]["ops"][k]
            op["kind"] = capi.OP_CX if rng.random() < 0.8 else capi.OP_MATS
            op["a"], op["b"] = a, b
            op["fd"][:4] = schedule.lane_order(6, a, b)
            if rng.random() < 0.8:
                op["flags"] |= capi.HAS_PA
                op["pa"][:] = [0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
            if rng.random() < 0.6:
                op["flags"] |= capi.HAS_PB
                op["pb"][:] = [0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
        order = (ctypes.c_int32 * capi.MAX_OPS)()
        chained = lib.dmb_emu_pass_chained(ctypes.c_void_p(P.ctypes.data), 12, order)
        assert chained >= 0
        total_chained += chained
        got = list(order[:n_ops])
        assert sorted(got) == list(range(n_ops))
        where = {op_index: position for position, op_index in enumerate(got)}
        for i in range(n_ops):
            for j in range(i + 1, n_ops):
                if set(pairs[i]) & set(pairs[j]):
                    assert where[i] < where[j], (trial, pairs, got)
    assert total_chained > 100
