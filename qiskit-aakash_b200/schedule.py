"""Gate-fusion scheduling: pack a stream of two-digit device ops into tile passes.

This generalises the reference's only fusion, the per-qubit U3 merge
(``basicaertools.py:251-307``): after the engine has folded *all* single-qubit work
(rotations, memory noise, projections, resets) into 4x4 matrices riding on the next
two-qubit op of their qubit, what is left is a sequence of two-digit ops.  One launch of
the tile kernel stages every 4^K-coefficient tile (K <= 6 digit positions, always
including digit 0 so that global accesses are >= 16-byte pairs, and digit 1 whenever
possible so that runs are >= 128 bytes) in shared memory and can apply up to
``MAX_OPS`` ops whose digits all lie inside the tile -- one HBM round trip for many gates.

``build_passes`` is a greedy list scheduler: it walks the op stream in order, admits an op
into the current pass when its digits fit the tile (growing the tile set while there is
room) and none of its digits is *blocked* by an earlier op that had to be skipped (ops on
disjoint digits commute, so hoisting past skipped ops is exact).
"""
from __future__ import annotations

import os

import numpy as np

from . import capi


class DevOp:
    """A two-digit op on physical digit positions ``da`` (matrix ``pa``) and ``db`` (``pb``)."""
    __slots__ = ("kind", "da", "db", "pa", "pb", "coef")

    def __init__(self, kind, da, db, pa=None, pb=None, coef=None):
        self.kind = kind
        self.da = da
        self.db = db          # None for a lone single-digit op (partner chosen per pass)
        self.pa = pa
        self.pb = pb
        self.coef = coef

    def digits(self):
        return (self.da,) if self.db is None else (self.da, self.db)


def lane_order(K, a, b):
    """Order in which thread-index bit pairs are dealt to the free tile-local digits so that
    shared-memory accesses are conflict-free under ``dmb_swz`` (csrc/dm_device.h):
    mode (A) (digit 0 free): digit 0 first; mode (B) (digit 0 is an op digit): digit 1 (else
    2) first, then digit 3 (else 5) whose low bit becomes lane bit 2."""
    free = [d for d in range(K) if d not in (a, b)]
    if not free:
        return []
    if 0 in free:
        first = [0]
    else:
        o1 = 1 if 1 in free else (2 if 2 in free else free[0])
        rest = [d for d in free if d != o1]
        o2 = 3 if 3 in rest else (5 if 5 in rest else (rest[0] if rest else None))
        first = [o1] + ([o2] if o2 is not None else [])
    return first + [d for d in free if d not in first]


def _rows13(m):
    m = np.asarray(m, dtype=np.float64)
    if m.shape != (4, 4):
        raise ValueError("single-qubit map must be 4x4")
    if not (m[0, 0] == 1.0 and m[0, 1] == 0.0 and m[0, 2] == 0.0 and m[0, 3] == 0.0):
        raise ValueError("single-qubit map is not trace preserving (row 0 != e0)")
    return m[1:4, :].reshape(12)


def build_passes(ops, n_digits, max_tile=capi.MAX_TILE_DIGITS, max_ops=capi.MAX_OPS, window=256,
                 reserve_low=2):
    """ops: list of DevOp in program order -> numpy array of ``capi.PASS_DTYPE``.

    ``reserve_low`` digit positions 0..reserve_low-1 are part of every tile, which makes the
    contiguous global-memory runs of a tile 4^reserve_low doubles long (2 -> 128 bytes)."""
    K = min(max_tile, n_digits)
    if K < 2:
        raise ValueError("state must have at least 2 digit positions")
    max_ops = min(max_ops, capi.MAX_OPS)
    remaining = list(ops)
    plans = []
    while remaining:
        tile = set(range(max(1, min(reserve_low, K))))
        blocked = set()
        chosen, keep = [], []
        scanned = 0
        for idx, op in enumerate(remaining):
            if scanned >= window or len(chosen) >= max_ops or len(blocked) >= n_digits:
                keep.extend(remaining[idx:])
                break
            scanned += 1
            dg = op.digits()
            if any(d in blocked for d in dg):
                blocked.update(dg)
                keep.append(op)
                continue
            need = tile.union(dg)
            if len(dg) == 1 and len(need) < 2:
                need = need.union({1 if dg[0] == 0 else 0})
            if len(need) <= K:
                tile = need
                chosen.append(op)
            else:
                blocked.update(dg)
                keep.append(op)
        if not chosen:
            raise RuntimeError("scheduler made no progress")
        # spend spare tile slots on the lowest unused digits: longer contiguous runs
        d = 0
        while len(tile) < K:
            if d not in tile:
                tile.add(d)
            d += 1
        plans.append((sorted(tile), chosen))
        remaining = keep
    return encode_passes(plans)


def _greedy_select(remaining, start_set, cap, max_ops, window, n_qubits):
    """One pass of list scheduling over QUBIT ids: returns (chosen indices, tile qubit set)."""
    tile = set(start_set)
    blocked = set()
    chosen = []
    for idx, op in enumerate(remaining):
        if idx >= window or len(chosen) >= max_ops or len(blocked) >= n_qubits:
            break
        dg = op.digits()
        if any(d in blocked for d in dg):
            blocked.update(dg)
            continue
        need = tile.union(dg)
        if len(need) <= cap:
            tile = need
            chosen.append(idx)
        else:
            blocked.update(dg)
    return chosen, tile


def pack_qops(qops):
    """[DevOp on qubit ids] -> ``capi.QOP_DTYPE`` array for ``dmb_schedule`` (vectorised: the
    matrices are stacked once instead of being copied field by field)."""
    n = len(qops)
    arr = np.zeros(n, dtype=capi.QOP_DTYPE)
    if not n:
        return arr
    arr["kind"] = [op.kind for op in qops]
    arr["qa"] = [op.da for op in qops]
    arr["qb"] = [-1 if op.db is None else op.db for op in qops]
    flags = np.zeros(n, dtype=np.int32)
    for field, bit in (("pa", capi.HAS_PA), ("pb", capi.HAS_PB)):
        idx = [i for i, op in enumerate(qops) if getattr(op, field) is not None]
        if idx:
            mats = np.asarray([getattr(qops[i], field) for i in idx], dtype=np.float64)
            if mats.shape[1:] != (4, 4):
                raise ValueError("single-qubit map must be 4x4")
            if not np.array_equal(mats[:, 0, :], np.broadcast_to([1.0, 0.0, 0.0, 0.0], (len(idx), 4))):
                raise ValueError("single-qubit map is not trace preserving (row 0 != e0)")
            arr[field][idx] = mats[:, 1:4, :].reshape(len(idx), 12)
            flags[idx] |= bit
    arr["flags"] = flags
    for i, op in enumerate(qops):
        if op.coef is not None:
            c = np.asarray(op.coef, dtype=np.float64).reshape(-1)
            arr["coef"][i, :c.size] = c
    return arr


#: The native scheduler (``dmb_schedule``, csrc/dm_schedule.h) is the product path; the Python
#: ``build_passes_relabel`` below is its executable specification (tests compare them byte for
#: byte) and serves the experimental knob the native one does not carry (swap_weight).
NATIVE_DEFAULT = bool(int(os.environ.get("DMB_NATIVE_SCHEDULE", "1")))


def relabel_passes(lib, qops, pos, n_digits, max_ops=capi.MAX_OPS, min_tail=0, final_moves=None, native=None,
                   strategy=capi.SCHED_PROGRAM_ORDER):
    """``build_passes_relabel`` through ``dmb_schedule`` of ``lib``.  Same return convention:
    passes | (passes, leftover ops) with ``min_tail`` | (passes, leftover moves) with ``final_moves``.
    ``strategy``: ``capi.SCHED_PROGRAM_ORDER`` (what the Python specification does) or
    ``capi.SCHED_TILE_SEARCH`` (native only)."""
    native = NATIVE_DEFAULT if native is None else native
    if not native:
        return build_passes_relabel(qops, pos, n_digits, max_ops=max_ops, min_tail=min_tail, final_moves=final_moves)
    passes, left, moves_left = capi.schedule(lib, pack_qops(qops), pos, n_digits, max_ops=max_ops, min_tail=min_tail,
                                             final_moves=final_moves, strategy=strategy)
    if final_moves is not None:
        return passes, moves_left
    if min_tail > 0:
        return passes, [qops[i] for i in left]
    return passes


def build_passes_relabel(qops, pos, n_digits, max_tile=capi.MAX_TILE_DIGITS, max_ops=capi.MAX_OPS, window=256,
                         swap_weight=0.0, min_tail=0, final_moves=None):
    """Like ``build_passes`` but with dynamic relabelling of the two low digit positions.

    ``final_moves`` = [(qubit, digit position)]: after the last op, move these qubits to the given
    positions (the sharded engine parks the qubits it is about to evict in the top local slots).
    Moves whose two digit positions fit into the last pass's tile ride along as trailing swaps
    (free: the library folds them into the write-back); the rest are returned for a pass of
    their own.  The return value is then ``(passes, leftover moves)``.

    With ``min_tail`` > 0 scheduling stops as soon as no more than ``min_tail`` ops are left and
    ``(passes, leftover ops)`` is returned: a caller that is still producing ops keeps the tail
    queued, so that pass boundaries do not depend on where the stream was cut.

    ``qops`` are DevOps whose ``da``/``db`` are QUBIT ids; ``pos[q]`` is the digit position of
    qubit q (mutated to the layout after the last pass).  Digit positions 0 and 1 are part of
    every tile (they give the 128-byte contiguous runs in HBM), so whatever qubits live there
    ride along in every pass for free.  After choosing the ops of pass k the scheduler looks at
    what pass k+1 would like to work on and, if that set shares qubits with pass k's tile,
    moves up to two of them into positions 0/1 with SWAP ops appended to pass k (both digits
    are in pass k's tile, so the swap is one more fused op, not another HBM round trip).  For
    nearest-neighbour circuits this turns the useful tile from 4 fresh qubits into a sliding
    window of 6."""
    n_qubits = len(pos)
    K = min(max_tile, n_digits)
    if K < 2:
        raise ValueError("state must have at least 2 digit positions")
    max_ops = min(max_ops, capi.MAX_OPS)
    real_ops_cap = max(1, max_ops - 2) if K >= 4 else max_ops     # leave room for two swaps
    owner = {pos[q]: q for q in range(n_qubits)}                  # digit position -> qubit (phantoms absent)
    remaining = list(qops)
    plans = []

    def lows():
        return [owner[d] for d in (0, 1) if d in owner and d < n_digits]

    while len(remaining) > min_tail:
        chosen, tile_q = _greedy_select(remaining, lows(), K, real_ops_cap, window, n_qubits)
        if not chosen:
            raise RuntimeError("scheduler made no progress")
        chosen_set = set(chosen)
        ops_now = [remaining[i] for i in chosen]
        remaining = [op for i, op in enumerate(remaining) if i not in chosen_set]
        # tile digit positions: the chosen qubits' positions + positions 0,1 (+ filler)
        tile_d = {pos[q] for q in tile_q} | {0, 1} if n_digits >= 2 else {pos[q] for q in tile_q}
        moves_here = []
        if final_moves and not remaining:
            for q, target in final_moves:
                if pos[q] == target:
                    continue
                if len(tile_d | {pos[q], target}) <= K:
                    tile_d |= {pos[q], target}
                    moves_here.append((q, target))
        d = 0
        while len(tile_d) < K:
            if d not in tile_d:
                tile_d.add(d)
            d += 1
        devops = []
        for op in ops_now:
            devops.append(DevOp(op.kind, pos[op.da], None if op.db is None else pos[op.db], op.pa, op.pb, op.coef))
        # look ahead: which two qubits of this tile should sit in positions 0/1 for the next
        # pass?  Try every pair (plus "leave as is") and keep the one whose greedy next pass
        # executes the most ops; ties prefer fewer swaps.
        if remaining and K >= 4:
            in_tile = sorted(owner[dd] for dd in tile_d if dd in owner)
            cur = lows()
            active = set()
            for op in remaining[:window]:
                active.update(op.digits())
            cands = [q for q in in_tile if q in active]
            best = (len(_greedy_select(remaining, cur, K, real_ops_cap, window, n_qubits)[0]), 0, tuple(cur))
            score = lambda cnt, swaps: (cnt - swap_weight * swaps, -swaps)
            for i in range(len(cands)):
                for j in range(i + 1, len(cands)):
                    pair = (cands[i], cands[j])
                    swaps = sum(1 for q in pair if q not in cur)
                    cnt = len(_greedy_select(remaining, pair, K, real_ops_cap, window, n_qubits)[0])
                    if score(cnt, swaps) > score(best[0], best[1]):
                        best = (cnt, swaps, pair)
            targets = list(best[2])
            keep = [q for q in cur if q in targets]
            new = [q for q in targets if q not in cur]
            free_low = [dd for dd in (0, 1) if owner.get(dd) not in keep]
            for q, dd in zip(new, free_low):
                other = owner.get(dd)
                src = pos[q]
                devops.append(DevOp(capi.OP_SWAP, dd, src))
                owner[dd] = q
                pos[q] = dd
                if other is not None:
                    owner[src] = other
                    pos[other] = src
                else:
                    owner.pop(src, None)
        for q, target in moves_here:                      # trailing swaps of the last pass
            src = pos[q]
            if src == target:
                continue
            other = owner.get(target)
            devops.append(DevOp(capi.OP_SWAP, src, target))
            owner[target], pos[q] = q, target
            if other is not None:
                owner[src], pos[other] = other, src
            else:
                owner.pop(src, None)
        plans.append((sorted(tile_d), devops))
    if final_moves is not None:
        return encode_passes(plans), [(q, t) for q, t in final_moves if pos[q] != t]
    if min_tail > 0:
        return encode_passes(plans), remaining
    return encode_passes(plans)


def encode_passes(plans):
    """[(sorted tile digits, [DevOp...])] -> PASS_DTYPE array."""
    out = np.zeros(len(plans), dtype=capi.PASS_DTYPE)
    for pi, (tile, chosen) in enumerate(plans):
        K = len(tile)
        rec = out[pi]
        rec["n_tile_digits"] = K
        rec["n_ops"] = len(chosen)
        rec["tile_digit"][:K] = tile
        local = {d: j for j, d in enumerate(tile)}
        for oi, op in enumerate(chosen):
            o = rec["ops"][oi]
            a = local[op.da]
            if op.db is None:
                b = 1 if a == 0 else 0
            else:
                b = local[op.db]
            o["kind"] = op.kind
            flags = 0
            if op.pa is not None:
                o["pa"] = _rows13(op.pa)
                flags |= capi.HAS_PA
            if op.pb is not None:
                o["pb"] = _rows13(op.pb)
                flags |= capi.HAS_PB
            o["flags"] = flags
            o["a"], o["b"] = a, b
            fd = lane_order(K, a, b)
            o["fd"][:len(fd)] = fd
            if op.coef is not None:
                c = np.asarray(op.coef, dtype=np.float64).reshape(-1)
                o["coef"][:c.size] = c
    return out
