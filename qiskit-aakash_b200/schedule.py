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

import numpy as np

from . import capi

IDENT = None


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
