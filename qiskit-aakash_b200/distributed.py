"""Multi-GPU layer: the Pauli-coefficient vector sharded over G = 2, 4 or 8 B200s of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink 5 / NVSwitch).  Rank r owns the
contiguous slice ``[r * 4^n / G, (r+1) * 4^n / G)`` of the reference's vector, i.e. the top
``g = log2 G`` bits of the 2n-bit flat index are the rank: for G = 4
qubit 0's whole digit, for G = 2 the high bit of that digit, for G = 8 that digit plus the
high bit of qubit 1's digit.

* Qubits sit in *slots* (digit positions).  The top ``m = ceil(g/2)`` slots are (partly)
  global; all others are ordinary local digits and the single-GPU tile kernel runs on them
  unchanged -- ops on local qubits are embarrassingly parallel, no collective.
* Single-qubit maps never force communication: they stay in the host-side pending matrix
  of their qubit (``PauliEngine`` docstring), also for global qubits, including the
  per-level memory noise with its I->Z mixing (SURVEY.md section 7 "hard parts").
* Before a two-qubit op touches a global qubit, the m global slots are swapped with the m
  top local slots: a permutation of index bits that leaves the low B bits alone, executed
  out of place as an exchange of 2^B-element blocks between ranks (NCCL grouped
  send/recv = all-to-all); ``ExchangePlan`` enumerates the block moves.  Which qubits are
  sent out is chosen Belady-style (farthest next use) and they are first moved into the top
  local slots with SWAP ops fused into the preceding tile passes.
* Readout: every rank gathers its part of the I/B marginal (``dmb_marginal`` drops terms
  whose global bits are not this rank's and folds pending maps of global qubits in as
  weights), one all-reduce of 2^n doubles, then the Walsh-Hadamard transform.

Every readout of the reference works on a sharded state: ensemble X/Y/Z/N, Expect and Bell (owner rank
reads, all-reduce), ``store_densitymatrix`` / ``compare`` / ``stored_density_matrix`` as per-rank slice files
(``canonicalise`` + ``store_shards`` / ``load_slice``: nothing is gathered), Pauli->matrix after a device-side
all-gather (small registers only).
"""
from __future__ import annotations

import os

import numpy as np

from . import capi, schedule
from .engine import PauliEngine, TorchCudaAllocator, shared_context
from .exceptions import BasicAerError


def join_shard_files(stem, world, out=None):
    """Concatenate the per-rank files of a sharded dump (``ShardedPauliEngine.store_shards``) into the single
    ``<stem>.npy`` the reference writes and reads (``dm_simulator.py:1271-1282``)."""
    parts = [np.load(ShardedPauliEngine.shard_file(stem, r, world), mmap_mode="r") for r in range(world)]
    vec = np.concatenate(parts)
    np.save(out or stem, vec)
    return vec


def log2_exact(x):
    g = int(x).bit_length() - 1
    if x < 1 or (1 << g) != x:
        raise BasicAerError("world size must be a power of two, got %r" % (x,))
    return g


class ExchangePlan:
    """Block moves of the global<->top-local slot swap for one rank.

    Index bits above B = 2*(n_loc - m) form two groups of 2m bits (upper: the m global
    slots, lower: the m top local slots); the swap exchanges the groups.  A block is the
    2^B elements sharing those 4m bits; (rank, block) pairs are the 4m-bit numbers
    H = rank << (4m-g) | block."""

    def __init__(self, n_qubits, world, rank):
        self.g = log2_exact(world)
        self.m = (self.g + 1) // 2
        self.n_loc = n_qubits - self.m
        self.n_bits_local = 2 * n_qubits - self.g
        self.block_bits = 4 * self.m - self.g
        self.B = self.n_bits_local - self.block_bits
        if self.n_loc < 2 * self.m or self.n_loc < 2 or self.B < 2:
            raise BasicAerError("too few qubits (%d) to shard over %d ranks" % (n_qubits, world))
        self.block_elems = 1 << self.B
        self.rank = rank
        self.world = world
        self.moves = []          # (src_block, dst_rank, dst_block)
        for s in range(1 << self.block_bits):
            dr, db = self.image(rank, s)
            self.moves.append((s, dr, db))

    def image(self, rank, block):
        two_m = 2 * self.m
        H = (rank << self.block_bits) | block
        hi, lo = H >> two_m, H & ((1 << two_m) - 1)
        H2 = (lo << two_m) | hi
        return H2 >> self.block_bits, H2 & ((1 << self.block_bits) - 1)

    def bytes_sent(self):
        return sum(8 * self.block_elems for (_, dr, _) in self.moves if dr != self.rank)


class TorchCommunicator:
    """torch.distributed plumbing (NCCL on GPUs; gloo in the CPU unit tests)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange(self, plan, src, dst):
        """dst[dst_block] on dst_rank <- src[src_block] for every move (out of place)."""
        dist, be = self.dist, plan.block_elems
        sends, recvs = [], []
        for (s, dr, db) in plan.moves:
            if dr == self.rank:
                dst[db * be:(db + 1) * be].copy_(src[s * be:(s + 1) * be])
            else:
                sends.append((dr, s))
        # the swap is an involution: block d of this rank is filled from image(rank, d)
        for d in range(1 << plan.block_bits):
            sr, sb = plan.image(self.rank, d)
            if sr != self.rank:
                recvs.append((sr, sb, d))
        ops = []
        for (dr, s) in sorted(sends):
            ops.append(dist.P2POp(dist.isend, src[s * be:(s + 1) * be], dr, group=self.group, tag=s))
        for (sr, sb, d) in sorted(recvs):
            ops.append(dist.P2POp(dist.irecv, dst[d * be:(d + 1) * be], sr, group=self.group, tag=sb))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def peer_addresses(self, ctx, ptrs):
        """Addresses, valid in THIS process, of every rank's buffers ``ptrs`` (one list per
        buffer): CUDA IPC handles exchanged through torch.distributed, peers mapped over
        NVLink.  Returns ``(addresses, opened handles)``, or ``(None, [])`` when the buffers cannot
        be mapped on EVERY rank -- not CUDA memory (gloo CPU tests), an allocator whose memory has
        no IPC handle (``expandable_segments`` / ``cudaMallocAsync``), no libcuda, no P2P between two
        ranks -- in which case the engine uses the NCCL block exchange instead.  The decision is
        collective (all-reduced flag), so all ranks take the same path."""
        if not self.torch.cuda.is_available() or self.dist.get_backend(self.group) != "nccl":
            return None, []
        mine, ok = None, 1
        try:
            mine = [ctx.ipc_export(p) for p in ptrs]
        except capi.DmbError as exc:
            ok, self.fallback_reason = 0, "export: %s" % exc
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, mine, group=self.group)
        opened, out = [], []
        if ok and all(e is not None for e in everyone):
            try:
                for b in range(len(ptrs)):
                    row = []
                    for r in range(self.world):
                        if r == self.rank:
                            row.append(ptrs[b])
                        else:
                            row.append(ctx.ipc_open(*everyone[r][b]))
                            opened.append(everyone[r][b][0])
                    out.append(row)
            except capi.DmbError as exc:
                ok, self.fallback_reason = 0, "open: %s" % exc
        else:
            ok = 0
        flag = self.torch.tensor([ok], device="cuda", dtype=self.torch.int32)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            for h in opened:
                ctx.ipc_close(h)
            return None, []
        return out, opened

    def all_reduce_sum(self, tensor):
        self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)

    def barrier(self):
        self.dist.barrier(group=self.group)

    def all_gather_device(self, out, shard):
        """out[r * len(shard) : (r + 1) * len(shard)] = rank r's ``shard`` (device tensors; NCCL over NVLink)."""
        k = shard.numel()
        self.dist.all_gather([out[r * k:(r + 1) * k] for r in range(self.world)], shard, group=self.group)

    def all_gather_host(self, arr):
        out = [None] * self.world
        self.dist.all_gather_object(out, arr, group=self.group)
        return out


class ShardedPauliEngine(PauliEngine):
    """``PauliEngine`` over a sharded state.  Every rank runs the same host logic on the same
    instruction stream, so schedules and collectives line up without negotiation."""

    def __init__(self, n_qubits, comm, lib=None, allocator=None, device=0, max_ops_per_pass=None,
                 reserve_low=None):
        self.comm = comm
        self.rank, self.world = comm.rank, comm.world
        self.plan_x = ExchangePlan(n_qubits, self.world, self.rank)
        self.g, self.m, self.n_loc = self.plan_x.g, self.plan_x.m, self.plan_x.n_loc
        self.n = int(n_qubits)
        self.nd = self.n_loc                     # digit positions the tile kernel may touch
        self.n_bits = self.plan_x.n_bits_local
        self.size = 1 << self.n_bits
        self.lib = lib if lib is not None else capi.load_library()
        self.alloc = allocator if allocator is not None else TorchCudaAllocator(device)
        self.ctx = shared_context(self.lib, getattr(self.alloc, "index", 0))
        self.ctx.set_stream(self.alloc.stream())
        self.ctx.reset_stats()
        self.state = self.alloc.empty(self.size)
        self.scratch = self.alloc.empty(self.size)
        self.pos = [self.n - 1 - q for q in range(self.n)]      # qubit -> slot
        self.pending = [None] * self.n
        self.queue = []
        self.max_ops_per_pass = int(max_ops_per_pass or os.environ.get("DMB_MAX_OPS_PER_PASS", capi.MAX_OPS))
        self.strategy = int(os.environ.get("DMB_SCHED_STRATEGY", capi.SCHED_TILE_SEARCH))
        self.reserve_low = int(reserve_low if reserve_low is not None else os.environ.get("DMB_RESERVE_LOW", 2))
        self.passes_run = 0
        self.h2d_bytes = 0
        self.drain_threshold = 0         # exchanges are collectives: compile whole queues
        self.relabel = False
        self.relabel_local = bool(int(os.environ.get("DMB_RELABEL", "1")))
        self.park_in_last_pass = bool(int(os.environ.get("DMB_PARK_IN_LAST_PASS", "1")))
        self.exchanges = 0
        self.direct_exchanges = 0        # exchanges that swapped a global slot with the evictee's own slot (no parking)
        self.nvlink_bytes_sent = 0
        # fused exchange: the pass after a slot swap pulls its tiles from the peers' buffers
        self.peers, self._ipc_handles = None, []
        if bool(int(os.environ.get("DMB_FUSED_EXCHANGE", "1"))) and callable(getattr(comm, "peer_addresses", None)):
            got = comm.peer_addresses(self.ctx, [self.alloc.ptr(self.state), self.alloc.ptr(self.scratch)])
            self.peers, self._ipc_handles = got if isinstance(got, tuple) else (got, [])
        self._cur = 0                    # which of the two registered buffers is `state`
        self._peers_may_read_scratch = False
        self.exchange_mode = os.environ.get("DMB_EXCHANGE", "pull")      # pull | push | nccl
        self.plain_exchange = bool(int(os.environ.get("DMB_EXCHANGE_PLAIN", "0")))   # op-free pull pass
        if self.exchange_mode == "nccl":
            self.peers = None
        # fused pull only: a global slot is swapped with the evicted qubit's OWN local slot (the pull pass gathers its
        # source table index from scattered index bits), so no pass is spent on parking the evictee on the top local
        # slots first.  DMB_DIRECT_SLOTS=0: always park (what the push / NCCL exchanges need).
        self.direct_slots = bool(int(os.environ.get("DMB_DIRECT_SLOTS", "1")))

    def recycle(self):
        """Make the engine ready for the next job on the same register size: the two shards and -- what costs 60 ms at
        8 ranks -- the peers' mappings of them stay, the per-job counters start over.  The layout, the pending maps and
        the queue are reset by ``init_product`` / ``upload``.  A pull that a slower peer may still be reading is
        remembered (``_peers_may_read_scratch``), so the next job's first write to the scratch shard still waits."""
        self.ctx.set_stream(self.alloc.stream())
        self.ctx.reset_stats()
        self.pos = [self.n - 1 - q for q in range(self.n)]
        self.pending = [None] * self.n
        self.queue = []
        self.passes_run = 0
        self.h2d_bytes = 0
        self.exchanges = 0
        self.direct_exchanges = 0
        self.nvlink_bytes_sent = 0
        if self.exchange_events is not None:
            self.exchange_events = []
        return self

    def close(self):
        """Release the peer mappings of this engine (reference counted in the library: the peer allocation is
        unmapped when its last user closes it).  Collective in effect: call it on every rank before the shards are
        freed; also runs when the engine is garbage collected."""
        handles, self._ipc_handles, self.peers = getattr(self, "_ipc_handles", []), [], None
        for h in handles:
            try:
                self.ctx.ipc_close(h)
            except Exception:
                pass

    def __del__(self):
        self.close()

    # -- layout -------------------------------------------------------------------------------
    def is_local(self, q):
        return self.pos[q] < self.n_loc

    def _hi_lo(self):
        return [2 * self.pos[q] + 1 for q in range(self.n)], [2 * self.pos[q] for q in range(self.n)]

    def init_product(self, vectors, scale):
        self.pos = [self.n - 1 - q for q in range(self.n)]
        hi, lo = self._hi_lo()
        self.ctx.init_product(self.sptr, self.n_bits, self.rank, hi, lo,
                              [list(map(float, v)) for v in vectors], scale)
        self.pending = [None] * self.n
        self.queue = []

    def upload(self, vec):
        """``stored_density_matrix``: every rank uploads its contiguous slice of the reference-order
        vector (the initial layout is the reference's, so slices are contiguous)."""
        vec = np.ascontiguousarray(vec, dtype=np.float64).reshape(-1)
        if vec.size != 4 ** self.n:
            raise BasicAerError("Wrong input stored density matrix")
        self.pos = [self.n - 1 - q for q in range(self.n)]
        part = np.ascontiguousarray(vec[self.rank * self.size:(self.rank + 1) * self.size])
        self.ctx.upload(self.sptr, part)
        self.h2d_bytes += part.nbytes
        self.pending = [None] * self.n
        self.queue = []

    # -- scheduling with exchanges ------------------------------------------------------------
    def compile(self, final=True):
        """Queue -> steps (see ``_compile_one``).  With the fused pull available the schedule is compiled up to three
        ways -- evictees parked on the top local slots (what every exchange mode can run), global slots swapped with the
        evictees' own slots, the same with the evictees chosen after the stretch is scheduled -- and the cheapest one
        by (passes + exchanges x their relative cost) is kept: which one wins depends on how the layout the exchange
        leaves behind suits the rest of the circuit (config 5: 25 -> 23 passes; QFT-16 on 8 ranks: parking wins).  An
        extra variant is only compiled while one compile (estimated from the op count, so that every rank decides
        alike) costs less than 1 % of the estimated device time."""
        forced = getattr(self, "force_exchange_variant", None) or os.environ.get("DMB_EXCHANGE_VARIANT")
        if forced:                                   # tests / A-B runs: "parked" | "direct" | "direct_late"
            steps = self._compile_one(final, forced)
            if any(st[0] == "exchange" for st in steps):
                self.last_compile_mode = forced
            return steps
        modes = ["parked"]
        if self._direct_exchange():
            modes += ["direct", "direct_late"]
        if len(modes) == 1:
            return self._compile_one(final, "parked")
        saved = (list(self.queue), list(self.pos), list(self.pending))
        pass_ms = 16.0 * 2.0 ** self.n_bits / 2.6e12 * 1e3            # a fused pass at ~0.4 of the HBM roofline
        ex_cost = 2.5 * (self.world - 1) / self.world                 # an exchange, in passes (NVLink ~600 GB/s)
        best = None
        for k, mode in enumerate(modes):
            self.queue, self.pos, self.pending = list(saved[0]), list(saved[1]), list(saved[2])
            steps = self._compile_one(final, mode)
            n_pass = sum(len(st[1]) for st in steps if st[0] == "passes")
            n_ex = sum(1 for st in steps if st[0] == "exchange")
            cost = n_pass + ex_cost * n_ex
            if best is None or cost < best[0]:
                best = (cost, steps, self.pos, self.pending, mode)
            # the decision must be the same on every rank: it depends on the schedule only, never on a measured time
            # (a compile takes ~10 us per op on the host; an extra one has to stay below ~1 % of the device time)
            n_ops = sum(int(p["n_ops"]) for st in steps if st[0] == "passes" for p in st[1])
            if n_ex == 0 or 0.010 * n_ops > 0.01 * cost * pass_ms:
                break
        self.queue = []
        self.pos, self.pending = best[2], best[3]
        if any(st[0] == "exchange" for st in best[1]):       # (a flush without exchanges says nothing about the choice)
            self.last_compile_mode = best[4]
        return best[1]

    def _compile_one(self, final, exchange_variant):
        """Queue (+ pending maps of LOCAL qubits if ``final``) -> list of steps
        ``("passes", PASS array)`` / ``("exchange",)``; updates ``self.pos`` to the layout
        after the last step and clears the queue.  Pending maps of qubits that end up global
        stay pending (readouts fold them in)."""
        steps = []
        queue = list(self.queue)
        pos = list(self.pos)
        n_loc, m = self.n_loc, self.m
        rounds = 0
        while True:
            rounds += 1
            if rounds > 2 * len(self.queue) + 8:          # every exchange must unblock at least one op
                raise BasicAerError("internal: sharded compile makes no progress")
            # ops runnable in the current layout: all qubits local, none blocked by a skipped op
            blocked, run, keep = set(), [], []
            for item in queue:
                _, kind, qa, qb, pa, pb, coef = item
                if qa in blocked or qb in blocked or pos[qa] >= n_loc or pos[qb] >= n_loc:
                    blocked.update((qa, qb))
                    keep.append(item)
                else:
                    run.append(item)
            use_relabel = self.relabel_local and self.nd >= 4
            chunks = []                      # PASS arrays of this exchange-free stretch
            victims = []
            if keep:
                # evict the local qubits whose next use is farthest away (never-used first)
                next_use = {}
                for idx, (_, _, qa, qb, _, _, _) in enumerate(keep):
                    next_use.setdefault(qa, idx)
                    next_use.setdefault(qb, idx)
                local = [q for q in range(self.n) if pos[q] < n_loc]
                local.sort(key=lambda q: (-next_use.get(q, len(keep) + 1), pos[q]))
                victims = local[:m]
            direct = bool(keep) and exchange_variant != "parked" and self._direct_exchange()
            moves = [] if direct else [(v, n_loc - m + i) for i, v in enumerate(victims)]
            if use_relabel:
                qops = [schedule.DevOp(kind, qa, qb, pa, pb, coef) for (_, kind, qa, qb, pa, pb, coef) in run]
                if not keep and final:
                    left = [q for q in range(self.n) if self.pending[q] is not None and pos[q] < n_loc]
                    left.sort(key=lambda q: pos[q])
                    for i in range(0, len(left) - 1, 2):
                        qops.append(schedule.DevOp(capi.OP_MATS, left[i], left[i + 1],
                                                   self.pending[left[i]], self.pending[left[i + 1]]))
                    if len(left) % 2:
                        qops.append(schedule.DevOp(capi.OP_MATS, left[-1], None, self.pending[left[-1]], None))
                    for q in left:
                        self.pending[q] = None
                if qops:
                    # the evictees are parked in the top local slots by trailing swaps of the stretch's
                    # last pass where its tile has room (folded into the write-back: free)
                    P, moves = schedule.relabel_passes(self.lib, qops, pos, self.nd, max_ops=self.max_ops_per_pass,
                                                       final_moves=moves if self.park_in_last_pass else [],
                                                       strategy=self.strategy)
                    if not self.park_in_last_pass and not direct:
                        moves = [(v, n_loc - m + i) for i, v in enumerate(victims)]
                    chunks.append(P)
                devops = []
            else:
                devops = [schedule.DevOp(kind, pos[qa], pos[qb], pa, pb, coef)
                          for (_, kind, qa, qb, pa, pb, coef) in run]
            queue = keep
            if queue:
                slot_owner = {pos[q]: q for q in range(self.n)}
                for v, target in moves:
                    if pos[v] != target:
                        other = slot_owner[target]
                        devops.append(schedule.DevOp(capi.OP_SWAP, pos[v], target))
                        slot_owner[pos[v]], slot_owner[target] = other, v
                        pos[other], pos[v] = pos[v], target
            elif final and not use_relabel:
                left = sorted((pos[q], q) for q in range(self.n) if self.pending[q] is not None and pos[q] < n_loc)
                for i in range(0, len(left) - 1, 2):
                    (da, qa), (db, qb) = left[i], left[i + 1]
                    devops.append(schedule.DevOp(capi.OP_MATS, da, db, self.pending[qa], self.pending[qb]))
                if len(left) % 2:
                    da, qa = left[-1]
                    devops.append(schedule.DevOp(capi.OP_MATS, da, None, self.pending[qa], None))
                for _, q in left:
                    self.pending[q] = None
            if devops:
                chunks.append(schedule.build_passes(devops, self.nd, max_ops=self.max_ops_per_pass,
                                                    reserve_low=self.reserve_low))
            if chunks:
                steps.append(("passes", np.concatenate(chunks) if len(chunks) > 1 else chunks[0]))
            if not queue:
                break
            # Slots 0 / 1 carry the 128-byte runs and cannot be swapped in place, and a qubit of the FIRST blocked op must
            # never be evicted (the exchange would not unblock it: no progress).  Nothing was parked, so evictees can
            # still be (re)chosen by the layout the stretch ended in.
            if direct and exchange_variant == "direct_late":
                late = [q for q in range(self.n) if 2 <= pos[q] < n_loc and next_use.get(q, 1) != 0]
                late.sort(key=lambda q: (-next_use.get(q, len(keep) + 1), pos[q]))
                if len(late) >= m:
                    victims = late[:m]
            elif direct and any(pos[v] < 2 for v in victims):
                spare = [q for q in range(self.n) if 2 <= pos[q] < n_loc and q not in victims and next_use.get(q, 1) != 0]
                spare.sort(key=lambda q: (-next_use.get(q, len(keep) + 1), pos[q]))
                victims = [v if pos[v] >= 2 else (spare.pop(0) if spare else v) for v in victims]
            if direct and all(pos[v] >= 2 for v in victims):
                # global slot n_loc + j <-> the slot victim j sits in (slots 0 and 1 hold the 128-byte runs: never)
                slots = [pos[v] for v in victims]
                steps.append(("exchange", slots))
                incoming = {pos[q] - n_loc: q for q in range(self.n) if pos[q] >= n_loc}
                for j, v in enumerate(victims):
                    pos[incoming[j]], pos[v] = slots[j], n_loc + j
                continue
            if direct:                           # a victim on slot 0 / 1: park this one time after all
                devops = []
                slot_owner = {pos[q]: q for q in range(self.n)}
                for i, v in enumerate(victims):
                    target = n_loc - m + i
                    if pos[v] != target:
                        other = slot_owner[target]
                        devops.append(schedule.DevOp(capi.OP_SWAP, pos[v], target))
                        slot_owner[pos[v]], slot_owner[target] = other, v
                        pos[other], pos[v] = pos[v], target
                if devops:
                    steps.append(("passes", schedule.build_passes(devops, self.nd, max_ops=self.max_ops_per_pass,
                                                                  reserve_low=self.reserve_low)))
            steps.append(("exchange",))
            for q in range(self.n):              # global slot s <-> local slot s - m
                if pos[q] >= n_loc:
                    pos[q] -= m
                elif pos[q] >= n_loc - m:
                    pos[q] += m
        self.queue = []
        self.pos = pos
        return steps

    def run_steps(self, steps, run_passes=None):
        """Execute compiled steps.  A slot swap is fused into an adjacent tile pass when the peers'
        buffers are mapped: ``pull`` fuses it into the pass AFTER the swap (remote loads), ``push``
        into the pass BEFORE it (remote stores).  ``run_passes`` lets a caller time the ordinary
        in-place passes."""
        rp = run_passes or self.run_passes
        mode = self.exchange_mode if self.peers is not None else "nccl"
        i, n = 0, len(steps)
        while i < n:
            st = steps[i]
            if st[0] == "passes":
                P = st[1]
                if mode == "push" and i + 1 < n and steps[i + 1][0] == "exchange" and len(P):
                    rp(P[:-1])
                    self.exchange(fused_pass=P[-1:])
                    i += 2
                    continue
                rp(P)
                i += 1
                continue
            slots = st[1] if len(st) > 1 else None
            if mode == "pull" and not self.plain_exchange and i + 1 < n and steps[i + 1][0] == "passes" \
                    and len(steps[i + 1][1]):
                nxt = steps[i + 1][1]
                self.exchange(fused_pass=nxt[:1], slots=slots)
                rp(nxt[1:])
                i += 2
                continue
            self.exchange(slots=slots)
            i += 1

    def _direct_exchange(self):
        """True when slot swaps run as the fused pull (peer-mapped buffers) and may therefore address arbitrary slots."""
        return (getattr(self, "direct_slots", False) and getattr(self, "peers", None) is not None
                and getattr(self, "exchange_mode", "pull") == "pull")

    def slot_swap_table(self, slots, bases):
        """Source table of a pull that swaps global slot n_loc + j with local slot slots[j].  Returns (sel_bits, tab,
        elements pulled from other ranks): element idx of the NEW local layout lies at  tab[k] + 8 * idx  where bit j of
        k is bit sel_bits[j] of idx -- the local index bits the swap moves; all other bits keep their place, so the
        source rank and the offset between old and new index depend on those bits only."""
        n_bits, n_loc = self.n_bits, self.n_loc
        sigma = {}
        for j, p in enumerate(slots):
            for half in (0, 1):
                a, b = 2 * (n_loc + j) + half, 2 * p + half
                sigma[a], sigma[b] = b, a
        sel = sorted(b for b in sigma if b < n_bits)
        tab = np.zeros(1 << len(sel), dtype=np.uint64)
        remote = 0
        for k in range(1 << len(sel)):
            idx = sum(((k >> j) & 1) << b for j, b in enumerate(sel))
            g_new = (self.rank << n_bits) | idx
            g_old = g_new
            for b, sb in sigma.items():          # bit b of the old index = bit sigma(b) of the new one
                g_old = (g_old & ~(1 << b)) | (((g_new >> sb) & 1) << b)
            src_rank, src_idx = g_old >> n_bits, g_old & ((1 << n_bits) - 1)
            tab[k] = (bases[src_rank] + 8 * (src_idx - idx)) % (1 << 64)
            remote += src_rank != self.rank
        return sel, tab, remote << (n_bits - len(sel))

    def exchange(self, fused_pass=None, slots=None):
        """Swap the m global slots with the m top local slots.  With peer-mapped buffers this is
        ONE tile-kernel launch: in ``pull`` mode it reads every tile from the rank holding it in
        the old layout (NVLink loads), applies ``fused_pass``'s ops and writes the new layout
        locally; in ``push`` mode it runs ``fused_pass`` in place on the old layout and stores
        every tile into the rank that owns it in the new layout (NVLink stores).  Without mapped
        peers: an out-of-place NCCL block exchange."""
        px = self.plan_x
        mode = self.exchange_mode if self.peers is not None else "nccl"
        if mode in ("pull", "push"):
            if fused_pass is None or len(fused_pass) == 0:
                fused_pass = np.zeros(1, dtype=capi.PASS_DTYPE)
                K = min(capi.MAX_TILE_DIGITS, self.nd)
                fused_pass[0]["n_tile_digits"] = K
                fused_pass[0]["tile_digit"][:K] = list(range(K))
            fused_pass = np.ascontiguousarray(fused_pass)
            tab = np.zeros(1 << px.block_bits, dtype=np.uint64)
            self.ctx.set_stream(self.alloc.stream())
            if mode == "pull":
                self.ctx.sync()
                self.comm.barrier()          # every rank's old buffer is final
                self._peers_may_read_scratch = False
                old = self.peers[self._cur]
                if slots is not None:        # global slots <-> the evictees' own slots
                    self.direct_exchanges += 1
                    sel, tab, pulled = self.slot_swap_table(slots, old)
                    ev = self._event_pair()
                    self.ctx.apply_pass_remote_sel(self.alloc.ptr(self.scratch), self.n_bits, fused_pass, tab, sel)
                    self._event_done(ev)
                    self.nvlink_bytes_sent += 8 * pulled - px.bytes_sent()     # (the common tail adds px.bytes_sent())
                else:
                    for d in range(1 << px.block_bits):
                        sr, sb = px.image(self.rank, d)
                        tab[d] = (old[sr] + ((sb - d) << px.B) * 8) % (1 << 64)
                    ev = self._event_pair()
                    self.ctx.apply_pass_remote(self.alloc.ptr(self.scratch), self.n_bits, fused_pass, tab, px.B)
                    self._event_done(ev)
            else:
                new = self.peers[self._cur ^ 1]      # the peers' idle buffers (free since the last barrier)
                for s_blk in range(1 << px.block_bits):
                    dr, db = px.image(self.rank, s_blk)
                    tab[s_blk] = (new[dr] + ((db - s_blk) << px.B) * 8) % (1 << 64)
                self.ctx.apply_pass_remote(self.sptr, self.n_bits, fused_pass, tab, px.B, push=True)
                self.ctx.sync()
                self.comm.barrier()          # all remote stores have landed
            self.passes_run += 1
            self._cur ^= 1
            self._peers_may_read_scratch = mode == "pull"     # a slower peer may still be pulling from the old buffer
        else:
            self.comm.exchange(px, self.state, self.scratch)
        self.state, self.scratch = self.scratch, self.state
        self.exchanges += 1
        self.nvlink_bytes_sent += px.bytes_sent()

    # optional CUDA-event timing of the exchange launches (bench.py: NVLink GB/s per exchange)
    exchange_events = None

    def alloc_has_events(self):
        t = getattr(self.alloc, "torch", None)
        return t is not None and t.cuda.is_available()

    def _event_pair(self):
        if self.exchange_events is None:
            return None
        t = self.alloc.torch
        a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a.record()
        return a, b

    def _event_done(self, ev):
        if ev is not None:
            ev[1].record()
            self.exchange_events.append(ev)

    def _own_scratch(self):
        """Call before writing the scratch shard outside an exchange.  After a fused pull the
        scratch shard is the buffer the peers read their new layout from; a rank that finishes its
        own pull first must not overwrite it until every rank has (a readout that uses the scratch
        shard as workspace would otherwise corrupt a slower peer's state)."""
        if self._peers_may_read_scratch:
            self.ctx.sync()
            self.comm.barrier()
            self._peers_may_read_scratch = False

    MAX_FOLDED_MAPS = 4      # DMB_MAX_MULTI of the marginal kernel: qubits whose pending map a readout can fold in

    def _drop_map_only_tail(self, steps, pending_before):
        """A flush that ends with a pass of single-qubit maps only (the pending maps no earlier pass had room for) pays
        a full round trip of the shard for them.  The I/B-marginal readout folds pending maps into its weights -- it
        does so for the global qubits anyway -- so when that readout comes next and at most MAX_FOLDED_MAPS qubits
        would carry one, the pass is dropped and its maps stay pending."""
        if not getattr(self, "fold_tail_maps", True) or not steps or steps[-1][0] != "passes" or not len(steps[-1][1]):
            return steps
        P = steps[-1][1]
        last = P[-1]
        ops = last["ops"][:int(last["n_ops"])]
        if any(int(o["kind"]) != capi.OP_MATS for o in ops):
            return steps
        owner = {self.pos[q]: q for q in range(self.n)}      # no swap in this pass: its layout is the final one
        restored = {}
        for o in ops:
            for flag, digit, rows in ((capi.HAS_PA, o["a"], o["pa"]), (capi.HAS_PB, o["b"], o["pb"])):
                if not (int(o["flags"]) & flag):
                    continue
                q = owner.get(int(last["tile_digit"][int(digit)]))
                before = None if q is None else pending_before[q]
                if before is None or q in restored or not np.array_equal(np.asarray(before)[1:4, :].reshape(12), rows):
                    return steps
                restored[q] = before
        carrying = sum(1 for q in range(self.n) if self.pending[q] is not None)
        if not restored or carrying + len(restored) > self.MAX_FOLDED_MAPS:
            return steps
        for q, M in restored.items():
            self.pending[q] = M
        return steps[:-1] + ([("passes", P[:-1])] if len(P) > 1 else [])

    def flush(self, keep_tail_maps=False):
        saved = list(self.pos)
        before = list(self.pending)
        steps = self.compile(final=True)
        if keep_tail_maps:
            steps = self._drop_map_only_tail(steps, before)
        final_pos = self.pos
        self.pos = saved                         # run_steps does not depend on pos; keep it coherent on error
        self.run_steps(steps)
        self.pos = final_pos

    def plan(self):
        raise BasicAerError("use compile() on a sharded engine")

    # -- readouts -----------------------------------------------------------------------------
    def marginal_probabilities(self, basis, err):
        self.flush(keep_tail_maps=True)
        b = {"X": 1, "Y": 2, "Z": 3}[basis]
        n = self.n
        hi, lo, wt = [], [], np.zeros((n, 2, 4))
        for k in range(n):
            q = n - 1 - k
            hi.append(2 * self.pos[q] + 1)
            lo.append(2 * self.pos[q])
            P = self.pending[q] if self.pending[q] is not None else np.eye(4)
            wt[k, 0, :] = P[0, :]
            wt[k, 1, :] = err * P[b, :]
        out = self.alloc.empty(2 ** n)
        self.ctx.marginal(self.sptr, self.n_bits, self.rank, hi, lo, wt, self.alloc.ptr(out))
        self.ctx.sync()
        self.comm.all_reduce_sum(out)
        self.ctx.fwht(self.alloc.ptr(out), n)
        host = np.empty(2 ** n)
        self.ctx.download(self.alloc.ptr(out), host)
        return host

    def n_basis_probabilities(self, nvec, err):
        """'N'-basis ensemble readout (``dm_simulator.py:446-458``) of a sharded state.

        Every rank contracts its local digit slots on the device, one ``contract_digit`` launch per
        slot: (I, n.(X,Y,Z)) -> 2 values, so the shard shrinks from 2^n_bits to 2^n_loc (x2 for the
        half digit when the rank count is an odd power of two).  The global slots have one fixed
        digit value per rank: their weights (with any pending single-qubit map folded in) scale
        the rank's partial vector into its share of the 2^n marginal, which is summed with one
        all-reduce and Walsh-Hadamard transformed on the device."""
        self.flush()
        n, n_loc = self.n, self.n_loc
        unit = np.asarray(nvec, dtype=float)
        owner = {self.pos[q]: q for q in range(n)}
        self._own_scratch()
        src = self.sptr
        base = self.alloc.ptr(self.scratch)
        offset, count = 0, self.size
        for slot in range(n_loc):
            if self.pending[owner[slot]] is not None:
                raise BasicAerError("internal: local qubit with a pending map after flush")
            count //= 2
            dst = base + 8 * offset
            self.ctx.contract_digit(src, dst, 1 << (self.n_bits - 2 * (slot + 1)), 1 << slot, unit * err)
            src = dst
            offset = count if offset == 0 else 0        # ping-pong inside the scratch shard
        rem = 1 << (self.n_bits - 2 * n_loc)             # 1, or 2 = low bit of the straddling digit
        partial = np.empty(rem << n_loc)
        self.ctx.download(src, partial)
        cur = partial.reshape(rem, 1 << n_loc)

        def weights(slot):
            P = self.pending[owner[slot]]
            P = np.eye(4) if P is None else np.asarray(P, dtype=float)
            return np.stack([P[0, :], err * (unit @ P[1:4, :])])           # [result bit][digit value]

        g_index = self.rank << self.n_bits
        slot = n_loc
        if rem == 2:                                     # digit = 2 * (rank bit) + (local bit)
            W = weights(slot)
            hi = (g_index >> (2 * slot + 1)) & 1
            cur = np.stack([W[c, 2 * hi] * cur[0] + W[c, 2 * hi + 1] * cur[1] for c in (0, 1)]).reshape(-1)
            slot += 1
        else:
            cur = cur.reshape(-1)
        while slot < n:
            W = weights(slot)
            d = (g_index >> (2 * slot)) & 3
            cur = np.concatenate([W[0, d] * cur, W[1, d] * cur])
            slot += 1
        # index bit p of `cur` <-> slot p; the result wants bit n-1-q <-> qubit q
        arr = cur.reshape([2] * n)                       # axis j <-> slot n-1-j
        arr = np.ascontiguousarray(np.transpose(arr, [n - 1 - self.pos[q] for q in range(n)])).reshape(-1)
        out = self.alloc.empty(2 ** n)
        self.ctx.upload(self.alloc.ptr(out), arr)
        self.ctx.sync()
        self.comm.all_reduce_sum(out)
        self.ctx.fwht(self.alloc.ptr(out), n)
        host = np.empty(2 ** n)
        self.ctx.download(self.alloc.ptr(out), host)
        return host

    def read_flat(self, flat_indices):
        """Coefficients at flat indices of the current (global) slot layout: the owning rank reads,
        one all-reduce (Expect / Bell readouts, ``dm_simulator.py:569-570, 744-760``; reduced states)."""
        self.settle()
        g = np.asarray(flat_indices, dtype=np.uint64)
        mine = (g >> np.uint64(self.n_bits)) == np.uint64(self.rank)
        local_idx = g & np.uint64((1 << self.n_bits) - 1)
        vals = self.ctx.read_coeffs(self.sptr, local_idx)
        vals = np.where(mine, vals, 0.0)
        t = self.alloc.empty(len(vals))
        self.ctx.upload(self.alloc.ptr(t), vals)
        self.comm.all_reduce_sum(t)
        out = np.empty(len(vals))
        self.ctx.download(self.alloc.ptr(t), out)
        return out

    def settle(self):
        self._localise_all_pending()

    def to_matrix(self):
        """``_compute_densitymatrix`` of a sharded state (registers small enough that the 2^n x 2^n matrix is wanted
        at all): the state is brought into the reference order, the shards are all-gathered DEVICE to device
        (NCCL over NVLink -- no host round trip, no pickling) and every rank's GPU runs the Pauli->matrix kernels
        on the full vector."""
        n = self.n
        if n > 14 or 3 * 16 * 4 ** n > (64 << 30):
            raise BasicAerError("compute_densitymatrix on %d sharded qubits would need a %d GiB matrix; pass "
                                "compute_densitymatrix=False" % (n, (16 * 4 ** n) >> 30))
        self.canonicalise()
        src = self.alloc.empty(4 ** n)
        if callable(getattr(self.comm, "all_gather_device", None)):
            self.ctx.sync()
            self.comm.all_gather_device(src, self.state)
        else:
            self.ctx.upload(self.alloc.ptr(src), self.download())
        work = self.alloc.empty(2 * 4 ** n)
        out = self.alloc.empty(2 * 4 ** n)
        self.ctx.to_matrix(self.alloc.ptr(src), n, self.alloc.ptr(work), self.alloc.ptr(out))
        host = np.empty(2 * 4 ** n)
        self.ctx.download(self.alloc.ptr(out), host)
        return host.view(np.complex128).reshape(2 ** n, 2 ** n)

    def _localise_all_pending(self):
        """Apply pending maps of global qubits too (costs exchanges): needed before the state
        itself (not just a marginal) is read."""
        self.flush()
        guard = 0
        while any(self.pending[q] is not None for q in range(self.n)):
            self.exchange()
            for q in range(self.n):
                if self.pos[q] >= self.n_loc:
                    self.pos[q] -= self.m
                elif self.pos[q] >= self.n_loc - self.m:
                    self.pos[q] += self.m
            self.flush()
            guard += 1
            if guard > 4:
                raise BasicAerError("internal: pending maps could not be localised")

    # -- canonical (reference-order) layout: rank r holds the contiguous slice [r * size, (r + 1) * size) ----
    def _exchange_and_relabel(self):
        self.exchange()
        for q in range(self.n):
            if self.pos[q] >= self.n_loc:
                self.pos[q] -= self.m
            elif self.pos[q] >= self.n_loc - self.m:
                self.pos[q] += self.m

    def _local_permute(self, wanted):
        """Move qubit q to LOCAL slot wanted[q] (dict; all slots < n_loc) with SWAP ops in ordinary tile passes;
        the displaced qubits take the vacated slots."""
        pos = list(self.pos)
        owner = {pos[q]: q for q in range(self.n)}
        ops = []
        for q, t in sorted(wanted.items(), key=lambda kv: -kv[1]):
            if pos[q] >= self.n_loc or t >= self.n_loc:
                raise BasicAerError("internal: _local_permute on a global slot")
            if pos[q] != t:
                other = owner[t]
                ops.append(schedule.DevOp(capi.OP_SWAP, pos[q], t))
                owner[pos[q]], pos[other] = other, pos[q]
                owner[t], pos[q] = q, t
        if ops:
            self.run_passes(schedule.build_passes(ops, self.nd, max_ops=self.max_ops_per_pass,
                                                  reserve_low=self.reserve_low))
        self.pos = pos

    def canonicalise(self):
        """Bring the sharded state into the reference's order (qubit q at slot n-1-q), so that rank r's shard is
        the contiguous slice [r * 4^n / G, (r+1) * 4^n / G) of ``DmSimulatorPy._densitymatrix`` -- what per-rank
        dumps, per-rank uploads and the device-side all-gather need.  At most two slot exchanges: one to bring
        wrongly placed global qubits home, one to send qubits 0 .. m-1 out in order; everything else is SWAP ops
        in ordinary tile passes."""
        self._localise_all_pending()
        n, n_loc, m = self.n, self.n_loc, self.m
        target = [n - 1 - q for q in range(n)]
        if self.pos == target:
            return
        outer = [q for q in range(n) if target[q] >= n_loc]          # the qubits that belong in the global slots
        if any(self.pos[q] != target[q] for q in outer):
            if any(self.pos[q] >= n_loc for q in outer):
                # some of them are global but in the wrong slot: bring the global slots home first, after making
                # sure none of `outer` sits in the top local slots (those go out in the same exchange)
                low = [s for s in range(n_loc - m) if all(self.pos[q] != s for q in outer)]
                parked = {}
                for q in outer:
                    if n_loc - m <= self.pos[q] < n_loc:
                        parked[q] = low.pop(0)
                self._local_permute(parked)
                self._exchange_and_relabel()
            self._local_permute({q: target[q] - m for q in outer})
            self._exchange_and_relabel()
        self._local_permute({q: target[q] for q in range(n) if target[q] < n_loc})
        if self.pos != target:
            raise BasicAerError("internal: canonicalise did not reach the reference layout")

    def download_shard(self, out=None):
        """This rank's contiguous slice of the reference-order coefficient vector (no gather)."""
        self.canonicalise()
        shard = out if out is not None else np.empty(self.size)
        self.ctx.download(self.sptr, shard)
        return shard

    @staticmethod
    def shard_file(stem, rank, world):
        """Per-rank file name of a sharded dump: ``<stem>.<rank>-of-<world>.npy`` (rank r holds elements
        [r * 4^n / world, (r+1) * 4^n / world) of the reference-order vector)."""
        return "%s.%03d-of-%03d.npy" % (stem, rank, world)

    def store_shards(self, stem):
        """``_store_density_matrix`` (``dm_simulator.py:1271-1275``) at sharded sizes: every rank writes its own
        slice; nothing is gathered.  ``join_shard_files`` rebuilds the single file the reference writes."""
        np.save(self.shard_file(stem, self.rank, self.world), self.download_shard())
        self.comm.barrier()

    def load_slice(self, stem):
        """This rank's slice of a stored vector: from its shard file when a sharded dump exists, else memory-mapped
        out of the single ``<stem>.npy`` the reference (or a one-GPU run) wrote.  Returns None when neither exists."""
        import os as _os
        name = self.shard_file(stem, self.rank, self.world)
        if _os.path.exists(name):
            part = np.load(name)
        elif _os.path.exists(stem + ".npy"):
            full = np.load(stem + ".npy", mmap_mode="r")
            if full.size != 4 ** self.n:
                raise BasicAerError("stored coefficients have the wrong length")
            part = np.ascontiguousarray(full.reshape(-1)[self.rank * self.size:(self.rank + 1) * self.size])
        else:
            return None
        if part.size != self.size:
            raise BasicAerError("stored coefficients have the wrong length")
        return np.ascontiguousarray(part, dtype=np.float64)

    def upload_slice(self, part):
        """Start from a stored state given as this rank's slice of the reference-order vector."""
        self.pos = [self.n - 1 - q for q in range(self.n)]
        self.ctx.upload(self.sptr, part)
        self.h2d_bytes += part.nbytes
        self.pending = [None] * self.n
        self.queue = []

    def overlap_with_slice(self, part):
        """dot(stored, state) from per-rank slices: canonical layout, per-shard device reduction, one all-reduce."""
        self.canonicalise()
        self._own_scratch()
        self.ctx.upload(self.alloc.ptr(self.scratch), part)
        local = self.ctx.dot(self.alloc.ptr(self.scratch), self.sptr, self.size)
        t = self.alloc.empty(1)
        self.ctx.upload(self.alloc.ptr(t), np.array([local]))
        self.comm.all_reduce_sum(t)
        out = np.empty(1)
        self.ctx.download(self.alloc.ptr(t), out)
        return float(out[0])

    def chop(self, thr):
        self._localise_all_pending()
        self.ctx.chop(self.sptr, self.size, thr)

    def download(self, out=None):
        """Gather the shards and undo the slot permutation on the host (small n only: this is
        result formatting for the 'coeffmatrix' entry, not part of the hot path)."""
        self._localise_all_pending()
        shard = np.empty(self.size)
        self.ctx.download(self.sptr, shard)
        shards = self.comm.all_gather_host(shard)
        full = np.concatenate(shards).reshape([4] * self.n)     # axis j <-> slot n-1-j
        # slot p holds qubit q with pos[q] == p; reference order wants qubit q on axis q
        axes = [self.n - 1 - self.pos[q] for q in range(self.n)]
        vec = np.ascontiguousarray(np.transpose(full, axes)).reshape(-1)
        if out is not None:
            out[:] = vec
            return out
        return vec

    def _to_current_layout(self, vec):
        """Reference-order 4^n vector -> this rank's slice in the current slot layout."""
        full = np.asarray(vec, dtype=np.float64).reshape([4] * self.n)          # axis q <-> qubit q
        # target axis j holds slot n-1-j, i.e. the qubit q with pos[q] == n-1-j
        owner = {self.pos[q]: q for q in range(self.n)}
        axes = [owner[self.n - 1 - j] for j in range(self.n)]
        flat = np.ascontiguousarray(np.transpose(full, axes)).reshape(-1)
        return np.ascontiguousarray(flat[self.rank * self.size:(self.rank + 1) * self.size])

    def overlap_with(self, other_vec):
        """dot(other, state) over all shards (``_state_overlap``, ``:1277-1282``): the stored vector
        is permuted to the current layout on the host, each rank reduces its slice on the device,
        one all-reduce."""
        self._localise_all_pending()
        other_vec = np.ascontiguousarray(other_vec, dtype=np.float64).reshape(-1)
        if other_vec.size != 4 ** self.n:
            raise BasicAerError("stored coefficients have the wrong length")
        self._own_scratch()
        self.ctx.upload(self.alloc.ptr(self.scratch), self._to_current_layout(other_vec))
        part = self.ctx.dot(self.alloc.ptr(self.scratch), self.sptr, self.size)
        t = self.alloc.empty(1)
        self.ctx.upload(self.alloc.ptr(t), np.array([part]))
        self.comm.all_reduce_sum(t)
        out = np.empty(1)
        self.ctx.download(self.alloc.ptr(t), out)
        return float(out[0])


class ShardedCircuitRunner:
    """bench.py helper: compile the circuit once (the backend's own level loop on an engine that does not launch),
    then time init -> steps -> marginal."""

    def __init__(self, n, circ, opts, device=0, comm=None, engine=None):
        import copy
        from . import engine as eng, hostpass
        from .dm_simulator import DmSimulatorB200
        self.n = n
        self.comm = comm or TorchCommunicator()
        self.engine = e = engine or ShardedPauliEngine(n, self.comm, device=device)
        be = DmSimulatorB200(device=device, _engine_factory=lambda nq: e)
        be._set_options(None, copy.deepcopy(opts))
        be._number_of_qubits = n
        be._initialize_errors()
        ops = hostpass.merge_single_qubit_gates(circ.instructions, n, be.MERGE)
        levels, self.n_levels = hostpass.partition_levels(ops, n)
        mem = be._error_params["memory"]
        noise = eng.memory_noise_matrix(mem["decoherence"], mem["thermalization"], mem["amplitude_decay"])
        noisy = not np.array_equal(noise, np.eye(4))
        for level in levels[:self.n_levels]:
            be._run_level(e, level, None, {})
            if noisy:
                e.apply_1q_all(noise)
        before = list(e.pending)
        self.steps = e._drop_map_only_tail(e.compile(final=True), before)      # what flush() does ahead of the marginal
        self.final_pos = list(e.pos)
        self.final_pending = list(e.pending)
        self.err = be._error_params["measurement"]
        self._ev = []

    def step(self):
        import torch
        e = self.engine
        e.init_product([[1, 0, 0, 1]] * self.n, 0.5 ** self.n)
        cuda = torch.cuda.is_available()

        def timed(passes):
            if len(passes) == 0:
                return
            if cuda:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                e.run_passes(passes)
                b.record()
                self._ev.append((a, b, len(passes)))
            else:
                e.run_passes(passes)

        e.run_steps(self.steps, run_passes=timed)
        e.pos = list(self.final_pos)
        e.pending = list(self.final_pending)
        self.probs = e.marginal_probabilities("Z", self.err)

    def reset_counters(self):
        self.engine.ctx.sync()
        self.engine.ctx.reset_stats()
        self.engine.exchanges = 0
        self.engine.nvlink_bytes_sent = 0
        self.engine.exchange_events = [] if self.engine.alloc_has_events() else None
        self._ev = []

    def counters(self):
        return self.engine.stats()

    def pass_ms_total(self):
        import torch
        torch.cuda.synchronize()
        return sum(ev[0].elapsed_time(ev[1]) for ev in self._ev)

    def timed_pass_launches(self):
        return sum(ev[2] for ev in self._ev)

    def exchange_stats(self):
        """NVLink traffic of the timed region: bytes each GPU pulled per slot swap and the rate it arrived at
        (CUDA events around the fused exchange launch, which also applies that pass's ops)."""
        e = self.engine
        if not e.exchanges:
            return {"exchanges": 0}
        out = {"exchanges": e.exchanges, "mode": e.exchange_mode if e.peers is not None else "nccl",
               "bytes_per_gpu_per_exchange": e.nvlink_bytes_sent // e.exchanges,
               "nvlink_bytes_sent_per_gpu": e.nvlink_bytes_sent,
               # which schedule compile() kept: evictees parked on the top local slots, or global slots swapped with
               # the evictees' own slots (direct / direct_late), and how many of the exchanges ran that way
               "schedule_variant": getattr(e, "last_compile_mode", "parked"),
               "direct_slot_exchanges": getattr(e, "direct_exchanges", 0)}
        evs = getattr(e, "exchange_events", None)
        if evs:
            import torch
            torch.cuda.synchronize()
            ms = [a.elapsed_time(b) for a, b in evs]
            out["avg_exchange_ms"] = sum(ms) / len(ms)
            out["inbound_gbps_per_gpu"] = out["bytes_per_gpu_per_exchange"] / (out["avg_exchange_ms"] * 1e-3) / 1e9
            out["nvlink5_line_rate_gbps"] = 900.0
        return out

    def close(self):
        """Unmap the peers' buffers on every rank, THEN free this rank's shards (CUDA IPC: exported memory must
        outlive its importers' mappings)."""
        import torch
        if self.engine is not None:
            self.engine.ctx.sync()
            self.engine.close()
            self.comm.barrier()
        self.engine = None
        self.steps = None
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
            self.comm.barrier()
