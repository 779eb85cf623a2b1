"""Device-resident Pauli-coefficient state and the lazy op queue in front of the CUDA kernels.

``PauliEngine`` owns what the reference keeps in ``DmSimulatorPy._densitymatrix``
(``dm_simulator.py:61-68``): a float64 vector of 4^n Pauli coefficients, here a PyTorch
CUDA tensor (PyTorch is used for buffer ownership and streams only).  Its methods mirror
the reference's state-mutating methods one to one but are *lazy*:

* every single-qubit map -- u1/u3 (a6-a8), per-level memory noise (a11), projective
  measurements (a19, a22), reset (a24) -- is multiplied on the host into a pending 4x4
  matrix of its qubit (exact: single-qubit maps on different qubits commute, and all of
  them are linear maps on one base-4 digit);
* a two-qubit op (CNOT a9/a10, Bell mask a23) takes the pending matrices of its two qubits
  with it and is appended to a queue;
* anything that must *see* the state (readouts, final state) first flushes: the queue is
  packed into tile passes by ``schedule.build_passes`` and executed by
  ``dmb_apply_passes`` -- one HBM round trip per pass instead of the reference's ~6 sweeps
  per U3, ~2 per CNOT and n per noise level.

Layout: qubit q sits at digit position ``pos[q]`` (reference layout: n-1-q, qubit 0 most
significant).  A 1-qubit register is padded with one phantom identity digit so that every
state has at least two digit positions (the kernels work on digit pairs).
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import capi, schedule
from .exceptions import BasicAerError


# --------------------------------------------------------------------------------------
# single-qubit maps on (I, X, Y, Z)
# --------------------------------------------------------------------------------------

_AXIS_PAIR = {"rz": (1, 2), "ry": (3, 1), "rx": (2, 3)}


def rotation_matrix(axis, angle, err):
    """``rot_gate_dm_matrix`` (``basicaertools.py:93-126``) as a 4x4 matrix:
    c = r cos(angle+delta), s = r sin(angle+delta), (r, delta) = err."""
    c = err[0] * np.cos(angle + err[1])
    s = err[0] * np.sin(angle + err[1])
    k0, k1 = _AXIS_PAIR[axis]
    m = np.eye(4)
    m[k0, k0], m[k0, k1] = c, -s
    m[k1, k1], m[k1, k0] = c, s
    return m


def gate_matrix(name, params, rotation_error):
    """u3(theta,phi,lam) = rz(lam), ry(theta), rz(phi) applied in that order; u1(lam) = rz(lam)
    (``single_gate_dm_matrix``, ``basicaertools.py:69-90``).  The product of the three
    ``rotation_matrix`` factors is written out (same association as multiplying them one after the
    other), because building and multiplying 4x4 arrays per gate dominated the lowering time."""
    p = [float(x) for x in params]
    rz, dz = rotation_error["rz"]
    if name == "u1":
        cl, sl = rz * math.cos(p[0] + dz), rz * math.sin(p[0] + dz)
        return np.array([[1.0, 0.0, 0.0, 0.0], [0.0, cl, -sl, 0.0], [0.0, sl, cl, 0.0], [0.0, 0.0, 0.0, 1.0]])
    if name not in ("U", "u3"):
        raise BasicAerError("Gate is not among the valid types: %s" % name)
    ry, dy = rotation_error["ry"]
    cl, sl = rz * math.cos(p[2] + dz), rz * math.sin(p[2] + dz)      # rz(lam):   X' = cX - sY, Y' = cY + sX
    ct, st = ry * math.cos(p[0] + dy), ry * math.sin(p[0] + dy)      # ry(theta): Z' = cZ - sX, X' = cX + sZ
    cp, sp = rz * math.cos(p[1] + dz), rz * math.sin(p[1] + dz)      # rz(phi)
    ax = (ct * cl, ct * -sl, st)              # rows of ry(theta) @ rz(lam)
    ay = (sl, cl, 0.0)
    return np.array([[1.0, 0.0, 0.0, 0.0],
                     [0.0, cp * ax[0] - sp * ay[0], cp * ax[1] - sp * ay[1], cp * ax[2]],
                     [0.0, sp * ax[0] + cp * ay[0], sp * ax[1] + cp * ay[1], sp * ax[2]],
                     [0.0, -st * cl, -st * -sl, ct]])


def memory_noise_matrix(f, p, g):
    """``_add_decoherence_and_amp_decay`` (``dm_simulator.py:397-425``) on one qubit."""
    off = np.sqrt(g) * f
    m = np.diag([1.0, off, off, g])
    m[3, 0] = (1 - g) * (2 * p - 1)
    return m


def measure_axis_matrix(basis, err):
    """``_add_qasm_measure_X/Y/Z`` (``dm_simulator.py:574-664``)."""
    m = np.zeros((4, 4))
    m[0, 0] = 1.0
    k = {"X": 1, "Y": 2, "Z": 3}[basis]
    m[k, k] = err
    return m


def measure_n_matrix(nvec, err):
    """``_add_qasm_measure_N`` (``dm_simulator.py:666-704``)."""
    nvec = np.asarray(nvec, dtype=float)
    m = np.zeros((4, 4))
    m[0, 0] = 1.0
    m[1:, 1:] = np.outer(nvec, nvec * err)
    return m


def reset_matrix():
    """``_add_qasm_reset`` (``dm_simulator.py:810-823``): X = Y = 0, Z = I."""
    m = np.zeros((4, 4))
    m[0, 0] = 1.0
    m[3, 0] = 1.0
    return m


_CX_COEF_CACHE = {}


def cx_coefficients(tsp):
    """(c, s, c2, s2, cs) of ``cx_gate_dm_matrix`` (``basicaertools.py:329-336``)."""
    key = (float(tsp[0]), float(tsp[1]))
    hit = _CX_COEF_CACHE.get(key)
    if hit is None:
        if len(_CX_COEF_CACHE) > 64:
            _CX_COEF_CACHE.clear()
        hit = _CX_COEF_CACHE[key] = _cx_coefficients(key)
    return hit


def _cx_coefficients(tsp):
    cav, e1 = float(tsp[0]), float(tsp[1])
    c2av = 4 * cav - 3
    c = cav * np.cos(e1)
    s = cav * np.sin(e1)
    c2 = 0.5 * (1 + c2av * np.cos(2 * e1))
    s2 = 0.5 * (1 - c2av * np.cos(2 * e1))
    cs = c2av * np.sin(e1) * np.cos(e1)
    return c, s, c2, s2, cs


# --------------------------------------------------------------------------------------
# buffers
# --------------------------------------------------------------------------------------

class TorchCudaAllocator:
    """Device buffers are PyTorch CUDA tensors; kernels run on torch's current stream."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise capi.DmbError("no CUDA device: the dm_simulator B200 backend needs a GPU (no CPU fallback)")
        self.torch = torch
        self.device = torch.device("cuda", int(device))
        self.index = int(device)

    def empty(self, count):
        return self.torch.empty(int(count), dtype=self.torch.float64, device=self.device)

    def ptr(self, buf):
        return buf.data_ptr()

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def pinned(self, count):
        t = self.torch.empty(int(count), dtype=self.torch.float64, pin_memory=True)
        return t, t.numpy()


# --------------------------------------------------------------------------------------
# engine
# --------------------------------------------------------------------------------------

_CONTEXTS = {}


def shared_context(lib, device_index):
    """One ``dmb_ctx`` per (library, device), reused by successive engines: creating one costs
    two cudaMallocs and a pinned allocation, which dominates the run time of small circuits."""
    key = (id(lib), int(device_index))
    ctx = _CONTEXTS.get(key)
    if ctx is None:
        ctx = _CONTEXTS[key] = capi.Context(lib, device_index)
    return ctx


class PauliEngine:
    # plan recording (dm_simulator.CompiledCircuit): when `tape` is a list every device-facing step is appended to it,
    # so that the same circuit can be re-run later without any host-side lowering (class defaults: off)
    tape = None
    tape_outputs = 0
    tape_valid = True
    last_output = None

    def __init__(self, n_qubits, lib=None, allocator=None, device=0, max_ops_per_pass=None,
                 reserve_low=None, relabel=None):
        if n_qubits < 1 or n_qubits > capi.MAX_QUBITS:
            raise BasicAerError("number of qubits out of range: %d" % n_qubits)
        self.n = int(n_qubits)
        self.nd = max(self.n, 2)                 # digit positions incl. phantom padding
        self.n_bits = 2 * self.nd
        self.size = 4 ** self.nd
        self.lib = lib if lib is not None else capi.load_library()
        self.alloc = allocator if allocator is not None else TorchCudaAllocator(device)
        self.ctx = shared_context(self.lib, getattr(self.alloc, "index", 0))
        self.ctx.set_stream(self.alloc.stream())
        self.ctx.reset_stats()
        self.state = self.alloc.empty(self.size)
        self.pos = [self.n - 1 - q for q in range(self.n)]     # qubit -> digit position
        self.pending = [None] * self.n
        self.queue = []
        # scheduling knobs (env overrides are for on-GPU experiments)
        self.max_ops_per_pass = int(max_ops_per_pass or os.environ.get("DMB_MAX_OPS_PER_PASS", capi.MAX_OPS))
        self.strategy = int(os.environ.get("DMB_SCHED_STRATEGY", capi.SCHED_TILE_SEARCH))
        self.reserve_low = int(reserve_low if reserve_low is not None else os.environ.get("DMB_RESERVE_LOW", 2))
        self.passes_run = 0
        self.h2d_bytes = 0
        # launch queued two-qubit ops as soon as this many have accumulated, so that the GPU
        # works while the host is still lowering later levels (0 = only at readouts)
        # (the threshold doubles after every drain: start the GPU early, fuse over long windows later)
        self.drain_threshold = int(os.environ.get("DMB_DRAIN_THRESHOLD", 64))
        self.drain_threshold_max = 1024
        # streaming drain (relabel scheduler): every `drain_chunk` new ops, launch all passes but
        # keep the newest `drain_tail` ops queued as look-ahead
        self.drain_tail = int(os.environ.get("DMB_DRAIN_TAIL", max(64, 6 * self.n)))
        self.drain_chunk = int(os.environ.get("DMB_DRAIN_CHUNK", 64))
        self._chunk_now = min(16, self.drain_chunk)      # small first chunks: the GPU starts early
        # dynamic relabelling of the two low digit positions (schedule.build_passes_relabel)
        self.relabel = bool(int(os.environ.get("DMB_RELABEL", "1"))) if relabel is None else bool(relabel)
        if os.environ.get("DMB_TILE_VARIANT"):
            self.ctx.set_tile_variant(int(os.environ["DMB_TILE_VARIANT"]))

    # -- helpers -----------------------------------------------------------------------
    @property
    def sptr(self):
        return self.alloc.ptr(self.state)

    def _hi_lo(self):
        hi = [2 * self.pos[q] + 1 for q in range(self.n)]
        lo = [2 * self.pos[q] for q in range(self.n)]
        return hi, lo

    # -- a4: initial states --------------------------------------------------------------
    def init_product(self, vectors, scale):
        """state = scale * kron_q vectors[q]  (``_initialize_densitymatrix``, ``:284-349``)."""
        self.pos = [self.n - 1 - q for q in range(self.n)]
        hi, lo = self._hi_lo()
        v = [list(map(float, vec)) for vec in vectors]
        if self.nd > self.n:                     # phantom digit: I component only
            hi.append(2 * self.n + 1)
            lo.append(2 * self.n)
            v.append([1.0, 0.0, 0.0, 0.0])
        self.ctx.init_product(self.sptr, self.n_bits, 0, hi, lo, v, scale)
        self._record(("init", list(hi), list(lo), [list(x) for x in v], float(scale)))
        self.pending = [None] * self.n
        self.queue = []

    # -- plan recording ---------------------------------------------------------------------
    def _record(self, entry, output=False):
        """Append a device-facing step to the tape; returns the index of its host output (if it has one)."""
        if self.tape is None:
            return None
        self.tape.append(entry)
        if output:
            self.tape_outputs += 1
            return self.tape_outputs - 1
        return None

    def _not_replayable(self):
        self.tape_valid = False

    def replay(self, tape):
        """Run a recorded tape on this engine's buffers: the C-ABI calls of the recording run, nothing else.  Device
        work buffers are allocated on the first replay and reused; every readout lands in a fresh pinned host array
        (PyTorch's caching host allocator recycles the blocks of results that have been dropped, so this is not a
        cudaHostAlloc per run and the caller owns what it gets); the copies are queued without waiting and the stream
        is synchronised once.  Returns the host arrays in recording order."""
        ctx = self.ctx
        ctx.set_stream(self.alloc.stream())
        bufs = getattr(self, "_replay_bufs", None)
        if bufs is None or bufs[0] is not tape:
            n, made = self.n, []
            for entry in tape:
                if entry[0] == "marginal":
                    made.append((self.alloc.empty(2 ** entry[4]), None))
                elif entry[0] == "to_matrix":
                    made.append((self.alloc.empty(2 * 4 ** n), self.alloc.empty(2 * 4 ** n)))
                else:
                    made.append(None)
            bufs = self._replay_bufs = (tape, made)
        pin = getattr(self.alloc, "pinned", None) or (lambda count: (None, np.empty(int(count))))
        landed = []
        for entry, b in zip(tape, bufs[1]):
            kind = entry[0]
            if kind == "init":
                ctx.init_product(self.sptr, self.n_bits, 0, entry[1], entry[2], entry[3], entry[4])
            elif kind == "passes":
                ctx.apply_passes(self.sptr, self.n_bits, entry[1])
            elif kind == "marginal":
                ctx.marginal(self.sptr, self.n_bits, 0, entry[1], entry[2], entry[3], self.alloc.ptr(b[0]))
                ctx.fwht(self.alloc.ptr(b[0]), entry[4])
                host = pin(2 ** entry[4])[1]
                ctx.download_async(self.alloc.ptr(b[0]), host)
                landed.append(host)
            elif kind == "chop":
                ctx.chop(self.sptr, self.size, entry[1])
            elif kind == "to_matrix":
                ctx.to_matrix(self.sptr, self.n, self.alloc.ptr(b[1]), self.alloc.ptr(b[0]))
                host = pin(2 * 4 ** self.n)[1]
                ctx.download_async(self.alloc.ptr(b[0]), host)
                landed.append(host.view(np.complex128).reshape(2 ** self.n, 2 ** self.n))
            elif kind == "download":
                host = pin(4 ** self.n)[1]
                ctx.download_async(self.sptr, host)
                landed.append(host)
            else:
                raise BasicAerError("internal: unknown tape entry %r" % (kind,))
        ctx.sync()
        return landed

    def upload(self, vec):
        self._not_replayable()
        vec = np.ascontiguousarray(vec, dtype=np.float64).reshape(-1)
        if vec.size != 4 ** self.n:
            raise BasicAerError("Wrong input stored density matrix")
        if self.nd > self.n:
            full = np.zeros(self.size)
            full[:vec.size] = vec
            vec = full
        self.ctx.upload(self.sptr, vec)
        self.pos = [self.n - 1 - q for q in range(self.n)]
        self.h2d_bytes += vec.nbytes
        self.pending = [None] * self.n
        self.queue = []

    # -- lazy op queue -------------------------------------------------------------------
    def apply_1q(self, q, m):
        """Left-multiply a single-qubit map onto qubit q's pending matrix."""
        cur = self.pending[q]
        self.pending[q] = np.array(m, dtype=float) if cur is None else np.asarray(m, dtype=float) @ cur

    def apply_1q_all(self, m):
        for q in range(self.n):
            self.apply_1q(q, m)

    def _take(self, q):
        m = self.pending[q]
        self.pending[q] = None
        return m

    def apply_cx(self, ctrl, tgt, tsp=(1.0, 0.0)):
        """``_add_unitary_two`` / ``cx_gate_dm_matrix`` (``dm_simulator.py:386-395``)."""
        if ctrl == tgt or not (0 <= ctrl < self.n) or not (0 <= tgt < self.n):
            raise BasicAerError("Qubit Labels out of bound in CX Gate")
        if float(tsp[0]) == 1.0 and float(tsp[1]) == 0.0:
            kind, coef = capi.OP_CX, None
        else:
            kind, coef = capi.OP_CX_TSP, cx_coefficients(tsp)
        self.queue.append(("2q", kind, ctrl, tgt, self._take(ctrl), self._take(tgt), coef))
        if self.drain_threshold:
            if getattr(self, "relabel", False) and getattr(self, "drain_tail", 0) > 0:
                if len(self.queue) >= self.drain_tail + getattr(self, "_chunk_now", self.drain_chunk):
                    self.drain()
                    self._chunk_now = min(2 * getattr(self, "_chunk_now", self.drain_chunk), self.drain_chunk)
            elif len(self.queue) >= self.drain_threshold:
                self.drain()

    def _schedule(self, final):
        """Outstanding work -> PASS array.  With relabelling the ops are scheduled on qubit ids
        and ``self.pos`` is advanced to the layout after the last pass."""
        if getattr(self, "relabel", False):
            qops = [schedule.DevOp(kind, qa, qb, pa, pb, coef) for (_, kind, qa, qb, pa, pb, coef) in self.queue]
            if final:
                left = [q for q in range(self.n) if self.pending[q] is not None]
                left.sort(key=lambda q: self.pos[q])
                for i in range(0, len(left) - 1, 2):
                    qops.append(schedule.DevOp(capi.OP_MATS, left[i], left[i + 1],
                                               self.pending[left[i]], self.pending[left[i + 1]]))
                if len(left) % 2:
                    qops.append(schedule.DevOp(capi.OP_MATS, left[-1], None, self.pending[left[-1]], None))
            if not qops:
                return np.zeros(0, dtype=capi.PASS_DTYPE)
            return schedule.relabel_passes(self.lib, qops, self.pos, self.nd, max_ops=self.max_ops_per_pass,
                                           strategy=getattr(self, "strategy", capi.SCHED_PROGRAM_ORDER))
        ops = self.device_ops(final=final)
        if not ops:
            return np.zeros(0, dtype=capi.PASS_DTYPE)
        return schedule.build_passes(ops, self.nd, max_ops=self.max_ops_per_pass, reserve_low=self.reserve_low)

    def drain(self):
        """Schedule and launch queued two-qubit ops now (pending matrices stay pending).  With
        relabelling the last ``drain_tail`` ops stay queued: the scheduler's look-ahead then sees
        the same ops it would see in a one-shot schedule, so streaming costs no extra passes."""
        tail = getattr(self, "drain_tail", 0)
        if getattr(self, "relabel", False) and tail > 0:
            if len(self.queue) <= tail:
                return
            qops = [schedule.DevOp(kind, qa, qb, pa, pb, coef) for (_, kind, qa, qb, pa, pb, coef) in self.queue]
            item_of = {id(op): item for op, item in zip(qops, self.queue)}
            passes, left = schedule.relabel_passes(self.lib, qops, self.pos, self.nd, max_ops=self.max_ops_per_pass,
                                                   min_tail=tail, strategy=getattr(self, "strategy", 0))
            self.queue = [item_of[id(op)] for op in left]
            self.run_passes(passes)
            return
        passes = self._schedule(final=False)
        self.queue = []
        self.run_passes(passes)
        self.drain_threshold = min(2 * self.drain_threshold, getattr(self, "drain_threshold_max", 1024))

    def apply_diag2(self, qa, qb, weights):
        """v[digit(qa)][digit(qb)] *= weights[i][j] (Bell mask, ``dm_simulator.py:749-756``)."""
        self.queue.append(("2q", capi.OP_DIAG2, qa, qb, self._take(qa), self._take(qb),
                           np.asarray(weights, dtype=float).reshape(16)))

    def device_ops(self, final=True):
        """Queue (+ pending matrices if ``final``) -> list of schedule.DevOp on digit positions."""
        ops = []
        for _, kind, qa, qb, pa, pb, coef in self.queue:
            ops.append(schedule.DevOp(kind, self.pos[qa], self.pos[qb], pa, pb, coef))
        if final:
            left = sorted((self.pos[q], q) for q in range(self.n) if self.pending[q] is not None)
            for i in range(0, len(left) - 1, 2):
                (da, qa), (db, qb) = left[i], left[i + 1]
                ops.append(schedule.DevOp(capi.OP_MATS, da, db, self.pending[qa], self.pending[qb]))
            if len(left) % 2:
                da, qa = left[-1]
                ops.append(schedule.DevOp(capi.OP_MATS, da, None, self.pending[qa], None))
        return ops

    def plan(self):
        """Schedule everything outstanding into passes (does not execute; advances ``pos``)."""
        return self._schedule(final=True)

    def flush(self):
        passes = self.plan()
        self.queue = []
        self.pending = [None] * self.n
        self.run_passes(passes)

    def run_passes(self, passes):
        if len(passes):
            self.ctx.set_stream(self.alloc.stream())
            self.ctx.apply_passes(self.sptr, self.n_bits, passes)
            self._record(("passes", passes.copy()))
            self.passes_run += len(passes)
            self.h2d_bytes += passes.nbytes

    # -- readouts (all flush first) --------------------------------------------------------
    def marginal_probabilities(self, basis, err):
        """2^n probabilities of ``_add_ensemble_measure`` for basis X/Y/Z (``:427-481``):
        I/B-marginal gather with weights err^wt(c), then a Walsh-Hadamard transform.
        Entry r has bit n-1-q = outcome of qubit q (= the reference's key order)."""
        self.flush()
        b = {"X": 1, "Y": 2, "Z": 3}[basis]
        n = self.n
        hi, lo, wt = [], [], np.zeros((n, 2, 4))
        for k in range(n):                       # result bit k <-> qubit n-1-k
            q = n - 1 - k
            hi.append(2 * self.pos[q] + 1)
            lo.append(2 * self.pos[q])
            wt[k, 0, 0] = 1.0
            wt[k, 1, b] = err
        out = self.alloc.empty(2 ** n)
        self.ctx.marginal(self.sptr, self.n_bits, 0, hi, lo, wt, self.alloc.ptr(out))
        self.ctx.fwht(self.alloc.ptr(out), n)
        host = np.empty(2 ** n)
        self.ctx.download(self.alloc.ptr(out), host)
        self.last_output = self._record(("marginal", list(hi), list(lo), wt.copy(), n), output=True)
        return host

    def n_basis_probabilities(self, nvec, err):
        """Basis 'N' of ``_add_ensemble_measure`` (``:457-462``): contract every digit
        (I, X, Y, Z) -> (I, err * n.(X,Y,Z)), then the same Walsh-Hadamard transform."""
        self._not_replayable()
        self.flush()
        self._require_reference_layout()
        n = self.n
        nv = np.asarray(nvec, dtype=float) * err
        src, src_ptr = self.state, self.sptr
        for d in range(n):                       # digit d: [H = 4^(nd-1-d)][4][L = 2^d]
            H, L = 4 ** (self.nd - 1 - d), 2 ** d
            dst = self.alloc.empty(H * 2 * L)
            self.ctx.contract_digit(src_ptr, self.alloc.ptr(dst), H, L, nv)
            src, src_ptr = dst, self.alloc.ptr(dst)
        # phantom digits (if any) are the leading axis with only index 0 populated: the first
        # 2^n entries are the marginal
        self.ctx.fwht(src_ptr, n)
        host = np.empty(2 ** n)
        self.ctx.download(src_ptr, host)
        return host

    def read_coefficients(self, digit_tuples):
        """Coefficients a[p_0, ..., p_{n-1}] (tuple index = qubit) -> numpy array."""
        self.settle()                                        # scheduling may relabel: settles self.pos
        pos = self.pos
        idx = []
        for tup in digit_tuples:
            flat = 0
            for q, p in enumerate(tup):
                flat |= int(p) << (2 * pos[q])
            idx.append(flat)
        return self.read_flat(idx)

    def settle(self):
        """Apply everything outstanding so that the buffer holds the state itself (layout ``self.pos``)."""
        self.flush()

    def read_flat(self, flat_indices):
        """Coefficients at flat indices of the CURRENT layout (digit of qubit q at bits 2*pos[q])."""
        self._not_replayable()
        self.settle()
        return self.ctx.read_coeffs(self.sptr, flat_indices)

    def reduced_coefficients(self, qubits):
        """Pauli-coefficient vector (4^k real numbers, ``qubits[0]`` = most significant digit, trace
        normalisation r[0] = 2^-k) of the reduced state Tr_{other qubits}(rho).  In the Pauli basis a
        partial trace is a gather: every Pauli string with a non-identity letter on a traced-out
        qubit is traceless, so r[P] = 2^(n-k) * a[P on `qubits`, I elsewhere].  Generalises the
        reduced two-qubit matrix of ``_add_bell_basis_measure`` (``dm_simulator.py:744-747``) to any
        subset of qubits; the state is not modified."""
        qubits = [int(q) for q in qubits]
        k = len(qubits)
        if k < 1 or len(set(qubits)) != k or not all(0 <= q < self.n for q in qubits):
            raise BasicAerError("reduced state: qubit list invalid: %r" % (qubits,))
        if k > 12:
            raise BasicAerError("reduced state on %d qubits is too large to gather" % k)
        self.settle()                                        # scheduling may relabel: settles self.pos
        j = np.arange(4 ** k, dtype=np.uint64)
        flat = np.zeros(4 ** k, dtype=np.uint64)
        for i, q in enumerate(qubits):                       # digit i of j (most significant first) -> qubit q
            digit = (j >> np.uint64(2 * (k - 1 - i))) & np.uint64(3)
            flat |= digit << np.uint64(2 * self.pos[q])
        return self.read_flat(flat) * float(2 ** (self.n - k))

    def reduced_densitymatrix(self, qubits):
        """2^k x 2^k complex matrix of the reduced state (``qubits[0]`` = most significant bit of the
        row / column index): ``reduced_coefficients`` + the Pauli->matrix kernels of
        ``_compute_densitymatrix`` (``:1198-1255``) on the 4^k gathered coefficients."""
        r = self.reduced_coefficients(qubits)
        k = len(qubits)
        src = self.alloc.empty(4 ** k)
        self.ctx.upload(self.alloc.ptr(src), r)
        work = self.alloc.empty(2 * 4 ** k)
        out = self.alloc.empty(2 * 4 ** k)
        self.ctx.to_matrix(self.alloc.ptr(src), k, self.alloc.ptr(work), self.alloc.ptr(out))
        host = np.empty(2 * 4 ** k)
        self.ctx.download(self.alloc.ptr(out), host)
        return host.view(np.complex128).reshape(2 ** k, 2 ** k)

    def _require_reference_layout(self):
        """Bring every qubit q back to digit position n-1-q (SWAP ops in ordinary tile passes)."""
        n = self.n
        pos = list(self.pos)
        if pos == [n - 1 - q for q in range(n)]:
            return
        owner = {pos[q]: q for q in range(n)}
        ops = []
        for q in range(n):
            t = n - 1 - q
            if pos[q] != t:
                other = owner[t]
                ops.append(schedule.DevOp(capi.OP_SWAP, pos[q], t))
                owner[pos[q]], pos[other] = other, pos[q]
                owner[t], pos[q] = q, t
        self.run_passes(schedule.build_passes(ops, self.nd, max_ops=self.max_ops_per_pass,
                                              reserve_low=self.reserve_low))
        self.pos = pos

    def download(self, out=None):
        """Flat 4^n coefficient vector in the reference order (host numpy array)."""
        self.flush()
        self._require_reference_layout()
        host = out if out is not None else np.empty(4 ** self.n)
        self.ctx.download(self.sptr, host)
        self.last_output = self._record(("download",), output=True)
        return host

    def chop(self, thr):
        self.flush()
        self.ctx.chop(self.sptr, self.size, thr)
        self._record(("chop", float(thr)))

    def to_matrix(self):
        """``_compute_densitymatrix`` (``:1198-1255``) -> complex 2^n x 2^n numpy array."""
        n = self.n
        if n > 14:
            # 2 x 16 * 4^n bytes of work / output next to the state (n = 15: 34 GB + the 17 GB state) and as much
            # pinned host memory; the reference cannot do it either (a Python loop over 4^n indices, :1233-1253)
            raise BasicAerError("compute_densitymatrix on %d qubits would need a %d GiB matrix; pass "
                                "compute_densitymatrix=False (the default is True, dm_simulator.py:259-260)"
                                % (n, (16 * 4 ** n) >> 30))
        self.flush()
        self._require_reference_layout()
        if self.nd > n:
            # 1-qubit register: convert on the first 4 coefficients (phantom digit is identity)
            self._not_replayable()
            tmp = self.alloc.empty(4 ** n)
            self.ctx.upload(self.alloc.ptr(tmp), self.ctx.read_coeffs(self.sptr, list(range(4 ** n))))
            src_ptr = self.alloc.ptr(tmp)
        else:
            src_ptr = self.sptr
        work = self.alloc.empty(2 * 4 ** n)
        out = self.alloc.empty(2 * 4 ** n)
        self.ctx.to_matrix(src_ptr, n, self.alloc.ptr(work), self.alloc.ptr(out))
        host = np.empty(2 * 4 ** n)
        self.ctx.download(self.alloc.ptr(out), host)
        self.last_output = self._record(("to_matrix",), output=True)
        return host.view(np.complex128).reshape(2 ** n, 2 ** n)

    def overlap_with(self, other_vec):
        """dot(other, state) (``_state_overlap``, ``:1277-1282``, without the 2^n factor)."""
        self._not_replayable()
        self.flush()
        self._require_reference_layout()
        other_vec = np.ascontiguousarray(other_vec, dtype=np.float64).reshape(-1)
        if other_vec.size != 4 ** self.n:
            raise BasicAerError("stored coefficients have the wrong length")
        buf = self.alloc.empty(4 ** self.n)
        self.ctx.upload(self.alloc.ptr(buf), other_vec)
        return self.ctx.dot(self.alloc.ptr(buf), self.sptr, 4 ** self.n)

    def sync(self):
        self.ctx.sync()

    def stats(self):
        return self.ctx.stats()
