"""CPU test of the fused exchange (tile pass that pulls its input from peer buffers): the ranks
are threads of one process, so "peer-mapped" pointers are ordinary pointers and the CPU
emulation of the kernels can follow them.  Covers the source-table arithmetic of
``ShardedPauliEngine.exchange`` and ``dmb_apply_pass_remote`` for world sizes 2 / 4 / 8
(including half-digit sharding) against the oracle, and against the NCCL-style block exchange."""
import copy
import threading

import numpy as np
import pytest

import cases
from emu_backend import NumpyAllocator, emu_lib
from oracle import dm_oracle
from qiskit_aakash_b200 import assemble, circuits as C, distributed
from qiskit_aakash_b200.dm_simulator import DmSimulatorB200


class ThreadCluster:
    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world, timeout=120)
        self.slots = [None] * world


class ThreadComm:
    """In-process stand-in for TorchCommunicator (same interface)."""

    def __init__(self, cluster, rank, fused):
        self.c, self.rank, self.world, self.fused = cluster, rank, cluster.world, fused
        if not fused:
            self.peer_addresses = None

    def barrier(self):
        self.c.bar.wait()

    def _share(self, obj):
        self.c.slots[self.rank] = obj
        self.c.bar.wait()
        got = list(self.c.slots)
        self.c.bar.wait()
        return got

    def exchange(self, plan, src, dst):
        bufs = self._share(src)
        be = plan.block_elems
        for d in range(1 << plan.block_bits):
            sr, sb = plan.image(self.rank, d)
            dst[d * be:(d + 1) * be] = bufs[sr][sb * be:(sb + 1) * be]
        self.c.bar.wait()

    def all_reduce_sum(self, tensor):
        total = np.sum(np.stack([np.array(x, copy=True) for x in self._share(tensor)]), axis=0)
        self.c.bar.wait()                 # everyone has read before anyone overwrites
        tensor[:] = total
        self.c.bar.wait()

    def all_gather_host(self, arr):
        return self._share(arr)

    def peer_addresses(self, ctx, ptrs):
        everyone = self._share(list(ptrs))
        return [[everyone[r][b] for r in range(self.world)] for b in range(len(ptrs))]


def _run_world(world, n, circ, opts, fused, mode="pull", first_job=None, direct_slots=None):
    """``first_job``: a circuit that runs BEFORE ``circ`` on the same engines, which are then recycled for ``circ``
    (``ShardedPauliEngine.recycle``: what the product's sharded factory does between jobs of one register size)."""
    cluster = ThreadCluster(world)
    out, errs = [None] * world, []
    direct_seen = [0] * world
    _run_world.direct_seen = direct_seen

    def work(rank):
        try:
            comm = ThreadComm(cluster, rank, fused)
            engines = []

            def factory(nq):
                if first_job is not None and engines:
                    return engines[0].recycle()
                e = distributed.ShardedPauliEngine(nq, comm, lib=emu_lib(), allocator=NumpyAllocator(), max_ops_per_pass=4)
                e.exchange_mode = mode
                if direct_slots is not None:     # "parked" | "direct" | "direct_late": force that schedule variant
                    e.force_exchange_variant = direct_slots
                engines.append(e)
                return e

            be = DmSimulatorB200(_engine_factory=factory)
            if first_job is not None:
                c0 = C.Circuit(n)
                c0.instructions = copy.deepcopy(first_job.instructions)
                be.run(assemble(c0), backend_options=copy.deepcopy(opts)).result()
            c2 = C.Circuit(n)
            c2.instructions = copy.deepcopy(circ.instructions)
            res = be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]
            out[rank] = (res, engines[0].exchanges, engines[0].peers is not None)
            direct_seen[rank] = engines[0].direct_exchanges
        except Exception as exc:                     # pragma: no cover - surfaced below
            errs.append(exc)
            cluster.bar.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("world,n,seed,mode", [(2, 6, 1, "rand"), (2, 7, 2, "layered"), (4, 7, 3, "rand"),
                                               (4, 8, 4, "layered"), (8, 8, 5, "rand"), (8, 8, 6, "layered"),
                                               (8, 7, 7, "rand")])
def test_fused_exchange_matches_oracle(world, n, seed, mode):
    circ = cases._rand_circuit(n, 50, seed) if mode == "rand" else C.random_layered(n, 6, seed, readout=False)
    circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="XYZ"[seed % 3])
    opts = dict(cases.FULL_NOISE, compute_densitymatrix=False)
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    p_ref = np.array(list(ref["data"]["ensemble_probability"].values()))
    for fused, xmode in ((True, "pull"), (True, "push"), (False, "nccl")):
        outs = _run_world(world, n, circ, opts, fused, xmode)
        for rank, (res, exchanges, has_peers) in enumerate(outs):
            assert has_peers == fused
            assert exchanges >= 1
            p = np.array(list(res["data"]["ensemble_probability"].values()))
            assert np.max(np.abs(p - p_ref)) <= 1e-10
            assert np.max(np.abs(res["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"])) <= 1e-10


@pytest.mark.parametrize("world,n,seed", [(2, 6, 21), (4, 7, 22), (8, 8, 23)])
def test_recycled_engine_runs_the_next_job_correctly(world, n, seed):
    """Two jobs on one set of engines (shards and peer tables kept, ``recycle()`` in between): the second job starts
    from whichever of the two buffers the first one ended in, possibly while a slower peer still pulls from the other
    one, and must still match the oracle in every exchange mode."""
    first = cases._rand_circuit(n, 40, seed + 100)
    first.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Y")
    circ = cases._rand_circuit(n, 50, seed)
    circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="XYZ"[seed % 3])
    opts = dict(cases.FULL_NOISE, compute_densitymatrix=False)
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    p_ref = np.array(list(ref["data"]["ensemble_probability"].values()))
    for fused, xmode in ((True, "pull"), (True, "push"), (False, "nccl")):
        outs = _run_world(world, n, circ, opts, fused, xmode, first_job=first)
        for rank, (res, exchanges, has_peers) in enumerate(outs):
            assert has_peers == fused
            p = np.array(list(res["data"]["ensemble_probability"].values()))
            assert np.max(np.abs(p - p_ref)) <= 1e-10
            assert np.max(np.abs(res["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"])) <= 1e-10


@pytest.mark.parametrize("world,n", [(2, 7), (4, 8), (8, 8)])
def test_sharded_qft_with_chained_cnots_matches_oracle(world, n):
    """QFT on a sharded state: the two CNOTs of a cu1 are chained (one shared-memory round trip, dmb_chain_ops) also in
    the pass that pulls its tiles from the peers, and cu1 gates on global qubits force exchanges in between."""
    circ = C.qft(n)
    opts = {"compute_densitymatrix": False}
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    p_ref = np.array(list(ref["data"]["ensemble_probability"].values()))
    for fused, xmode in ((True, "pull"), (True, "push"), (False, "nccl")):
        outs = _run_world(world, n, circ, opts, fused, xmode)
        for rank, (res, exchanges, has_peers) in enumerate(outs):
            assert exchanges >= 1
            p = np.array(list(res["data"]["ensemble_probability"].values()))
            assert np.max(np.abs(p - p_ref)) <= 1e-10
            assert np.max(np.abs(res["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"])) <= 1e-10


@pytest.mark.parametrize("world,n,seed,kind", [(2, 7, 31, "rand"), (4, 8, 32, "rand"), (8, 8, 33, "rand"), (8, 9, 34, "layered"),
                                               (4, 9, 35, "layered"), (2, 8, 36, "layered")])
def test_direct_slot_exchange_needs_no_parking_and_matches_oracle(world, n, seed, kind):
    """Fused pull with the global slots swapped with the evictees' OWN local slots (scattered-bit source table,
    ``dmb_apply_pass_remote_sel``): same results as the oracle and as the parking variant, the direct path is the one
    that ran, and it never runs more passes than parking."""
    circ = cases._rand_circuit(n, 60, seed) if kind == "rand" else C.random_layered(n, 8, seed, readout=False)
    circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="XYZ"[seed % 3])
    opts = dict(cases.FULL_NOISE, compute_densitymatrix=False)
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    p_ref = np.array(list(ref["data"]["ensemble_probability"].values()))
    for variant in ("direct", "direct_late", "parked"):
        outs = _run_world(world, n, circ, opts, True, "pull", direct_slots=variant)
        seen = list(_run_world.direct_seen)
        for rank, (res, exchanges, has_peers) in enumerate(outs):
            assert has_peers and exchanges >= 1
            assert (seen[rank] >= 1) == (variant != "parked"), (variant, seen)
            p = np.array(list(res["data"]["ensemble_probability"].values()))
            assert np.max(np.abs(p - p_ref)) <= 1e-10
            assert np.max(np.abs(res["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"])) <= 1e-10


@pytest.mark.parametrize("world,n", [(2, 7), (4, 8), (8, 8)])
def test_map_only_tail_pass_is_folded_into_the_marginal_readout(world, n):
    """A job whose last pass would only apply pending single-qubit maps (here: a u3 on two qubits after the last CNOT) and
    that ends in an X / Y / Z ensemble readout: the pass is dropped, the maps stay pending and the readout folds them into
    its weights (``_drop_map_only_tail``) -- same probabilities, one round trip of the shard less."""
    circ = C.Circuit(n)
    for q in range(n):
        circ.u3(0.3 + 0.1 * q, 0.2, 0.5, q)
    for q in range(0, n - 1, 2):
        circ.cx(q, q + 1)
    circ.cx(n - 1, 0)
    circ.u3(0.7, -0.4, 1.1, 2)
    circ.u3(-0.2, 0.9, 0.3, 3)
    circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Y")
    opts = {"compute_densitymatrix": False}
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    p_ref = np.array(list(ref["data"]["ensemble_probability"].values()))
    passes = {}
    for fold in (True, False):
        cluster_out = []

        def run(fold=fold):
            cluster = ThreadCluster(world)
            out, errs = [None] * world, []

            def work(rank):
                try:
                    comm = ThreadComm(cluster, rank, True)
                    e = distributed.ShardedPauliEngine(n, comm, lib=emu_lib(), allocator=NumpyAllocator(), max_ops_per_pass=4)
                    e.fold_tail_maps = fold
                    be = DmSimulatorB200(_engine_factory=lambda nq: e)
                    be.SHOW_FINAL_STATE = False
                    c2 = C.Circuit(n)
                    c2.instructions = copy.deepcopy(circ.instructions)
                    res = be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]
                    out[rank] = (res, e.passes_run)
                except Exception as exc:                 # pragma: no cover
                    errs.append(exc)
                    cluster.bar.abort()

            threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errs:
                raise errs[0]
            return out

        outs = run()
        for res, n_passes in outs:
            p = np.array(list(res["data"]["ensemble_probability"].values()))
            assert np.max(np.abs(p - p_ref)) <= 1e-10
        passes[fold] = outs[0][1]
    assert passes[True] <= passes[False]
