"""world_size 2 / 4 / 8 tests of the multi-GPU layer on CPU: gloo backend, host buffers and
the CPU emulation of the kernels.  Covers the exchange plan (bit-permutation block moves),
the exchange itself, the sharded scheduler with evictions and the sharded marginal readout,
against the oracle."""
import copy
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_exchange_plan_is_the_slot_swap():
    """Every (rank, block) lands where the global<->top-local digit swap sends it."""
    from qiskit_aakash_b200.distributed import ExchangePlan
    for world in (2, 4, 8):
        for n in (6, 7):
            g = world.bit_length() - 1
            m = (g + 1) // 2
            plans = [ExchangePlan(n, world, r) for r in range(world)]
            B = plans[0].B
            seen = set()
            for r, p in enumerate(plans):
                assert p.n_loc == n - m and p.n_bits_local == 2 * n - g
                for (s, dr, db) in p.moves:
                    gidx = ((r << p.block_bits) | s) << B            # first element of the block
                    digits = [(gidx >> (2 * k)) & 3 for k in range(n)]
                    for i in range(m):                               # swap slot n-m+i <-> n-2m+i
                        a, b = n - m + i, n - 2 * m + i
                        digits[a], digits[b] = digits[b], digits[a]
                    want = sum(d << (2 * k) for k, d in enumerate(digits))
                    assert want == ((dr << p.block_bits) | db) << B
                    seen.add((dr, db))
            assert len(seen) == world * (1 << plans[0].block_bits)


def _worker(rank, world, port, n, seed, mode, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    import cases
    from emu_backend import emu_lib
    from oracle import dm_oracle
    from qiskit_aakash_b200 import assemble, circuits as C, distributed
    from qiskit_aakash_b200.dm_simulator import DmSimulatorB200

    class CpuAlloc:
        index = 0

        def empty(self, count):
            return torch.empty(int(count), dtype=torch.float64)

        def ptr(self, buf):
            return buf.data_ptr()

        def stream(self):
            return 0

    comm = distributed.TorchCommunicator()
    circ = cases._rand_circuit(n, 45, seed) if mode != "layered" else C.random_layered(n, 5, seed, readout=False)
    if mode == "expect":
        circ.measure(0, 0, basis="Expect", add_param=("XZIY" * n)[:n])
        circ.u3(0.1, 0.2, 0.3, 0); circ.cx(0, n - 1)
    elif mode == "bell":
        circ.measure(0, 0, basis="Bell", add_param="0%d" % (n - 1))
        circ.u3(0.1, 0.2, 0.3, 1); circ.cx(1, 0)
    if mode == "nbasis":
        def ensemble_n(vec):
            circ.barrier()
            for q in range(n):
                circ.instructions.append(C.instr("measure", [q], ["Ensemble", ["N", np.array(vec)]], memory=[q]))
            circ.barrier()
        ensemble_n([0.3, -1.1, 0.7])
        for q in range(n):
            circ.u3(0.4 + q, 0.1, 0.2, q)                       # pending maps on global qubits at the readout
        ensemble_n([1.0, 2.0, -0.5])
    else:
        circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="XYZ"[seed % 3])
    opts = dict(cases.FULL_NOISE, compute_densitymatrix=False)
    if mode == "matrix":
        opts["compute_densitymatrix"] = True
    if mode == "reduced":
        opts["reduced_state"] = [n - 1, 0, 2]            # one global, two local qubits
    if mode == "stored":
        # start from a stored state and compare against stored coefficients (a4, a27) on shards
        os.chdir(out_dir)
        start = dm_oracle.run_oracle(n, copy.deepcopy(cases._rand_circuit(n, 20, seed + 100).instructions),
                                     {"compute_densitymatrix": False})["data"]["coeffmatrix"]
        other = dm_oracle.run_oracle(n, copy.deepcopy(cases._rand_circuit(n, 20, seed + 200).instructions),
                                     {"compute_densitymatrix": False})["data"]["coeffmatrix"]
        if rank == 0:
            np.save("stored_density_matrix.npy", start)
            np.save("stored_coefficients.npy", other)
        dist.barrier()
        opts.update(custom_densitymatrix="stored_density_matrix", initial_densitymatrix=True, compare=True)
    engines = []

    def factory(nq):
        if mode == "recycle" and engines:                  # what the product's sharded factory does between jobs
            return engines[0].recycle()
        e = distributed.ShardedPauliEngine(nq, comm, lib=emu_lib(), allocator=CpuAlloc(), max_ops_per_pass=4)
        engines.append(e)
        return e

    be = DmSimulatorB200(_engine_factory=factory, comm=comm)
    if mode == "recycle":                                  # an earlier job leaves its layout, counters and buffers behind
        c0 = cases._rand_circuit(n, 30, seed + 50)
        c0.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Y")
        be.run(assemble(c0), backend_options=copy.deepcopy(opts)).result()
        first_exchanges = engines[0].exchanges
    c2 = C.Circuit(n)
    c2.instructions = copy.deepcopy(circ.instructions)
    res = be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]
    if mode == "recycle":
        assert len(engines) == 1 and first_exchanges >= 0
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    d_red = 0.0
    if mode == "reduced":
        import test_reduced_state as trs
        full = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), dict(copy.deepcopy(opts), compute_densitymatrix=True))
        want = trs.numpy_partial_trace(np.asarray(full["data"]["densitymatrix"]), n, opts["reduced_state"])
        d_red = max(float(np.max(np.abs(res["data"].pop("reduced_densitymatrix") - want))),
                    float(np.max(np.abs(res["data"].pop("reduced_coeffmatrix") - trs.pauli_vector(want, 3)))))
    p_got = np.array(list(res["data"]["ensemble_probability"].values()))
    p_ref = np.array(list(ref["data"]["ensemble_probability"].values()))
    d_p = float(np.max(np.abs(p_got - p_ref)))
    d_c = max(d_red, float(np.max(np.abs(res["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"]))))
    assert set(res["data"]) == set(ref["data"]), (set(res["data"]), set(ref["data"]))
    for key, val in ref["data"].items():
        if key.startswith(("Pauli_string", "reduced_bell")):
            d_c = max(d_c, float(np.max(np.abs(np.asarray(val) - np.asarray(res["data"][key])))))
        elif key.startswith("bell_prob"):
            d_p = max(d_p, max(abs(val[k] - res["data"][key][k]) for k in val))
        elif key == "densitymatrix":
            d_c = max(d_c, float(np.max(np.abs(np.asarray(val) - np.asarray(res["data"][key])))))
        elif key == "fidelity":
            d_c = max(d_c, float(abs(val - res["data"][key])))
    with open(os.path.join(out_dir, "r%d.txt" % rank), "w") as f:
        f.write("%r %r %d %d\n" % (d_p, d_c, engines[0].exchanges, res["number_of_clock_cycles"] - ref["number_of_clock_cycles"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,seed,mode", [(2, 5, 1, "rand"), (2, 6, 2, "layered"), (4, 6, 3, "rand"),
                                               (4, 7, 4, "layered"), (8, 7, 5, "rand"), (8, 7, 6, "layered"),
                                               (2, 5, 7, "expect"), (4, 6, 8, "bell"), (8, 7, 9, "expect"),
                                               (2, 5, 10, "stored"), (8, 7, 11, "stored"),
                                               (4, 6, 15, "matrix"), (2, 5, 16, "reduced"), (8, 7, 17, "reduced"),
                                               (2, 5, 12, "nbasis"), (4, 6, 13, "nbasis"), (8, 7, 14, "nbasis"),
                                               (4, 6, 18, "recycle"), (8, 7, 19, "recycle")])
def test_sharded_backend_matches_oracle(world, n, seed, mode, tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu
    build_emu.build()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, seed, mode, str(tmp_path)), nprocs=world, join=True)
    exchanges = 0
    for r in range(world):
        d_p, d_c, ex, dl = open(os.path.join(str(tmp_path), "r%d.txt" % r)).read().split()
        assert float(d_p) <= 1e-10 and float(d_c) <= 1e-10 and int(dl) == 0, (r, d_p, d_c)
        exchanges = int(ex)
    assert exchanges >= 1, "test circuit never needed a global-qubit exchange"



def _dump_worker(rank, world, port, n, seed, out_dir):
    """store_densitymatrix / compare / stored_density_matrix as PER-RANK slice files (no gather anywhere)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    import cases
    from emu_backend import emu_lib
    from oracle import dm_oracle
    from qiskit_aakash_b200 import assemble, circuits as C, distributed
    from qiskit_aakash_b200.dm_simulator import DmSimulatorB200

    class CpuAlloc:
        index = 0

        def empty(self, count):
            return torch.empty(int(count), dtype=torch.float64)

        def ptr(self, buf):
            return buf.data_ptr()

        def stream(self):
            return 0

    os.chdir(out_dir)
    comm = distributed.TorchCommunicator()
    engines = []

    def factory(nq):
        e = distributed.ShardedPauliEngine(nq, comm, lib=emu_lib(), allocator=CpuAlloc(), max_ops_per_pass=4)
        engines.append(e)
        return e

    def run(circ, opts):
        be = DmSimulatorB200(_engine_factory=factory, comm=comm)
        c2 = C.Circuit(n)
        c2.instructions = copy.deepcopy(circ.instructions)
        return be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]

    worst = 0.0
    # 1. store: every rank writes its slice of the reference-order vector
    first = cases._rand_circuit(n, 40, seed)
    first.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Z")
    opts = dict(cases.FULL_NOISE, compute_densitymatrix=False)
    run(first, dict(opts, store_densitymatrix=True))
    ref1 = dm_oracle.run_oracle(n, copy.deepcopy(first.instructions), copy.deepcopy(opts))["data"]["coeffmatrix"]
    # the state at the (final) ensemble readout, before the post-level noise of the readout level and the chop:
    # compare slices with a second oracle run that stores at the same point
    size = 4 ** n // world
    mine = np.load(distributed.ShardedPauliEngine.shard_file("stored_coefficients", rank, world))
    assert mine.size == size
    assert not os.path.exists("stored_coefficients.npy")
    dist.barrier()
    if rank == 0:
        joined = distributed.join_shard_files("stored_coefficients", world, out="joined_coefficients")
        sub = os.path.join(out_dir, "oracle_store")
        os.makedirs(sub)
        os.chdir(sub)
        dm_oracle.run_oracle(n, copy.deepcopy(first.instructions), dict(copy.deepcopy(opts), store_densitymatrix=True))
        want = np.load("stored_coefficients.npy")
        os.chdir(out_dir)
        worst = max(worst, float(np.max(np.abs(joined - want))))
    dist.barrier()
    worst = max(worst, 0.0 * float(ref1[0]))
    # 2. compare: a different circuit against the sharded dump (each rank reads its own slice file)
    second = cases._rand_circuit(n, 40, seed + 1)
    second.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Z")
    got = run(second, dict(opts, compare=True))["data"]["fidelity"]
    if rank == 0:
        sub = os.path.join(out_dir, "oracle_cmp")
        os.makedirs(sub)
        os.chdir(sub)
        np.save("stored_coefficients.npy", np.load(os.path.join(out_dir, "joined_coefficients.npy")))
        want_f = dm_oracle.run_oracle(n, copy.deepcopy(second.instructions), dict(copy.deepcopy(opts), compare=True))["data"]["fidelity"]
        os.chdir(out_dir)
        worst = max(worst, abs(got - want_f))
    dist.barrier()
    # 3. start from the sharded dump: rename the slice files to the name 'stored_density_matrix' reads
    os.replace(distributed.ShardedPauliEngine.shard_file("stored_coefficients", rank, world),
               distributed.ShardedPauliEngine.shard_file("stored_density_matrix", rank, world))
    dist.barrier()
    third = cases._rand_circuit(n, 30, seed + 2)
    third.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="X")
    o3 = dict(opts, custom_densitymatrix="stored_density_matrix", initial_densitymatrix=True)
    res = run(third, o3)
    if rank == 0:
        sub = os.path.join(out_dir, "oracle_init")
        os.makedirs(sub)
        os.chdir(sub)
        np.save("stored_density_matrix.npy", np.load(os.path.join(out_dir, "joined_coefficients.npy")))
        ref3 = dm_oracle.run_oracle(n, copy.deepcopy(third.instructions), copy.deepcopy(o3))
        os.chdir(out_dir)
        worst = max(worst, float(np.max(np.abs(res["data"]["coeffmatrix"] - ref3["data"]["coeffmatrix"]))))
        p = np.array(list(res["data"]["ensemble_probability"].values()))
        worst = max(worst, float(np.max(np.abs(p - np.array(list(ref3["data"]["ensemble_probability"].values()))))))
    # 4. canonicalise() from whatever layout the last run ended in: the shard IS the reference slice
    e = engines[-1]
    shard = e.download_shard()
    full = res["data"]["coeffmatrix"]
    worst = max(worst, float(np.max(np.abs(shard - full[rank * size:(rank + 1) * size]))))
    with open(os.path.join(out_dir, "d%d.txt" % rank), "w") as f:
        f.write("%r %d\n" % (float(worst), sum(x.exchanges for x in engines)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,seed", [(2, 5, 31), (4, 6, 32), (8, 7, 33)])
def test_sharded_dumps_are_per_rank_slices(world, n, seed, tmp_path):
    """(f)2: ``store_densitymatrix`` writes one slice file per rank (no rank-0 gather), ``compare`` and
    ``stored_density_matrix`` read them back per rank; the joined file is what the oracle stores at the same point,
    the fidelity and the restarted run match the oracle, and ``canonicalise`` leaves every rank with the contiguous
    slice of the reference-order vector."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu
    build_emu.build()
    port = _free_port()
    mp.spawn(_dump_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        worst, ex = open(os.path.join(str(tmp_path), "d%d.txt" % r)).read().split()
        assert float(worst) <= 1e-10, (r, worst)
