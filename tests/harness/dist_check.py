"""torchrun --nproc-per-node N tests/harness/dist_check.py : sharded engine on real GPUs vs the oracle."""
import copy
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    dist.barrier()
    import cases
    from oracle import dm_oracle
    from qiskit_aakash_b200 import assemble, circuits as C, distributed
    from qiskit_aakash_b200.dm_simulator import DmSimulatorB200
    comm = distributed.TorchCommunicator()
    worst = 0.0
    for n, seed, mode in ((8, 1, "rand"), (9, 2, "layered"), (10, 3, "rand"), (10, 4, "layered"), (9, 5, "nbasis"),
                          (8, 6, "matrix"), (10, 7, "qft"), (10, 8, "qft_noisy")):
        if mode.startswith("qft"):               # every cu1 = two chained CNOTs (dmb_chain_ops), also in the pull pass
            circ = C.qft(n, readout=False)
        else:
            circ = cases._rand_circuit(n, 60, seed) if mode != "layered" else C.random_layered(n, 6, seed, readout=False)
        if mode == "nbasis":                     # N-basis ensemble readout with pending maps on global qubits
            for q in range(n):
                circ.u3(0.4 + q, 0.1, 0.2, q)
            circ.barrier()
            for q in range(n):
                circ.instructions.append(C.instr("measure", [q], ["Ensemble", ["N", np.array([1.0, 2.0, -0.5])]],
                                                 memory=[q]))
            circ.barrier()
        else:
            circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="XYZ"[seed % 3])
        opts = dict({} if mode == "qft" else cases.FULL_NOISE, compute_densitymatrix=(mode == "matrix"))
        engines = []

        def factory(nq):
            e = distributed.ShardedPauliEngine(nq, comm, device=local)
            engines.append(e)
            return e

        be = DmSimulatorB200(_engine_factory=factory, comm=comm)
        c2 = C.Circuit(n)
        c2.instructions = copy.deepcopy(circ.instructions)
        res = be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]
        if rank == 0:
            ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
            p_got = np.array(list(res["data"]["ensemble_probability"].values()))
            p_ref = np.array(list(ref["data"]["ensemble_probability"].values()))
            d_p = float(np.max(np.abs(p_got - p_ref)))
            d_c = float(np.max(np.abs(res["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"])))
            if mode == "matrix":
                d_c = max(d_c, float(np.max(np.abs(res["data"]["densitymatrix"] - ref["data"]["densitymatrix"]))))
            worst = max(worst, d_p, d_c)
            print(json.dumps({"check": "sharded_vs_oracle", "world": world, "n": n, "mode": mode, "d_prob": d_p,
                              "d_coeff": d_c, "exchanges": engines[0].exchanges, "fused_exchange": engines[0].peers is not None,
                              "nvlink_bytes_sent_per_rank": engines[0].nvlink_bytes_sent,
                              "chained_ops": engines[0].stats().get("chained_ops")}))
            if mode.startswith("qft"):
                assert engines[0].stats().get("chained_ops", 0) > 0
        dist.barrier()
    # sharded dumps: store -> per-rank slice files (no gather), compare and restart from them, canonical layout;
    # through the PUBLIC constructor (DmSimulatorB200(comm=...): what get_backend returns inside a process group)
    import tempfile
    n = 9
    tmp = [tempfile.mkdtemp(prefix="dmb_dist_") if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    os.chdir(tmp[0])
    opts = dict(cases.FULL_NOISE, compute_densitymatrix=False)
    first = cases._rand_circuit(n, 50, 71)
    first.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Z")
    assert DmSimulatorB200(comm=comm, device=local).configuration().n_qubits == {2: 17, 4: 17, 8: 18}.get(world, 16)
    backends = []

    def run_public(circ, o):
        # a fresh backend per job: store_densitymatrix / compare stick to the instance like in the reference
        backends.append(DmSimulatorB200(comm=comm, device=local))
        c2 = C.Circuit(n)
        c2.instructions = copy.deepcopy(circ.instructions)
        return backends[-1].run(assemble(c2), backend_options=copy.deepcopy(o)).result()["results"][0]

    run_public(first, dict(opts, store_densitymatrix=True))
    assert os.path.exists(distributed.ShardedPauliEngine.shard_file("stored_coefficients", rank, world))
    assert not os.path.exists("stored_coefficients.npy")
    second = cases._rand_circuit(n, 50, 72)
    second.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Z")
    fid = run_public(second, dict(opts, compare=True))["data"]["fidelity"]
    be = backends[-1]
    shard = be._engine.download_shard()
    full = be._engine.download()
    size = 4 ** n // world
    d_dump = float(np.max(np.abs(shard - full[rank * size:(rank + 1) * size])))
    if rank == 0:
        joined = distributed.join_shard_files("stored_coefficients", world, out="joined")
        os.makedirs("oracle")
        os.chdir("oracle")
        dm_oracle.run_oracle(n, copy.deepcopy(first.instructions), dict(copy.deepcopy(opts), store_densitymatrix=True))
        d_dump = max(d_dump, float(np.max(np.abs(joined - np.load("stored_coefficients.npy")))))
        want = dm_oracle.run_oracle(n, copy.deepcopy(second.instructions), dict(copy.deepcopy(opts), compare=True))["data"]["fidelity"]
        d_dump = max(d_dump, abs(fid - want))
        os.chdir(tmp[0])
        worst = max(worst, d_dump)
        print(json.dumps({"check": "sharded_dumps", "world": world, "n": n, "d": d_dump, "fidelity": fid}))
    t = torch.tensor([d_dump], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t.item()) <= 1e-10, float(t.item())
    for b in backends:
        b.release_engines()
    dist.barrier()
    # consecutive jobs on ONE public backend: the sharded engine is pooled (same register size -> recycle(): shards and
    # peer mappings kept; another size -> closed and rebuilt)
    be = DmSimulatorB200(comm=comm, device=local)
    d_pool, reused = 0.0, []
    for job, (nq, seed) in enumerate(((9, 81), (9, 82), (9, 83), (8, 84), (9, 85))):
        circ = cases._rand_circuit(nq, 60, seed)
        circ.measure(list(range(nq)), list(range(nq)), basis="Ensemble", add_param="XYZ"[seed % 3])
        c2 = C.Circuit(nq)
        c2.instructions = copy.deepcopy(circ.instructions)
        before = be._sharded_pool
        res = be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]
        reused.append(be._sharded_pool is before)
        ref = dm_oracle.run_oracle(nq, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
        p_got = np.array(list(res["data"]["ensemble_probability"].values()))
        p_ref = np.array(list(ref["data"]["ensemble_probability"].values()))
        d_pool = max(d_pool, float(np.max(np.abs(p_got - p_ref))),
                     float(np.max(np.abs(res["data"]["coeffmatrix"] - ref["data"]["coeffmatrix"]))))
    assert reused == [False, True, True, False, False], reused
    worst = max(worst, d_pool)
    if rank == 0:
        print(json.dumps({"check": "pooled_engine_jobs", "world": world, "jobs": len(reused), "reused": reused, "d": d_pool}))
    t = torch.tensor([d_pool], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t.item()) <= 1e-10, float(t.item())
    be.release_engines()
    dist.barrier()
    # random programs (tests/harness/fuzz_emu.py's generator: every measurement mode mid-circuit, resets, random
    # options) -- includes the pattern that exposed the scratch-shard race after a fused pull (a readout
    # that uses the scratch shard as workspace right after an exchange)
    sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
    import contextlib
    import io
    import fuzz_emu
    lo = {2: 3, 4: 4, 8: 6}.get(world, 6)
    n_fuzz, bad = int(os.environ.get("DIST_CHECK_FUZZ", "40")), []
    for seed in range(n_fuzz):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(lo, 10))
        circ = fuzz_emu.random_circuit(rng, n, 80)
        if seed % 4 == 0:                       # the minimal trigger: global cx, then an N-basis partial readout
            circ = C.Circuit(n)
            circ.cx(0, 1)
            circ.measure([1, n - 1], [1, n - 1], basis="N", add_param=np.array([0.67, -0.43, -0.57]))
            circ.cx(1, 0)
            circ.measure(list(range(n)), list(range(n)), basis="Ensemble", add_param="Y")
        opts = fuzz_emu.random_options(rng)
        fuzz_emu.random_init(rng, n, opts)
        be = DmSimulatorB200(_engine_factory=lambda nq: distributed.ShardedPauliEngine(nq, comm, device=local))
        c2 = C.Circuit(n)
        c2.instructions = copy.deepcopy(circ.instructions)
        with contextlib.redirect_stdout(io.StringIO()):
            res = be.run(assemble(c2), backend_options=copy.deepcopy(opts)).result()["results"][0]
            ref = dm_oracle.run_oracle(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
        d = 0.0
        for k, v in ref["data"].items():
            d = max(d, float(np.max(np.abs(fuzz_emu.as_arr(v) - fuzz_emu.as_arr(res["data"][k])))))
        if set(ref["data"]) != set(res["data"]) or d > 1e-10:
            bad.append((seed, n, d))
        worst = max(worst, d)
        dist.barrier()
    flag = torch.tensor([len(bad)], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print(json.dumps({"check": "random_programs", "world": world, "programs": n_fuzz, "failed_on_any_rank": int(flag.item()),
                          "bad_rank0": bad[:5]}))
        assert int(flag.item()) == 0
        assert worst <= 1e-10, worst
        print("DIST_CHECK_OK world=%d worst=%.3e" % (world, worst))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
