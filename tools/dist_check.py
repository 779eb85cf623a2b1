"""torchrun --nproc-per-node N tools/dist_check.py : sharded engine on real GPUs vs the oracle."""
import copy
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
                          (8, 6, "matrix")):
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
        opts = dict(cases.FULL_NOISE, compute_densitymatrix=(mode == "matrix"))
        engines = []

        def factory(nq):
            e = distributed.ShardedPauliEngine(nq, comm, device=local)
            engines.append(e)
            return e

        be = DmSimulatorB200(_engine_factory=factory)
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
                              "nvlink_bytes_sent_per_rank": engines[0].nvlink_bytes_sent}))
        dist.barrier()
    if rank == 0:
        assert worst <= 1e-10, worst
        print("DIST_CHECK_OK world=%d worst=%.3e" % (world, worst))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
