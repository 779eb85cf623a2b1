"""torchrun --nproc-per-node 2 tools/nvlink_probe.py : what can the slot-swap exchange reach over NVLink?

Every rank times, on its own stream with CUDA events (max over ranks):
  * an op-free tile pass that PULLS every tile from the peer's buffer (``dmb_apply_pass_remote``), for tile
    shapes with 32 KiB / 512 B / 128 B contiguous runs;
  * the same pass PUSHING every tile into the peer's buffer;
  * the same passes with 8 fused CX+maps ops (what the exchange pass usually carries);
  * an NCCL send/recv of the whole shard, and a local op-free pass, as yardsticks.
Prints one JSON line per probe (GB/s = bytes crossing NVLink per rank and direction / time).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from qiskit_aakash_b200 import capi, distributed, engine as eng, schedule
    n_bits = int(os.environ.get("PROBE_BITS", 28))          # 2 GiB per buffer
    nd = n_bits // 2
    size = 1 << n_bits
    comm = distributed.TorchCommunicator()
    alloc = eng.TorchCudaAllocator(local)
    ctx = eng.shared_context(capi.load_library(), local)
    ctx.set_stream(alloc.stream())
    a, b = alloc.empty(size), alloc.empty(size)
    a.fill_(0.5)
    b.fill_(0.25)
    peers = comm.peer_addresses(ctx, [a.data_ptr(), b.data_ptr()])
    peer = (rank + 1) % world
    tab_bits = 3
    B = n_bits - tab_bits
    tab_pull = np.full(1 << tab_bits, peers[0][peer], dtype=np.uint64)       # read the peer's `a`
    tab_push = np.full(1 << tab_bits, peers[1][peer], dtype=np.uint64)       # write the peer's `b`
    tab_self = np.full(1 << tab_bits, a.data_ptr(), dtype=np.uint64)

    def make_pass(tile_digits, n_ops):
        P = np.zeros(1, dtype=capi.PASS_DTYPE)
        K = len(tile_digits)
        P[0]["n_tile_digits"] = K
        P[0]["tile_digit"][:K] = tile_digits
        rng = np.random.default_rng(5)
        for k in range(n_ops):
            op = P[0]["ops"][k]
            la, lb = [(2, 3), (4, 5), (3, 4), (2, 5)][k % 4]
            op["kind"] = capi.OP_CX_TSP
            op["a"], op["b"] = la, lb
            op["flags"] = 3
            m = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            full = np.zeros((3, 4))
            full[:, 1:] = m
            op["pa"] = full.ravel()
            op["pb"] = full.ravel()
            op["coef"][:5] = eng.cx_coefficients((0.999, 0.0))
            op["fd"][:K - 2] = schedule.lane_order(K, la, lb)
        P[0]["n_ops"] = n_ops
        return P

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(t.item())

    shapes = {"run32KiB": [0, 1, 2, 3, 4, 5], "run512B": [0, 1, 2, nd - 3, nd - 2, nd - 1],
              "run128B": [0, 1, nd - 4, nd - 3, nd - 2, nd - 1]}
    nbytes = 8.0 * size
    out = []
    for n_ops in (0, 8):
        for name, td in shapes.items():
            P = make_pass(td, n_ops)
            ms = timed(lambda: ctx.apply_pass_remote(b.data_ptr(), n_bits, P, tab_pull, B))
            out.append({"probe": "pull", "shape": name, "ops": n_ops, "ms": ms, "GBps": nbytes / ms / 1e6})
            ms = timed(lambda: ctx.apply_pass_remote(a.data_ptr(), n_bits, P, tab_push, B, push=True))
            out.append({"probe": "push", "shape": name, "ops": n_ops, "ms": ms, "GBps": nbytes / ms / 1e6})
            ms = timed(lambda: ctx.apply_pass_remote(b.data_ptr(), n_bits, P, tab_self, B))
            out.append({"probe": "local_out_of_place", "shape": name, "ops": n_ops, "ms": ms, "GBps": nbytes / ms / 1e6})

    def nccl():
        ops = [dist.P2POp(dist.isend, a, peer), dist.P2POp(dist.irecv, b, (rank - 1) % world)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    ms = timed(nccl)
    out.append({"probe": "nccl_sendrecv", "ms": ms, "GBps": nbytes / ms / 1e6})
    if rank == 0:
        for o in out:
            o["world"] = world
            o["bytes"] = nbytes
            print(json.dumps(o))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
