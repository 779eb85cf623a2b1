"""Single-GPU probe: cost of a tile pass whose trailing swaps are folded into the write-back
(dmb_make_lean_pass(fold_swaps)) versus the same swaps run as shared-memory ops (DMB_FOLD_SWAPS=0)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from qiskit_aakash_b200 import capi, engine as eng, schedule
    n_bits = 28
    alloc = eng.TorchCudaAllocator(0)
    ctx = eng.shared_context(capi.load_library(), 0)
    ctx.set_stream(alloc.stream())
    state = alloc.empty(1 << n_bits)
    state.fill_(0.5)
    rng = np.random.default_rng(5)

    def make_pass(tile_digits, n_ops, swaps):
        P = np.zeros(1, dtype=capi.PASS_DTYPE)
        K = len(tile_digits)
        P[0]["n_tile_digits"] = K
        P[0]["tile_digit"][:K] = tile_digits
        k = 0
        for k in range(n_ops):
            op = P[0]["ops"][k]
            la, lb = [(2, 3), (4, 5), (3, 4), (2, 5), (0, 2), (1, 3)][k % 6]
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
        for j, (a, b) in enumerate(swaps):
            op = P[0]["ops"][n_ops + j]
            op["kind"] = capi.OP_SWAP
            op["a"], op["b"] = a, b
            op["fd"][:K - 2] = schedule.lane_order(K, a, b)
        P[0]["n_ops"] = n_ops + len(swaps)
        return P

    def timed(P, reps=10):
        for _ in range(3):
            ctx.apply_passes(state.data_ptr(), n_bits, P)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.apply_passes(state.data_ptr(), n_bits, P)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    td = [0, 1, 5, 6, 7, 8]
    for n_ops in (0, 4, 8):
        for name, swaps in (("none", []), ("swap(1,x)", [(1, 3)]), ("swap(0,x)", [(0, 3)]),
                            ("swap(0,x)+swap(1,y)", [(0, 3), (1, 4)]), ("swap(1,x)+swap(0,y)", [(1, 3), (0, 4)])):
            ms = timed(make_pass(td, n_ops, swaps))
            print(json.dumps({"probe": "store", "fold": os.environ.get("DMB_FOLD_SWAPS", "1"), "ops": n_ops,
                              "swaps": name, "ms": round(ms, 4)}))


if __name__ == "__main__":
    main()
