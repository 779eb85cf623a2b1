"""On-GPU micro-experiments for the tile kernel (writes JSON lines to stdout).

Measures, at n = 14 (2 GiB state): a torch copy for reference, then the tile pass as a
function of (tile digit set, number of fused ops, op kind/digits).  Used to choose the
scheduler knobs (reserve_low, max ops per pass) with data; results are summarised under
profiles/."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from qiskit_aakash_b200 import capi, engine, schedule  # noqa: E402

N = int(os.environ.get("PROBE_N", 14))


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    e = engine.PauliEngine(N)
    e.init_product([[1, 0, 0, 1]] * N, 0.5 ** N)
    bytes_pass = 16.0 * 4 ** N
    src = torch.empty(4 ** N, dtype=torch.float64, device="cuda")
    ms = timed(lambda: src.copy_(e.state))
    print(json.dumps({"probe": "torch_copy", "ms": ms, "GBps": bytes_pass / ms / 1e6}))

    rng = np.random.default_rng(0)

    def rand_rot():
        return engine.gate_matrix("u3", rng.uniform(0, 6, 3), {"rz": [0.999, 0], "ry": [0.999, 0]})

    tiles = {"low6": [0, 1, 2, 3, 4, 5], "r2_high": [0, 1, 10, 11, 12, 13], "r2_mid": [0, 1, 5, 6, 7, 8],
             "r1_mid": [0, 4, 5, 6, 7, 8]}
    tsp = engine.cx_coefficients([0.999, 0.0])
    variants = [int(x) for x in os.environ.get("PROBE_VARIANTS", "0,2,3,1").split(",")]
    for variant, (tname, tile) in [(v, it) for v in variants for it in tiles.items()]:
        e.ctx.set_tile_variant(variant)
        for n_ops in (0, 2, 3, 4, 5, 6, 8):
            for kind in ("cx_tsp_mats", "cx_ideal_nomats"):
                if n_ops == 0 and kind != "cx_tsp_mats":
                    continue
                ops = []
                for k in range(n_ops):
                    a, b = rng.choice(tile[2:] if k % 3 else tile, size=2, replace=False)
                    if kind == "cx_tsp_mats":
                        ops.append(schedule.DevOp(capi.OP_CX_TSP, int(a), int(b), rand_rot(), rand_rot(), tsp))
                    else:
                        ops.append(schedule.DevOp(capi.OP_CX, int(a), int(b), None, None, None))
                passes = schedule.encode_passes([(tile, ops)])
                ms = timed(lambda: e.ctx.apply_passes(e.sptr, e.n_bits, passes))
                print(json.dumps({"probe": "tile_pass", "variant": variant, "tile": tname, "digits": tile, "n_ops": n_ops, "kind": kind,
                                  "ms": round(ms, 4), "GBps": round(bytes_pass / ms / 1e6, 1)}))
                sys.stdout.flush()


if __name__ == "__main__":
    main()
