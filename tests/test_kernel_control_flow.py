"""The tile kernel's control flow on the CPU (tests/emu: dmb_emu_run_half_kernel).

``dmb_half_kernel_body`` (csrc/dm_device.h) is the function the CUDA kernel ``k_tile_pass6`` wraps: tile
loop, one or two staging stages, asynchronous 16-byte copies, CTA barriers, the two virtual threads per real
thread and the paired op bodies.  Here it runs with 128 real host threads per CTA, a CTA barrier and cp.async
emulated as copies that only land at ``wait<N>()``, on passes produced by the real scheduler (relabelling
stores included), and must reproduce the sequential emulation of the kernel's per-thread bodies bit for bit."""
import ctypes

import numpy as np
import pytest

from emu_backend import emu_lib
from qiskit_aakash_b200 import capi, circuits as C, engine as eng, schedule


def _passes(n, layers, seed, max_ops):
    circ = C.random_layered(n, layers, seed, readout=False)
    rng = np.random.default_rng(seed)
    qops = []
    for ins in circ.instructions:
        if ins.name != "cx":
            continue
        pa = eng.gate_matrix("u3", list(rng.uniform(0, 6, 3)), {"rx": [0.99, 0.01], "ry": [0.98, 0.0], "rz": [0.97, 0.02]})
        pb = eng.gate_matrix("u3", list(rng.uniform(0, 6, 3)), {"rx": [1, 0], "ry": [1, 0], "rz": [1, 0]})
        kind = capi.OP_CX_TSP if rng.random() < 0.7 else capi.OP_CX
        coef = eng.cx_coefficients([0.995, 0.0 if rng.random() < 0.5 else 0.01]) if kind == capi.OP_CX_TSP else None
        qops.append(schedule.DevOp(kind, ins.qubits[0], ins.qubits[1], pa, pb, coef))
    pos = [n - 1 - q for q in range(n)]
    return schedule.relabel_passes(capi.load_library(), qops, pos, n, max_ops=max_ops, native=True,
                                   strategy=capi.SCHED_TILE_SEARCH)


@pytest.mark.parametrize("paired,stages,grid", [(1, 1, 3), (1, 1, 2), (1, 2, 3), (1, 2, 1), (1, 1, 5)])
def test_threaded_kernel_body_equals_sequential_emulation(paired, stages, grid):
    n = 8                                              # 16 tiles of 4^6 coefficients
    P = _passes(n, 6, 40 + paired + 2 * stages, 8)
    assert len(P) >= 3 and all(int(p["n_tile_digits"]) == 6 for p in P)
    lib = emu_lib()
    raw = ctypes.CDLL(lib._name)
    rng = np.random.default_rng(7)
    start = rng.standard_normal(4 ** n)
    want = start.copy()
    ctx = capi.Context(lib, 0)
    ctx.set_tile_variant(0)
    ctx.apply_passes(want.ctypes.data, 2 * n, P)
    got = start.copy()
    rc = raw.dmb_emu_run_half_kernel(ctypes.c_void_p(got.ctypes.data), ctypes.c_int(2 * n), P.ctypes.data_as(ctypes.c_void_p),
                                     ctypes.c_size_t(len(P)), ctypes.c_int(paired), ctypes.c_int(stages), ctypes.c_int(grid))
    assert rc == 0
    assert np.array_equal(got, want)
    assert not np.array_equal(got, start)
