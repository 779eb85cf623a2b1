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


def test_pull_pass_of_the_threaded_body_equals_the_sequential_emulation():
    """Fused exchange, pull flavour (REMOTE = 1 of dmb_half_kernel_body): every 16-byte pair of a tile is read through
    the source table; its index is gathered from up to five SELECTED index bits and composed from a per-tile, a
    per-thread and a per-pair part.  Real host threads per CTA against the element-by-element emulation behind
    ``dmb_apply_pass_remote_sel``, for high-bit tables (top-slot swap) and scattered ones (direct-slot swap)."""
    import ctypes
    from emu_backend import emu_lib
    from qiskit_aakash_b200 import capi, schedule
    lib = emu_lib()
    raw = ctypes.CDLL(lib._name)
    rng = np.random.default_rng(11)
    n_bits = 15                                   # 7.5 digits: 8 tiles, one straddling half digit
    size = 1 << n_bits
    for trial in range(6):
        n_sel = int(rng.integers(1, 6))
        sel = sorted(int(b) for b in rng.choice(np.arange(4, n_bits), n_sel, replace=False)) if trial % 2 else \
            list(range(n_bits - n_sel, n_bits))
        sources = [rng.standard_normal(size) for _ in range(1 << n_sel)]       # one "peer buffer" per table entry
        # entry k is pre-offset so that element idx lies at tab[k] + 8 * idx: here a per-entry shift of the index
        shifts = [int(rng.integers(0, 4)) * 16 for _ in range(1 << n_sel)]
        padded = [np.concatenate([np.zeros(64), s_, np.zeros(64)]) for s_ in sources]
        tab = np.array([(p.ctypes.data + 8 * 64 + 8 * sh) % (1 << 64) for p, sh in zip(padded, shifts)], dtype=np.uint64)
        P = np.zeros(1, dtype=capi.PASS_DTYPE)
        P[0]["n_tile_digits"] = 6
        digits = [0, 1] + sorted(int(d) for d in rng.choice(np.arange(2, 7), 4, replace=False))
        P[0]["tile_digit"][:6] = digits
        P[0]["n_ops"] = 2
        for k, (a, b) in enumerate(((2, 4), (0, 3))):
            op = P[0]["ops"][k]
            op["kind"], op["a"], op["b"] = capi.OP_CX, a, b
            op["fd"][:4] = schedule.lane_order(6, a, b)
            op["flags"] = capi.HAS_PA | capi.HAS_PB
            op["pa"][:] = rng.standard_normal(12)
            op["pb"][:] = rng.standard_normal(12)
        sel_arr = np.array(sel, dtype=np.int32)
        want = np.zeros(size)
        capi.Context(lib, 0).apply_pass_remote_sel(want.ctypes.data, n_bits, P, tab, sel_arr)
        got = np.zeros(size)
        rc = raw.dmb_emu_run_half_kernel_pull(ctypes.c_void_p(got.ctypes.data), n_bits, ctypes.c_void_p(P.ctypes.data),
                                              ctypes.c_void_p(tab.ctypes.data), n_sel, ctypes.c_void_p(sel_arr.ctypes.data), 3)
        assert rc == 0
        assert got.tobytes() == want.tobytes(), (trial, sel)
        assert np.any(want != 0)

