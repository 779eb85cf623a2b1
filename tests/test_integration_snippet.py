"""INTEGRATION.md section 3 as an executable test: the raw ctypes calls a maintainer of the reference would
write (no helper of this package except the struct mirrors and lane_order), run against the emulated-kernel
build of the same C ABI with NumPy buffers, checked against the oracle.  Keeps the document honest."""
import ctypes

import numpy as np

from emu_backend import emu_lib
from oracle import dm_oracle
from qiskit_aakash_b200 import capi, circuits as C, engine as eng, schedule


def test_raw_ctypes_calls_of_the_integration_guide():
    lib = ctypes.CDLL(emu_lib()._name)                     # same symbols as libdmb200.so
    lib.dmb_last_error.restype = ctypes.c_char_p
    vp, i32, u64, f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_double
    n = 7
    ctx = vp()
    assert lib.dmb_create(0, ctypes.byref(ctx)) == 0
    state = np.empty(4 ** n)                               # replaces self._densitymatrix (a torch CUDA tensor on the GPU)
    sp = vp(state.ctypes.data)

    # _initialize_densitymatrix: product state, qubit q at digit n-1-q
    hi = np.array([2 * (n - 1 - q) + 1 for q in range(n)], np.int32)
    lo = np.array([2 * (n - 1 - q) for q in range(n)], np.int32)
    v = np.ascontiguousarray(np.tile(np.array([1., 0., 0., 1.]), (n, 1)))
    assert lib.dmb_init_product(ctx, sp, i32(2 * n), u64(0), i32(n), hi.ctypes.data_as(vp), lo.ctypes.data_as(vp),
                                v.ctypes.data_as(vp), f64(0.5 ** n)) == 0

    # one hand-built pass: u3 on qubits 0 and 1 folded into a CNOT(0 -> 1)
    ang0, ang1 = (0.3, 0.2, 0.1), (1.1, -0.4, 0.7)
    M0 = eng.gate_matrix("u3", list(ang0), {"rx": [1, 0], "ry": [1, 0], "rz": [1, 0]})
    M1 = eng.gate_matrix("u3", list(ang1), {"rx": [1, 0], "ry": [1, 0], "rz": [1, 0]})
    K = 6
    passes = np.zeros(1, dtype=capi.PASS_DTYPE)
    p = passes[0]
    p["n_tile_digits"] = K
    p["tile_digit"][:K] = [0, 1, 2, 3, n - 2, n - 1]       # ascending digit positions, 0 always in the tile
    op = p["ops"][0]
    op["kind"] = capi.OP_CX
    op["a"], op["b"] = 5, 4                                # control qubit 0 = digit n-1 = local 5, target qubit 1 = local 4
    op["flags"] = 1 | 2
    op["pa"] = np.asarray(M0)[1:4].ravel()
    op["pb"] = np.asarray(M1)[1:4].ravel()
    op["fd"][:K - 2] = schedule.lane_order(K, 5, 4)
    p["n_ops"] = 1
    rc = lib.dmb_apply_passes(ctx, sp, i32(2 * n), passes.ctypes.data_as(vp), ctypes.c_size_t(1))
    assert rc == 0, lib.dmb_last_error().decode()

    # ... and a scheduled stream: two more CNOTs packed by dmb_schedule
    qops = np.zeros(2, dtype=capi.QOP_DTYPE)
    qops["kind"] = capi.OP_CX
    qops["qa"], qops["qb"] = [1, 2], [2, 6]
    pos = np.array([n - 1 - q for q in range(n)], np.int32)
    out = np.zeros(2, dtype=capi.PASS_DTYPE)
    n_out = ctypes.c_size_t()
    rc = lib.dmb_schedule(qops.ctypes.data_as(vp), ctypes.c_size_t(2), pos.ctypes.data_as(vp), i32(n), i32(n), i32(6), i32(16),
                          i32(256), i32(1), ctypes.c_size_t(0), None, None, out.ctypes.data_as(vp), ctypes.c_size_t(2),
                          ctypes.byref(n_out), None, None)
    assert rc == 0, lib.dmb_last_error().decode()
    assert lib.dmb_apply_passes(ctx, sp, i32(2 * n), out.ctypes.data_as(vp), n_out) == 0

    # _get_densitymatrix: chop + download (layout: pos[] after the last pass)
    assert lib.dmb_chop(ctx, sp, u64(4 ** n), f64(1e-15)) == 0
    got = state.reshape([4] * n)
    # digit position p is axis n-1-p; qubit q sits at pos[q] -> bring the axes back to qubit order
    got = np.transpose(got, [n - 1 - int(pos[q]) for q in range(n)]).reshape(-1)

    circ = C.Circuit(n)
    circ.u3(*ang0, 0); circ.u3(*ang1, 1); circ.cx(0, 1); circ.cx(1, 2); circ.cx(2, 6)
    ref = dm_oracle.run_oracle(n, circ.instructions, {"compute_densitymatrix": False})["data"]["coeffmatrix"]
    assert np.max(np.abs(got - ref)) <= 1e-12
    assert lib.dmb_destroy(ctx) == 0
