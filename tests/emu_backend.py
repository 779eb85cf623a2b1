"""TEST INFRASTRUCTURE: a DmSimulatorB200 whose engine is bound to the CPU emulation of the
kernels (tests/emu) and host (NumPy) buffers -- so that the host logic (merge, partition,
lowering, pass scheduling, readout plumbing) and the kernels' index arithmetic can be tested
without a GPU.  Never used by the product."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

import build_emu  # noqa: E402
from qiskit_aakash_b200 import capi, engine  # noqa: E402
from qiskit_aakash_b200.dm_simulator import DmSimulatorB200  # noqa: E402

_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        _lib = capi.load_library(build_emu.build())
    return _lib


class NumpyAllocator:
    index = 0

    def empty(self, count):
        return np.empty(int(count), dtype=np.float64)

    def ptr(self, buf):
        return buf.ctypes.data

    def stream(self):
        return 0


def emu_engine(n, **kw):
    return engine.PauliEngine(n, lib=emu_lib(), allocator=NumpyAllocator(), **kw)


def emu_backend(**kw):
    return DmSimulatorB200(_engine_factory=lambda n: emu_engine(n, **kw))
