"""B200-native implementation of the qiskit-aakash ``dm_simulator`` hot path.

    from qiskit_aakash_b200 import QuantumCircuit, BasicAer, execute
    qc = QuantumCircuit(2); qc.x(1); qc.cx(0, 1)
    backend = BasicAer.get_backend('dm_simulator')
    result = execute(qc, backend, **options).result()

See DESIGN.md for the path, the data layout and the kernels; INTEGRATION.md for how the
C ABI (include/dmb200.h) plugs into the reference.
"""
from .exceptions import BasicAerError, QiskitError          # noqa: F401
from .circuits import Circuit                               # noqa: F401
from .dm_simulator import BasicAer, DmSimulatorB200, execute, assemble   # noqa: F401
from .frontend import (QuantumCircuit, QuantumRegister, ClassicalRegister, transpile, pi,   # noqa: F401
                       install_as_qiskit)

__version__ = "0.1.0"
