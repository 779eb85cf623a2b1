"""Errors of this backend, mirroring ``qiskit/exceptions.py`` / ``providers/basicaer/exceptions.py:22-32``."""


class QiskitError(Exception):
    """Base class for errors raised by the simulator tools."""

    def __init__(self, *message):
        super().__init__(" ".join(str(m) for m in message))
        self.message = " ".join(str(m) for m in message)

    def __str__(self):
        return repr(self.message)


class BasicAerError(QiskitError):
    """Base class for errors raised by Basic Aer backends."""
