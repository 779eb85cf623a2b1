"""TEST / BENCH INFRASTRUCTURE -- stage the reference's hot path for the GPU box.

``/root/reference`` exists only in the build container.  The reference's dm_simulator path is three Python
files; this recipe copies exactly those, unmodified, from where they lie into ``oracle/_ref/`` (git-ignored, NOT
gpurun-ignored, so the copy travels with the snapshot like a built ``.so``) so that ``oracle/ref_harness.py`` can
load the REAL reference on the GPU box and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs time it there
(``cpu_baseline.kind == "reference"``) instead of the faster oracle port.  Nothing under ``oracle/_ref`` is ever
committed, imported by the product package or used by the ``-m gpu`` tests.

    python oracle/stage_ref.py            # no-op when /root/reference is absent
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/qiskit/providers/basicaer"
DST = os.path.join(HERE, "_ref", "qiskit", "providers", "basicaer")
FILES = ("exceptions.py", "basicaertools.py", "dm_simulator.py")


def stage():
    """Returns 'staged' / 'fresh' / 'present' (no reference tree here, staged copy exists) / 'absent'."""
    have = all(os.path.isfile(os.path.join(DST, f)) for f in FILES)
    if not all(os.path.isfile(os.path.join(SRC, f)) for f in FILES):
        return "present" if have else "absent"
    if have and all(filecmp.cmp(os.path.join(SRC, f), os.path.join(DST, f), shallow=False) for f in FILES):
        return "fresh"
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return "staged"


if __name__ == "__main__":
    print(stage())
