"""Generate ``tests/golden/golden.npz`` by running the reference's *own*, unmodified
``dm_simulator.py`` / ``basicaertools.py`` (loaded by path through ``oracle/ref_harness.py``)
on every case in ``tests/cases.py``.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The fixture travels with the repo; nothing at test time reads /root/reference.
Keys: ``<case>/levels``, ``<case>/coeff`` (full vector for n <= 7, else a strided sample
``vec[::stride]`` with ``<case>/coeff_sum``, ``<case>/coeff_sumsq``), ``<case>/dm`` (complex
matrix when the case computes it and n <= 5), and one entry per result ``data`` key
(dict-valued probabilities are stored as the value array in key order plus ``.../keys``).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import ref_harness  # noqa: E402


def pack(name, n, result, out):
    out[name + "/levels"] = np.array(result["number_of_clock_cycles"])
    for key, val in result["data"].items():
        if key == "coeffmatrix":
            vec = np.asarray(val, dtype=float).reshape(-1)
            if n <= 7:
                out[name + "/coeff"] = vec
            else:
                out[name + "/coeff"] = vec[::cases.SAMPLE_STRIDE[n]].copy()
                out[name + "/coeff_sum"] = np.array(vec.sum())
                out[name + "/coeff_sumsq"] = np.array(np.dot(vec, vec))
        elif key == "densitymatrix":
            if n <= 5:
                out[name + "/dm"] = np.asarray(val)
        elif isinstance(val, dict):
            out[name + "/" + key] = np.array(list(val.values()), dtype=float)
            out[name + "/" + key + "/keys"] = np.array(list(val.keys()))
        else:
            out[name + "/" + key] = np.asarray(val, dtype=float)


def main():
    import tempfile
    out = {}
    for name in cases.CASES:
        case = cases.get(name)
        tmp = tempfile.mkdtemp(prefix="dmb_golden_")
        cases.write_files(case, tmp)
        os.chdir(tmp)
        res = ref_harness.run_reference(case["n"], case["instrs"], case["options"], name=name)
        pack(name, case["n"], res, out)
        print("%-32s n=%-2d levels=%-3d keys=%s" % (name, case["n"], res["number_of_clock_cycles"],
                                                    sorted(res["data"].keys())))
    path = os.path.join(HERE, "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
