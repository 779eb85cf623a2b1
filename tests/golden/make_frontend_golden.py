"""Generates tests/golden/frontend_golden.json by running tests/frontend_cases.py through the
UNMODIFIED reference front-end (oracle/frontend_harness.py; needs /root/reference).

    python tests/golden/make_frontend_golden.py
"""
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    warnings.simplefilter("ignore")
    from oracle import frontend_harness
    import frontend_cases
    api = frontend_harness.load()
    out = {}
    for name, build in frontend_cases.CASES.items():
        out[name] = api.lower(build(api))
        out[name].pop("name")
        print("%-30s %3d qubits %4d instructions" % (name, out[name]["n_qubits"], len(out[name]["instructions"])))
    with open(os.path.join(HERE, "frontend_golden.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
