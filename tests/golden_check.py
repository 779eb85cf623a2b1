"""Compare a result dict (reference result-dict schema, dm_simulator.py:1189-1196) with the
committed golden fixture of a case.  Used for the oracle and for the CUDA backend alike."""
import numpy as np

import cases

TOL = 1e-10          # BASELINE.json north_star: max |delta| <= 1e-10 on coefficients and probabilities
TRACE_TOL = 1e-12    # trace preserved to 1e-12


def check_against_golden(golden, name, result, tol=TOL, check_trace=True):
    n = cases.CASES[name]["n"]
    assert int(golden[name + "/levels"]) == result["number_of_clock_cycles"], "level count"
    data = result["data"]
    prefix = name + "/"
    expected_keys = set()
    for k in golden.files:
        if k.startswith(prefix):
            expected_keys.add(k[len(prefix):].split("/")[0])
    expected_keys -= {"levels", "coeff_sum", "coeff_sumsq"}
    got = set(data.keys())
    got = {("coeff" if k == "coeffmatrix" else "dm" if k == "densitymatrix" else k) for k in got}
    if n > 5:
        got.discard("dm")
    assert expected_keys == got, (expected_keys, got)
    worst = 0.0
    vec = np.asarray(data["coeffmatrix"], dtype=float).reshape(-1)
    assert vec.size == 4 ** n
    ref = golden[name + "/coeff"]
    if n <= 7:
        worst = max(worst, float(np.max(np.abs(vec - ref))))
    else:
        worst = max(worst, float(np.max(np.abs(vec[::cases.SAMPLE_STRIDE[n]] - ref))))
        assert abs(vec.sum() - float(golden[name + "/coeff_sum"])) <= 1e-9
        assert abs(np.dot(vec, vec) - float(golden[name + "/coeff_sumsq"])) <= 1e-9
    if check_trace and not cases.CASES[name]["options"].get("chop_threshold"):
        assert abs(vec[0] * 2 ** n - 1.0) <= TRACE_TOL, "trace"
    if name + "/dm" in golden.files:
        worst = max(worst, float(np.max(np.abs(np.asarray(data["densitymatrix"]) - golden[name + "/dm"]))))
    for key, val in data.items():
        if key in ("coeffmatrix", "densitymatrix"):
            continue
        g = golden[prefix + key]
        if isinstance(val, dict):
            assert list(val.keys()) == [str(x) for x in golden[prefix + key + "/keys"]], key
            arr = np.array(list(val.values()), dtype=float)
        else:
            arr = np.asarray(val, dtype=float)
        worst = max(worst, float(np.max(np.abs(arr - g))))
    assert worst <= tol, "max |delta| = %g" % worst
    return worst
