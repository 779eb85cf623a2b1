"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): n = 7 noisy layered circuit
through the backend (lean tile kernel, swaps, marginal, FWHT, matrix conversion, chop, download)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from qiskit_aakash_b200 import BasicAer, execute, circuits  # noqa: E402

circ = circuits.random_layered(7, 4, 7)
opts = dict(circuits.noisy_options(), **circuits.grover_options())
res = execute(circ, BasicAer.get_backend("dm_simulator"), **opts).result()
d = res["results"][0]["data"]
print("sanitize_smoke ok: prob sum %.15f trace %.15f" % (sum(d["ensemble_probability"].values()), d["coeffmatrix"][0] * 2 ** 7))
