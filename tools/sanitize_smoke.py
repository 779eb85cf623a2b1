"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): n = 7 / 9 / 5 noisy layered circuits
through the backend (lean tile kernel, swaps, marginal, FWHT, matrix conversion, chop, download), then QFT-9 ideal and
noisy (the chained instantiations of the tile kernel: two CNOTs of a cu1 in one round trip) and a compiled replay."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from qiskit_aakash_b200 import BasicAer, execute, circuits  # noqa: E402

opts = dict(circuits.noisy_options(), **circuits.grover_options())
for n in (7, 9, 5):        # 7: single-launch cluster path (4 tiles); 9: k_tile_pass6 (64 tiles); 5: below one tile
    circ = circuits.random_layered(n, 4, 7)
    res = execute(circ, BasicAer.get_backend("dm_simulator"), **dict(opts, compute_densitymatrix=(n <= 7))).result()
    d = res["results"][0]["data"]
    print("sanitize_smoke ok n=%d: prob sum %.15f trace %.15f" % (n, sum(d["ensemble_probability"].values()),
                                                                    d["coeffmatrix"][0] * 2 ** n))

for noisy in (False, True):
    circ = circuits.qft(9)
    o = dict(opts if noisy else {}, compute_densitymatrix=False)
    be = BasicAer.get_backend("dm_simulator")
    res = execute(circ, be, **o).result()
    d = res["results"][0]["data"]
    print("sanitize_smoke ok qft9 noisy=%s: prob sum %.15f trace %.15f chained ops %d"
          % (noisy, sum(d["ensemble_probability"].values()), d["coeffmatrix"][0] * 2 ** 9, be.last_engine_stats["chained_ops"]))
from qiskit_aakash_b200 import assemble  # noqa: E402
be = BasicAer.get_backend("dm_simulator")
compiled = be.compile(assemble(circuits.qft(8)), backend_options={"compute_densitymatrix": False})
for _ in range(2):
    d = compiled.run()["results"][0]["data"]
print("sanitize_smoke ok replay qft8: prob sum %.15f" % sum(d["ensemble_probability"].values()))
