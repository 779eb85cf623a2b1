"""cProfile of one end-to-end backend.run(...).result() of the bench workload (host-side view)."""
import cProfile
import copy
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from qiskit_aakash_b200 import BasicAer, assemble, circuits
    n, depth, _ = bench.workload(1)
    opts = dict(circuits.noisy_options(), compute_densitymatrix=False)
    if len(sys.argv) > 1 and sys.argv[1] == "grover12":
        qobj = assemble(circuits.grover(7, "1011001", 1))
        opts = dict(circuits.grover_options(), compute_densitymatrix=False)
    elif len(sys.argv) > 1 and sys.argv[1] == "qft8":
        qobj = assemble(circuits.qft(8))
        opts = {}
    else:
        if len(sys.argv) > 2:
            n, depth = int(sys.argv[1]), int(sys.argv[2])
        qobj = assemble(circuits.random_layered(n, depth, 100 * n))
    backend = BasicAer.get_backend("dm_simulator")
    for _ in range(2):
        backend.run(qobj, backend_options=copy.deepcopy(opts)).result()
    for _ in range(3):
        t0 = time.perf_counter()
        backend.run(qobj, backend_options=copy.deepcopy(opts)).result()
        dt = time.perf_counter() - t0
        print("wall %.1f ms" % (1e3 * dt), {k: round(1e3 * v, 1) for k, v in backend.last_engine_stats.items()
                                          if k.startswith("t_")})
    pr = cProfile.Profile()
    pr.enable()
    backend.run(qobj, backend_options=copy.deepcopy(opts)).result()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)


if __name__ == "__main__":
    main()
