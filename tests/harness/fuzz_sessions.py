"""Randomised check of option STICKINESS across consecutive jobs on one backend object (TEST TOOL;
needs /root/reference): the reference keeps several option values from earlier runs
(dm_simulator.py:177-271 mutates instance / class state), and the backend
mirror must reproduce that number for number.  A session = several jobs with random option dicts
on ONE reference ``DmSimulatorPy`` (``_set_options`` in the parent, ``run_experiment`` on a copy,
like the forked worker of basicaerjob.py:51-54) and on ONE emulated-kernel backend.

    python tests/harness/fuzz_sessions.py [--seeds 100] [--start 0]
"""
import argparse
import contextlib
import copy
import io
import os
import sys
from types import SimpleNamespace as NS

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--start", type=int, default=0)
    a = ap.parse_args()
    from oracle import ref_harness
    import fuzz_emu
    from emu_backend import emu_backend
    from qiskit_aakash_b200 import circuits as C
    from qiskit_aakash_b200.dm_simulator import assemble
    dm, _ = ref_harness.load()
    import tempfile
    base_dir = tempfile.mkdtemp(prefix="dmb_fuzz_sessions_")
    counts = {}
    jobs = {'compared': 0, 'both-raise': 0}
    for seed in range(a.start, a.start + a.seeds):
        rng = np.random.default_rng(seed)
        cls = dm.DmSimulatorPy
        cls.DEFAULT_OPTIONS = dict(cls.DEFAULT_OPTIONS)
        cls.DEFAULT_OPTIONS["rotation_error"] = {"rx": [1., 0.], "ry": [1., 0.], "rz": [1., 0.]}
        sim = cls()
        be = emu_backend()
        dirs = [os.path.join(base_dir, "ref_%d" % seed), os.path.join(base_dir, "own_%d" % seed)]   # one cwd per side:
        for d in dirs:                                                                            # each keeps its own
            os.makedirs(d)                                                                        # stored_coefficients.npy
        st, msg = "ok", ""
        for job in range(int(rng.integers(2, 6))):
            n = int(rng.integers(1, 6))
            n_exp = 1 if rng.random() < 0.6 else int(rng.integers(2, 4))      # experiments of this job (same register size)
            circs = [fuzz_emu.random_circuit(rng, n, 30) for _ in range(n_exp)]
            opts = fuzz_emu.random_options(rng)
            fuzz_emu.random_init(rng, n, opts)
            if rng.random() < 0.3:
                opts["compute_densitymatrix"] = bool(rng.random() < 0.5)
            # store / compare / stored initial state go through stored_density_matrix.npy in the working directory
            # (dm_simulator.py:476-480, 1271-1282); both sides write and read the same file name
            r = rng.random()
            if r < 0.15:
                opts["store_densitymatrix"] = True
            elif r < 0.3:
                opts["compare"] = True
            elif r < 0.4:
                opts["custom_densitymatrix"] = "stored_density_matrix"      # np.load("stored_density_matrix.npy"): absent -> both raise
            ref = got = e_ref = e_got = None
            with contextlib.redirect_stdout(io.StringIO()):
                try:
                    os.chdir(dirs[0])
                    sim._set_options(qobj_config=NS(n_qubits=n), backend_options=copy.deepcopy(opts))
                    worker = copy.deepcopy(sim)
                    ref = []
                    for circ in circs:              # one worker runs all experiments of the job (dm_simulator.py:934-935)
                        exp = NS(config=NS(n_qubits=n, memory_slots=n),
                                 instructions=ref_harness.to_reference_instructions(copy.deepcopy(circ.instructions)),
                                 header=NS(name="fuzz", as_dict=lambda: {"name": "fuzz"}))
                        ref.append(worker.run_experiment(exp))
                except Exception as e:  # noqa: BLE001
                    e_ref = e
                try:
                    os.chdir(dirs[1])
                    mine = []
                    for circ in circs:
                        c2 = C.Circuit(n)
                        c2.instructions = copy.deepcopy(circ.instructions)
                        mine.append(c2)
                    got = be.run(assemble(mine), backend_options=copy.deepcopy(opts)).result()["results"]
                except Exception as e:  # noqa: BLE001
                    e_got = e
            if e_ref or e_got:
                if not (e_ref and e_got):
                    st, msg = "FAIL", "job %d n=%d opts=%r reference: %r / backend: %r" % (job, n, opts, e_ref, e_got)
                    break
                jobs['both-raise'] += 1
                continue
            jobs['compared'] += 1
            for e, (r_e, g_e) in enumerate(zip(ref, got)):
                if r_e["number_of_clock_cycles"] != g_e["number_of_clock_cycles"] or set(r_e["data"]) != set(g_e["data"]):
                    st, msg = "FAIL", "job %d exp %d: levels/keys %s vs %s (opts %r)" % (job, e, sorted(r_e["data"]),
                                                                                        sorted(g_e["data"]), opts)
                    break
                for k, v in r_e["data"].items():
                    x, y = fuzz_emu.as_arr(v), fuzz_emu.as_arr(g_e["data"][k])
                    if x.shape != y.shape or (x.size and float(np.max(np.abs(x - y))) > 1e-10):
                        st, msg = "FAIL", "job %d exp %d n=%d %s differs (opts %r)" % (job, e, n, k, opts)
                        break
                if st != "ok":
                    break
            if st != "ok" or len(ref) != len(got):
                st = "FAIL"
                break
        counts[st] = counts.get(st, 0) + 1
        if st == "FAIL":
            print("seed %d: %s" % (seed, msg), flush=True)
    print("summary:", counts, jobs)
    return 1 if counts.get("FAIL") else 0


if __name__ == "__main__":
    sys.exit(main())
