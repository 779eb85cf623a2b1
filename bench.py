#!/usr/bin/env python
"""bench.py -- headline benchmark of the dm_simulator hot path on B200.

Metric (BASELINE.json): effective HBM GB/s per gate (and circuit wall time) on the n = 14..18 circuits of
BASELINE.json's configs.  One "step" = one complete execution of the circuit (state init, every gate / noise
level, I/Z-marginal probability readout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (``--workload``; the default depends on ``--gpus``):
    config3    random layered U3+CX, n = 14, depth 200, per-gate noise          BASELINE configs[2]   default at N = 1
    layered16  the same circuit family at n = 16, depth 60 (34 GB state)        fits 1, 2, 4 and 8 GPUs: THE scaling
                                                                                curve; default at N = 2 and N = 4
    qft16      QFT n = 16, no noise, ensemble readout                           BASELINE configs[3]
    config5    random layered U3+CX, n = 18, depth 20 (550 GB state)            BASELINE configs[4]   default at N = 8
Whatever the headline workload is, every line also carries ``scaling_curve``: layered16 timed in the same
invocation (same circuit at every N -> same-workload strong scaling), and ``config4_qft16`` where it is cheap.

Prints ONE JSON line (rank 0).  ``value`` times the device path with the plan and the state resident in HBM (CUDA
events on the launching stream, max over ranks); ``e2e`` times the public API call
``BasicAer.get_backend('dm_simulator').run(qobj, backend_options).result()`` with host buffers: host-side merge /
partition / scheduling, kernel launches and the device->host copies of the results are inside its timed region.
``parity_max_abs``: before anything is timed, the SAME runner class that is timed executes the same circuit
generator at n = 10 and its probabilities and coefficients are compared with the CPU oracle (checker only).
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "effective_hbm_gbps_per_gate"
UNIT = "GB/s"

WORKLOADS = {
    # name: (n_qubits, description, circuit builder(readout), backend options)
    "config3": (14, "random layered U3+CX n=14 depth 200, rotation/tsp error r=0.999, depolarization 0.99, "
                    "ensemble-Z readout (BASELINE configs[2])"),
    "layered16": (16, "random layered U3+CX n=16 depth 60, rotation/tsp error r=0.999, depolarization 0.99, "
                      "ensemble-Z readout (BASELINE configs[2] family at 34 GB; same circuit at every GPU count)"),
    "qft16": (16, "QFT n=16, no noise, ensemble-Z readout (BASELINE configs[3])"),
    "config5": (18, "random layered U3+CX n=18 depth 20 (550 GB Pauli vector), rotation/tsp error r=0.999, "
                    "depolarization 0.99, 2^18-probability I/Z-marginal readout (BASELINE configs[4])"),
}
DEFAULT_WORKLOAD = {1: "config3", 2: "layered16", 4: "layered16", 8: "config5"}
PARITY_N, PARITY_DEPTH, PARITY_SEED = 10, 6, 1000


def build_workload(name, readout=True, qubits=0, layers=0):
    """(circuit, backend options, description)."""
    from qiskit_aakash_b200 import circuits
    if qubits:
        depth = layers or 20
        return (circuits.random_layered(qubits, depth, 100 * qubits, readout=readout), circuits.noisy_options(),
                "random layered U3+CX n=%d depth %d, rotation/tsp error r=0.999, depolarization 0.99 (--qubits override)"
                % (qubits, depth))
    desc = WORKLOADS[name][1]
    if name == "config3":
        return circuits.random_layered(14, 200, 1400, readout=readout), circuits.noisy_options(), desc
    if name == "layered16":
        return circuits.random_layered(16, 60, 1600, readout=readout), circuits.noisy_options(), desc
    if name == "qft16":
        return circuits.qft(16, readout=readout), {}, desc
    if name == "config5":
        return circuits.random_layered(18, 20, 1800, readout=readout), circuits.noisy_options(), desc
    raise SystemExit("unknown workload %r" % name)


def count_gates(circ):
    return sum(1 for i in circ.instructions if i.name in ("u1", "u2", "u3", "cx"))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own dm_simulator (oracle/_ref, staged by build()) or, when that
# is absent, the oracle port of its algorithm -- on host cores, on a WHOLE circuit at a size the CPU finishes
# --------------------------------------------------------------------------------------------

CPU_N, CPU_DEPTH, CPU_SEED = 11, 4, 1100      # the whole circuit the CPU arm times: 60 gates, 8 noisy levels


def cpu_implementation():
    """('reference', runner) when the reference's three hot-path files are staged under oracle/_ref (or
    /root/reference exists), else ('port', runner).  runner(n, instrs, opts) -> result dict."""
    from oracle import ref_harness
    if ref_harness.available():
        return "reference", ref_harness.run_reference
    from oracle import dm_oracle
    return "port", dm_oracle.run_oracle


def time_cpu_circuit(n=CPU_N, depth=CPU_DEPTH, seed=CPU_SEED, size_probe=None):
    """Times ONE whole circuit of the benchmark's family (random layered U3+CX with the workload's noise options,
    ensemble-Z readout) on the CPU at n = 11 and returns the metric exactly as the GPU arm defines it:
    gates * 16 * 4^n / wall time.  The value is for the size it ran at (stated in `sample`); an extrapolation to
    the GPU workload is reported separately and never folded into `value`.
    size_probe = (n_big, gates_big, levels_big): additionally time 2 gates + 1 noise level at n_big to scale the
    n = 11 circuit time to the GPU workload's size (`extrapolation`)."""
    from qiskit_aakash_b200 import circuits
    kind, run = cpu_implementation()
    circ = circuits.random_layered(n, depth, seed)
    opts = dict(circuits.noisy_options(), compute_densitymatrix=False)
    gates = count_gates(circ)
    t0 = time.perf_counter()
    res = run(n, copy.deepcopy(circ.instructions), copy.deepcopy(opts))
    dt = time.perf_counter() - t0
    levels = res["number_of_clock_cycles"]
    psum = float(sum(res["data"]["ensemble_probability"].values()))
    assert abs(psum - 1) < 1e-9
    value = gates * 16.0 * 4 ** n / dt / 1e9
    out = {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
           "sample": "WHOLE circuit on the CPU: random layered U3+CX n=%d depth %d (%d gates, %d noisy levels, "
                     "ensemble-Z readout) through %s, 1 thread (NumPy slicing is single-threaded): %.2f s"
                     % (n, depth, gates, levels,
                        "the reference's own dm_simulator.py (oracle/_ref)" if kind == "reference"
                        else "the oracle port (skips the reference's full-state copies: faster than the reference)", dt)}
    if size_probe is not None:
        n_big, gates_big, levels_big = size_probe
        n_probe = min(n_big, 14)                       # 8 * 4^14 = 2 GiB: what the CPU leg can hold and finish
        def sample(nn):
            full = circuits.random_layered(nn, 1, seed, readout=False)
            u3s = [i for i in full.instructions if i.name == "u3"]
            t = time.perf_counter()
            run(nn, copy.deepcopy([u3s[0], u3s[nn - 1]]), copy.deepcopy(opts))
            return time.perf_counter() - t
        t_small, t_big = min(sample(n) for _ in range(2)), sample(n_probe)
        per_size = t_big / t_small * 4.0 ** (n_big - n_probe)      # beyond n_probe: proportional to the state size
        est = dt * per_size * (levels_big / float(levels))
        out["extrapolation"] = {"to": "n=%d, %d gates, %d levels" % (n_big, gates_big, levels_big),
                                "size_factor": per_size, "estimated_seconds": est,
                                "estimated_value": gates_big * 16.0 * 4 ** n_big / est / 1e9,
                                "how": "2 gates + 1 noise level timed at n=%d (%.2f s) and n=%d (%.1f s); circuit time "
                                       "scaled by that ratio%s and by the level count"
                                       % (n, t_small, n_probe, t_big,
                                          "" if n_probe == n_big else " x 4^(%d-%d)" % (n_big, n_probe))}
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    g.stage_reference()
    for _ in range(min(args.warmup, 1)):
        time_cpu_circuit()
    runs = [time_cpu_circuit() for _ in range(args.steps)]
    secs = [r["seconds"] for r in runs]
    value = len(runs) / sum(1.0 / r["value"] for r in runs)          # equal work per step: harmonic mean of the rates
    cpu = dict(runs[-1], value=value)
    ours, _, ours_desc = build_workload(args.workload or DEFAULT_WORKLOAD.get(args.gpus, "config3"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
            "higher_is_better": True, "scaling": "strong" if args.gpus in (2, 4) else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "random layered U3+CX n=%d depth %d, rotation/tsp error r=0.999, depolarization 0.99, "
                                   "ensemble-Z readout -- the GPU arm's circuit family at the size the CPU finishes in "
                                   "seconds; value is the same metric (gates x 16 x 4^n / wall time) AT THAT SIZE"
                                   % (CPU_N, CPU_DEPTH),
                       "gpu_arm_workload": ours_desc, "same_size_as_gpu_arm": False},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

class SingleGpuRunner:
    """Device path with the compiled plan resident: init -> tile passes -> marginal -> FWHT.  The plan is built by
    the backend's own level loop (``DmSimulatorB200._run_level`` + per-level noise) on an engine that does not
    launch, i.e. exactly what ``backend.run`` would execute, minus the host work."""

    def __init__(self, n, circ, opts, device=0):
        import torch
        from qiskit_aakash_b200 import DmSimulatorB200, engine as eng, hostpass
        self.torch = torch
        self.n = n
        self.engine = e = eng.PauliEngine(n, device=device)
        e.drain_threshold = 0                    # compile the whole circuit into one resident plan
        be = DmSimulatorB200(device=device, _engine_factory=lambda nq: e)
        be._set_options(None, copy.deepcopy(opts))
        be._number_of_qubits = n
        be._initialize_errors()
        ops = hostpass.merge_single_qubit_gates(circ.instructions, n, be.MERGE)
        levels, self.n_levels = hostpass.partition_levels(ops, n)
        mem = be._error_params["memory"]
        noise = eng.memory_noise_matrix(mem["decoherence"], mem["thermalization"], mem["amplitude_decay"])
        noisy = not np.array_equal(noise, np.eye(4))
        for level in levels[:self.n_levels]:
            be._run_level(e, level, None, {})
            if noisy:
                e.apply_1q_all(noise)
        self.passes = e.plan()
        self.final_pos = list(e.pos)
        e.queue, e.pending = [], [None] * n
        self.err = be._error_params["measurement"]
        self._ev = []

    def step(self):
        torch, e = self.torch, self.engine
        e.init_product([[1, 0, 0, 1]] * self.n, 0.5 ** self.n)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e.run_passes(self.passes)
        b.record()
        self._ev.append((a, b, len(self.passes)))
        e.pos = list(self.final_pos)
        self.probs = e.marginal_probabilities("Z", self.err)

    def reset_counters(self):
        self.torch.cuda.synchronize()
        self.engine.ctx.reset_stats()
        self._ev = []

    def counters(self):
        return self.engine.stats()

    def pass_ms_total(self):
        self.torch.cuda.synchronize()
        return sum(ev[0].elapsed_time(ev[1]) for ev in self._ev)

    def timed_pass_launches(self):
        return sum(ev[2] for ev in self._ev)

    def exchange_stats(self):
        return None

    def close(self):
        self.engine = None
        self.torch.cuda.empty_cache()


def make_runner(world, n, circ, opts, device):
    if world > 1:
        from qiskit_aakash_b200 import distributed
        return distributed.ShardedCircuitRunner(n, circ, opts, device=device)
    return SingleGpuRunner(n, circ, opts, device=device)


def parity_check(world, device):
    """The runner class that is about to be timed, on the benchmark's circuit generator at n = 10, against the CPU
    oracle (checker only): max |delta| over the 2^n probabilities and the 4^n coefficients, and |trace - 1|."""
    from oracle import dm_oracle
    from qiskit_aakash_b200 import circuits
    n = PARITY_N
    opts = circuits.noisy_options()
    runner = make_runner(world, n, circuits.random_layered(n, PARITY_DEPTH, PARITY_SEED, readout=False), opts, device)
    runner.step()
    probs = np.array(runner.probs)
    vec = runner.engine.download()
    ref = dm_oracle.run_oracle(n, copy.deepcopy(circuits.random_layered(n, PARITY_DEPTH, PARITY_SEED).instructions),
                               dict(copy.deepcopy(opts), compute_densitymatrix=False, chop_threshold=0.0))
    ref_p = np.array(list(ref["data"]["ensemble_probability"].values()))
    d_prob = float(np.max(np.abs(probs - ref_p)))
    d_coeff = float(np.max(np.abs(vec - ref["data"]["coeffmatrix"])))
    out = {"parity_max_abs": max(d_prob, d_coeff), "prob": d_prob, "coeff": d_coeff,
           "trace": float(abs(vec[0] * 2 ** n - 1.0)),
           "circuit": "random layered U3+CX n=%d depth %d seed %d, same noise options, through %s"
                      % (n, PARITY_DEPTH, PARITY_SEED, type(runner).__name__)}
    runner.close()
    return out


def time_runner(runner, steps, warmup, world, sampler_index=None):
    """W warm-up steps, then K steps bracketed by barrier + synchronize, CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        runner.step()
    barrier()
    sampler = ClockSampler(sampler_index) if sampler_index is not None else None
    if sampler:
        sampler.start()
    runner.reset_counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        runner.step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop() if sampler else None
    return ms_total / steps, clocks


def side_workload(name, world, device, steps=2, warmup=1):
    """A second workload timed in the same invocation (scaling_curve / config4_qft16): same timing discipline, fewer
    steps."""
    circ, opts, desc = build_workload(name, readout=False)
    n = circ.n_qubits
    gates = count_gates(circ)
    runner = make_runner(world, n, circ, opts, device)
    ms, _ = time_runner(runner, steps, warmup, world)
    counters = runner.counters()
    out = {"workload": desc, "ms_per_step": ms, "value": gates * 16.0 * 4 ** n / (ms * 1e-3) / 1e9, "unit": UNIT,
           "gates": gates, "steps": steps, "warmup": warmup, "prob_sum": float(np.sum(runner.probs)),
           "passes_per_step": counters["tile_pass_launches"] / steps}
    ex = runner.exchange_stats()
    if ex:
        out["nvlink"] = ex
    runner.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from qiskit_aakash_b200 import BasicAer, assemble

    name = args.workload or DEFAULT_WORKLOAD.get(args.gpus, "config3")
    circ, opts, desc = build_workload(name, readout=False, qubits=args.qubits, layers=args.layers)
    if args.qubits:
        name = "custom"
    n = circ.n_qubits
    n_gates = count_gates(circ)
    state_bytes = 8 * 4 ** n

    parity = None if args.no_parity else parity_check(world, local_rank)

    runner = make_runner(world, n, circ, opts, local_rank)
    sched_desc = "dmb_schedule strategy %d (%s), <= %d ops per pass" % (
        runner.engine.strategy, "tile search" if runner.engine.strategy == 1 else "program order",
        runner.engine.max_ops_per_pass)
    ms_step, clocks = time_runner(runner, args.steps, args.warmup, world, local_rank if rank == 0 else None)
    counters = runner.counters()
    value = n_gates * 16.0 * 4 ** n / (ms_step * 1e-3) / 1e9
    prob_sum = float(np.sum(runner.probs))

    # dominant kernel: the tile pass; average launch duration measured live (events around the back-to-back
    # in-place pass launches of every step, accumulated by the runner)
    peak, peak_src = measured_peak()
    launches = max(1, counters["tile_pass_launches"])
    pass_ms = runner.pass_ms_total() / max(1, runner.timed_pass_launches())
    per_launch_bytes = 16.0 * 4 ** n / world
    achieved = per_launch_bytes / (pass_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_tile_pass6<5 CTAs/SM, paired>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "avg_launch_ms": pass_ms, "algorithmic_bytes_per_launch": per_launch_bytes,
                "fused_ops_per_launch": counters["fused_ops"] / launches,
                "swaps_folded_into_store_per_launch": counters.get("folded_swaps", 0) / launches,
                "ops_chained_to_their_predecessor_per_launch": counters.get("chained_ops", 0) / launches}
    if world == 1 and n <= 14:
        roofline.update(unfused_launch_points(runner.engine, n, peak))
        roofline["note"] = ("launches that fuse more than ~3 ops are shared-memory-bandwidth bound (one 64 KiB round "
                            "trip per tile and op), so frac against HBM falls as fusion grows while circuit time "
                            "improves; see staging_only / one_gate_per_launch for the HBM-bound operating points")
    # second roofline of the same launches: shared-memory traffic of the op phase.  Every op executed in shared
    # memory reads and writes the whole tile (2 x 8 B per coefficient), staging adds one write (cp.async) and one
    # read (write-back); peak = 128 B/clk/SM x SMs x the SM clock sampled during the run.
    try:
        smem_ops = (counters["fused_ops"] - counters.get("folded_swaps", 0) - counters.get("chained_ops", 0)) / launches
        smem_bytes = (smem_ops + 1.0) * 16.0 * 4 ** n / world
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
        if mhz:
            smem_peak = 128.0 * sm_count * mhz * 1e6 / 1e9
            smem_ach = smem_bytes / (pass_ms * 1e-3) / 1e9
            roofline["shared_memory"] = {"bound": "smem", "achieved": smem_ach, "peak": smem_peak, "unit": "GB/s",
                                         "frac": smem_ach / smem_peak, "ops_in_smem_per_launch": smem_ops,
                                         "peak_source": "128 B/clk/SM x %d SMs x %.0f MHz (sampled)" % (sm_count, mhz)}
    except Exception:
        pass
    # DRAM bytes per launch from the latest committed `ncu --set full` capture of this kernel
    for fname in ("r02f_tile_pass_ncu_summary.json", "r02_tile_pass_ncu_summary.json", "r01b_tile_pass_ncu_summary.json"):
        prof = os.path.join(ROOT, "profiles", fname)
        if os.path.exists(prof) and world == 1 and n == 14:
            try:
                roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
                roofline["traffic_source"] = "profiles/" + fname
                break
            except Exception:
                pass
    nv = runner.exchange_stats()
    passes_per_step = counters["tile_pass_launches"] / args.steps
    gpu_launches = int(counters["tile_pass_launches"] + counters["other_launches"])
    levels = runner.n_levels
    runner.close()
    del runner

    # end to end through the public API (host buffers; host lowering + D2H of the results inside)
    e2e = None
    if not args.no_e2e:
        e2e = time_e2e(args, world, local_rank, lambda: build_workload(name, qubits=args.qubits, layers=args.layers)[0],
                       opts, n, n_gates, state_bytes)

    scaling_curve = config4 = None
    if not args.no_side:
        if name == "layered16":
            scaling_curve = {"workload": desc, "ms_per_step": ms_step, "value": value, "unit": UNIT, "gates": n_gates,
                             "steps": args.steps, "warmup": args.warmup, "note": "this line's headline workload"}
        else:
            scaling_curve = side_workload("layered16", world, local_rank)
        if name != "qft16":
            config4 = side_workload("qft16", world, local_rank)

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        g.stage_reference()
        cpu = time_cpu_circuit(size_probe=(n, n_gates, levels))

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if name == "layered16" else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "workload_name": name, "gates": n_gates, "levels": levels,
                           "state_bytes": state_bytes, "passes_per_step": passes_per_step, "scheduler": sched_desc,
                           "l2": "state (%.1f GiB/GPU) >> 126 MB L2, no flush needed" % (state_bytes / world / 2 ** 30),
                           "parallelism": "1 GPU" if world == 1 else "%d GPUs, high-order Pauli digits sharded" % world,
                           "scaling_note": "the headline workload differs between GPU counts (n=14 / 16 / 16 / 18: what "
                                           "BASELINE.json names for each); `scaling_curve` is ONE circuit (layered16) "
                                           "at every GPU count"},
                "circuit_ms": ms_step, "prob_sum": prob_sum, "parity_max_abs": parity and parity["parity_max_abs"],
                "parity": parity, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "scaling_curve": scaling_curve, "config4_qft16": config4,
                "gpu_launches": gpu_launches, "clocks": clocks}
        if nv:
            line["nvlink"] = nv
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def time_e2e(args, world, device, circ_fn, opts, n, n_gates, state_bytes):
    """``BasicAer.get_backend('dm_simulator').run(qobj, backend_options).result()`` -- the call a user makes.  On one
    GPU the result carries the 4^n-coefficient vector (D2H into pinned memory inside the timed region) when it is
    at most 2 GiB; a sharded state (N > 1) and n = 16 return the 2^n probabilities only (the backend does not
    gather 4^n doubles to every host above 14 qubits unless asked to) -- so the N = 1 and N > 1 figures are NOT
    comparable with each other."""
    import torch
    import torch.distributed as dist
    from qiskit_aakash_b200 import BasicAer, assemble
    backend = BasicAer.get_backend("dm_simulator")          # sharded automatically inside a torch.distributed group
    full_state = world == 1 and n <= 14
    backend.SHOW_FINAL_STATE = full_state
    run_opts = dict(opts, compute_densitymatrix=False)
    qobj = assemble(circ_fn())                               # the backend never mutates it
    reps = max(1, min(args.steps, 3 if n <= 14 else 2))
    times = []
    for it in range(reps + 1):                               # first call = warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = backend.run(qobj, backend_options=copy.deepcopy(run_opts)).result()
        assert res["success"]
        probs = res["results"][0]["data"]["ensemble_probability"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it:
            times.append(dt)
        h2d = backend.last_engine_stats["h2d_bytes"]
        breakdown = {k: round(1e3 * v, 2) for k, v in backend.last_engine_stats.items() if k.startswith("t_")}
        del res
        backend._engine = None
    backend.release_engines()            # the sharded backend keeps shards + peer mappings for the next job
    dt = sum(times) / len(times)
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        dist.barrier()                   # every rank has unmapped its peers' shards before anyone frees them
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    return {"value": n_gates * 16.0 * 4 ** n / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int((state_bytes if full_state else 0) + 8 * 2 ** n),
            "prob_sum": float(sum(probs.values())), "breakdown_ms": breakdown,
            "returns": "coeffmatrix (4^n) + 2^n probabilities" if full_state else
                       "2^n probabilities only (no 4^n gather: not comparable with the 1-GPU n=14 figure)"}


def unfused_launch_points(e, n, peak):
    """The same kernel at its HBM-bound operating points, measured live: a launch that only stages
    the tiles (0 ops) and a launch that applies ONE gate (TSP CNOT + two single-qubit maps) -- the
    reference's own granularity of one sweep per gate."""
    import torch
    from qiskit_aakash_b200 import capi, engine as eng, schedule
    rng = np.random.default_rng(0)
    rot = lambda: eng.gate_matrix("u3", rng.uniform(0, 6, 3), {"rz": [0.999, 0.0], "ry": [0.999, 0.0]})
    tile = list(range(6))
    out = {}
    for name, ops in (("staging_only", []),
                      ("one_gate_per_launch", [schedule.DevOp(capi.OP_CX_TSP, 2, 3, rot(), rot(),
                                                              eng.cx_coefficients([0.999, 0.0]))])):
        passes = schedule.encode_passes([(tile, ops)])
        for _ in range(3):
            e.ctx.apply_passes(e.sptr, e.n_bits, passes)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record()
        for _ in range(reps):
            e.ctx.apply_passes(e.sptr, e.n_bits, passes)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        ach = 16.0 * 4 ** n / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "achieved": ach, "frac": ach / peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="", choices=[""] + sorted(WORKLOADS),
                    help="default: config3 at 1 GPU, layered16 at 2 / 4, config5 (n=18) at 8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--qubits", type=int, default=0, help="experiments: layered circuit at another size (--layers)")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip scaling_curve / config4_qft16")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
