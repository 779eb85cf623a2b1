#!/usr/bin/env python
"""bench.py -- headline benchmark of the dm_simulator hot path on B200.

Metric (BASELINE.json): effective HBM GB/s per gate (and circuit wall time) on the
n=14..18 noisy random U3+CX circuits.  One "step" = one complete execution of the circuit
(state init, every gate/noise level, I/Z-marginal probability readout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N=1 workload: BASELINE.json configs[2] -- random layered U3+CX, n=14, depth 200, per-gate
noise (rotation_error / tsp_model_error r=0.999, depolarization 0.99), ensemble-Z readout;
the 2 GiB state is 17x the 126 MB L2, so no L2 flush between steps is needed.
N>1: the same circuit family sharded over the high-order Pauli digits.

Prints ONE JSON line (rank 0).  `value` times the device path with the plan and the state
resident in HBM (CUDA events on the launching stream); `e2e` times the public API call
``backend.run(qobj, backend_options).result()`` with host buffers: host-side merge /
partition / scheduling, kernel launches, and the device->host copy of the 4^n coefficient
vector + 2^n probabilities into pinned memory are all inside its timed region.
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


def workload(n_gpus):
    """(n_qubits, depth, seed) per GPU count; per-GPU state 2-4 GiB."""
    return {1: (14, 200, 1400), 2: (15, 100, 1500), 4: (15, 100, 1500), 8: (16, 60, 1600)}[n_gpus]


def workload_config(n, depth, n_gates=None, extra=None):
    cfg = {"workload": "random layered U3+CX n=%d depth %d, rotation/tsp error r=0.999, depolarization 0.99, "
                       "ensemble-Z readout (BASELINE configs[2] shape)" % (n, depth)}
    if n_gates is not None:
        cfg["gates"] = n_gates
    cfg.update(extra or {})
    return cfg


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
# CPU baseline / reference arm: the oracle port of the reference's algorithm on host cores
# --------------------------------------------------------------------------------------------

def cpu_sample_instructions(n, seed):
    """Bounded sample of the workload at full size: the first level of the circuit restricted to its two
    outermost qubits (u3 on qubit 0 and on qubit n-1) followed by that level's per-qubit memory-noise sweep
    -- i.e. 2 gates + 1 level of the reference's run_experiment loop."""
    from qiskit_aakash_b200 import circuits
    full = circuits.random_layered(n, 1, seed, readout=False)
    u3s = [i for i in full.instructions if i.name == "u3"]
    return [u3s[0], u3s[n - 1]]


N_CALIBRATION = 11      # register size at which one full layer (u3 level + CX level) of the workload is timed


def _time_oracle(n, instrs, levels):
    from oracle import dm_oracle
    from qiskit_aakash_b200 import circuits
    opts = dict(circuits.noisy_options(), compute_densitymatrix=False)
    t0 = time.perf_counter()
    res = dm_oracle.run_oracle(n, copy.deepcopy(instrs), opts)
    dt = time.perf_counter() - t0
    assert res["number_of_clock_cycles"] == levels
    return dt


def time_cpu_sample(n, seed, depth=None):
    """CPU (oracle port of the reference) figure for the workload's metric, from a bounded sample.

    The full circuit cannot be run on the CPU at n = 14 (hours), and its cost is not proportional to the gate
    count: the reference sweeps the memory-noise channel over all n qubits after EVERY level
    (dm_simulator.py:1173-1177), which is most of its time.  So
      1. at the full size n: 2 gates + 1 noise level are timed (t_big) -- pins the absolute speed at the real
         state size (cache / memory-bandwidth regime);
      2. at n = 11: the same sample (t_small) and ONE full layer of the workload -- u3 on every qubit, the CX
         brick, two noise levels -- are timed (t_layer): pins the workload's true mix of gates and noise sweeps;
      3. circuit time at n is extrapolated as depth * t_layer * (t_big / t_small) and the metric is
         gates * 16 * 4^n / that, the same definition the GPU arm uses.
    Returns (GB/s per gate, seconds spent, description)."""
    from qiskit_aakash_b200 import circuits
    if depth is None:
        depth = 200
    t_big = _time_oracle(n, cpu_sample_instructions(n, seed), 1)
    ns = min(n, N_CALIBRATION)
    t_small = min(_time_oracle(ns, cpu_sample_instructions(ns, seed), 1) for _ in range(3))
    layer = [i for i in circuits.random_layered(ns, 1, seed, readout=False).instructions]
    t_layer = min(_time_oracle(ns, layer, 2) for _ in range(2))
    full = circuits.random_layered(n, depth, seed, readout=False)
    gates = sum(1 for i in full.instructions if i.name in ("u3", "cx"))
    t_circuit = depth * t_layer * (t_big / t_small)
    value = gates * 16.0 * 4 ** n / t_circuit / 1e9
    spent = t_big + 3 * t_small + 2 * t_layer
    desc = ("n=%d: u3 on qubits 0 and %d + one memory-noise level timed at full size (%.1f s); one full layer of the "
            "workload (u3 on all qubits, CX brick, 2 noise levels) timed at n=%d (%.3f s) for the gate / noise-sweep mix; "
            "circuit time extrapolated to %.0f s for %d gates over %d levels.  The oracle port skips the reference's "
            "full-state copies, so it is faster than the reference itself"
            % (n, n - 1, t_big, ns, t_layer, t_circuit, gates, 2 * depth))
    return value, spent, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, depth, seed = workload(args.gpus)
    n_cpu = min(n, 14)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        time_cpu_sample(n_cpu, seed, depth)
    vals, times, desc = [], [], ""
    for _ in range(args.steps):
        v, dt, desc = time_cpu_sample(n_cpu, seed, depth)
        vals.append(v)
        times.append(dt)
    value = len(vals) / sum(1.0 / v for v in vals)           # steps are equal work: harmonic mean of the rates
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(n, depth, extra={"cpu_sample": desc}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

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
    from qiskit_aakash_b200 import BasicAer, DmSimulatorB200, assemble, circuits, engine, hostpass

    n, depth, seed = workload(args.gpus)
    if args.qubits:
        n, depth, seed = args.qubits, (args.layers or depth), 100 * args.qubits
    circ = circuits.random_layered(n, depth, seed)
    opts = circuits.noisy_options()
    n_gates = sum(1 for i in circ.instructions if i.name in ("u3", "cx"))
    state_bytes = 8 * 4 ** n

    if world > 1:
        from qiskit_aakash_b200 import distributed
        runner = distributed.ShardedCircuitRunner(n, circ, opts, device=local_rank)
    else:
        runner = SingleGpuRunner(n, circ, opts, device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sched_desc = "dmb_schedule strategy %d (%s), <= %d ops per pass" % (
        runner.engine.strategy, "tile search" if runner.engine.strategy == 1 else "program order",
        runner.engine.max_ops_per_pass)
    for _ in range(args.warmup):
        runner.step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    runner.reset_counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        runner.step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop() if sampler else None
    counters = runner.counters()
    ms_step = ms_total / args.steps
    value = n_gates * 16.0 * 4 ** n / (ms_step * 1e-3) / 1e9

    # dominant kernel: tile pass; average launch duration measured live (events around the
    # back-to-back pass launches of every step, accumulated by the runner)
    peak, peak_src = measured_peak()
    timed_launches = runner.timed_pass_launches() if hasattr(runner, "timed_pass_launches") \
        else counters["tile_pass_launches"]
    pass_ms = runner.pass_ms_total() / max(1, timed_launches)
    per_launch_bytes = 16.0 * 4 ** n / world
    achieved = per_launch_bytes / (pass_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_tile_pass<6>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "avg_launch_ms": pass_ms, "algorithmic_bytes_per_launch": per_launch_bytes,
                "fused_ops_per_launch": counters["fused_ops"] / max(1, counters["tile_pass_launches"]),
                "swaps_folded_into_store_per_launch": counters.get("folded_swaps", 0) / max(1, counters["tile_pass_launches"])}
    if world == 1:
        roofline.update(unfused_launch_points(runner.engine, n, peak))
        roofline["note"] = ("launches that fuse more than ~3 ops are shared-memory-bandwidth bound (one 64 KiB round "
                            "trip per tile and op), so frac against HBM falls as fusion grows while circuit time "
                            "improves; see staging_only / one_gate_per_launch for the HBM-bound operating points")
    # second roofline of the same launches: shared-memory traffic of the op phase.  Every op executed in
    # shared memory reads and writes the whole tile (2 x 8 B per coefficient), staging adds one write
    # (cp.async) and one read (write-back); peak = 128 B/clk/SM x SMs x the SM clock sampled during the run.
    try:
        launches = max(1, counters["tile_pass_launches"])
        smem_ops = (counters["fused_ops"] - counters.get("folded_swaps", 0)) / launches
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
    for name in ("r01b_tile_pass_ncu_summary.json", "r01_tile_pass_ncu_summary.json"):
        prof = os.path.join(ROOT, "profiles", name)
        if os.path.exists(prof) and world == 1:
            try:
                roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
                roofline["traffic_source"] = "profiles/" + name
                break
            except Exception:
                pass

    # end to end through the public API (host buffers; D2H of the coefficient vector inside)
    e2e = None
    if world == 1 and args.no_e2e:
        pass
    elif world == 1:
        backend = BasicAer.get_backend("dm_simulator")
        run_opts = dict(opts, compute_densitymatrix=False)
        qobj = assemble(circuits.random_layered(n, depth, seed))    # the backend never mutates it
        qobj_fn = lambda: qobj
        for _ in range(2):
            res = backend.run(qobj_fn(), backend_options=copy.deepcopy(run_opts)).result()
            del res
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 3))
        for _ in range(reps):
            res = backend.run(qobj_fn(), backend_options=copy.deepcopy(run_opts)).result()
            assert res["success"]
            probs = res["results"][0]["data"]["ensemble_probability"]
            h2d = backend.last_engine_stats["h2d_bytes"]
            breakdown = {k: round(1e3 * v, 2) for k, v in backend.last_engine_stats.items() if k.startswith("t_")}
            del res
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        e2e = {"value": n_gates * 16.0 * 4 ** n / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(state_bytes + 8 * 2 ** n),
               "prob_sum": float(sum(probs.values())), "breakdown_ms": breakdown}
    else:
        nv = {"exchanges_per_step": runner.engine.exchanges / args.steps,
              "nvlink_bytes_sent_per_gpu_per_step": runner.engine.nvlink_bytes_sent / args.steps}
        e2e = None if args.no_e2e else runner.e2e(args, lambda: circuits.random_layered(n, depth, seed), opts,
                                                 n_gates, device=local_rank)

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        v, dt, desc = time_cpu_sample(min(n, 14), seed, depth)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(n, depth, n_gates, {
                           "levels": runner.n_levels, "state_bytes": state_bytes,
                           "passes_per_step": counters["tile_pass_launches"] / args.steps,
                           "scheduler": sched_desc,
                           "tile_variant": int(os.environ.get("DMB_TILE_VARIANT", "0") or 0),
                           "l2": "state (%.1f GiB/GPU) >> 126 MB L2, no flush needed" % (state_bytes / world / 2 ** 30),
                           "parallelism": "1 GPU" if world == 1 else "%d GPUs, high-order Pauli digits sharded" % world}),
                "circuit_ms": ms_step, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(counters["tile_pass_launches"] + counters["other_launches"]),
                "clocks": clocks}
        if world > 1:
            line["nvlink"] = nv
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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


class SingleGpuRunner:
    """Device path with the compiled plan resident: init -> passes -> marginal -> FWHT."""

    def __init__(self, n, circ, opts, device=0):
        import torch
        from qiskit_aakash_b200 import DmSimulatorB200, engine as eng, hostpass
        self.torch = torch
        self.n = n
        self.engine = eng.PauliEngine(n, device=device)
        self.engine.drain_threshold = 0          # compile the whole circuit into one resident plan
        be = DmSimulatorB200(device=device)
        be._set_options(None, copy.deepcopy(opts))
        be._initialize_errors()
        ops = hostpass.merge_single_qubit_gates(circ.instructions, n, True)
        levels, self.n_levels = hostpass.partition_levels(ops, n)
        e = self.engine
        noise = eng.memory_noise_matrix(1., 1., 1.)
        for level in levels[:self.n_levels]:
            for op in level:
                if op.name in ("u1", "u3"):
                    e.apply_1q(op.qubits[0], eng.gate_matrix(op.name, op.params, be._error_params["one_qubit_gates"]))
                elif op.name == "cx":
                    e.apply_cx(op.qubits[0], op.qubits[1], be._error_params["two_qubit_gates"])
        self.passes = e.plan()
        self.final_pos = list(e.pos)
        e.queue, e.pending = [], [None] * n
        self.err = be._error_params["measurement"]
        self._pass_ms = 0.0
        self._ev = []

    def step(self):
        torch, e = self.torch, self.engine
        e.init_product([[1, 0, 0, 1]] * self.n, 0.5 ** self.n)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e.run_passes(self.passes)
        b.record()
        self._ev.append((a, b))
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
        return sum(a.elapsed_time(b) for a, b in self._ev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--qubits", type=int, default=0, help="override the qubit count (experiments: --qubits 18 --layers 20)")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
