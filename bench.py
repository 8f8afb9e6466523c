#!/usr/bin/env python
"""Benchmark of the MOC transport-sweep hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5g7-2d|simple-lattice|pin-cell]
  python bench.py --impl reference ...      # the reference's own CPUSolver on the host cores

metric   segment-group integrations / s (the reference's own count W = 2*F*N_seg per
         sweep, src/Solver.cpp:1901-1902), whole job over all N GPUs.
step     one full source iteration of Solver::computeEigenvalue (sources -> transport
         sweep -> closure -> k_eff -> normalisation -> residual -> store) on synthetic
         tracks of the named deck (openmoc_b200.synth), everything resident in HBM.
e2e      the same iteration driven through the public host API with HOST buffers, the
         way openmoc.krylov drives a solver (setFluxes(numpy) -> sweep -> getFluxes(),
         openmoc/krylov.py:160-240): H2D of the scalar flux from pinned memory and D2H
         of the new flux + k_eff every step, wall-clock timed.
One JSON line on stdout (rank 0).  Nothing here reads /root/reference at run time.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synth model, num_azim, spacing cm, num_polar, BASELINE.json config it is)
    "c5g7-2d": ("c5g7-2d", 128, 0.01, 6, "configs[2]: 2D C5G7 quarter core, 128 azim, 0.01 cm"),
    "simple-lattice": ("simple-lattice", 128, 0.01, 6, "configs[1]: simple-lattice 2D, 128 azim, 0.01 cm"),
    "pin-cell": ("pin-cell", 128, 0.01, 6, "configs[0]: pin-cell 2D, 128 azim, 0.01 cm"),
}
# bounded CPU sample of the same deck for the reference arm (same per-integration work,
# coarser track laydown so that ~10-30 s of host time suffice)
CPU_SAMPLE = {"c5g7-2d": (32, 0.05), "simple-lattice": (64, 0.02), "pin-cell": (128, 0.01)}
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append((time.time(), l)) for l in self.proc.stdout],
                                      daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ts, l in self.lines:
            # keep the samples taken while the timed region ran
            if self.t0 is not None and not (self.t0 <= ts <= self.t1 + 0.12):
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------- reference arm
def run_reference_cpu(workload, iters, threads):
    """Times the UNMODIFIED reference CPUSolver (oracle/_ref/ref_driver, built from
    /root/reference by oracle/Makefile) on a bounded sample of the workload's deck;
    falls back to the plain-C oracle port when the reference build is absent."""
    model, _, _, num_polar, _ = WORKLOADS[workload]
    azim, spacing = CPU_SAMPLE[workload]
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    sample = f"{model} deck, {azim} azim, {spacing} cm, {iters} source iterations"
    if os.path.exists(driver):
        with tempfile.TemporaryDirectory() as td:
            js = os.path.join(td, "ref.json")
            env = dict(os.environ, OMP_NUM_THREADS=str(threads))
            t0 = time.perf_counter()
            subprocess.run([driver, "--model", model, "--azim", str(azim), "--spacing", str(spacing),
                            "--polar", str(num_polar), "--max-iters", str(iters), "--threads", str(threads),
                            "--quiet", "--no-fluxes", "--json", js], check=True, env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=td)
            wall = time.perf_counter() - t0
            d = json.load(open(js))
        return {"kind": "reference", "value": d["integrations"] / d["sweep_time_s"], "cores": threads,
                "sample": sample + f" (N_seg={d['n_segments']}, sweep timer {d['sweep_time_s']:.2f} s, "
                                   f"whole run incl. ray tracing {wall:.1f} s)",
                "sweep_time_s": d["sweep_time_s"], "iterations": d["iterations"],
                "integrations": d["integrations"], "total_time_s": d["total_time_s"]}
    from openmoc_b200.synth import make_tracks
    from oracle.oracle_py import OracleSolver
    ft = make_tracks(model, num_azim=azim, spacing=spacing, num_polar=num_polar)
    s = OracleSolver(ft)
    s.setNumThreads(threads)
    t0 = time.perf_counter()
    s.computeEigenvalue(iters, 1e-30)
    wall = time.perf_counter() - t0
    W = 2.0 * ft.fluxes_per_track * ft.n_segments * iters
    return {"kind": "port", "value": W / s.sweepSeconds(), "cores": threads,
            "sample": sample + f" (N_seg={ft.n_segments}, oracle port, sweep {s.sweepSeconds():.2f} s)",
            "sweep_time_s": s.sweepSeconds(), "iterations": iters, "integrations": W, "total_time_s": wall}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    iters = args.steps + args.warmup
    cb = run_reference_cpu(args.workload, iters, threads)
    model, azim, spacing, num_polar, cfgname = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "segment-group integrations/s", "value": cb["value"],
        "unit": "integrations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cb["total_time_s"] / max(cb["iterations"], 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "ns_per_integration_thread": 1e9 * cb["sweep_time_s"] * cb["cores"] / cb["integrations"],
        "config": {"workload": args.workload, "deck": cfgname, "solver": "CPUSolver (OpenMP)",
                   "timed": "reference 'Transport Sweep' timer split over all iterations"},
        "cpu_baseline": {k: cb[k] for k in ("value", "cores", "kind", "sample")} | {"unit": "integrations/s"},
        "e2e": {"value": cb["value"], "unit": "integrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5g7-2d", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="double", choices=["double", "mixed"])
    ap.add_argument("--azim", type=int, default=None)
    ap.add_argument("--spacing", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--partition", default="pair", choices=["pair", "chain", "track"])
    ap.add_argument("--deterministic", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    from openmoc_b200 import capi
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    model, azim, spacing, num_polar, cfgname = WORKLOADS[args.workload]
    azim = args.azim or azim
    spacing = args.spacing or spacing
    ncpu = os.cpu_count() or 1
    t0 = time.perf_counter()
    ft = make_tracks(model, num_azim=azim, spacing=spacing, num_polar=num_polar,
                     num_threads=max(1, ncpu // world))
    t_gen = time.perf_counter() - t0
    F, G = ft.fluxes_per_track, ft.num_groups
    W_sweep = 2.0 * F * ft.n_segments                     # whole job, all ranks
    precision = capi.PRECISION_MIXED if args.precision == "mixed" else capi.PRECISION_DOUBLE

    t0 = time.perf_counter()
    solver = B200Solver(ft, device=local_rank, precision=precision,
                        process_group=(dist.group.WORLD if world > 1 else None),
                        partition=args.partition, deterministic=args.deterministic)
    solver.useTorchStream()
    t_setup = time.perf_counter() - t0
    local_W = 2.0 * F * solver.tracks.n_segments

    # initial state of Solver::computeEigenvalue: psi = 0, phi = 1 normalised, stored
    solver.zeroTrackFluxes()
    solver.flattenFSRFluxes(0.0); solver.storeFSRFluxes()
    solver.flattenFSRFluxes(1.0); solver.normalizeFluxes(); solver.storeFSRFluxes()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    solver.iterate(args.warmup)
    barrier()
    solver.resetSweepStats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.t0 = time.time()
    ev0.record()
    solver.iterate(args.steps)
    ev1.record()
    torch.cuda.synchronize()
    sampler.t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    sweep_ms, n_sweeps, launches = solver.getSweepStats()
    k_dev = solver.getKeff()
    value = W_sweep * args.steps / (ms * 1e-3)

    # ---------------- end to end through the host API ----------------
    n_phi = ft.n_fsrs * G
    host_phi = torch.empty(n_phi, dtype=torch.float64).pin_memory()
    host_np = host_phi.numpy()
    host_np[:] = solver.getFluxes()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(i):
        solver.setFluxes(host_np)                 # H2D, pinned
        solver.computeFSRSources(1000 + i)
        solver.transportSweep()
        solver.addSourceToScalarFlux()
        k = solver.computeKeff()                   # D2H scalar
        solver.normalizeFluxes(fetch=False)
        host_np[:] = solver.getFluxes()            # D2H
        return k
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = W_sweep * e2e_steps / t_e2e

    # ---------------- roofline of the dominant kernel (the sweep) ----------------
    peak, peak_src = measured_peak()
    b_alg = (12.0 * ft.n_segments + 16.0 * F * ft.n_tracks + 16.0 * G * ft.n_fsrs) / W_sweep
    sweep_avg_ms = sweep_ms / max(n_sweeps, 1)
    achieved = b_alg * local_W / (sweep_avg_ms * 1e-3) / 1e9       # GB/s of algorithmic bytes, this rank
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cb = run_reference_cpu(args.workload, 6, ncpu)
            cpu_baseline = {"value": cb["value"], "unit": "integrations/s", "cores": cb["cores"],
                            "kind": cb["kind"], "sample": cb["sample"]}
        line = {
            "metric": "segment-group integrations/s", "value": value, "unit": "integrations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if precision == capi.PRECISION_DOUBLE else "f32 segment math, f64 tally",
            "data": "synthetic",
            "ns_per_integration": 1e9 / value,
            "config": {"workload": args.workload, "deck": cfgname, "num_azim": azim, "spacing_cm": spacing,
                       "num_polar": num_polar, "n_tracks": ft.n_tracks, "n_segments": ft.n_segments,
                       "n_fsrs": ft.n_fsrs, "groups": G, "fluxes_per_track": F,
                       "integrations_per_sweep": W_sweep,
                       "parallelism": "1 GPU" if world == 1 else f"{args.partition} partition x{world} + NCCL all-reduce of the FSR tally"
                                      + (" + NCCL send/recv of cross-rank boundary fluxes" if args.partition == "track" else ""),
                       "deterministic_tally": bool(args.deterministic),
                       "l2": "segment stream (%.2f GB) is larger than the 126 MB L2; no flush" % (12.0 * ft.n_segments / 1e9)
                             if 12.0 * ft.n_segments > 2.5e8 else "inputs fit in L2; not flushed (launch-bound shape)",
                       "k_eff_after_timed_steps": k_dev, "track_generation_s": round(t_gen, 2),
                       "upload_and_setup_s": round(t_setup, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "b200::sweep_kernel", "bytes_per_integration": b_alg,
                         "kernel_ms": sweep_avg_ms, "kernel_share_of_step": sweep_ms / ms if ms > 0 else None,
                         "note": "FP64-issue bound, not HBM bound, for G*P/2=21 (SURVEY 8d): 22 FP64 instr per integration vs "
                                 "0.34 algorithmic bytes; see DESIGN.md 4.1 and profiles/"},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "integrations/s", "h2d_bytes_per_step": n_phi * 8,
                    "d2h_bytes_per_step": n_phi * 8 + 8, "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
