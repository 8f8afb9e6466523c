#!/usr/bin/env python
"""Benchmark of the MOC transport-sweep hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--also NAME,NAME|none]
  python bench.py --impl reference ...      # the reference's own CPUSolver on the host cores
  python bench.py --impl refgpu ...         # the reference's own GPUSolver recompiled for sm_100a

metric   segment-group integrations / s (the reference's own count W = 2*F*N_seg per
         sweep, src/Solver.cpp:1901-1902), whole job over all N GPUs.
step     one full source iteration of Solver::computeEigenvalue (sources -> transport
         sweep -> closure -> k_eff -> normalisation -> residual -> store) on synthetic
         tracks of the named deck (openmoc_b200.synth), everything resident in HBM.
e2e      the same iteration driven through the public host API with HOST buffers, the
         way openmoc.krylov drives a solver (setFluxes(numpy) -> sweep -> getFluxes(),
         openmoc/krylov.py:160-240): H2D of the scalar flux from pinned memory and D2H
         of the new flux + k_eff every step, wall-clock timed.
workloads  the primary workload (default c5g7-2d = configs[2], the deck the north-star target is
         quoted on) fills the top-level keys; every workload named by --also (default: c5g7-3d =
         configs[4]'s deck at the reference's own 3D parameters) is measured the same way in the
         same run and reported under "workloads" - so a 1/2/4/8-GPU series of this command also
         is the 3D strong-scaling series.
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
    # 2D decks: TY polar quadrature (what the reference's decks really run)
    "c5g7-2d": dict(dims=2, model="c5g7-2d", azim=128, spacing=0.01, polar=6,
                    desc="configs[2]: 2D C5G7 quarter core, 128 azim, 0.01 cm",
                    cpu_sample=dict(azim=128, spacing=0.05),
                    # CMFD as in sample-input/benchmarks/c5g7/c5g7-2d.py:51-56 (51 x 51, two groups, SOR 1.5)
                    cmfd=dict(num_z=1, sample=dict(azim=32, spacing=0.1))),
    "simple-lattice": dict(dims=2, model="simple-lattice", azim=128, spacing=0.01, polar=6,
                           desc="configs[1]: simple-lattice 2D, 128 azim, 0.01 cm",
                           cpu_sample=dict(azim=128, spacing=0.01)),
    "pin-cell": dict(dims=2, model="pin-cell", azim=128, spacing=0.01, polar=6,
                     desc="configs[0]: pin-cell 2D, 128 azim, 0.01 cm", cpu_sample=dict(azim=128, spacing=0.01)),
    # configs[4]: the parameters of profile/models/c5g7/c5g7-3d-cmfd.cpp:15-27,537-573 (16 azim, 0.1 cm,
    # 8 polar equal-angle, 1.0 cm axial spacing, 9 x 15 = 135 axial layers of 0.476 cm, OTF tracks)
    "c5g7-3d": dict(dims=3, model="c5g7-2d", azim=16, spacing=0.1, polar=8, zspacing=1.0, n_axial=135,
                    quad="equal-angle",
                    desc="configs[4]: 3D C5G7 extruded core, 16 azim, 0.1 cm, 8 polar (equal angle), 1.0 cm axial "
                         "spacing, 135 axial layers, on-the-fly axial ray tracing on the device",
                    cpu_sample=dict(azim=4, spacing=0.5, polar=4, zspacing=4.0, n_axial=9),
                    # CMFD as in profile/models/c5g7/c5g7-3d-cmfd.cpp:537-556 (51 x 51 x 9*axial_refines = 135)
                    cmfd=dict(num_z=135, sample=dict(azim=4, spacing=0.8, polar=4, zspacing=6.0, n_axial=9, num_z=9))),
}
CMFD_GROUPS = [[1, 2, 3], [4, 5, 6, 7]]
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md
# FP64-pipe instructions per integration in the flat sweep kernels (cuobjdump of libb200moc.so,
# profiles/r02_sass.md): 11 Horner DFMA + 4 for the quotient + tau, L*q, delta-psi, psi update,
# tally (+ F2F conversions), per polar angle
FP64_PER_INTEGRATION = {2: 65.0 / 3.0, 3: 21.0}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.windows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append((time.time(), l)) for l in self.proc.stdout],
                                      daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ts, l in self.lines:
            # keep the samples taken while a timed region ran (nvidia-smi reports with some lag)
            if not any(t0 <= ts <= t1 + 0.1 for t0, t1 in self.windows):
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
def _driver_args(wl, sample):
    """ref_driver command line of the bounded sample of a workload's deck"""
    a = ["--model", wl["model"], "--azim", str(sample["azim"]), "--spacing", str(sample["spacing"]),
         "--polar", str(sample.get("polar", wl["polar"]))]
    if wl["dims"] == 3:
        a += ["--dims", "3", "--zspacing", str(sample["zspacing"]), "--axial", str(sample["n_axial"]),
              "--quad", wl.get("quad", "gl"), "--formation", "otf-stacks"]
    return a


def sample_text(wl, sample):
    s = f"{wl['model']} deck, {sample['azim']} azim, {sample['spacing']} cm"
    if wl["dims"] == 3:
        s += (f", {sample.get('polar', wl['polar'])} polar, {sample['zspacing']} cm axial spacing, "
              f"{sample['n_axial']} axial layers, OTF_STACKS")
    return s


def run_reference(workload, iters, threads, solver="cpu", keep_fluxes=False):
    """Times the UNMODIFIED reference (oracle/_ref/ref_driver, built from /root/reference by
    oracle/Makefile) - CPUSolver on the host cores, or its own GPUSolver recompiled for sm_100a -
    on a bounded sample of the workload's deck: same geometry, materials and angles per track,
    coarser track laydown.  Falls back to the plain-C oracle port when the reference build is absent."""
    wl = WORKLOADS[workload]
    sample = wl["cpu_sample"]
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    text = sample_text(wl, sample) + f", {iters} source iterations"
    if os.path.exists(driver):
        with tempfile.TemporaryDirectory() as td:
            js = os.path.join(td, "ref.json")
            env = dict(os.environ, OMP_NUM_THREADS=str(threads))
            cmd = [driver] + _driver_args(wl, sample) + ["--max-iters", str(iters), "--threads", str(threads),
                                                         "--tol", "1e-30", "--quiet", "--json", js, "--solver", solver]
            if not keep_fluxes:
                cmd.append("--no-fluxes")
            t0 = time.perf_counter()
            subprocess.run(cmd, check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=td)
            wall = time.perf_counter() - t0
            d = json.load(open(js))
        out = {"kind": "reference", "value": d["integrations"] / d["sweep_time_s"], "cores": threads,
               "sample": text + f" (N_seg={d['n_segments']}, sweep timer {d['sweep_time_s']:.2f} s, "
                                f"whole run incl. ray tracing {wall:.1f} s)",
               "sweep_time_s": d["sweep_time_s"], "iterations": d["iterations"],
               "integrations": d["integrations"], "total_time_s": d["total_time_s"], "keff": d.get("keff"),
               "n_segments": d["n_segments"], "n_fsrs": d.get("n_fsrs")}
        if keep_fluxes:
            out["fluxes"] = d.get("fluxes")
        return out
    if solver != "cpu":
        return None
    from oracle.oracle_py import OracleSolver
    ft = make_sample_tracks(workload)
    s = OracleSolver(ft)
    s.setNumThreads(threads)
    t0 = time.perf_counter()
    s.computeEigenvalue(iters, 1e-30)
    wall = time.perf_counter() - t0
    W = 2.0 * ft.fluxes_per_track * ft.n_segments * iters
    return {"kind": "port", "value": W / s.sweepSeconds(), "cores": threads,
            "sample": text + f" (N_seg={ft.n_segments}, oracle port, sweep {s.sweepSeconds():.2f} s)",
            "sweep_time_s": s.sweepSeconds(), "iterations": iters, "integrations": W, "total_time_s": wall,
            "keff": s.getKeff(), "n_segments": ft.n_segments, "n_fsrs": ft.n_fsrs,
            "fluxes": list(map(float, s.getFluxes())) if keep_fluxes else None}


def make_sample_tracks(workload, expand=True):
    from openmoc_b200.synth import make_tracks, make_tracks_3d
    wl = WORKLOADS[workload]
    s = wl["cpu_sample"]
    if wl["dims"] == 2:
        return make_tracks(wl["model"], num_azim=s["azim"], spacing=s["spacing"], num_polar=wl["polar"])
    return make_tracks_3d(wl["model"], num_azim=s["azim"], spacing=s["spacing"], num_polar=s.get("polar", wl["polar"]),
                          z_spacing=s["zspacing"], n_axial=s["n_axial"], polar_quad=_quad(wl), expand=expand)


def _quad(wl):
    from openmoc_b200 import synth
    return {"ty": synth.QUAD_TY, "equal-angle": synth.QUAD_EQUAL_ANGLE, "gl": synth.QUAD_GAUSS_LEGENDRE,
            "equal-weight": synth.QUAD_EQUAL_WEIGHT}[wl.get("quad", "gl")]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    iters = args.steps + args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == "refgpu":
        cb = run_reference(args.workload, iters, threads, solver="refgpu") if wl["dims"] == 2 else None
        if cb is None:
            print(json.dumps({"impl": "refgpu", "unavailable": "the reference GPUSolver handles 2D flat-source decks only "
                                                               "and needs oracle/_ref/libopenmoc_refgpu.so"}))
            return
        solver, timed = "GPUSolver (reference CUDA, recompiled for sm_100a)", "reference 'Transport Sweep' timer split"
    else:
        cb = run_reference(args.workload, iters, threads)
        solver, timed = "CPUSolver (OpenMP)", "reference 'Transport Sweep' timer split over all iterations"
    line = {
        "impl": args.impl, "metric": "segment-group integrations/s", "value": cb["value"],
        "unit": "integrations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cb["total_time_s"] / max(cb["iterations"], 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "ns_per_integration_thread": 1e9 * cb["sweep_time_s"] * cb["cores"] / cb["integrations"],
        "config": {"workload": args.workload, "deck": wl["desc"], "solver": solver, "timed": timed,
                   "sample": cb["sample"],
                   "same_deck_coarser_tracks": "the rate is per segment-group integration; the sample keeps the deck, "
                                               "materials, groups and polar angles per track and lays tracks coarser so "
                                               "that the reference's own ray tracing fits the time budget"},
        "cpu_baseline": {k: cb[k] for k in ("value", "cores", "kind", "sample")} | {"unit": "integrations/s"},
        "e2e": {"value": cb["value"], "unit": "integrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------ our arm
class Env:
    pass


def build_tracks(name, args, ncpu, world):
    from openmoc_b200.synth import make_tracks, make_tracks_3d
    wl = WORKLOADS[name]
    azim = args.azim if (args.azim and name == args.workload) else wl["azim"]
    spacing = args.spacing if (args.spacing and name == args.workload) else wl["spacing"]
    if wl["dims"] == 2:
        ft = make_tracks(wl["model"], num_azim=azim, spacing=spacing, num_polar=wl["polar"],
                         num_threads=max(1, ncpu // world))
    else:
        ft = make_tracks_3d(wl["model"], num_azim=azim, spacing=spacing, num_polar=wl["polar"],
                            z_spacing=wl["zspacing"], n_axial=wl["n_axial"], polar_quad=_quad(wl),
                            num_threads=max(1, ncpu // world), expand=False)
    return ft, azim, spacing


def measure(name, args, env, primary):
    """One workload: device-resident timing, end-to-end timing, roofline.  Returns the result dict."""
    import numpy as np
    import torch
    from openmoc_b200 import capi
    from openmoc_b200.solver import B200Solver
    dist, rank, world, local_rank = env.dist, env.rank, env.world, env.local_rank
    wl = WORKLOADS[name]
    ncpu = os.cpu_count() or 1
    t0 = time.perf_counter()
    ft, azim, spacing = build_tracks(name, args, ncpu, world)
    t_gen = time.perf_counter() - t0
    F, G = ft.fluxes_per_track, ft.num_groups
    precision = {"mixed": capi.PRECISION_MIXED, "table": capi.PRECISION_TABLE}.get(args.precision, capi.PRECISION_DOUBLE)
    partition = args.partition if wl["dims"] == 2 else args.partition_3d

    t0 = time.perf_counter()
    solver = B200Solver(ft, device=local_rank, precision=precision,
                        process_group=(dist.group.WORLD if world > 1 else None),
                        partition=partition, deterministic=args.deterministic,
                        balance_domains=args.balance_domains)
    solver.useTorchStream()
    t_setup = time.perf_counter() - t0
    local_seg = solver.num_segments
    n_seg = local_seg
    if dist is not None:
        t = torch.tensor([local_seg], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        n_seg = int(t.item())
    W_sweep = 2.0 * F * n_seg                              # whole job, all ranks
    local_W = 2.0 * F * local_seg

    # initial state of Solver::computeEigenvalue: psi = 0, phi = 1 normalised, stored
    solver.zeroTrackFluxes()
    solver.flattenFSRFluxes(0.0); solver.storeFSRFluxes()
    solver.flattenFSRFluxes(1.0); solver.normalizeFluxes(); solver.storeFSRFluxes()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = args.steps, args.warmup
    if not primary:
        # secondary workloads share the run's time budget: bounded so that a sweep of ~0.1 s still fits
        steps, warmup = max(5, min(args.steps, 20)), 3
    # ---------------- device-resident timing ----------------
    solver.iterate(warmup)
    barrier()
    solver.resetSweepStats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    ev0.record()
    solver.iterate(steps)
    ev1.record()
    torch.cuda.synchronize()
    env.sampler.windows.append((w0, time.time()))
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    _, _, launches = solver.getSweepStats()
    k_dev = solver.getKeff()
    value = W_sweep * steps / (ms * 1e-3)
    # the sweep kernel on its own (CUDA events around the launch on the engine's stream): the multi-GPU
    # loop replays a CUDA graph, inside which no events are recorded, so it is timed separately here
    solver.resetSweepStats()
    for _ in range(max(3, min(steps, 10))):
        solver.transportSweep()
    solver.synchronize()
    barrier()
    sweep_ms, n_sweeps, _ = solver.getSweepStats()

    # ---------------- end to end through the host API ----------------
    n_phi = ft.n_fsrs * G
    host_phi = torch.empty(n_phi, dtype=torch.float64).pin_memory()
    host_np = host_phi.numpy()
    host_np[:] = solver.getFluxes()
    e2e_steps = max(3, min(steps, 10))

    def e2e_step(i):
        solver.setFluxes(host_np)                 # H2D, pinned
        solver.computeFSRSources(1000 + i)
        solver.transportSweep()
        solver.addSourceToScalarFlux()
        solver.computeKeff(fetch=False)
        solver.normalizeFluxes(fetch=False)
        solver.getFluxes(out=host_np)              # D2H into the pinned buffer; the one host sync of the step
        return solver.getKeffNoSync()
    for i in range(2):
        e2e_step(i)
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    env.sampler.windows.append((w0, time.time()))
    if dist is not None:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = W_sweep * e2e_steps / t_e2e

    # ---------------- roofline of the dominant kernel (the sweep) ----------------
    peak, peak_src = measured_peak()
    n_trk_total = ft.n_tracks
    b_alg = (12.0 * n_seg + 16.0 * F * n_trk_total + 16.0 * G * ft.n_fsrs) / W_sweep
    sweep_avg_ms = sweep_ms / max(n_sweeps, 1)
    achieved = b_alg * local_W / (sweep_avg_ms * 1e-3) / 1e9       # GB/s of algorithmic bytes, this rank
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(name, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    if traffic is not None:
        if azim != wl["azim"] or spacing != wl["spacing"]:
            traffic = None                      # the ncu capture is of the workload's own shape
        elif world > 1:
            traffic *= local_W / W_sweep        # one rank's launch streams its share of the captured deck
    rate_local = local_W / (sweep_avg_ms * 1e-3)                    # integrations/s of this rank's sweep kernel
    fp64_ceiling = env.fp64_rate / FP64_PER_INTEGRATION[wl["dims"]]
    # flat 2D tracks merge the tally over the NP polar angles and over runs of segments in one FSR;
    # 3D tracks change FSR at (nearly) every segment: one RED.ADD.F64 per integration at worst
    red_ceiling = env.red_rate if wl["dims"] == 3 else env.red_rate * (F / G)
    binding = min(("fp64_issue", fp64_ceiling), ("red_f64", red_ceiling), key=lambda x: x[1])

    res = {
        "value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
        "ns_per_integration": 1e9 / value,
        "config": {"workload": name, "deck": wl["desc"], "num_azim": azim, "spacing_cm": spacing,
                   "num_polar": wl["polar"], "n_tracks": ft.n_tracks, "n_segments": n_seg,
                   "n_fsrs": ft.n_fsrs, "groups": G, "fluxes_per_track": F,
                   "integrations_per_sweep": W_sweep,
                   "parallelism": "1 GPU" if world == 1 else f"{partition} partition x{world} + NCCL all-reduce of the FSR tally"
                                  + (" + NCCL send/recv of cross-rank boundary fluxes" if partition in ("track", "block", "domain") else ""),
                   "deterministic_tally": bool(args.deterministic),
                   "l2": "segment stream (%.2f GB) is larger than the 126 MB L2; no flush" % (16.0 * n_seg / world / 1e9)
                         if 16.0 * n_seg / world > 2.5e8 else "inputs fit in L2; not flushed (launch-bound shape)",
                   "k_eff_after_timed_steps": k_dev, "track_generation_s": round(t_gen, 2),
                   "upload_and_setup_s": round(t_setup, 2)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel": "b200::sweep_kernel", "bytes_per_integration": b_alg,
                     "kernel_ms": sweep_avg_ms, "kernel_share_of_step": sweep_avg_ms / (ms / steps) if ms > 0 else None,
                     "binding_bound": binding[0], "binding_peak": binding[1], "binding_unit": "integrations/s",
                     "binding_achieved": rate_local, "binding_frac": rate_local / binding[1],
                     "ceilings": {"fp64_instr_per_s": env.fp64_rate, "fp64_instr_per_integration": FP64_PER_INTEGRATION[wl["dims"]],
                                  "red_f64_per_s": env.red_rate, "fp64_issue_integrations_per_s": fp64_ceiling,
                                  "red_integrations_per_s": red_ceiling,
                                  "how": "b200_measure_ceilings (csrc/microbench.cuh) run on this GPU before the timed region"},
                     "note": "HBM is not what binds this kernel (SURVEY 8d): the machine-readable fraction of the binding "
                             "resource is binding_frac"},
        "e2e": {"value": e2e_value, "unit": "integrations/s", "h2d_bytes_per_step": n_phi * 8,
                "d2h_bytes_per_step": n_phi * 8 + 8, "steps": e2e_steps, "ms_per_step": 1e3 * t_e2e / e2e_steps},
        "gpu_launches": int(launches),
    }
    solver.close()
    del solver
    torch.cuda.empty_cache()
    if world == 1 and not args.no_cmfd:
        res["cmfd"] = measure_cmfd(name, ft, args, local_rank, precision)
    if world > 1 and not args.no_group:
        res["group"] = measure_group(name, ft, args, env, steps, warmup, W_sweep, precision)
    return res


def measure_cmfd(name, ft, args, device, precision, devices=None, max_iters=100):
    """Time to solution of the CMFD-accelerated eigenvalue solve of the workload's deck (tolerance 1e-5, the
    reference's default): the whole source iteration - sweep with current tally, closure, CMFD collapse / diffusion
    eigenvalue solve / prolongation (b200_cmfd_*), normalisation, residual, stopping rule - runs on the device."""
    import torch
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import cmfd_mesh
    wl = WORKLOADS[name]
    if "cmfd" not in wl:
        return None
    try:
        mesh = cmfd_mesh(ft, wl["model"], num_z=wl["cmfd"]["num_z"], group_structure=CMFD_GROUPS)
        t0 = time.perf_counter()
        s = B200Solver(ft, device=device, precision=precision, devices=devices, cmfd=mesh)
        t_setup = time.perf_counter() - t0
        s.setConvergenceThreshold(1e-5)
        s.resetSweepStats()
        t0 = time.perf_counter()
        s.computeEigenvalue(max_iters)
        s.synchronize()
        dt = time.perf_counter() - t0
        iters = s.getNumIterations()
        sweep_ms, n_sweeps, _ = s.getSweepStats()
        W = 2.0 * ft.fluxes_per_track * s.num_segments
        out = {"mesh": "%d x %d x %d CMFD cells, %d groups, SOR %.1f" % (mesh.num_x, mesh.num_y, mesh.num_z, len(CMFD_GROUPS), mesh.sor_factor),
               "iterations_to_1e-5": iters, "converged": iters < max_iters, "k_eff": s.getKeff(),
               "time_to_solution_s": dt, "ms_per_iteration": 1e3 * dt / max(iters, 1),
               "sweep_ms_per_iteration": sweep_ms / max(n_sweeps, 1),
               "cmfd_and_fsr_ms_per_iteration": 1e3 * dt / max(iters, 1) - sweep_ms / max(n_sweeps, 1),
               "integrations_per_s_with_cmfd": W * iters / dt, "setup_s": round(t_setup, 2),
               "timed": "host clock around b200_compute_eigenvalue (fused device loop, polled every 8 iterations)"}
        s.close()
        torch.cuda.empty_cache()
        return out
    except Exception as e:                     # the CMFD block must not lose the main measurement
        return {"error": repr(e)}


def cmfd_parity(workload, env, threads):
    """CMFD-accelerated solve of a bounded sample of the deck: the CUDA path against the reference's CPUSolver + Cmfd
    (oracle/_ref/ref_driver --cmfd, k-nearest updating off: its stencils need a Geometry), both to convergence."""
    import numpy as np
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks, make_tracks_3d, cmfd_mesh
    wl = WORKLOADS[workload]
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if "cmfd" not in wl or not os.path.exists(driver):
        return None
    sm = wl["cmfd"]["sample"]
    nz = sm.get("num_z", 1)
    max_iters = 60 if wl["dims"] == 2 else 20
    with tempfile.TemporaryDirectory() as td:
        js = os.path.join(td, "ref.json")
        cmd = [driver] + _driver_args(wl, sm) + ["--cmfd", "51x51" + ("x%d" % nz if wl["dims"] == 3 else ""), "--no-knearest",
                                                 "--max-iters", str(max_iters), "--threads", str(threads), "--quiet",
                                                 "--json", js, "--solver", "cpu"]
        if wl["dims"] == 3:
            cmd.append("--no-fluxes")
        t0 = time.perf_counter()
        subprocess.run(cmd, check=True, env=dict(os.environ, OMP_NUM_THREADS=str(threads)), stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, cwd=td)
        wall = time.perf_counter() - t0
        ref = json.load(open(js))
    if wl["dims"] == 2:
        ft = make_tracks(wl["model"], num_azim=sm["azim"], spacing=sm["spacing"], num_polar=wl["polar"])
    else:
        ft = make_tracks_3d(wl["model"], num_azim=sm["azim"], spacing=sm["spacing"], num_polar=sm["polar"],
                            z_spacing=sm["zspacing"], n_axial=sm["n_axial"], polar_quad=_quad(wl), expand=False)
    s = B200Solver(ft, device=env.local_rank, cmfd=cmfd_mesh(ft, wl["model"], num_z=nz, group_structure=CMFD_GROUPS))
    if wl["dims"] == 3:
        s.setConvergenceThreshold(1e-30)
    t0 = time.perf_counter()
    s.computeEigenvalue(max_iters)
    s.synchronize()
    dt = time.perf_counter() - t0
    out = {"deck": sample_text(wl, sm) + ", CMFD 51 x 51" + (" x %d" % nz if wl["dims"] == 3 else "") + ", 2 groups",
           "iterations_b200": s.getNumIterations(), "iterations_reference": ref["iterations"],
           "k_eff_b200": s.getKeff(), "k_eff_reference": ref["keff"], "dk_pcm": abs(s.getKeff() - ref["keff"]) * 1e5,
           "solve_s_b200": dt, "solve_s_reference": ref["total_time_s"], "reference_cores": threads,
           "reference_wall_incl_ray_tracing_s": round(wall, 1), "tolerance": "north star: 1 pcm, 1e-4"}
    if wl["dims"] == 2 and ref.get("fluxes") and len(ref["fluxes"]) == ft.n_fsrs * ft.num_groups:
        phi, rp = s.getFluxes(), np.asarray(ref["fluxes"])
        out["max_rel_phi_err"] = float(np.max(np.abs(phi - rp) / np.maximum(np.abs(rp), 1e-300)))
    s.close()
    return out


def measure_group(name, ft, args, env, steps, warmup, W_sweep, precision):
    """The same workload with ALL GPUs of the node behind ONE solver handle in ONE process
    (b200_set_devices: the library shards the tracks and sums the tallies with its own two-shot
    all-reduce kernels over NVLink peer memory, no NCCL) - the mode the reference-facing plug-in
    (B200Solver::setDevices) uses.  Rank 0 drives every GPU while the other ranks, whose own solvers
    are closed, wait at a barrier."""
    import torch
    from openmoc_b200.solver import B200Solver
    dist, rank, world = env.dist, env.rank, env.world
    out = None
    # the waiting ranks must not sit in an NCCL barrier: its kernel spins ON THEIR GPU and would time-slice
    # with the shard rank 0 runs there.  A gloo barrier waits on the CPU.
    torch.cuda.synchronize()
    dist.barrier(group=env.cpu_group)
    if rank == 0:
        try:
            t0 = time.perf_counter()
            s = B200Solver(ft, device=0, precision=precision, devices=list(range(world)),
                           deterministic=args.deterministic)
            t_setup = time.perf_counter() - t0
            s.zeroTrackFluxes()
            s.flattenFSRFluxes(0.0); s.storeFSRFluxes()
            s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
            s.iterate(warmup)
            s.synchronize()
            s.resetSweepStats()
            w0 = time.time()
            t0 = time.perf_counter()
            s.iterate(steps)
            s.synchronize()
            dt = time.perf_counter() - t0
            env.sampler.windows.append((w0, time.time()))
            sweep_ms, n_sweeps, launches = s.getSweepStats()
            out = {"value": W_sweep * steps / dt, "ms_per_step": 1e3 * dt / steps, "steps": steps,
                   "timed": "host clock around steps enqueued on all devices + synchronize of every device "
                            "(one process: CUDA events of one stream do not span the group)",
                   "sweep_kernel_ms_slowest_shard": sweep_ms / max(n_sweeps, 1), "gpu_launches": int(launches),
                   "k_eff_after_timed_steps": s.getKeff(), "setup_s": round(t_setup, 2),
                   "collective": "library's own two-shot all-reduce over peer memory (csrc/group.cuh)"}
            s.close()
            torch.cuda.empty_cache()
            if args.group_cmfd and not args.no_cmfd:
                out["cmfd"] = measure_cmfd(name, ft, args, 0, precision, devices=list(range(world)))
        except Exception as e:
            out = {"error": repr(e)}
        torch.cuda.empty_cache()
    dist.barrier(group=env.cpu_group)
    return out


def parity_check(workload, cb, env):
    """The CUDA path on the very deck the CPU baseline just ran (synthetic tracks of the same
    parameters), same number of source iterations from the same initial state: k_eff and scalar
    fluxes against the reference's.  The reference is the checker here, not the thing measured."""
    import numpy as np
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.capi import FISSION_SOURCE
    wl = WORKLOADS[workload]
    ft = make_sample_tracks(workload, expand=(wl["dims"] == 2))
    s = B200Solver(ft, device=env.local_rank)
    s.setConvergenceThreshold(1e-30)
    s.computeEigenvalue(int(cb["iterations"]), FISSION_SOURCE)
    out = {"deck": sample_text(wl, wl["cpu_sample"]), "iterations": int(cb["iterations"]),
           "k_eff_b200": s.getKeff(), "k_eff_reference": cb["keff"],
           "dk_pcm": abs(s.getKeff() - cb["keff"]) * 1e5 if cb.get("keff") is not None else None,
           "n_segments_b200": s.num_segments, "n_segments_reference": cb.get("n_segments"),
           "tolerance": "north star: 1 pcm, 1e-4"}
    ref_phi = cb.get("fluxes")
    if ref_phi is not None and wl["dims"] == 2 and len(ref_phi) == ft.n_fsrs * ft.num_groups:
        phi = s.getFluxes()
        ref_phi = np.asarray(ref_phi)
        out["max_rel_phi_err"] = float(np.max(np.abs(phi - ref_phi) / np.maximum(np.abs(ref_phi), 1e-300)))
    elif ref_phi is not None:
        # 3D: the reference numbers its FSRs in hash-map order; compare the sorted flux spectra
        # and the synthetic deck keeps the FSRs no track crosses, which the reference never creates: the
        # fission-source normalisation (sum = number of FSRs, CPUSolver.cpp:1910) differs by that ratio
        vol = s.getVolumes() if ft.n_segments == 0 else ft.arrays["fsr_volume"]
        n_ref = len(ref_phi) // ft.num_groups
        phi = np.sort(s.getFluxes()[np.repeat(vol > 0, ft.num_groups)]) * (n_ref / ft.n_fsrs)
        ref_sorted = np.sort(np.asarray(ref_phi))
        if phi.size == ref_sorted.size:
            out["max_rel_sorted_phi_err"] = float(np.max(np.abs(phi - ref_sorted) / np.maximum(np.abs(ref_sorted), 1e-300)))
            out["flux_comparison"] = "order statistics of the scalar flux, rescaled by N_FSR(reference) / N_FSR(synthetic)"
    s.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "refgpu"])
    ap.add_argument("--workload", default="c5g7-2d", choices=sorted(WORKLOADS))
    ap.add_argument("--also", default="c5g7-3d", help="comma-separated secondary workloads, or 'none'")
    ap.add_argument("--precision", default="double", choices=["double", "mixed", "table"])
    ap.add_argument("--azim", type=int, default=None)
    ap.add_argument("--spacing", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--partition", default="pair", choices=["pair", "chain", "track", "domain"])
    ap.add_argument("--partition-3d", default="block", choices=["chain", "track", "block"])
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--balance-domains", action="store_true",
                    help="--partition domain: box faces at the quantiles of the segment count instead of equal boxes")
    ap.add_argument("--no-group", action="store_true", help="skip the one-process all-GPU measurement at N > 1")
    ap.add_argument("--no-cmfd", action="store_true", help="skip the CMFD-accelerated time-to-solution blocks")
    ap.add_argument("--group-cmfd", action="store_true",
                    help="N > 1: also time the CMFD-accelerated solve on the one-process all-GPU group (the CMFD solve "
                         "runs replicated on every GPU)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl != "b200":
        return reference_arm(args)

    import torch
    from openmoc_b200 import capi

    env = Env()
    env.rank = int(os.environ.get("RANK", "0"))
    env.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(env.local_rank)
    env.dist = None
    if env.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", env.local_rank))
        env.dist = dist
        env.cpu_group = dist.new_group(backend="gloo")
    env.fp64_rate, env.red_rate = capi.measure_ceilings(env.local_rank, 23869)
    env.sampler = ClockSampler(env.local_rank)
    if env.rank == 0:
        env.sampler.start()

    also = [w for w in args.also.split(",") if w and w != "none" and w != args.workload]
    for w in also:
        if w not in WORKLOADS:
            raise SystemExit(f"unknown workload {w!r}")
    main_res = measure(args.workload, args, env, primary=True)
    others = {w: measure(w, args, env, primary=False) for w in also}
    clocks = env.sampler.stop() if env.rank == 0 else None

    if env.rank == 0:
        ncpu = os.cpu_count() or 1
        cpu_baseline = None
        if env.world == 1 and not args.no_cpu_baseline:
            for w in [args.workload] + also:
                cb = run_reference(w, 6, ncpu, keep_fluxes=True)
                block = {"value": cb["value"], "unit": "integrations/s", "cores": cb["cores"],
                         "kind": cb["kind"], "sample": cb["sample"]}
                try:
                    block["parity"] = parity_check(w, cb, env)
                except Exception as e:                     # a failed check must not lose the measurement
                    block["parity"] = {"error": repr(e)}
                if not args.no_cmfd:
                    try:
                        block["parity_cmfd"] = cmfd_parity(w, env, ncpu)
                    except Exception as e:
                        block["parity_cmfd"] = {"error": repr(e)}
                if w == args.workload:
                    cpu_baseline = block
                else:
                    others[w]["cpu_baseline"] = block
        precision = args.precision
        line = {
            "metric": "segment-group integrations/s", "value": main_res["value"], "unit": "integrations/s",
            "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"double": "f64", "mixed": "f32 segment math, f64 tally",
                      "table": "f64, exponential from an fp32 shared-memory table"}[precision],
            "data": "synthetic",
            "ns_per_integration": main_res["ns_per_integration"],
            "config": main_res["config"],
            "roofline": main_res["roofline"],
            "cpu_baseline": cpu_baseline,
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["gpu_launches"],
            "clocks": clocks,
            "group": main_res.get("group"),
            "cmfd": main_res.get("cmfd"),
            "workloads": others,
        }
        print(json.dumps(line))
    if env.dist is not None:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
