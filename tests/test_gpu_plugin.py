"""The C++ drop-in: `B200Solver : public Solver` (openmoc_b200/cpp/B200Solver.cpp) running
inside the reference's own process, next to the unmodified CPUSolver, on the same
TrackGenerator (oracle/_ref/ref_driver --solver both; built by oracle/Makefile in the CPU
container, travels to the GPU box as a prebuilt binary)."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def run(args):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    out = subprocess.run([DRIVER] + args + ["--quiet"], check=True, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.parametrize("args,iters", [
    (["--model", "pin-cell", "--azim", "4", "--spacing", "0.1"], 261),                       # test_forward_pin_cell
    (["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12"], 187),                # test_forward_simple_lattice
    (["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "2.1",
      "--zspacing", "2.8", "--groups70", "--tol", "5e-3"], 258),                              # test_forward_3D_lattice_70g
    (["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
      "--zspacing", "0.9", "--formation", "otf-stacks"], None),                               # OTF_STACKS flattening
    (["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
      "--zspacing", "0.9", "--formation", "explicit"], None),                                 # EXPLICIT_3D flattening
    # configs[4] shape at coarse tracks: extruded 3D C5G7 core, OTF_STACKS
    (["--model", "c5g7-2d", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "1.0", "--zspacing", "10",
      "--formation", "otf-stacks", "--max-iters", "25", "--threads", "4"], None),
])
def test_b200solver_matches_cpusolver_in_process(args, iters):
    r = run(args + ["--solver", "both"])
    assert r["dk_pcm"] < 1.0                     # north_star
    assert r["max_rel_flux_err"] < 1e-4          # north_star
    assert r["dk_pcm"] < 1e-3 and r["max_rel_flux_err"] < 1e-7   # what it actually achieves
    assert r["b200_iters"] == r["cpu_iters"]
    if iters is not None:
        assert r["cpu_iters"] == iters


def test_fused_loop_through_the_plugin(tmp_path):
    js = os.path.join(tmp_path, "r.json")
    if not os.path.exists(DRIVER):
        pytest.skip("ref_driver not built")
    subprocess.run([DRIVER, "--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--solver", "b200-fused",
                    "--quiet", "--json", js, "--results", os.path.join(tmp_path, "res.dat")], check=True)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_goldens.json")))["test_forward_pin_cell"]
    assert open(os.path.join(tmp_path, "res.dat")).read() == golden


# ---------------------------------------------------------------- linear source (CPULSSolver)
@pytest.mark.parametrize("args,iters", [
    (["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.6",
      "--zspacing", "2.8", "--groups70", "--tol", "5e-3"], 186),          # test_forward_3D_lattice_linear_70g
    (["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12"], None),               # 2D, 3 polar, rings+sectors
    (["--model", "pin-cell", "--azim", "8", "--spacing", "0.05"], None),
    (["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
      "--zspacing", "0.9"], None),                                                            # test_forward_3D_lattice_linear shape
    (["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
      "--zspacing", "0.9", "--formation", "otf-stacks"], None),
])
def test_b200lssolver_matches_cpulssolver_in_process(args, iters):
    """B200LSSolver next to the unmodified CPULSSolver on the same TrackGenerator."""
    r = run(args + ["--solver", "both", "--ls"])
    assert r["dk_pcm"] < 1.0 and r["max_rel_flux_err"] < 1e-4          # north_star
    assert r["dk_pcm"] < 1e-3 and r["max_rel_flux_err"] < 1e-7         # achieved
    assert r["b200_iters"] == r["cpu_iters"]
    if iters is not None:
        assert r["cpu_iters"] == iters


def test_linear_source_reference_golden_from_gpu(tmp_path):
    # tests/test_forward_3D_lattice_linear_70g/results_true.dat: 186 iterations, keff 8.71566E-01
    if not os.path.exists(DRIVER):
        pytest.skip("ref_driver not built")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([DRIVER, "--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2",
                    "--spacing", "0.6", "--zspacing", "2.8", "--groups70", "--tol", "5e-3", "--solver", "b200ls",
                    "--quiet", "--no-fluxes", "--results", res], check=True, capture_output=True)
    assert open(res).read() == "# Iterations: 186\nkeff:  8.71566E-01\n"


# ---------------------------------------------------------------- CMFD acceleration
CMFD_CASES = [
    ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--cmfd", "2x2"],
    ["--model", "simple-lattice", "--azim", "8", "--spacing", "0.05", "--cmfd", "4x4"],
    ["--model", "c5g7-2d", "--azim", "8", "--spacing", "0.2", "--cmfd", "51x51", "--threads", "1", "--max-iters", "60"],
    ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
     "--zspacing", "0.9", "--cmfd", "2x2x2"],
    ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
     "--zspacing", "0.9", "--cmfd", "2x2x2", "--ls", "--formation", "otf-stacks"],
]


@pytest.mark.parametrize("where", ["device", "host"])
@pytest.mark.parametrize("args", CMFD_CASES)
def test_cmfd_accelerated_solve_matches_reference(args, where, monkeypatch):
    """CMFD-accelerated eigenvalue solves against CPUSolver/CPULSSolver + Cmfd on the same tracks: k_eff, fluxes
    and the iteration count must match.  `device`: the sweep tallies the surface currents and the library runs the
    whole Cmfd::computeKeff (current splitting, collapse, diffusion eigenvalue solve, prolongation:
    b200_cmfd_solve).  `host`: the reference's own Cmfd object, fed with the device's fluxes and currents
    (B200_HOST_CMFD=1 / B200Solver::setCmfdOnDevice(false))."""
    if where == "host":
        monkeypatch.setenv("B200_HOST_CMFD", "1")
    r = run(args + ["--solver", "both"])
    assert r["cmfd_on_device"] == (where == "device")
    assert r["b200_iters"] == r["cpu_iters"]
    assert r["dk_pcm"] < 1.0 and r["max_rel_flux_err"] < 1e-4          # north_star
    # CMFD prolongation amplifies summation-order noise (the reference itself moves by ~5e-6 between
    # 1 and 8 OpenMP threads); still far inside the tolerance
    assert r["dk_pcm"] < 1e-2 and r["max_rel_flux_err"] < 2e-5


@pytest.mark.parametrize("args", [
    # every MOC group its own CMFD group (no condensation): 7-group kernel, k-nearest off / on
    ["--model", "pin-cell", "--azim", "8", "--spacing", "0.05", "--cmfd", "3x3", "--no-knearest"],
    ["--model", "hom-inf", "--azim", "4", "--spacing", "0.1", "--cmfd", "2x2"],
    # VACUUM sides: the boundary branch of the surface diffusion coefficients
    ["--model", "gradient-1d", "--azim", "4", "--spacing", "0.1", "--cmfd", "5x1"],
    ["--model", "gradient-2d", "--azim", "4", "--spacing", "0.1", "--cmfd", "3x3"],
    # 70 CMFD groups: the generic (run-time group count) kernel, cooperative grid forced below
    ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "2.1",
     "--zspacing", "2.8", "--groups70", "--tol", "5e-3", "--cmfd", "2x2x2"],
    # flux limiting off
    ["--model", "simple-lattice", "--azim", "8", "--spacing", "0.05", "--cmfd", "4x4", "--no-flux-limiting"],
    # 3D C5G7 (configs[4] shape, coarse tracks): 51x51x9 cells, cooperative-grid solve, vertex and edge currents
    ["--model", "c5g7-2d", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "1.0", "--zspacing", "5",
     "--axial", "9", "--formation", "otf-stacks", "--cmfd", "51x51x9", "--max-iters", "12", "--threads", "8"],
])
def test_device_cmfd_options_and_boundaries(args):
    r = run(args + ["--solver", "both"])
    assert r["cmfd_on_device"]
    assert r["b200_iters"] == r["cpu_iters"]
    assert r["dk_pcm"] < 1e-2 and r["max_rel_flux_err"] < 2e-5


@pytest.mark.parametrize("mode,cluster", [("0", None), ("1", None), ("2", None), ("2", "3"), ("2", "16")])
@pytest.mark.parametrize("args", [
    ["--model", "simple-lattice", "--azim", "8", "--spacing", "0.05", "--cmfd", "4x4"],
    ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
     "--zspacing", "0.9", "--cmfd", "5x4x3", "--no-knearest"],
])
def test_device_cmfd_launch_shapes_agree(args, mode, cluster, monkeypatch):
    """The eigenvalue kernel as one CTA (flux in shared memory, __syncthreads), as a cooperative grid (flux in L2,
    grid barrier) and as one thread-block cluster (flux in distributed shared memory, cluster barrier; also with
    more CTAs than the mesh needs): all against the reference."""
    monkeypatch.setenv("B200_CMFD_MODE", mode)
    if cluster is not None:
        monkeypatch.setenv("B200_CMFD_CLUSTER", cluster)
    r = run(args + ["--solver", "both"])
    assert r["cmfd_on_device"] and r["b200_iters"] == r["cpu_iters"]
    assert r["dk_pcm"] < 1e-4 and r["max_rel_flux_err"] < 1e-9


@pytest.mark.parametrize("args", [
    ["--model", "simple-lattice", "--azim", "8", "--spacing", "0.05", "--cmfd", "4x4"],
    ["--model", "c5g7-2d", "--azim", "8", "--spacing", "0.2", "--cmfd", "51x51", "--max-iters", "60"],
])
def test_fused_cmfd_loop_matches_cpusolver(args, tmp_path):
    """computeEigenvalueFused with CMFD: the whole CMFD-accelerated source iteration (sweep with current tally,
    closure, CMFD solve, prolongation, normalisation, residual, stopping rule) runs on the device without a host
    round trip per step; compared with CPUSolver + Cmfd run in a separate process."""
    import numpy as np
    if not os.path.exists(DRIVER):
        pytest.skip("ref_driver not built")
    res = {}
    for solver in ("cpu", "b200-fused"):
        js = os.path.join(tmp_path, solver + ".json")
        subprocess.run([DRIVER] + args + ["--solver", solver, "--threads", "1", "--quiet", "--json", js],
                       check=True, capture_output=True)
        res[solver] = json.load(open(js))
    a, b = res["cpu"], res["b200-fused"]
    assert a["iterations"] == b["iterations"]
    fa, fb = np.array(a["fluxes"]), np.array(b["fluxes"])
    assert abs(a["keff"] - b["keff"]) * 1e5 < 1e-2
    assert np.max(np.abs(fa - fb) / np.abs(fa)) < 2e-5


CMFD_GOLDEN_ARGS = ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "4", "--spacing", "0.12",
                    "--zspacing", "0.5", "--formation", "otf-stacks", "--cmfd", "4x4x4", "--cmfd-relax", "1.0", "--tol", "1e-4",
                    "--quiet", "--no-fluxes", "--results-fsrs"]


@pytest.mark.parametrize("where", ["device", "host"])
def test_cmfd_reference_golden_from_gpu(where, tmp_path, monkeypatch):
    """tests/test_forward_3D_lattice_CMFD/results_true.dat (CPULSSolver, OTF_STACKS, CMFD 4 x 4 x 4, two groups,
    k-nearest 3, relaxation 1.0: 24 iterations, keff 8.37390E-01, 2048 FSRs) byte for byte from B200LSSolver, with
    the CMFD on the device and with the reference's host Cmfd fed by the device."""
    if not os.path.exists(DRIVER):
        pytest.skip("ref_driver not built")
    if where == "host":
        monkeypatch.setenv("B200_HOST_CMFD", "1")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([DRIVER] + CMFD_GOLDEN_ARGS + ["--solver", "b200ls", "--threads", "4", "--results", res],
                   check=True, capture_output=True)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_goldens.json")))["test_forward_3D_lattice_CMFD"]
    assert open(res).read() == golden


def test_3d_c5g7_linear_source_cmfd_in_separate_processes(tmp_path):
    """configs[4] shape at coarse tracks: extruded 3D C5G7, OTF_STACKS, CPULSSolver, CMFD 51x51x3.
    Run in two fresh processes: `--solver both` shares one Cmfd object between the two solves and
    the second solve (whichever solver runs it) inherits state from the first - measured 0.05 pcm on
    this strongly oscillating, unconverged deck against 2e-11 pcm between fresh processes."""
    import numpy as np
    args = ["--model", "c5g7-2d", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "1.0",
            "--zspacing", "10", "--formation", "otf-stacks", "--cmfd", "51x51x3", "--max-iters", "20",
            "--threads", "1", "--quiet", "--no-fluxes"]      # one thread: same FSR numbering in both runs
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = {}
    for solver in ("cpuls", "b200ls"):
        js = os.path.join(tmp_path, solver + ".json")
        subprocess.run([DRIVER] + args + ["--solver", solver, "--json", js], check=True, capture_output=True)
        res[solver] = json.load(open(js))
    a, b = res["cpuls"], res["b200ls"]
    assert a["iterations"] == b["iterations"] == 20
    dk_pcm = abs(a["keff"] - b["keff"]) * 1e5
    fa, fb = np.array(a["fluxes"]), np.array(b["fluxes"])
    err = np.max(np.abs(fa - fb) / np.abs(fa))
    print("3D C5G7 LS + CMFD: dk = %.3e pcm, max rel flux err = %.3e" % (dk_pcm, err))
    assert dk_pcm < 1.0 and err < 1e-4           # north_star
    assert dk_pcm < 1e-3 and err < 1e-6          # achieved: ~1e-11 pcm


# ---------------------------------------------------------------- several devices behind the plug-in
def _devices():
    import torch
    n = torch.cuda.device_count()
    return ["0,0", "0,0,0"] + (["0,1"] if n >= 2 else []) + ([",".join(map(str, range(n)))] if n >= 4 else [])


@pytest.mark.parametrize("devices", _devices() if os.path.exists(DRIVER) else ["0,0"])
@pytest.mark.parametrize("args", [
    ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12"],                                   # flat 2D
    ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--ls"],                          # linear source
    ["--model", "simple-lattice", "--azim", "8", "--spacing", "0.05", "--cmfd", "4x4"],                 # CMFD currents reduced across shards
    ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
     "--zspacing", "0.9", "--cmfd", "2x2x2", "--ls", "--formation", "otf-stacks"],                       # 3D + LS + CMFD
    ["--model", "c5g7-2d", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "1.0", "--zspacing", "10",
     "--formation", "otf-stacks", "--max-iters", "25", "--threads", "4"],                                # configs[4] shape, coarse
])
def test_multi_device_b200solver_matches_cpusolver_in_process(args, devices):
    """B200Solver::setDevices: ONE solver object inside the reference's process drives several shards
    (b200_set_devices); flat and linear source; with CMFD the summed currents feed the device CMFD, which runs
    replicated on every shard (same bits on each).  Same tolerances as the single-device cases above."""
    r = run(args + ["--solver", "both", "--devices", devices])
    assert r["b200_iters"] == r["cpu_iters"]
    assert r["dk_pcm"] < 1.0 and r["max_rel_flux_err"] < 1e-4          # north_star
    tight = (1e-2, 2e-5) if "--cmfd" in args else (1e-3, 1e-7)
    assert r["dk_pcm"] < tight[0] and r["max_rel_flux_err"] < tight[1]


# ---------------------------------------------------------------- OTF decks through the device tracer
@pytest.mark.parametrize("formation", ["otf-stacks", "otf-tracks"])
def test_otf_formations_are_traced_on_the_device(formation, monkeypatch):
    """OTF_TRACKS / OTF_STACKS decks: the plug-in hands the 2D segments, the ExtrudedFSR axial meshes and
    the start point of every 3D track to the device tracer (b200_upload_tracks_otf) - no 3D segment is made
    on the host.  Same answer as CPUSolver, and as the host expansion of round 1 (B200_HOST_OTF=1)."""
    args = ["--model", "c5g7-2d", "--dims", "3", "--azim", "4", "--polar", "4", "--spacing", "0.8", "--zspacing", "6",
            "--axial", "3", "--quad", "equal-angle", "--formation", formation, "--max-iters", "15", "--threads", "4",
            "--solver", "both"]
    dev = run(args)
    assert dev["b200_iters"] == dev["cpu_iters"] == 15
    assert dev["dk_pcm"] < 1e-3 and dev["max_rel_flux_err"] < 1e-7
    monkeypatch.setenv("B200_HOST_OTF", "1")
    host = run(args)
    assert abs(host["b200_keff"] - dev["b200_keff"]) < 1e-10


@pytest.mark.parametrize("args", [
    ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24", "--zspacing", "0.9",
     "--cmfd", "2x2x2", "--formation", "otf-stacks"],
    ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24", "--zspacing", "0.9",
     "--cmfd", "4x4x3", "--formation", "otf-tracks", "--max-tau", "0.5"],
    ["--model", "c5g7-2d", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "1.0", "--zspacing", "5", "--axial", "9",
     "--formation", "otf-stacks", "--cmfd", "51x51x9", "--max-iters", "12", "--threads", "8"],
    ["--model", "c5g7-2d", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "1.0", "--zspacing", "5", "--axial", "9",
     "--formation", "otf-stacks", "--cmfd", "51x51x9", "--max-iters", "12", "--threads", "8", "--devices", "0,0"],
])
def test_otf_decks_with_cmfd_are_traced_on_the_device(args, monkeypatch):
    """OTF decks WITH CMFD: the device tracer also produces the CMFD surface of every 3D segment
    (b200_upload_otf_cmfd: TraverseSegments.cpp:429-457 + Lattice::getLatticeSurfaceOTF restated in otf.cuh), so
    the current tally needs no host expansion either.  Against CPUSolver + Cmfd, and against the host expansion
    through the reference's own traversal (B200_HOST_OTF=1)."""
    dev = run(args + ["--solver", "both"])
    assert dev["cmfd_on_device"] and dev["b200_iters"] == dev["cpu_iters"]
    assert dev["dk_pcm"] < 1e-2 and dev["max_rel_flux_err"] < 2e-5
    monkeypatch.setenv("B200_HOST_OTF", "1")
    host = run(args + ["--solver", "both"])
    assert abs(host["b200_keff"] - dev["b200_keff"]) < 5e-10


def test_fixed_linear_source_golden_through_the_plugin(tmp_path):
    """B200LSSolver inside the reference's process: setFixedSourceByCell + setFixedSourceMomentsByCell +
    allowNegativeFluxes + computeFlux reproduce tests/test_fixed_linear_source/results_true.dat"""
    if not os.path.exists(DRIVER):
        pytest.skip("ref_driver not built")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([DRIVER, "--model", "water-box", "--azim", "4", "--spacing", "0.1", "--solver", "b200ls", "--mode", "flux",
                    "--res", "flux", "--allow-negative", "--fixed-source", "1:1.0,2:0.5,3:0.25,4:1.0,5:0.5,6:0.25,7:1.0",
                    "--fixed-moments", "1:0.01:0.1:0.2,2:-0.1:0:-0.04,3:0.02:0:0", "--quiet", "--results", res],
                   check=True, capture_output=True)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_goldens.json")))["test_fixed_linear_source"]
    assert open(res).read() == golden


@pytest.mark.parametrize("stab", ["0.5:2", "0.7:0"])
def test_stabilised_linear_source_through_the_plugin(stab):
    """B200LSSolver::stabilizeTransport next to CPULSSolver in one process (moments stabilised too)"""
    r = run(["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--solver", "both", "--ls",
             "--stabilize", stab, "--max-iters", "1000"])
    assert r["b200_iters"] == r["cpu_iters"]
    assert r["dk_pcm"] < 1e-3 and r["max_rel_flux_err"] < 1e-7


@pytest.mark.parametrize("formation", ["otf-stacks", "otf-tracks", "explicit"])
def test_maximum_optical_length_cuts(formation):
    """Solver::setMaxOpticalLength(0.5): the reference's on-the-fly kernels cut every 3D segment longer than
    that (MOCKernel.cpp:216-268, 353-410); the device tracer cuts the same pieces (b200_set_max_optical_length)."""
    r = run(["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
             "--zspacing", "0.9", "--formation", formation, "--max-tau", "0.5", "--solver", "both"])
    assert r["b200_iters"] == r["cpu_iters"]
    assert r["dk_pcm"] < 1e-3 and r["max_rel_flux_err"] < 1e-7


@pytest.mark.parametrize("devices", [None, "0,0"])
def test_cmfd_sigma_t_rebalance_starting_currents(devices):
    """Cmfd::rebalanceSigmaT(true): every sweep starts by tallying the currents the starting angular fluxes
    carry into the boundary CMFD cells (CPUSolver::tallyStartingCurrents, src/CPUSolver.cpp:498-537); the
    plug-in feeds Cmfd::tallyStartingCurrent from the device's start fluxes."""
    args = ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
            "--zspacing", "0.9", "--cmfd", "2x2x2", "--rebalance", "--solver", "both"]
    if devices:
        args += ["--devices", devices]
    r = run(args)
    assert r["b200_iters"] == r["cpu_iters"]
    assert r["dk_pcm"] < 1e-2 and r["max_rel_flux_err"] < 2e-5
