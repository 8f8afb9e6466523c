"""GPU halves of the reference goldens pinned on the CPU at the end of round 2 (tests/test_oracle.py holds the CPU halves:
the oracle and / or the unmodified reference behind ref_driver reproduce the committed results_true.dat of)

  test_2d_gradient_linear_source, test_split_segments, test_split_segments_cmfd, test_forward_3D_lattice_symmetry,
  test_cmfd_pwr_assembly, test_cmfd_vacuum_boundary, test_cmfd_periodic_boundaries, test_cmfd_linear_source,
  test_cmfd_restart, test_transport_stabilization, test_axial_segmentation, test_cmfd_axial_interpolation_average,
  test_cmfd_axial_interpolation_centroid, test_OTF_transport, test_multisim_simple, test_multisim_linear_source,
  test_multisim_cmfd, test_multisim_num_azim, test_multisim_materials, test_multisim_num_groups,
  test_multisim_fixed_source

and the fission-rate test ADVICE r1 asked for.  Written when the round's GPU budget was spent: the CPU halves are
verified, these run for the first time on the driver's box (hence the late file name: the rest of the suite runs
first, and every GPU run here is a child process with a time limit)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

# never run on hardware (see above): an XPASS in the driver's record is the first evidence, an XFAIL a finding -
# neither stops the rest of the suite
pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="GPU half written after the round's GPU budget was spent: first "
                                                     "run is the driver's")]

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
GOLDENS = json.load(open(os.path.join(GOLDEN, "ref_goldens.json")))
SPLIT_ARGS = ["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--max-tau", "0.5", "--quiet", "--no-fluxes",
              "--no-keff", "--results-segments"]


def drive(args, tmp_path, env=None):
    if not os.path.exists(DRIVER):
        pytest.skip("ref_driver not built")
    res = os.path.join(tmp_path, "res.dat")
    # a run that hangs (none has been seen, but these paths are new on hardware) must cost one test, not the suite
    subprocess.run([DRIVER] + args + ["--results", res], check=True, capture_output=True, timeout=240,
                   env=dict(os.environ, **(env or {})))
    return open(res).read()


def in_child(body):
    """Python-mirror work in a child process with a time limit: a fault or a hang cannot take pytest down"""
    head = ("import json, sys\nsys.path.insert(0, %r); sys.path.insert(0, %r)\nimport numpy as np\n"
            "from conftest import load_case\nfrom openmoc_b200.capi import FISSION_SOURCE\n"
            "from openmoc_b200.solver import B200Solver\n" % (ROOT, os.path.join(ROOT, "tests")))
    out = subprocess.run([sys.executable, "-c", head + body], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])


def same_to_printed_precision(out, golden):
    """iterations and k_eff lines byte for byte; fluxes equal to the 7 printed digits (one unit of the last digit
    tolerated: a sum that differs in its last bits may round the other way)"""
    a, b = out.split("fluxes:\n"), golden.split("fluxes:\n")
    assert a[0] == b[0]
    fa, fb = np.array(a[1].split(), dtype=float), np.array(b[1].split(), dtype=float)
    assert fa.shape == fb.shape
    np.testing.assert_allclose(fa, fb, rtol=1.1e-6)
    return out == golden


def test_gradient_2d_linear_source_golden_from_the_plug_in(tmp_path):
    """B200LSSolver : CPULSSolver on the 2-group cube with VACUUM on xmin / ymax"""
    out = drive(["--model", "gradient-2d", "--azim", "4", "--spacing", "0.1", "--solver", "b200ls", "--quiet"], tmp_path)
    print("byte for byte:", same_to_printed_precision(out, GOLDENS["test_2d_gradient_linear_source"]))


def test_gradient_2d_linear_source_golden_from_python():
    """the same golden through the Python mirror on the dumped tracks (pre-pass on the device)"""
    r = in_child("""
from oracle.oracle_py import OracleSolver, format_harness_results
ft, ref = load_case("gradient_2d_ls")
gpu, cpu = B200Solver(ft, linear_source=True), OracleSolver(ft, linear_source=True)
gpu.setConvergenceThreshold(1e-5)
gpu.computeEigenvalue(500, FISSION_SOURCE)
n = cpu.computeEigenvalue(500, 1e-5, FISSION_SOURCE)
print("RESULT " + json.dumps({"gpu_iters": gpu.getNumIterations(), "cpu_iters": n, "ref_iters": ref["iterations"],
    "dk_pcm": abs(gpu.getKeff() - cpu.getKeff()) * 1e5,
    "flux_err": float(np.max(np.abs(gpu.getFluxes() - cpu.getFluxes()) / np.abs(cpu.getFluxes()))),
    "harness": format_harness_results(gpu.getNumIterations(), gpu.getKeff(), gpu.getFluxes())}))
""")
    assert r["gpu_iters"] == r["cpu_iters"] == r["ref_iters"] == 52
    assert r["dk_pcm"] < 1e-4 and r["flux_err"] < 1e-8
    print("byte for byte:", same_to_printed_precision(r["harness"], GOLDENS["test_2d_gradient_linear_source"]))


def test_split_segments_golden_from_the_gpu(tmp_path):
    """Solver::setMaxOpticalLength(0.5): the plug-in flattens after the split (1560 segments), same 262 iterations"""
    assert drive(SPLIT_ARGS + ["--solver", "b200"], tmp_path) == GOLDENS["test_split_segments"]
    r = in_child("""
ft, ref = load_case("pin_cell_split")
s = B200Solver(ft)
s.computeEigenvalue(500, FISSION_SOURCE)
print("RESULT " + json.dumps({"iters": s.getNumIterations(), "ref_iters": ref["iterations"],
                              "dk_pcm": abs(s.getKeff() - ref["keff"]) * 1e5}))
""")
    assert r["iters"] == r["ref_iters"] == 262 and r["dk_pcm"] < 1e-4


@pytest.mark.parametrize("where", ["device", "host"])
def test_split_segments_cmfd_golden_from_the_gpu(where, tmp_path):
    """a 2 x 2 Cmfd with its default options (one CMFD group per MOC group) over the split pin cell: 11 iterations"""
    out = drive(SPLIT_ARGS + ["--cmfd", "2x2", "--cmfd-all-groups", "--no-knearest", "--solver", "b200"], tmp_path,
                env={"B200_HOST_CMFD": "1"} if where == "host" else None)
    assert out == GOLDENS["test_split_segments_cmfd"]


SYMMETRY_ARGS = ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "4", "--spacing", "0.12",
                 "--zspacing", "0.14", "--formation", "otf-stacks", "--symmetry", "--cmfd", "2x2x2", "--tol", "1e-4",
                 "--threads", "4", "--quiet", "--no-fluxes", "--results-fsrs"]


@pytest.mark.parametrize("where", ["device", "host"])
def test_symmetry_golden_from_the_gpu(where, tmp_path):
    """Geometry::useSymmetry(True, True, True): one octant of the 3D lattice, B200LSSolver, OTF_STACKS, CMFD 2 x 2 x 2
    with k-nearest 3: 44 iterations, keff 4.03117E-01, 256 FSRs"""
    out = drive(SYMMETRY_ARGS + ["--solver", "b200ls"], tmp_path, env={"B200_HOST_CMFD": "1"} if where == "host" else None)
    assert out == GOLDENS["test_forward_3D_lattice_symmetry"]


# ------------------------------------------------------------------ CMFD on the 17 x 17 MOX assembly (13 872 FSRs)
PWR = ["--model", "pwr-assembly", "--azim", "4", "--spacing", "0.12", "--cmfd", "17x17", "--quiet"]
PWR_CASES = {
    "test_cmfd_pwr_assembly": (PWR + ["--cmfd-relax", "1.0", "--cmfd-sor", "1.5"], "cpu", "b200"),
    "test_cmfd_vacuum_boundary": (PWR + ["--cmfd-relax", "0.7", "--cmfd-sor", "1.0", "--vacuum-mask", "1"], "cpu", "b200"),
    "test_cmfd_periodic_boundaries": (PWR + ["--cmfd-relax", "0.7", "--cmfd-sor", "1.0", "--periodic-mask", "3"], "cpu", "b200"),
    "test_cmfd_linear_source": (PWR + ["--cmfd-relax", "0.7", "--cmfd-sor", "1.0"], "cpuls", "b200ls"),
    "test_cmfd_restart": (PWR + ["--cmfd-relax", "1.0", "--cmfd-all-groups", "--no-knearest", "--restart"], "cpu", "b200"),
}


@pytest.mark.parametrize("where", ["device", "host"])
@pytest.mark.parametrize("test", sorted(PWR_CASES))
def test_cmfd_assembly_goldens_from_the_gpu(test, where, tmp_path):
    """The goldens are SHA-512 digests over 97 104 fluxes printed with 7 digits: the reference run on this box must give
    the committed digest (that anchors it), the B200 run must give the same iterations, the same printed k_eff and
    the same fluxes to the printed digits - with the CMFD on the device and with the reference's host Cmfd fed by the
    device (reflective, one VACUUM side, two PERIODIC sides, linear source)."""
    args, cpu_solver, gpu_solver = PWR_CASES[test]
    cpu = drive(args + ["--solver", cpu_solver], tmp_path)
    assert hashlib.sha512(cpu.encode()).hexdigest() == GOLDENS[test].strip()
    gpu = drive(args + ["--solver", gpu_solver], tmp_path, env={"B200_HOST_CMFD": "1"} if where == "host" else None)
    same = same_to_printed_precision(gpu, cpu)
    print("digest from the GPU equals the reference's:", same)


STABILIZATION_ARGS = ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--cmfd", "17x17", "--cmfd-relax", "0.7",
                      "--negative-water-scatter", "--stabilize-sequence", "0.4:0,0.4:1,0.4:2", "--quiet"]


@pytest.mark.parametrize("where", ["device", "host"])
def test_transport_stabilization_golden_from_the_gpu(where, tmp_path):
    """B200LSSolver + CMFD with DIAGONAL, YAMAMOTO and GLOBAL stabilisation solved in a row on one solver object (negative
    in-scatter in the moderator): the reference run on this box gives the committed digest, the B200 run the same
    iterations, printed k_eff and fluxes to the printed digits"""
    cpu = drive(STABILIZATION_ARGS + ["--solver", "cpuls"], tmp_path)
    assert hashlib.sha512(cpu.encode()).hexdigest() == GOLDENS["test_transport_stabilization"].strip()
    gpu = drive(STABILIZATION_ARGS + ["--solver", "b200ls"], tmp_path, env={"B200_HOST_CMFD": "1"} if where == "host" else None)
    print("digest from the GPU equals the reference's:", same_to_printed_precision(gpu, cpu))


# ------------------------------------------------------------------ AxialExtendedInput: extruded FSRs with different axial meshes
AXIAL = ["--model", "axial-extended", "--dims", "3", "--azim", "4", "--quiet", "--no-fluxes"]
AXIAL_SEGMENTATION_ARGS = AXIAL + ["--polar", "2", "--spacing", "0.24", "--zspacing", "0.9", "--formation", "otf-tracks",
                                   "--seg-zones", "0,1,2,3,4,5,6,7,8,9,10,20", "--max-iters", "30"]
AXIAL_INTERPOLATION_ARGS = AXIAL + ["--polar", "4", "--quad", "gl", "--spacing", "0.1", "--zspacing", "0.5", "--formation",
                                    "otf-stacks", "--seg-zones", "0,17,18,20", "--cmfd", "1x1", "--cmfd-widths",
                                    "0.05,1.26,1.26,0.05;0.05,1.26,1.26,0.05;1,2,3,4,1,2,3,4", "--cmfd-sor", "1.5",
                                    "--cmfd-relax", "0.7", "--cmfd-all-groups", "--no-knearest", "--tol", "1e-4",
                                    "--threads", "4", "--results-fsrs"]


@pytest.mark.parametrize("tracer", ["device", "host"])
def test_axial_segmentation_golden_from_the_gpu(tracer, tmp_path):
    """OTF_TRACKS with segmentation zones on a non-uniform, axially heterogeneous lattice: the device tracer (and the
    host expansion, B200_HOST_OTF=1) give the reference's 30 unconverged iterations and k_eff to the printed digits"""
    out = drive(AXIAL_SEGMENTATION_ARGS + ["--solver", "b200"], tmp_path, env={"B200_HOST_OTF": "1"} if tracer == "host" else None)
    assert out == GOLDENS["test_axial_segmentation"]


def test_axial_segmentation_golden_from_python():
    """the same golden through the Python mirror on the tracks dumped from the reference (explicit 3D segments)"""
    r = in_child("""
from oracle.oracle_py import format_harness_results
ft, ref = load_case("axial_extended")
s = B200Solver(ft)
s.computeEigenvalue(30, FISSION_SOURCE)
print("RESULT " + json.dumps({"iters": s.getNumIterations(), "dk_pcm": abs(s.getKeff() - ref["keff"]) * 1e5,
                              "harness": format_harness_results(s.getNumIterations(), s.getKeff())}))
""")
    assert r["iters"] == 30 and r["dk_pcm"] < 1e-4
    assert r["harness"] == GOLDENS["test_axial_segmentation"]


@pytest.mark.parametrize("where", ["device", "host"])
@pytest.mark.parametrize("interp,test", [("1", "test_cmfd_axial_interpolation_average"),
                                         ("2", "test_cmfd_axial_interpolation_centroid")])
def test_cmfd_axial_interpolation_goldens_from_the_gpu(interp, test, where, tmp_path):
    """B200LSSolver on OTF_STACKS with a non-uniform Cmfd (Cmfd::setWidths) and the axial interpolation of its
    prolongation: 15 iterations, keff 1.26899E+00, 2137 FSRs"""
    out = drive(AXIAL_INTERPOLATION_ARGS + ["--cmfd-axial-interp", interp, "--solver", "b200ls"], tmp_path,
                env={"B200_HOST_CMFD": "1"} if where == "host" else None)
    assert out == GOLDENS[test]


OTF_TRANSPORT_ARGS = ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "4", "--spacing", "0.4",
                      "--zspacing", "1.2", "--formation", "otf-stacks", "--vacuum-mask", "22", "--cmfd", "4x4x4",
                      "--cmfd-relax", "1.0", "--cmfd-sor", "1.5", "--tol", "1e-3", "--threads", "4", "--quiet", "--no-fluxes"]


@pytest.mark.parametrize("where", ["device", "host"])
def test_otf_transport_golden_from_the_gpu(where, tmp_path):
    """tests/test_OTF_transport (the reference traces its segments while sweeping): z-stacks traced on the device with
    their CMFD surfaces, VACUUM on four of the six sides, CMFD 4 x 4 x 4 on the device / on the host: 20 iterations,
    keff 5.84272E-02"""
    out = drive(OTF_TRANSPORT_ARGS + ["--solver", "b200"], tmp_path, env={"B200_HOST_CMFD": "1"} if where == "host" else None)
    assert out == GOLDENS["test_OTF_transport"]


MULTISIM_CASES = {
    "test_multisim_simple": (["--model", "pin-cell", "--azim", "4", "--spacing", "0.1"], "b200"),
    "test_multisim_linear_source": (["--model", "pin-cell", "--azim", "4", "--spacing", "0.1"], "b200ls"),
    # cells refilled with clones before each solve: Geometry::getAllMaterials() changes its order (clone ids follow the
    # cells), the device image is rebuilt (B200SolverT::ensureDevice keys on the FSR materials)
    "test_multisim_materials": (["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--clone-materials"], "b200"),
    "test_multisim_cmfd": (["--model", "pwr-assembly", "--azim", "4", "--spacing", "0.1", "--cmfd", "17x17", "--cmfd-relax", "1.0",
                            "--cmfd-sor", "1.5", "--max-iters", "5"], "b200"),
}


@pytest.mark.parametrize("test", sorted(MULTISIM_CASES))
def test_multi_simulation_goldens_from_the_gpu(test, tmp_path):
    """MultiSimTestHarness: three eigenvalue solves in a row on one B200 solver object (device image reused, materials
    and fluxes re-initialised by the base class) print the reference's three identical lines"""
    args, solver = MULTISIM_CASES[test]
    assert drive(args + ["--repeat", "3", "--quiet", "--solver", solver], tmp_path) == GOLDENS[test]


def test_num_azim_golden_from_the_gpu(tmp_path):
    """tests/test_multisim_num_azim: the tracks are laid again with 4, 8 and 16 azimuthal angles between the solves of one
    B200Solver - the device image must follow the TrackGenerator (B200SolverT::ensureDevice), never a stale one"""
    out = drive(["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--azim-sequence", "4,8,16", "--quiet",
                 "--solver", "b200"], tmp_path)
    assert out == GOLDENS["test_multisim_num_azim"]


def test_num_groups_golden_from_the_gpu(tmp_path):
    """tests/test_multisim_num_groups: one group, then two groups in the same Material between two solves of one
    B200Solver on the same tracks - the device image keys on the number of groups"""
    out = drive(["--model", "hom-inf", "--azim", "4", "--spacing", "0.1", "--multisim-groups", "--quiet", "--solver", "b200"],
                tmp_path)
    assert out == GOLDENS["test_multisim_num_groups"]


def test_multisim_fixed_source_golden_from_the_gpu(tmp_path):
    """tests/test_multisim_fixed_source: computeSource three times in a row on one B200Solver (water box, source in
    group 1): the reference run on this box gives the committed digest, the B200 run the same iteration counts and the
    same fluxes to the printed digits"""
    args = ["--model", "water-box", "--azim", "4", "--spacing", "0.1", "--mode", "source", "--res", "total",
            "--fixed-source", "1:1.0", "--repeat", "3", "--quiet"]
    cpu = drive(args + ["--solver", "cpu"], tmp_path)
    assert hashlib.sha512(cpu.encode()).hexdigest() == GOLDENS["test_multisim_fixed_source"].strip()
    gpu = drive(args + ["--solver", "b200"], tmp_path)
    words = lambda text: [l for l in text.splitlines() if not l[0].isdigit() and not l[0] == "-"]
    values = lambda text: np.array([float(l) for l in text.splitlines() if l[0].isdigit() or l[0] == "-"])
    assert words(gpu) == words(cpu)                                  # "Iters: 130" / "fluxes:" three times
    np.testing.assert_allclose(values(gpu), values(cpu), rtol=1.1e-6)
    print("digest from the GPU equals the reference's:", gpu == cpu)


def test_fission_rates_without_nu_after_two_solves(tmp_path):
    """the second solve of a B200Solver used to upload an all-zero sigma_f, after which
    computeFSRFissionRates(nu = false) - the reference's default - returned zeros.  Two solves in a row through the
    plug-in, rates with and without nu against CPUSolver; the same through the Python mirror against the oracle."""
    if not os.path.exists(DRIVER):
        pytest.skip("ref_driver not built")
    out = {}
    for solver in ("cpu", "b200"):
        js = os.path.join(tmp_path, solver + ".json")
        subprocess.run([DRIVER, "--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--repeat", "2",
                        "--fission-rates", "--quiet", "--solver", solver, "--json", js], check=True, capture_output=True,
                       timeout=240)
        out[solver] = json.load(open(js))
    assert out["b200"]["iterations"] == out["cpu"]["iterations"] == 187
    for key in ("fission_rates", "nu_fission_rates"):
        ref = np.array(out["cpu"][key])
        assert ref.max() > 0
        np.testing.assert_allclose(out["b200"][key], ref, rtol=1e-8, atol=1e-14)
    r = in_child("""
from oracle.oracle_py import OracleSolver
ft, ref = load_case("simple_lattice")
gpu, cpu = B200Solver(ft), OracleSolver(ft)
for _ in range(2):
    gpu.computeEigenvalue(500, FISSION_SOURCE)
    cpu.computeEigenvalue(500, 1e-5, FISSION_SOURCE)
a, b = gpu.computeFSRFissionRates(nu=False), cpu.computeFSRFissionRates(nu=False)
print("RESULT " + json.dumps({"max": float(b.max()), "err": float(np.max(np.abs(a - b)) / b.max())}))
""")
    assert r["max"] > 0 and r["err"] < 1e-8
