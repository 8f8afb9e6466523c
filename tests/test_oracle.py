"""Pins the plain-C oracle (oracle/moc_oracle.c) to the reference.

Two anchors, neither needs /root/reference at test time:
  * the reference's own committed goldens (tests/golden/ref_goldens.json holds
    tests/test_forward_*/results_true.dat verbatim), compared byte-for-byte in
    the format of tests/testing_harness.py:158-207;
  * full-precision results of the unmodified reference CPUSolver run on the same
    decks (tests/golden/<case>.json, written by tests/golden/make_fixtures.py).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_case
from oracle.oracle_py import (OracleSolver, format_harness_results, lib,
                              FISSION_SOURCE, SCALAR_FLUX, TOTAL_SOURCE)

GOLDENS = json.load(open(os.path.join(GOLDEN, "ref_goldens.json")))


def solve(name, tol=1e-5, max_iters=500):
    ft, ref = load_case(name)
    ft.validate()
    s = OracleSolver(ft)
    n = s.computeEigenvalue(max_iters, tol, FISSION_SOURCE)
    return s, n, ref


def test_pin_cell_golden_bytes():
    s, n, ref = solve("pin_cell")
    out = format_harness_results(n, s.getKeff(), s.getFluxes())
    assert out == GOLDENS["test_forward_pin_cell"]
    assert n == ref["iterations"] == 261
    assert abs(s.getKeff() - ref["keff"]) < 1e-12
    np.testing.assert_allclose(s.getFluxes(), ref["fluxes"], rtol=1e-11)


def test_simple_lattice_golden_sha512():
    s, n, ref = solve("simple_lattice")
    out = format_harness_results(n, s.getKeff(), s.getFluxes())
    assert hashlib.sha512(out.encode()).hexdigest() == GOLDENS["test_forward_simple_lattice"].strip()
    assert n == ref["iterations"] == 187
    np.testing.assert_allclose(s.getFluxes(), ref["fluxes"], rtol=1e-10)


def test_hom_inf_medium_golden_bytes():
    s, n, ref = solve("hom_inf")
    out = format_harness_results(n, s.getKeff(), s.getFluxes())
    assert out == GOLDENS["test_forward_hom_inf_medium"]


def test_lattice3d_70g_golden_bytes():
    s, n, ref = solve("lattice3d_70g", tol=5e-3)
    assert format_harness_results(n, s.getKeff()) == GOLDENS["test_forward_3D_lattice_70g"]
    assert abs(s.getKeff() - ref["keff"]) < 1e-11


def test_lattice3d_7g_golden_bytes():
    s, n, ref = solve("lattice3d_7g")
    assert format_harness_results(n, s.getKeff()) == GOLDENS["test_forward_3D_lattice"]
    np.testing.assert_allclose(s.getFluxes(), ref["fluxes"], rtol=1e-9)


def test_c5g7_coarse_matches_reference_run():
    s, n, ref = solve("c5g7_2d_coarse", max_iters=40)
    assert n == ref["iterations"] == 40
    assert abs(s.getKeff() - ref["keff"]) < 1e-10


def test_expF1_known_answers():
    # tests/unit_tests/test_exponentials.py:13,74-77 (expF1_fractional)
    taus = [1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1, 10, 100]
    expect = [0.9999999950003422, 0.9999999500034228, 0.9999995000343784,
              0.9999950003587617, 0.9999500050851808, 0.9995002005718668,
              0.9950169413610763, 0.9516272442309586, 0.6321211831479621,
              0.09999547551656497, 0.010000005017389789]
    got = [lib().moc_oracle_expF1(t) for t in taus]
    assert np.all(np.abs(np.array(got) - np.array(expect)) < 1e-8)


def test_threads_do_not_change_answer():
    ft, ref = load_case("simple_lattice")
    a = OracleSolver(ft); a.computeEigenvalue(30, 1e-5)
    b = OracleSolver(ft); b.setNumThreads(4); b.computeEigenvalue(30, 1e-5)
    np.testing.assert_allclose(a.getFluxes(), b.getFluxes(), rtol=1e-11)


def test_fixed_source_flux_positive_and_converges():
    ft, _ = load_case("pin_cell")
    s = OracleSolver(ft)
    s.setFixedSourceByFSR(1, 1, 1.0)
    n = s.computeFlux(400, 1e-6)
    assert n < 400
    phi = s.getFluxes().reshape(-1, ft.num_groups)
    # computeFlux evaluates the source once (Solver.cpp:1390): only the fixed-source group is lit
    assert np.all(phi[:, 0] > 0) and np.all(phi[:, 1:] == 0)


# ------------------------------------------------------------------ linear source
def solve_ls(name, tol=1e-5, max_iters=500):
    ft, ref = load_case(name)
    ft.validate()
    s = OracleSolver(ft, linear_source=True)
    n = s.computeEigenvalue(max_iters, tol, FISSION_SOURCE)
    return s, n, ref


def test_linear_source_3d_70g_golden_bytes():
    # tests/test_forward_3D_lattice_linear_70g/results_true.dat: 186 iterations, keff 8.71566E-01
    s, n, ref = solve_ls("lattice3d_ls_70g", tol=5e-3)
    assert format_harness_results(n, s.getKeff()) == GOLDENS["test_forward_3D_lattice_linear_70g"]
    assert n == ref["iterations"] and abs(s.getKeff() - ref["keff"]) < 1e-11
    assert s.num_flat_fsrs == 200         # "Unable to form linear source components in 200 / 280 source regions"


def test_linear_source_3d_7g_golden_bytes():
    # tests/test_forward_3D_lattice_linear/results_true.dat: 156 iterations, keff 6.89615E-01, 480 FSRs
    s, n, ref = solve_ls("lattice3d_ls_7g")
    out = format_harness_results(n, s.getKeff()) + "# FSRs: {0}\n".format(s.ft.n_fsrs)
    assert out == GOLDENS["test_forward_3D_lattice_linear"]
    assert n == ref["iterations"] and abs(s.getKeff() - ref["keff"]) < 1e-11


def test_linear_source_2d_matches_reference_run():
    # CPULSSolver on the 2D simple lattice (rings and sectors: centroids away from the cell centres)
    s, n, ref = solve_ls("simple_lattice_ls")
    assert n == ref["iterations"] == 188
    assert abs(s.getKeff() - ref["keff"]) * 1e5 < 1e-6
    phi, rf = s.getFluxes(), np.array(ref["fluxes"])
    assert np.max(np.abs(phi - rf) / np.abs(rf)) < 1e-12
    # the linear source really is in play: the flat-source answer on the same tracks differs
    flat, n_flat, _ = solve("simple_lattice")
    assert abs(flat.getKeff() - s.getKeff()) * 1e5 > 50
    m = s.getFluxMoments().reshape(-1, 3, 7)
    assert np.abs(m[:, :2]).max() > 1e-3 and np.all(m[:, 2] == 0.0)      # no z moment in 2D


def test_linear_source_threads_do_not_change_answer():
    ft, _ = load_case("simple_lattice_ls")
    res = []
    for threads in (1, 4):
        s = OracleSolver(ft, linear_source=True)
        s.setNumThreads(threads)
        s.computeEigenvalue(40, 1e-30, FISSION_SOURCE)
        res.append((s.getKeff(), s.getFluxes()))
    assert abs(res[0][0] - res[1][0]) < 1e-12
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-11)


# ------------------------------------------------------------------ fixed-source drivers
def format_flux_results(num_iters, fluxes):
    """tests/testing_harness.py:158-207 for solution types without an eigenvalue"""
    return ("# Iterations: {0}\n".format(num_iters) + "fluxes:\n"
            + "\n".join("{0:12.6E}".format(f) for f in np.ravel(fluxes)) + "\n")


def test_compute_flux_golden_bytes():
    # tests/test_compute_flux: water box with VACUUM sides, fixed source 1.0 / 0.5 / 0.25 in groups 1-3
    ft, ref = load_case("water_box")
    s = OracleSolver(ft)
    for fsr in ref["source_fsrs"]:
        for group, value in ((1, 1.0), (2, 0.5), (3, 0.25)):
            s.setFixedSourceByFSR(fsr, group, value)
    n = s.computeFlux(500, 1e-5, True)
    assert format_flux_results(n, s.getFluxes()) == GOLDENS["test_compute_flux"]
    np.testing.assert_allclose(s.getFluxes(), ref["fluxes"], rtol=1e-12)


FIXED_FLAT = ((1, 1.0), (2, 0.5), (3, 0.25), (4, 1.0), (5, 0.5), (6, 0.25), (7, 1.0))
FIXED_MOMENTS = ((1, 0.01, 0.1, 0.2), (2, -0.1, 0.0, -0.04), (3, 0.02, 0.0, 0.0))


def test_fixed_linear_source_golden_bytes():
    # tests/test_fixed_linear_source: the water box, CPULSSolver::computeFlux, a flat fixed source in all seven
    # groups plus x, y, z moments in the first three, negative fluxes allowed (16th reference golden)
    ft, ref = load_case("water_box_ls")
    s = OracleSolver(ft, linear_source=True)
    s.allowNegativeFluxes(True)
    for fsr in ref["source_fsrs"]:
        for group, value in FIXED_FLAT:
            s.setFixedSourceByFSR(fsr, group, value)
        for group, sx, sy, sz in FIXED_MOMENTS:
            s.setFixedSourceMomentsByFSR(fsr, group, sx, sy, sz)
    n = s.computeFlux(500, 1e-5, True)
    assert format_flux_results(n, s.getFluxes()) == GOLDENS["test_fixed_linear_source"]
    np.testing.assert_allclose(s.getFluxes(), ref["fluxes"], rtol=1e-12)


def test_compute_source_golden_bytes():
    # tests/test_compute_source: same deck, source 1.0 in group 1, computeSource with TOTAL_SOURCE residual
    ft, ref = load_case("water_box")
    s = OracleSolver(ft)
    for fsr in ref["source_fsrs"]:
        s.setFixedSourceByFSR(fsr, 1, 1.0)
    n = s.computeSource(500, 1.0, 1e-5, TOTAL_SOURCE)
    assert n == 130
    assert format_flux_results(n, s.getFluxes()) == GOLDENS["test_compute_source"]


# ------------------------------------------------------------------ adjoint mode
def adjoint_tracks(name):
    """Solver::initializeMaterials(ADJOINT) transposes the scattering and fission matrices of every
    material (src/Solver.cpp:805-806, Material::transposeProductionMatrices); on flattened data that
    is a transpose of two arrays."""
    import copy
    ft, ref = load_case(name)
    ft = copy.deepcopy(ft)
    G = ft.num_groups
    for k in ("mat_sigma_s", "mat_fiss_matrix"):
        ft.arrays[k] = ft.arrays[k].reshape(-1, G, G).transpose(0, 2, 1).copy().ravel()
    return ft


@pytest.mark.parametrize("fixture,test,iters", [("pin_cell", "test_adjoint_pin_cell", 331),
                                                ("simple_lattice", "test_adjoint_simple_lattice", 315),
                                                ("hom_inf", "test_adjoint_hom_inf_medium", 163)])
def test_adjoint_goldens(fixture, test, iters):
    s = OracleSolver(adjoint_tracks(fixture))
    n = s.computeEigenvalue(500, 1e-5, FISSION_SOURCE)
    assert n == iters
    out = format_harness_results(n, s.getKeff(), s.getFluxes())
    gold = GOLDENS[test]
    assert out == gold or hashlib.sha512(out.encode()).hexdigest() == gold.strip()


def test_pin_cell_70g_golden_bytes():
    # tests/test_forward_pin_cell_70g: 70 groups in 2D (F = 210), SCALAR_FLUX residual, 8 iterations
    ft, ref = load_case("pin_cell_70g")
    s = OracleSolver(ft)
    n = s.computeEigenvalue(500, 1e-5, SCALAR_FLUX)
    assert n == 8
    assert format_harness_results(n, s.getKeff(), s.getFluxes()) == GOLDENS["test_forward_pin_cell_70g"]


@pytest.mark.parametrize("fixture,test,iters", [("gradient_1d", "test_1d_gradient", 46),
                                                ("gradient_2d", "test_2d_gradient", 52)])
def test_vacuum_gradient_goldens(fixture, test, iters):
    # tests/test_1d_gradient, tests/test_2d_gradient: 2-group cube with VACUUM on two sides
    s, n, _ = solve(fixture)
    assert n == iters
    assert format_harness_results(n, s.getKeff(), s.getFluxes()) == GOLDENS[test]


def test_linear_source_global_stabilisation_matches_reference():
    """CPULSSolver with stabilizeTransport(0.5, GLOBAL): the flux moments are damped like the scalar flux
    (src/CPULSSolver.cpp:888-1052); full-precision run of the unmodified reference on the same tracks"""
    ft, _ = load_case("simple_lattice_ls")
    ref = json.load(open(os.path.join(GOLDEN, "simple_lattice_ls_stab.json")))
    s = OracleSolver(ft, linear_source=True)
    s.stabilizeTransport(0.5, 2)
    n = s.computeEigenvalue(1000, 1e-5, FISSION_SOURCE)
    assert n == ref["iterations"] == 234
    assert abs(s.getKeff() - ref["keff"]) * 1e5 < 1e-6
    np.testing.assert_allclose(s.getFluxes(), ref["fluxes"], rtol=1e-12)


def test_ref_driver_reproduces_the_cmfd_golden(tmp_path):
    """The checker of the CMFD path is the unmodified reference driven by oracle/_ref/ref_driver: its restatement of
    tests/test_forward_3D_lattice_CMFD (SimpleLatticeInput 3D, OTF_STACKS, CPULSSolver, CMFD 4 x 4 x 4) must give the
    reference's committed golden byte for byte."""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver, "--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "4", "--spacing", "0.12",
                    "--zspacing", "0.5", "--formation", "otf-stacks", "--cmfd", "4x4x4", "--cmfd-relax", "1.0", "--tol", "1e-4",
                    "--solver", "cpuls", "--threads", "4", "--quiet", "--no-fluxes", "--results-fsrs", "--results", res],
                   check=True, capture_output=True)
    assert open(res).read() == GOLDENS["test_forward_3D_lattice_CMFD"]


def test_gradient_2d_linear_source_golden_bytes():
    # tests/test_2d_gradient_linear_source/results_true.dat: CPULSSolver, VACUUM on xmin / ymax, 52 iterations
    s, n, ref = solve_ls("gradient_2d_ls")
    assert n == ref["iterations"] == 52
    assert format_harness_results(n, s.getKeff(), s.getFluxes()) == GOLDENS["test_2d_gradient_linear_source"]


def test_split_segments_golden_bytes():
    # tests/test_split_segments/results_true.dat: Solver::setMaxOpticalLength(0.5) on the pin cell, 196 -> 1560 segments;
    # the fixture was dumped after the split, so the oracle sweeps the segments CPUSolver swept: 262 iterations, not 261
    s, n, ref = solve("pin_cell_split")
    assert s.ft.n_segments == 1560
    assert "# Iterations: {0}\n# segments: {1}\n".format(n, s.ft.n_segments) == GOLDENS["test_split_segments"]
    assert n == ref["iterations"] and abs(s.getKeff() - ref["keff"]) < 1e-11


SPLIT_ARGS = ["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--max-tau", "0.5", "--quiet", "--no-fluxes",
              "--no-keff", "--results-segments"]
SPLIT_CMFD_ARGS = SPLIT_ARGS + ["--cmfd", "2x2", "--cmfd-all-groups", "--no-knearest"]


@pytest.mark.parametrize("args,test", [(SPLIT_ARGS, "test_split_segments"), (SPLIT_CMFD_ARGS, "test_split_segments_cmfd")])
def test_ref_driver_reproduces_the_split_segment_goldens(args, test, tmp_path):
    """tests/test_split_segments and tests/test_split_segments_cmfd (a 2 x 2 Cmfd with its default options over the
    pin cell: 11 iterations, 1616 segments) from the unmodified reference through ref_driver"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver] + args + ["--solver", "cpu", "--results", res], check=True, capture_output=True)
    assert open(res).read() == GOLDENS[test]


SYMMETRY_ARGS = ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "4", "--spacing", "0.12",
                 "--zspacing", "0.14", "--formation", "otf-stacks", "--symmetry", "--cmfd", "2x2x2", "--tol", "1e-4",
                 "--threads", "4", "--quiet", "--no-fluxes", "--results-fsrs"]


def test_ref_driver_reproduces_the_symmetry_golden(tmp_path):
    """tests/test_forward_3D_lattice_symmetry: Geometry::useSymmetry(True, True, True) (one octant: 256 FSRs), CPULSSolver,
    OTF_STACKS, CMFD 2 x 2 x 2 with k-nearest 3: 44 iterations, keff 4.03117E-01"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver] + SYMMETRY_ARGS + ["--solver", "cpuls", "--results", res], check=True, capture_output=True)
    assert open(res).read() == GOLDENS["test_forward_3D_lattice_symmetry"]


# tests/test_cmfd_*: PwrAssemblyInput (17 x 17 MOX assembly, 13 872 FSRs), 4 azim, 0.12 cm, CMFD 17 x 17, two groups,
# k-nearest 3; the committed goldens are SHA-512 digests of iterations + k_eff + all fluxes
PWR = ["--model", "pwr-assembly", "--azim", "4", "--spacing", "0.12", "--cmfd", "17x17", "--quiet"]
PWR_CASES = {
    "test_cmfd_pwr_assembly": PWR + ["--cmfd-relax", "1.0", "--cmfd-sor", "1.5", "--solver", "cpu"],
    "test_cmfd_vacuum_boundary": PWR + ["--cmfd-relax", "0.7", "--cmfd-sor", "1.0", "--vacuum-mask", "1", "--solver", "cpu"],
    "test_cmfd_periodic_boundaries": PWR + ["--cmfd-relax", "0.7", "--cmfd-sor", "1.0", "--periodic-mask", "3", "--solver", "cpu"],
    "test_cmfd_linear_source": PWR + ["--cmfd-relax", "0.7", "--cmfd-sor", "1.0", "--solver", "cpuls"],
    # a Cmfd with its default group structure (7 groups) and k-nearest; Solver::setRestartStatus(True) and a second
    # computeEigenvalue that starts from the converged fluxes (2 iterations)
    "test_cmfd_restart": PWR + ["--cmfd-relax", "1.0", "--cmfd-all-groups", "--no-knearest", "--restart", "--solver", "cpu"],
}


@pytest.mark.parametrize("test", sorted(PWR_CASES))
def test_ref_driver_reproduces_the_cmfd_assembly_goldens(test, tmp_path):
    """The CMFD checker (the unmodified reference behind ref_driver) on the restated PwrAssemblyInput: the digest of its
    output equals the reference's committed one - reflective, one VACUUM side, two PERIODIC sides, linear source.  (The
    Python decks read sample-input/c5g7-mgxs.h5, whose MOX-4.3% total of group 7 is 0.682852, not the 0.68285 of the C++
    decks: the digests only match with it.)"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver] + PWR_CASES[test] + ["--results", res], check=True, capture_output=True)
    assert hashlib.sha512(open(res).read().encode()).hexdigest() == GOLDENS[test].strip()


STABILIZATION_ARGS = ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--cmfd", "17x17", "--cmfd-relax", "0.7",
                      "--negative-water-scatter", "--stabilize-sequence", "0.4:0,0.4:1,0.4:2", "--quiet"]


def test_ref_driver_reproduces_the_transport_stabilization_golden(tmp_path):
    """tests/test_transport_stabilization: CPULSSolver + CMFD 17 x 17 on the simple lattice with a large negative
    in-scatter in the moderator, DIAGONAL, YAMAMOTO and GLOBAL stabilisation (factor 0.4) solved one after the other
    on the same solver; the golden is the SHA-512 of the last solve (43 iterations, keff 5.67687E-01, all fluxes)"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver] + STABILIZATION_ARGS + ["--solver", "cpuls", "--results", res], check=True, capture_output=True)
    out = open(res).read()
    assert out.startswith("# Iterations: 43\nkeff:  5.67687E-01\n")
    assert hashlib.sha512(out.encode()).hexdigest() == GOLDENS["test_transport_stabilization"].strip()


def test_axial_segmentation_golden_bytes():
    # tests/test_axial_segmentation/results_true.dat: AxialExtendedInput (non-uniform 4 x 4 x 20 lattice, one layer
    # with two pins taken out), OTF_TRACKS with segmentation zones, computeEigenvalue(max_iters=30): not converged
    s, n, ref = solve("axial_extended", max_iters=30)
    assert n == ref["iterations"] == 30 and abs(s.getKeff() - ref["keff"]) < 1e-11
    assert format_harness_results(n, s.getKeff()) == GOLDENS["test_axial_segmentation"]


AXIAL = ["--model", "axial-extended", "--dims", "3", "--azim", "4", "--quiet", "--no-fluxes"]
AXIAL_SEGMENTATION_ARGS = AXIAL + ["--polar", "2", "--spacing", "0.24", "--zspacing", "0.9", "--formation", "otf-tracks",
                                   "--seg-zones", "0,1,2,3,4,5,6,7,8,9,10,20", "--max-iters", "30"]
AXIAL_INTERPOLATION_ARGS = AXIAL + ["--polar", "4", "--quad", "gl", "--spacing", "0.1", "--zspacing", "0.5", "--formation",
                                    "otf-stacks", "--seg-zones", "0,17,18,20", "--cmfd", "1x1", "--cmfd-widths",
                                    "0.05,1.26,1.26,0.05;0.05,1.26,1.26,0.05;1,2,3,4,1,2,3,4", "--cmfd-sor", "1.5",
                                    "--cmfd-relax", "0.7", "--cmfd-all-groups", "--no-knearest", "--tol", "1e-4",
                                    "--threads", "4", "--results-fsrs"]


@pytest.mark.parametrize("args,solver,test", [
    (AXIAL_SEGMENTATION_ARGS, "cpu", "test_axial_segmentation"),
    (AXIAL_INTERPOLATION_ARGS + ["--cmfd-axial-interp", "1"], "cpuls", "test_cmfd_axial_interpolation_average"),
    (AXIAL_INTERPOLATION_ARGS + ["--cmfd-axial-interp", "2"], "cpuls", "test_cmfd_axial_interpolation_centroid")])
def test_ref_driver_reproduces_the_axial_goldens(args, solver, test, tmp_path):
    """The restated AxialExtendedInput through ref_driver: OTF_TRACKS with segmentation zones; CPULSSolver on OTF_STACKS
    with a non-uniform Cmfd (Cmfd::setWidths, one CMFD group per MOC group) and its axial interpolation of the
    prolongation (FSR average / centroid): 15 iterations, keff 1.26899E+00, 2137 FSRs"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver] + args + ["--solver", solver, "--results", res], check=True, capture_output=True)
    assert open(res).read() == GOLDENS[test]


OTF_TRANSPORT_ARGS = ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "4", "--spacing", "0.4",
                      "--zspacing", "1.2", "--formation", "otf-stacks", "--vacuum-mask", "22", "--cmfd", "4x4x4",
                      "--cmfd-relax", "1.0", "--cmfd-sor", "1.5", "--tol", "1e-3", "--threads", "4", "--quiet", "--no-fluxes"]


def test_ref_driver_reproduces_the_otf_transport_golden(tmp_path):
    """tests/test_OTF_transport: the 3D lattice with VACUUM on xmax, ymin and zmin (and zmax), OTF_STACKS, CMFD 4 x 4 x 4,
    CPUSolver::setOTFTransport (segments traced while sweeping): 20 iterations, keff 5.84272E-02"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver] + OTF_TRANSPORT_ARGS + ["--otf-transport", "--solver", "cpu", "--results", res], check=True,
                   capture_output=True)
    assert open(res).read() == GOLDENS["test_OTF_transport"]


MULTISIM_CASES = {
    "test_multisim_simple": ["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--solver", "cpu"],
    "test_multisim_linear_source": ["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--solver", "cpuls"],
    # every material cell refilled with a clone of its Material before each solve
    "test_multisim_materials": ["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--clone-materials", "--solver", "cpu"],
    "test_multisim_cmfd": ["--model", "pwr-assembly", "--azim", "4", "--spacing", "0.1", "--cmfd", "17x17", "--cmfd-relax", "1.0",
                           "--cmfd-sor", "1.5", "--max-iters", "5", "--solver", "cpu"],
}


@pytest.mark.parametrize("test", sorted(MULTISIM_CASES))
def test_ref_driver_reproduces_the_multi_simulation_goldens(test, tmp_path):
    """MultiSimTestHarness (tests/testing_harness.py:398-425): the same eigenvalue solve three times on one solver
    object gives the same iteration count and k_eff three times - flat, linear source, with CMFD"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver] + MULTISIM_CASES[test] + ["--repeat", "3", "--quiet", "--results", res], check=True,
                   capture_output=True)
    assert open(res).read() == GOLDENS[test]


def test_oracle_repeated_solves_match_the_multi_simulation_golden():
    """the oracle solved three times in a row on the pin cell: tests/test_multisim_simple/results_true.dat"""
    ft, _ = load_case("pin_cell")
    s, out = OracleSolver(ft), ""
    for _ in range(3):
        n = s.computeEigenvalue(500, 1e-5, FISSION_SOURCE)
        out += "Iters: {0}\tkeff: {1:12.5E}\n".format(n, s.getKeff())
    assert out == GOLDENS["test_multisim_simple"]


def test_ref_driver_reproduces_the_num_azim_golden(tmp_path):
    """tests/test_multisim_num_azim: 4, 8 and 16 azimuthal angles one after the other on the same TrackGenerator and
    the same solver (tracks regenerated in between)"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver, "--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--azim-sequence", "4,8,16", "--quiet",
                    "--solver", "cpu", "--results", res], check=True, capture_output=True)
    assert open(res).read() == GOLDENS["test_multisim_num_azim"]


def test_ref_driver_reproduces_the_num_groups_golden(tmp_path):
    """tests/test_multisim_num_groups: the infinite medium with 1-group and then 2-group data set on the same Material,
    solved one after the other on the same tracks and solver"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver, "--model", "hom-inf", "--azim", "4", "--spacing", "0.1", "--multisim-groups", "--quiet",
                    "--solver", "cpu", "--results", res], check=True, capture_output=True)
    assert open(res).read() == GOLDENS["test_multisim_num_groups"]


def test_multisim_fixed_source_golden_from_the_oracle_and_the_reference(tmp_path):
    """tests/test_multisim_fixed_source: computeSource three times in a row on the water box (source 1.0 in group 1);
    the golden is the SHA-512 of the three 'Iters / fluxes' blocks - from the oracle on the dumped tracks and from the
    unmodified reference through ref_driver"""
    ft, ref = load_case("water_box")
    s, out = OracleSolver(ft), ""
    for fsr in ref["source_fsrs"]:
        s.setFixedSourceByFSR(fsr, 1, 1.0)
    for _ in range(3):
        n = s.computeSource(500, 1.0, 1e-5, TOTAL_SOURCE)
        out += "Iters: {0}\nfluxes:\n".format(n) + "\n".join("{0:12.6E}".format(f) for f in s.getFluxes()) + "\n"
    assert hashlib.sha512(out.encode()).hexdigest() == GOLDENS["test_multisim_fixed_source"].strip()
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    res = os.path.join(tmp_path, "res.dat")
    subprocess.run([driver, "--model", "water-box", "--azim", "4", "--spacing", "0.1", "--mode", "source", "--res", "total",
                    "--fixed-source", "1:1.0", "--repeat", "3", "--quiet", "--solver", "cpu", "--results", res], check=True,
                   capture_output=True)
    assert open(res).read() == out


def test_fission_rates_without_nu_match_the_reference(tmp_path):
    """Solver::computeFSRFissionRates with its default nu = false after TWO solves on one solver (the second solve
    re-initialises the materials): the oracle on the dumped tracks against the unmodified reference (only
    nu = true was covered, and a second solve used to wipe sigma_f on the device)"""
    import subprocess
    driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(driver):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    js = os.path.join(tmp_path, "ref.json")
    subprocess.run([driver, "--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--repeat", "2", "--fission-rates",
                    "--quiet", "--solver", "cpu", "--json", js], check=True, capture_output=True)
    ref = json.load(open(js))
    s, n, _ = solve("simple_lattice")
    s.computeEigenvalue(500, 1e-5, FISSION_SOURCE)                    # second solve
    assert n == ref["iterations"] == 187
    rates = np.array(ref["fission_rates"])
    assert rates.max() > 0 and np.count_nonzero(rates) < rates.size    # fuel only
    np.testing.assert_allclose(s.computeFSRFissionRates(nu=False), rates, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(s.computeFSRFissionRates(nu=True), ref["nu_fission_rates"], rtol=1e-10, atol=1e-14)
