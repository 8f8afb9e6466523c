"""The pybind11 module (openmoc_b200/cpp/pymodule.cpp + openmoc_b200/openmoc.py): an OpenMOC input script builds
its geometry and TrackGenerator with the reference's own classes and hands them to `B200Solver(track_generator)`,
as it would to the reference's CPUSolver / GPUSolver (openmoc/cuda/openmoc_cuda.i:52-56).  The example is the
reference's sample-input/pin-cell deck in the shape of tests/test_forward_pin_cell (golden: results_true.dat)."""
import glob
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

EXAMPLE = os.path.join(ROOT, "examples", "pin_cell_b200.py")
BUILT = bool(glob.glob(os.path.join(ROOT, "openmoc_b200", "_openmoc_b200*.so")))


def run_example(solver):
    if not BUILT:
        pytest.skip("openmoc_b200/_openmoc_b200 was not built (needs the reference headers: make -C oracle ref)")
    out = subprocess.run([sys.executable, EXAMPLE, "--solver", solver, "-a", "4", "-s", "0.1", "-t", "1"],
                         check=True, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("RESULT ")][-1]
    fields = dict(kv.split("=", 1) for kv in line.split()[1:4])
    fluxes = line.split("fluxes=", 1)[1].split()
    return int(fields["iterations"]), float(fields["keff"]), fluxes


def golden():
    text = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_goldens.json")))["test_forward_pin_cell"]
    lines = text.splitlines()
    assert lines[0].startswith("# Iterations:") and lines[1].startswith("keff:") and lines[2] == "fluxes:"
    return int(lines[0].split(":")[1]), float(lines[1].split(":")[1]), [l.strip() for l in lines[3:]]


def test_reference_cpusolver_through_the_module_reproduces_the_golden():
    """CPU only: the module loads, the script API works, CPUSolver gives tests/test_forward_pin_cell's result."""
    it, k, fluxes = run_example("cpu")
    g_it, g_k, g_fluxes = golden()
    assert it == g_it
    assert "%12.5E" % k == "%12.5E" % g_k
    assert fluxes == g_fluxes


@pytest.mark.gpu
def test_b200solver_from_a_track_generator_in_python():
    it, k, fluxes = run_example("b200")
    g_it, g_k, g_fluxes = golden()
    assert it == g_it
    assert "%12.5E" % k == "%12.5E" % g_k
    assert fluxes == g_fluxes


@pytest.mark.gpu
def test_b200lssolver_from_a_track_generator_in_python():
    it, k, _ = run_example("b200ls")
    assert it > 10 and abs(k - 1.0467) < 5e-3        # linear source on the same deck: close to the flat result


# ---------------------------------------------------------------- lattices, rings / sectors, Cmfd through the module
LATTICE = os.path.join(ROOT, "examples", "simple_lattice_cmfd_b200.py")


def run_lattice(solver, cmfd):
    if not BUILT:
        pytest.skip("openmoc_b200/_openmoc_b200 was not built (needs the reference headers: make -C oracle ref)")
    cmd = [sys.executable, LATTICE, "--solver", solver, "-a", "4", "-s", "0.12", "-t", "1"] + ([] if cmfd else ["--no-cmfd"])
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("RESULT ")][-1]
    f = dict(kv.split("=", 1) for kv in line.split()[1:])
    return int(f["iterations"]), float(f["keff"]), f["sha512"]


def test_simple_lattice_golden_hash_through_the_module():
    """CPU only: Lattice.setUniverses, rings and sectors through the module; CPUSolver reproduces the SHA-512 of
    tests/test_forward_simple_lattice/results_true.dat (187 iterations)."""
    it, _, sha = run_lattice("cpu", cmfd=False)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_goldens.json")))["test_forward_simple_lattice"]
    assert it == 187 and sha == golden.strip()


@pytest.mark.gpu
def test_b200solver_with_cmfd_from_python_matches_cpusolver():
    """The Cmfd object is built in Python, B200Solver runs its work on the device: same iterations and k_eff as
    CPUSolver + Cmfd from the same script; without CMFD the golden hash comes from the GPU."""
    it_c, k_c, _ = run_lattice("cpu", cmfd=True)
    it_g, k_g, _ = run_lattice("b200", cmfd=True)
    assert it_g == it_c and abs(k_g - k_c) * 1e5 < 1e-3
    it, _, sha = run_lattice("b200", cmfd=False)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_goldens.json")))["test_forward_simple_lattice"]
    assert it == 187 and sha == golden.strip()
