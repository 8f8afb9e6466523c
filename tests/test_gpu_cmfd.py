"""CMFD acceleration through the Python mirror (B200Solver(tracks, cmfd=CmfdMesh)): synthetic tracks with the CMFD
surfaces of their segments, the whole CMFD-accelerated source iteration fused on the device
(b200_compute_eigenvalue with b200_cmfd_* inside).  Checker: the unmodified reference, CPUSolver + Cmfd, on the same
deck parameters (oracle/_ref/ref_driver --cmfd ... --no-knearest: the k-nearest stencils need a Geometry)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
GROUPS = [[1, 2, 3], [4, 5, 6, 7]]          # sample-input/benchmarks/c5g7/c5g7-2d.py:54


def reference(args, tmp_path, threads=4):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_driver was not built (no /root/reference at build time)")
    js = os.path.join(tmp_path, "ref.json")
    subprocess.run([DRIVER] + args + ["--solver", "cpu", "--threads", str(threads), "--no-knearest", "--quiet", "--json", js],
                   check=True, capture_output=True)
    return json.load(open(js))


@pytest.mark.parametrize("model,azim,spacing,cmfd", [("simple-lattice", 8, 0.05, "4x4"), ("c5g7-2d", 8, 0.2, "51x51")])
def test_python_cmfd_2d_matches_reference(model, azim, spacing, cmfd, tmp_path):
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks, cmfd_mesh
    ref = reference(["--model", model, "--azim", str(azim), "--spacing", str(spacing), "--cmfd", cmfd, "--max-iters", "80"], tmp_path)
    ft = make_tracks(model, azim, spacing)
    s = B200Solver(ft, cmfd=cmfd_mesh(ft, model, group_structure=GROUPS))
    s.computeEigenvalue(80)
    assert s.getNumIterations() == ref["iterations"]
    dk_pcm = abs(s.getKeff() - ref["keff"]) * 1e5
    phi, ref_phi = s.getFluxes(), np.array(ref["fluxes"])
    err = float(np.max(np.abs(phi - ref_phi) / np.abs(ref_phi)))
    print("python CMFD %s: %d iterations, dk %.3e pcm, flux %.3e" % (model, s.getNumIterations(), dk_pcm, err))
    assert dk_pcm < 1.0 and err < 1e-4                # north star
    assert dk_pcm < 1e-2 and err < 2e-5               # tracks differ at 1e-10 (analytic vs CSG ray tracing)
    # without CMFD the same deck needs many more iterations: the acceleration is real
    plain = B200Solver(ft)
    plain.computeEigenvalue(400)
    assert plain.getNumIterations() > 3 * s.getNumIterations()


def test_python_cmfd_3d_on_the_fly_tracks_matches_reference(tmp_path):
    """configs[4] shape at coarse tracks: extruded 3D C5G7, axial on-the-fly tracing on the device WITH the CMFD
    surfaces (b200_upload_otf_cmfd), CMFD 51 x 51 x 9 solved on a cooperative grid."""
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks_3d, cmfd_mesh, QUAD_EQUAL_ANGLE
    ref = reference(["--model", "c5g7-2d", "--dims", "3", "--azim", "4", "--polar", "4", "--spacing", "0.8", "--zspacing", "6",
                     "--axial", "9", "--quad", "equal-angle", "--formation", "otf-stacks", "--cmfd", "51x51x9",
                     "--max-iters", "15", "--no-fluxes"], tmp_path, threads=8)
    ft = make_tracks_3d("c5g7-2d", 4, 0.8, 4, 6.0, 9, polar_quad=QUAD_EQUAL_ANGLE, expand=False)
    s = B200Solver(ft, cmfd=cmfd_mesh(ft, "c5g7-2d", num_z=9, group_structure=GROUPS))
    s.setConvergenceThreshold(1e-30)
    s.computeEigenvalue(15)
    dk_pcm = abs(s.getKeff() - ref["keff"]) * 1e5
    print("python CMFD 3D OTF: k %.10f vs %.10f, dk %.3e pcm" % (s.getKeff(), ref["keff"], dk_pcm))
    assert s.getNumIterations() == ref["iterations"] == 15
    assert dk_pcm < 1.0                               # north star
    assert dk_pcm < 5e-2


def test_python_cmfd_on_a_device_group_equals_one_device():
    """devices=[0, 0]: tracks sharded inside the library, currents summed over peer memory, the CMFD solve replicated on
    every shard - same answer as one device, to the summation order of the tallies."""
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks, cmfd_mesh
    ft = make_tracks("c5g7-2d", 8, 0.2)
    res = []
    for devices in (None, [0, 0], [0, 0, 0]):
        s = B200Solver(ft, cmfd=cmfd_mesh(ft, "c5g7-2d", group_structure=GROUPS), devices=devices)
        s.computeEigenvalue(80)
        res.append((s.getNumIterations(), s.getKeff(), s.getFluxes()))
        s.close()
    for it, k, phi in res[1:]:
        assert it == res[0][0]
        assert abs(k - res[0][1]) * 1e5 < 1e-3
        assert np.max(np.abs(phi - res[0][2]) / res[0][2]) < 1e-6


def test_cmfd_from_a_reference_track_file():
    """Tracks, CMFD surfaces and CMFD mesh dumped by the reference's own ray tracer (tests/golden/
    simple_lattice_cmfd.b2trk, ref_driver --cmfd 4x4 --dump-tracks): the device solve reproduces the CPUSolver + Cmfd
    run the dump came from (27 iterations) - same tracks, so to rounding."""
    from openmoc_b200.solver import B200Solver, CmfdMesh
    from openmoc_b200.trackfile import read_trackfile
    golden = os.path.join(ROOT, "tests", "golden")
    ft = read_trackfile(os.path.join(golden, "simple_lattice_cmfd.b2trk"))
    ref = json.load(open(os.path.join(golden, "simple_lattice_cmfd.json")))
    s = B200Solver(ft, cmfd=CmfdMesh.from_tracks(ft))
    s.computeEigenvalue(500)
    assert s.getNumIterations() == ref["iterations"] == 27
    assert abs(s.getKeff() - ref["keff"]) * 1e5 < 1e-4
    assert np.max(np.abs(s.getFluxes() - np.array(ref["fluxes"])) / np.array(ref["fluxes"])) < 1e-8
