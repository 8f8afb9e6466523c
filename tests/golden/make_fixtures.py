#!/usr/bin/env python
"""Regenerate tests/golden/ from the UNMODIFIED reference.

Runs oracle/_ref/ref_driver (the reference's CPUSolver built by oracle/Makefile
from /root/reference/src) on the reference's own regression decks and stores

  <case>.b2trk   flattened tracks (input of the hot path)
  <case>.json    reference results at full precision (k_eff, iterations, phi)
  ref_goldens.json  the reference's own committed golden files, verbatim
                    (tests/test_forward_*/results_true.dat), used to pin the
                    C oracle without needing /root/reference at test time.

Needs /root/reference; run in the CPU container:  python tests/golden/make_fixtures.py
"""
import json, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
REF = "/root/reference"

CASES = {
    # name: ref_driver args                                            (reference test it mirrors)
    "pin_cell": ["--model", "pin-cell", "--azim", "4", "--spacing", "0.1"],          # test_forward_pin_cell
    "simple_lattice": ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12"],  # test_forward_simple_lattice
    "lattice3d_70g": ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2",
                      "--spacing", "2.1", "--zspacing", "2.8", "--groups70", "--tol", "5e-3"],  # test_forward_3D_lattice_70g
    "lattice3d_7g": ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2",
                     "--spacing", "0.24", "--zspacing", "0.9"],                       # test_forward_3D_lattice
    "hom_inf": ["--model", "hom-inf", "--azim", "4", "--spacing", "0.1"],            # test_forward_hom_inf_medium
    # C3 deck, coarse tracks.  Default (TY) quadrature: that is what c5g7-2d.py really runs, its
    # EqualAnglePolarQuad is discarded by generateTracks for want of setNumAzimAngles.
    "c5g7_2d_coarse": ["--model", "c5g7-2d", "--azim", "4", "--spacing", "0.5", "--polar", "6",
                       "--max-iters", "40", "--no-fluxes"],
    "pin_cell_70g": ["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--groups70", "--res", "flux"],  # test_forward_pin_cell_70g
    "gradient_1d": ["--model", "gradient-1d", "--azim", "4", "--spacing", "0.1"],    # test_1d_gradient (VACUUM in x)
    "gradient_2d": ["--model", "gradient-2d", "--azim", "4", "--spacing", "0.1"],    # test_2d_gradient (VACUUM xmin, ymax)
    # fixed-source decks: the track file is shared by test_compute_flux and test_compute_source
    "water_box": ["--model", "water-box", "--azim", "4", "--spacing", "0.1", "--mode", "flux",
                  "--fixed-source", "1:1.0,2:0.5,3:0.25", "--res", "flux"],
    # linear source (CPULSSolver): the track files carry the centroid-relative segment starting
    # points and the quadrature factors the LinearExpansionGenerator pre-pass needs
    "simple_lattice_ls": ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--solver", "cpuls"],
    "lattice3d_ls_70g": ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2",
                         "--spacing", "0.6", "--zspacing", "2.8", "--groups70", "--tol", "5e-3",
                         "--solver", "cpuls"],                                  # test_forward_3D_lattice_linear_70g
    # tests/test_fixed_linear_source: the water box with a flat fixed source and its x, y, z moments, negative
    # fluxes allowed, CPULSSolver::computeFlux
    "water_box_ls": ["--model", "water-box", "--azim", "4", "--spacing", "0.1", "--solver", "cpuls", "--mode", "flux",
                     "--res", "flux", "--allow-negative",
                     "--fixed-source", "1:1.0,2:0.5,3:0.25,4:1.0,5:0.5,6:0.25,7:1.0",
                     "--fixed-moments", "1:0.01:0.1:0.2,2:-0.1:0:-0.04,3:0.02:0:0"],
    # linear source with GLOBAL transport stabilisation, moments included (CPULSSolver::computeStabilizingFlux /
    # stabilizeFlux, src/CPULSSolver.cpp:888-1052); same tracks as simple_lattice_ls: only the results are kept
    "simple_lattice_ls_stab": ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--solver", "cpuls",
                               "--stabilize", "0.5:2"],
    # CMFD from a track file: the dump carries the mesh (cmfd_* chunks, fsr_cmfd_cell) next to the surfaces of the segments
    "simple_lattice_cmfd": ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12", "--cmfd", "4x4", "--no-knearest"],
    # tests/test_2d_gradient_linear_source: vacuum on xmin / ymax with the linear source
    "gradient_2d_ls": ["--model", "gradient-2d", "--azim", "4", "--spacing", "0.1", "--solver", "cpuls"],
    # tests/test_split_segments: Solver::setMaxOpticalLength(0.5) splits the segments of the pin cell (196 -> 1560);
    # the dump is taken after the split, so the oracle sweeps what CPUSolver swept
    "pin_cell_split": ["--model", "pin-cell", "--azim", "4", "--spacing", "0.1", "--max-tau", "0.5", "--no-fluxes"],
    # tests/test_axial_segmentation: AxialExtendedInput (non-uniform lattice, axially heterogeneous), OTF_TRACKS with
    # segmentation zones, 30 iterations without convergence
    "axial_extended": ["--model", "axial-extended", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
                       "--zspacing", "0.9", "--formation", "otf-tracks", "--seg-zones", "0,1,2,3,4,5,6,7,8,9,10,20",
                       "--max-iters", "30", "--no-fluxes"],
    "lattice3d_ls_7g": ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2",
                        "--spacing", "0.24", "--zspacing", "0.9", "--solver", "cpuls"],  # test_forward_3D_lattice_linear
}

def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    only = sys.argv[1:]                      # optional: regenerate the named fixtures only
    for name, args in CASES.items():
        if only and name not in only:
            continue
        trk = os.path.join(HERE, name + ".b2trk")
        js = os.path.join(HERE, name + ".json")
        subprocess.check_call([DRIVER] + args + ["--quiet", "--dump-tracks", trk, "--json", js])
        d = json.load(open(js))
        if name.startswith(("c5g7", "lattice3d_70g", "lattice3d_ls", "axial_extended")):
            d.pop("fluxes", None)    # keep the fixture small; phi is compared through the oracle
        d.pop("sweep_time_s", None); d.pop("total_time_s", None)
        if name.endswith("_stab"):
            os.remove(trk)                     # identical to the fixture it is named after
        json.dump(d, open(js, "w"))
        print(name, d.get("iterations"), d.get("keff"), d["n_tracks"], d["n_segments"], d["n_fsrs"])
    gold = {}
    for t in ("test_forward_pin_cell", "test_forward_simple_lattice", "test_forward_3D_lattice_70g",
              "test_forward_3D_lattice", "test_forward_hom_inf_medium",
              "test_forward_3D_lattice_linear", "test_forward_3D_lattice_linear_70g",
              "test_compute_flux", "test_compute_source", "test_fixed_linear_source",
              "test_forward_pin_cell_70g", "test_1d_gradient", "test_2d_gradient", "test_adjoint_pin_cell", "test_adjoint_simple_lattice", "test_adjoint_hom_inf_medium",
              "test_forward_3D_lattice_CMFD", "test_2d_gradient_linear_source", "test_split_segments",
              "test_split_segments_cmfd", "test_forward_3D_lattice_symmetry", "test_cmfd_pwr_assembly",
              "test_cmfd_vacuum_boundary", "test_cmfd_periodic_boundaries", "test_cmfd_linear_source", "test_transport_stabilization", "test_axial_segmentation",
              "test_cmfd_axial_interpolation_average", "test_cmfd_axial_interpolation_centroid", "test_OTF_transport", "test_cmfd_restart", "test_multisim_simple",
              "test_multisim_linear_source", "test_multisim_cmfd", "test_multisim_num_azim", "test_multisim_materials", "test_multisim_num_groups", "test_multisim_fixed_source"):
        gold[t] = open(os.path.join(REF, "tests", t, "results_true.dat")).read()
    json.dump(gold, open(os.path.join(HERE, "ref_goldens.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
