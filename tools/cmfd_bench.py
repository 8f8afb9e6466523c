#!/usr/bin/env python
"""Device time of the CMFD solve (b200_cmfd_solve) on a synthetic C5G7 deck, step by step: SOR iterations and
microseconds per SOR iteration of the eigenvalue kernel's launch shape (B200_CMFD_MODE / _THREADS / _BLOCKS / _CLUSTER)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmoc_b200.solver import B200Solver
from openmoc_b200.synth import make_tracks, make_tracks_3d, cmfd_mesh, QUAD_EQUAL_ANGLE
from openmoc_b200.capi import FISSION_SOURCE

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, default=3)
ap.add_argument("--azim", type=int, default=4)
ap.add_argument("--spacing", type=float, default=0.5)
ap.add_argument("--polar", type=int, default=4)
ap.add_argument("--zspacing", type=float, default=4.0)
ap.add_argument("--axial", type=int, default=135)
ap.add_argument("--nz", type=int, default=135)
ap.add_argument("--iters", type=int, default=6)
args = ap.parse_args()
if args.dims == 3:
    ft = make_tracks_3d("c5g7-2d", args.azim, args.spacing, args.polar, args.zspacing, args.axial,
                        polar_quad=QUAD_EQUAL_ANGLE, expand=False)
    mesh = cmfd_mesh(ft, "c5g7-2d", num_z=args.nz, group_structure=[[1, 2, 3], [4, 5, 6, 7]])
else:
    ft = make_tracks("c5g7-2d", args.azim, args.spacing)
    mesh = cmfd_mesh(ft, "c5g7-2d", group_structure=[[1, 2, 3], [4, 5, 6, 7]])
s = B200Solver(ft, cmfd=mesh)
s.zeroTrackFluxes(); s.flattenFSRFluxes(0.0); s.storeFSRFluxes()
s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
tot_ms, tot_it = 0.0, 0
res = 1e-4
for i in range(args.iters):
    s.computeFSRSources(i)
    s.transportSweep()
    s.addSourceToScalarFlux()
    k, st = s.cmfdSolve(i, 1e-6 if i == 0 else 0.01 * res)
    s.normalizeFluxes()
    res = s.computeResidual(FISSION_SOURCE)
    s.storeFSRFluxes()
    print("it %d k %.8f res %.3e: cmfd %.3f ms, %d power its, %d SOR its, %.2f us/SOR, failed %d" % (
        i, k, res, st.device_ms, st.cmfd_iters, st.linear_iters_total, 1e3 * st.device_ms / max(st.linear_iters_total, 1), st.failed))
    if i > 0:
        tot_ms += st.device_ms; tot_it += st.linear_iters_total
print("MODE=%s THREADS=%s BLOCKS=%s cells=%d: %.2f us per SOR iteration, %.2f ms per solve" % (
    os.environ.get("B200_CMFD_MODE", "auto"), os.environ.get("B200_CMFD_THREADS", "auto"), os.environ.get("B200_CMFD_BLOCKS", "auto"),
    mesh.num_cells, 1e3 * tot_ms / max(tot_it, 1), tot_ms / max(args.iters - 1, 1)))
