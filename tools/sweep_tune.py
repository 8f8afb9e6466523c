#!/usr/bin/env python
"""Times the sweep kernel alone on a synthetic deck; knobs come from the environment
(B200_GPL, B200_IPC) so launch geometries can be compared in one GPU session."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmoc_b200 import capi
from openmoc_b200.solver import B200Solver
from openmoc_b200.synth import make_tracks

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="c5g7-2d")
ap.add_argument("--azim", type=int, default=64)
ap.add_argument("--spacing", type=float, default=0.02)
ap.add_argument("--sweeps", type=int, default=10)
ap.add_argument("--precision", default="double")
ap.add_argument("--groups70", action="store_true")
ap.add_argument("--as3d", action="store_true")
ap.add_argument("--deterministic", action="store_true")
ap.add_argument("--ls", action="store_true", help="linear source (2D synthetic decks)")
args = ap.parse_args()
ft = make_tracks(args.model, num_azim=args.azim, spacing=args.spacing, groups70=args.groups70, as_3d=args.as3d,
                 linear_source=args.ls)
s = B200Solver(ft, precision={"mixed": capi.PRECISION_MIXED, "table": capi.PRECISION_TABLE}.get(args.precision, capi.PRECISION_DOUBLE),
               deterministic=args.deterministic, linear_source=args.ls)
s.zeroTrackFluxes(); s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
s.computeFSRSources(0)
for _ in range(3):
    s.transportSweep()
s.synchronize(); s.resetSweepStats()
for _ in range(args.sweeps):
    s.transportSweep()
s.synchronize()
ms, n, _ = s.getSweepStats()
W = s.integrationsPerSweep()
print(f"GPL={os.environ.get('B200_GPL','auto')} IPC={os.environ.get('B200_IPC','auto')} {args.precision} G={ft.num_groups} 3d={ft.solve_3d} det={args.deterministic}: "
      f"N_seg={ft.n_segments} sweep {ms/n:.3f} ms  {W/(ms/n*1e-3):.3e} integrations/s")
