#!/usr/bin/env python
"""Times the 3D sweep kernel alone on the synthetic 3D C5G7 deck (device-side axial tracing);
knobs (B200_ORDER, B200_GPL, B200_CTA, ...) come from the environment."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmoc_b200.solver import B200Solver
from openmoc_b200.synth import make_tracks_3d, QUAD_EQUAL_ANGLE

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="c5g7-2d")
ap.add_argument("--azim", type=int, default=16)
ap.add_argument("--spacing", type=float, default=0.1)
ap.add_argument("--polar", type=int, default=8)
ap.add_argument("--zspacing", type=float, default=1.0)
ap.add_argument("--axial", type=int, default=135)
ap.add_argument("--sweeps", type=int, default=5)
ap.add_argument("--groups70", action="store_true")
args = ap.parse_args()
t0 = time.time()
ft = make_tracks_3d(args.model, args.azim, args.spacing, args.polar, args.zspacing, args.axial,
                    polar_quad=QUAD_EQUAL_ANGLE, expand=False, groups70=args.groups70)
t1 = time.time()
s = B200Solver(ft)
t2 = time.time()
s.zeroTrackFluxes(); s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
s.computeFSRSources(0)
for _ in range(2):
    s.transportSweep()
s.synchronize(); s.resetSweepStats()
for _ in range(args.sweeps):
    s.transportSweep()
s.synchronize()
ms, n, _ = s.getSweepStats()
W = s.integrationsPerSweep()
print(f"ORDER={os.environ.get('B200_ORDER','default')} GPL={os.environ.get('B200_GPL','auto')} CTA={os.environ.get('B200_CTA','auto')} "
      f"G={ft.num_groups}: tracks={ft.n_tracks} N_seg={s.num_segments} n_fsrs={ft.n_fsrs} gen {t1-t0:.2f}s setup {t2-t1:.2f}s "
      f"sweep {ms/n:.3f} ms  {W/(ms/n*1e-3):.3e} integrations/s", flush=True)
