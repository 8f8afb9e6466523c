#!/usr/bin/env python
"""Time-to-solution of small decks with and without CUDA-graph replay of the fused iteration."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmoc_b200.solver import B200Solver
from openmoc_b200.synth import make_tracks

for model, azim, spacing in (("pin-cell", 4, 0.1), ("simple-lattice", 32, 0.05), ("simple-lattice", 128, 0.01)):
    ft = make_tracks(model, num_azim=azim, spacing=spacing)
    out = {}
    for g in ("0", "1"):
        os.environ["B200_GRAPH"] = g
        s = B200Solver(ft)
        s.setConvergenceThreshold(1e-5)
        s.computeEigenvalue(1000)            # warm-up (module load, first launches)
        s.synchronize()
        t0 = time.perf_counter()
        s.computeEigenvalue(1000)
        s.synchronize()
        dt = time.perf_counter() - t0
        out[g] = (s.getNumIterations(), s.getKeff(), s.getFluxes(), dt)
        print("%s azim %d spacing %g: N_seg=%d graph=%s iterations=%d k=%.10f  %.3f ms per iteration (%.1f ms total)"
              % (model, azim, spacing, ft.n_segments, g, out[g][0], out[g][1], 1e3 * dt / out[g][0], 1e3 * dt))
    a, b = out["0"], out["1"]
    print("   same iterations: %s, |dk| = %.2e, max rel dphi = %.2e, speed-up %.2fx"
          % (a[0] == b[0], abs(a[1] - b[1]), np.max(np.abs(a[2] - b[2]) / np.abs(a[2])), a[3] / b[3]))
