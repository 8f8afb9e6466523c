"""sample-input/pin-cell/pin-cell.py + geometry.py of the reference on the B200: the import line, the materials
line and the solver class are the only edits (no plotting: matplotlib is not in this image).

    python examples/pin_cell_b200.py [-a 4 -s 0.1] [--solver cpu|b200|b200ls]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openmoc_b200.openmoc as openmoc

which = "b200"
if "--solver" in sys.argv:
    which = sys.argv[sys.argv.index("--solver") + 1]
opts = openmoc.options.Options()

openmoc.log.set_log_level('NORMAL')
materials = openmoc.materialize.load_c5g7()

zcylinder = openmoc.ZCylinder(x=0.0, y=0.0, radius=1.0, name='pin')
xmin = openmoc.XPlane(x=-2.0, name='xmin')
ymin = openmoc.YPlane(y=-2.0, name='ymin')
xmax = openmoc.XPlane(x=2.0, name='xmax')
ymax = openmoc.YPlane(y=2.0, name='ymax')
for s in (xmin, ymin, xmax, ymax):
    s.setBoundaryType(openmoc.REFLECTIVE)

fuel = openmoc.Cell(name='fuel')
fuel.setFill(materials['UO2'])
fuel.addSurface(halfspace=-1, surface=zcylinder)
moderator = openmoc.Cell(name='moderator')
moderator.setFill(materials['Water'])
moderator.addSurface(halfspace=+1, surface=zcylinder)
moderator.addSurface(halfspace=+1, surface=xmin)
moderator.addSurface(halfspace=-1, surface=xmax)
moderator.addSurface(halfspace=+1, surface=ymin)
moderator.addSurface(halfspace=-1, surface=ymax)

root_universe = openmoc.Universe(name='root universe')
root_universe.addCell(fuel)
root_universe.addCell(moderator)
geometry = openmoc.Geometry()
geometry.setRootUniverse(root_universe)
geometry.initializeFlatSourceRegions()

track_generator = openmoc.TrackGenerator(geometry, opts.num_azim, opts.azim_spacing)
track_generator.setNumThreads(1)
track_generator.generateTracks()

solver = {"cpu": openmoc.CPUSolver, "b200": openmoc.B200Solver, "b200ls": openmoc.B200LSSolver}[which](track_generator)
solver.setNumThreads(opts.num_omp_threads)
solver.setConvergenceThreshold(opts.tolerance)
solver.computeEigenvalue(opts.max_iters)
solver.printTimerReport()
n = geometry.getNumFSRs() * geometry.getNumEnergyGroups()
print("RESULT solver=%s iterations=%d keff=%.10f fluxes=%s" % (
    which, solver.getNumIterations(), solver.getKeff(), " ".join("%.6E" % v for v in solver.getFluxes(n))))
