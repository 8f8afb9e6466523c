"""sample-input/simple-lattice (tests/input_set.py: SimpleLatticeInput) with CMFD acceleration, from Python, on the
B200: geometry, lattices, Cmfd and TrackGenerator are the reference's own classes (pybind11 module), the solver is
B200Solver with the CMFD collapse / diffusion solve / prolongation on the device.

    python examples/simple_lattice_cmfd_b200.py [-a 4 -s 0.12] [--solver cpu|b200] [--no-cmfd]
"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openmoc_b200.openmoc as openmoc

which = sys.argv[sys.argv.index("--solver") + 1] if "--solver" in sys.argv else "b200"
use_cmfd = "--no-cmfd" not in sys.argv
opts = openmoc.options.Options()
if "-s" not in sys.argv and "--azim-spacing" not in sys.argv:
    opts.azim_spacing = 0.12                                 # tests/test_forward_simple_lattice

openmoc.log.set_log_level('NORMAL')
materials = openmoc.materialize.load_c5g7()

xmin, xmax = openmoc.XPlane(x=-2.0, name='xmin'), openmoc.XPlane(x=2.0, name='xmax')
ymin, ymax = openmoc.YPlane(y=-2.0, name='ymin'), openmoc.YPlane(y=2.0, name='ymax')
for s in (xmin, xmax, ymin, ymax):
    s.setBoundaryType(openmoc.REFLECTIVE)

pins = []
for radius in (0.4, 0.3, 0.2):
    cylinder = openmoc.ZCylinder(x=0.0, y=0.0, radius=radius)
    fuel = openmoc.Cell()
    fuel.setNumRings(3)
    fuel.setNumSectors(8)
    fuel.setFill(materials['UO2'])
    fuel.addSurface(halfspace=-1, surface=cylinder)
    moderator = openmoc.Cell()
    moderator.setNumSectors(8)
    moderator.setFill(materials['Water'])
    moderator.addSurface(halfspace=+1, surface=cylinder)
    pin = openmoc.Universe()
    pin.addCell(fuel)
    pin.addCell(moderator)
    pins.append(pin)

lattice_cell, root_cell = openmoc.Cell(), openmoc.Cell()
root_cell.addSurface(halfspace=+1, surface=xmin)
root_cell.addSurface(halfspace=-1, surface=xmax)
root_cell.addSurface(halfspace=+1, surface=ymin)
root_cell.addSurface(halfspace=-1, surface=ymax)
assembly, root = openmoc.Universe(), openmoc.Universe()
assembly.addCell(lattice_cell)
root.addCell(root_cell)

lattice = openmoc.Lattice()
lattice.setWidth(width_x=1.0, width_y=1.0)
lattice.setUniverses([[[pins[0], pins[1]], [pins[0], pins[2]]]])
lattice_cell.setFill(lattice)
core = openmoc.Lattice()
core.setWidth(width_x=2.0, width_y=2.0)
core.setUniverses([[[assembly, assembly], [assembly, assembly]]])
root_cell.setFill(core)

geometry = openmoc.Geometry()
geometry.setRootUniverse(root)
if use_cmfd:
    cmfd = openmoc.Cmfd()
    cmfd.setSORRelaxationFactor(1.5)
    cmfd.setLatticeStructure(4, 4)
    cmfd.setGroupStructure([[1, 2, 3], [4, 5, 6, 7]])
    cmfd.setKNearest(3)
    geometry.setCmfd(cmfd)
geometry.initializeFlatSourceRegions()

track_generator = openmoc.TrackGenerator(geometry, opts.num_azim, opts.azim_spacing)
track_generator.setNumThreads(1)
track_generator.generateTracks()

solver = {"cpu": openmoc.CPUSolver, "b200": openmoc.B200Solver}[which](track_generator)
solver.setNumThreads(opts.num_omp_threads)
solver.setConvergenceThreshold(opts.tolerance)
solver.computeEigenvalue(opts.max_iters)
n = geometry.getNumFSRs() * geometry.getNumEnergyGroups()
fluxes = solver.getFluxes(n)
# the string tests/testing_harness.py:158-207 hashes for the regression goldens
text = "# Iterations: %d\n" % solver.getNumIterations() + "keff: %12.5E\n" % solver.getKeff() + "fluxes:\n" \
       + "".join("%12.6E\n" % v for v in fluxes)
print("RESULT solver=%s cmfd=%s iterations=%d keff=%.10f sha512=%s" % (
    which, use_cmfd, solver.getNumIterations(), solver.getKeff(), hashlib.sha512(text.encode()).hexdigest()))
