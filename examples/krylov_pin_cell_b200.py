"""Eigenmodes of a pin cell with vacuum boundaries by the implicitly restarted Arnoldi method - the deck of the
reference's tests/test_krylov_forward with `openmoc.krylov.IRAMSolver` replaced by `openmoc_b200.krylov.IRAMSolver`
(same constructor and `computeEigenmodes`; the reference's module needs SWIG and an older scipy).

    python examples/krylov_pin_cell_b200.py [--solver cpu|b200] [--modes 2]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openmoc_b200.openmoc as openmoc
from openmoc_b200.krylov import IRAMSolver

which = sys.argv[sys.argv.index("--solver") + 1] if "--solver" in sys.argv else "b200"
modes = int(sys.argv[sys.argv.index("--modes") + 1]) if "--modes" in sys.argv else 2

openmoc.log.set_log_level('WARNING')
materials = openmoc.materialize.load_c5g7()
zcylinder = openmoc.ZCylinder(x=0.0, y=0.0, radius=1.0, name='pin')
xmin = openmoc.XPlane(x=-2.0, name='xmin')
xmax = openmoc.XPlane(x=+2.0, name='xmax')
ymin = openmoc.YPlane(y=-2.0, name='ymin')
ymax = openmoc.YPlane(y=+2.0, name='ymax')
for s in (xmin, xmax, ymin, ymax):
    s.setBoundaryType(openmoc.VACUUM)
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
track_generator = openmoc.TrackGenerator(geometry, 4, 0.1)
track_generator.setNumThreads(1)
track_generator.generateTracks()

moc = {"cpu": openmoc.CPUSolver, "b200": openmoc.B200Solver}[which](track_generator)
moc.setNumThreads(1)
moc.setConvergenceThreshold(1e-5)
solver = IRAMSolver(moc)
solver.computeEigenmodes(num_modes=modes, solver_mode=openmoc.FORWARD)

# the same operators applied to unit vectors: the dense eigenvalues the Arnoldi iteration must find
n = geometry.getNumFSRs() * geometry.getNumEnergyGroups()
eye = np.eye(n)
A = np.column_stack([solver._A(eye[:, j]) for j in range(n)])
M = np.column_stack([solver._M(eye[:, j]) for j in range(n)])
dense = np.linalg.eigvals(np.linalg.solve(A, M))
dense = dense[np.argsort(-np.abs(dense))][:modes]
print("RESULT solver=%s eigenvalues=%s dense=%s a_sweeps=%d m_sweeps=%d" % (
    which, ",".join("%.12e" % v.real for v in solver._eigenvalues), ",".join("%.12e" % v.real for v in dense),
    solver._a_count, solver._m_count))
