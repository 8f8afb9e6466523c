"""Eigenmodes of the criticality problem by the implicitly restarted Arnoldi method, on top of any solver with the
Krylov helpers of `Solver` - the role of `openmoc/krylov.py:14-272` (`IRAMSolver`, after C. Josey) in the reference.

The reference's module drives `Solver::fissionTransportSweep` / `scatterTransportSweep` (`src/Solver.cpp:1282-1306`)
through `setFluxes(numpy)` / `getFluxes()`: F x = A^-1 M x with M x = transport sweep of the fission source of x and
A x = x - transport sweep of the scattering source of x, `scipy.sparse.linalg.eigs` outside and a Krylov solve
(`lgmres` by default) inside.  It cannot be imported in this image (no SWIG module; and the scipy here no longer
accepts the `tol=` keyword it passes), so the driver is restated against the same solver methods.  Works with the
solvers of the pybind11 module (`openmoc_b200.openmoc.CPUSolver / B200Solver`, built from a TrackGenerator) and with
the ctypes mirror `openmoc_b200.solver.B200Solver` (built from flattened tracks).  Vacuum boundaries only, like the
reference's (`krylov.py:23`).
"""
from __future__ import annotations

import numpy as np

FORWARD, ADJOINT = 0, 1


class IRAMSolver:
    """`openmoc.krylov.IRAMSolver`: same constructor, `computeEigenmodes` signature and result attributes."""

    def __init__(self, moc_solver):
        self._moc_solver = moc_solver
        self._precision = np.float64                       # double-precision build of the reference
        moc_solver.allowNegativeFluxes(True)               # krylov.py:61
        if hasattr(moc_solver, "getGeometry"):
            g = moc_solver.getGeometry()
            self._op_size = g.getNumFSRs() * g.getNumEnergyGroups()
        else:                                              # ctypes mirror: sizes from the flattened tracks
            self._op_size = moc_solver.tracks.n_fsrs * moc_solver.tracks.num_groups
        self._num_modes = self._interval = self._outer_tol = self._inner_tol = None
        self._a_count = self._m_count = 0
        self._eigenvalues = self._eigenvectors = None

    def computeEigenmodes(self, solver_mode=FORWARD, num_modes=5, inner_method="lgmres", outer_tol=1e-5,
                          inner_tol=1e-6, interval=10):
        """krylov.py:84-150: `num_modes` dominant eigenvalues (k_eff first) and eigenvectors (FSR scalar fluxes)."""
        import scipy.sparse.linalg as linalg
        if inner_method not in ("gmres", "lgmres", "bicgstab", "cgs"):
            raise ValueError("Unable to use %s to solve Ax=b" % inner_method)
        self._num_modes, self._inner_method = int(num_modes), inner_method
        self._outer_tol, self._inner_tol, self._interval = outer_tol, inner_tol, interval
        self._a_count = self._m_count = 0
        if hasattr(self._moc_solver, "initializeSolver"):
            self._moc_solver.initializeSolver(solver_mode)             # krylov.py:126
        elif solver_mode != FORWARD:
            raise ValueError("adjoint eigenmodes need a solver with initializeSolver(mode)")
        shape = (self._op_size, self._op_size)
        self._A_op = linalg.LinearOperator(shape, self._A, dtype=self._precision)
        self._M_op = linalg.LinearOperator(shape, self._M, dtype=self._precision)
        self._F_op = linalg.LinearOperator(shape, self._F, dtype=self._precision)
        vals, vecs = linalg.eigs(self._F_op, k=self._num_modes, tol=self._outer_tol)
        self._eigenvalues, self._eigenvectors = vals, vecs
        if hasattr(self._moc_solver, "resetMaterials"):
            self._moc_solver.resetMaterials(solver_mode)               # krylov.py:150
        return vals

    # ---- operators (krylov.py:153-272)
    def _sweep(self, flux, sweep):
        # a private copy: CPUSolver::setFluxes adopts the caller's buffer (src/CPUSolver.cpp:190) and sweeps into it
        flux = np.array(np.real(flux), dtype=self._precision, order="C", copy=True)
        self._moc_solver.setFluxes(flux)
        sweep()
        return np.array(self._moc_solver.getFluxes(self._op_size), dtype=self._precision)

    def _A(self, flux):
        """x - (transport sweep of the scattering source of x)"""
        self._a_count += 1
        old = np.real(flux).astype(self._precision)
        return old - self._sweep(old, self._moc_solver.scatterTransportSweep)

    def _M(self, flux):
        """transport sweep of the fission source of x"""
        self._m_count += 1
        return self._sweep(flux, self._moc_solver.fissionTransportSweep)

    def _F(self, flux):
        import scipy.sparse.linalg as linalg
        rhs = self._M_op * flux
        solve = getattr(linalg, self._inner_method)
        x, info = solve(self._A_op, rhs, rtol=self._inner_tol)
        if info != 0:
            raise RuntimeError("Unable to solve Ax=b with %s" % self._inner_method)
        return x
