"""ctypes wrapper around oracle/libmoc_oracle.so (the plain-C CPU restatement).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under openmoc_b200/
may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmoc_oracle.so")

SCALAR_FLUX, FISSION_SOURCE, TOTAL_SOURCE = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "moc_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "oracle", "CC=gcc"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.moc_oracle_create.restype = vp
        L.moc_oracle_create.argtypes = [i32, i32, i32, i32, i64, i64, i64, i32] + [vp] * 21
        for name in ("destroy", "zero_track_fluxes", "store_fsr_fluxes", "compute_fsr_fission_sources",
                     "compute_fsr_scatter_sources", "transport_sweep", "add_source_to_scalar_flux",
                     "compute_keff", "compute_stabilizing_flux", "stabilize_flux"):
            f = getattr(L, "moc_oracle_" + name); f.restype = None; f.argtypes = [vp]
        L.moc_oracle_flatten_fsr_fluxes.restype = None
        L.moc_oracle_flatten_fsr_fluxes.argtypes = [vp, dbl]
        L.moc_oracle_normalize_fluxes.restype = dbl
        L.moc_oracle_normalize_fluxes.argtypes = [vp]
        L.moc_oracle_compute_fsr_sources.restype = None
        L.moc_oracle_compute_fsr_sources.argtypes = [vp, i32]
        L.moc_oracle_compute_residual.restype = dbl
        L.moc_oracle_compute_residual.argtypes = [vp, i32]
        L.moc_oracle_compute_eigenvalue.restype = i32
        L.moc_oracle_compute_eigenvalue.argtypes = [vp, i32, dbl, i32]
        L.moc_oracle_compute_flux.restype = i32
        L.moc_oracle_compute_flux.argtypes = [vp, i32, dbl, i32]
        L.moc_oracle_compute_source.restype = i32
        L.moc_oracle_compute_source.argtypes = [vp, i32, dbl, dbl, i32]
        L.moc_oracle_get_keff.restype = dbl
        L.moc_oracle_get_keff.argtypes = [vp]
        L.moc_oracle_set_keff.restype = None
        L.moc_oracle_set_keff.argtypes = [vp, dbl]
        for name in ("get_fluxes", "set_fluxes", "get_sources", "set_sources",
                     "get_start_fluxes", "set_start_fluxes"):
            f = getattr(L, "moc_oracle_" + name); f.restype = None; f.argtypes = [vp, vp]
        L.moc_oracle_set_fixed_source.restype = None
        L.moc_oracle_set_fixed_source.argtypes = [vp, i64, i32, dbl]
        L.moc_oracle_set_fixed_source_moments.restype = None
        L.moc_oracle_set_fixed_source_moments.argtypes = [vp, i64, i32, dbl, dbl, dbl]
        L.moc_oracle_allow_negative_fluxes.restype = None
        L.moc_oracle_allow_negative_fluxes.argtypes = [vp, i32]
        L.moc_oracle_stabilize_transport.restype = None
        L.moc_oracle_stabilize_transport.argtypes = [vp, dbl, i32]
        L.moc_oracle_compute_fission_rates.restype = None
        L.moc_oracle_compute_fission_rates.argtypes = [vp, vp, i32]
        L.moc_oracle_set_num_threads.restype = None
        L.moc_oracle_set_num_threads.argtypes = [vp, i32]
        L.moc_oracle_set_keff_from_neutron_balance.restype = None
        L.moc_oracle_set_keff_from_neutron_balance.argtypes = [vp, i32]
        L.moc_oracle_sweep_seconds.restype = dbl
        L.moc_oracle_sweep_seconds.argtypes = [vp, i32]
        L.moc_oracle_expF1.restype = dbl
        L.moc_oracle_expF1.argtypes = [dbl]
        L.moc_oracle_enable_linear_source.restype = i32
        L.moc_oracle_enable_linear_source.argtypes = [vp] * 8
        L.moc_oracle_get_flux_moments.restype = None
        L.moc_oracle_get_flux_moments.argtypes = [vp, vp]
        L.moc_oracle_get_linear_source_tables.restype = None
        L.moc_oracle_get_linear_source_tables.argtypes = [vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleSolver:
    """Method names follow the reference Solver (src/Solver.h)."""

    def __init__(self, ft, linear_source=False):
        L = lib()
        a = ft.arrays
        c = lambda k, dt: np.ascontiguousarray(a[k], dtype=dt)
        self._keep = [
            c("seg_length", "f8"), c("seg_fsr", "i4"), c("trk_seg_offset", "i8"),
            c("trk_azim", "i4"), c("trk_polar", "i4"), c("trk_next_fwd", "i8"),
            c("trk_next_bwd", "i8"), c("trk_flags", "u1"), c("trk_bc_fwd", "u1"),
            c("trk_bc_bwd", "u1"), c("quad_weight", "f8"), c("quad_sin_theta", "f8"),
            c("fsr_volume", "f8"), c("fsr_mat", "i4"), c("mat_sigma_t", "f8"),
            c("mat_sigma_s", "f8"), c("mat_fiss_matrix", "f8"), c("mat_nu_sigma_f", "f8"),
            c("mat_sigma_f", "f8"), c("mat_chi", "f8"), c("mat_fissionable", "u1")]
        self.ft = ft
        self.G = ft.num_groups
        self.F = ft.fluxes_per_track
        self.h = L.moc_oracle_create(ft.num_groups, ft.num_azim, ft.num_polar, ft.solve_3d,
                                     ft.n_tracks, ft.n_segments, ft.n_fsrs, ft.n_materials,
                                     *[_p(x) for x in self._keep])
        self.num_iterations = 0
        self.linear_source = bool(linear_source)
        if linear_source:
            # CPULSSolver: needs the LS chunks of the track file (dumped after a CPULSSolver run)
            ls = [c("seg_start", "f8"), c("trk_phi", "f8"), c("trk_theta", "f8"), c("quad_azim_spacing", "f8"),
                  c("quad_azim_weight", "f8"), c("quad_polar_spacing", "f8"), c("quad_polar_weight", "f8")]
            assert ls[0].size == 3 * ft.n_segments, "track file has no segment starting points"
            self.num_flat_fsrs = L.moc_oracle_enable_linear_source(self.h, *[_p(x) for x in ls])

    def getFluxMoments(self):
        """[n_fsrs][3][G] as CPULSSolver stores them (src/CPULSSolver.h:22-26)"""
        out = np.empty(self.ft.n_fsrs * 3 * self.G)
        lib().moc_oracle_get_flux_moments(self.h, _p(out))
        return out

    def getLinearSourceTables(self):
        nc = 6 if self.ft.solve_3d else 3
        a, b = np.empty(self.ft.n_fsrs * nc), np.empty(self.ft.n_fsrs * self.G * nc)
        lib().moc_oracle_get_linear_source_tables(self.h, _p(a), _p(b))
        return a, b

    def __del__(self):
        if getattr(self, "h", None):
            lib().moc_oracle_destroy(self.h)
            self.h = None

    # --- Solver virtuals ---
    def zeroTrackFluxes(self): lib().moc_oracle_zero_track_fluxes(self.h)
    def flattenFSRFluxes(self, v): lib().moc_oracle_flatten_fsr_fluxes(self.h, float(v))
    def storeFSRFluxes(self): lib().moc_oracle_store_fsr_fluxes(self.h)
    def normalizeFluxes(self): return lib().moc_oracle_normalize_fluxes(self.h)
    def computeFSRSources(self, iteration): lib().moc_oracle_compute_fsr_sources(self.h, iteration)
    def computeFSRFissionSources(self): lib().moc_oracle_compute_fsr_fission_sources(self.h)
    def computeFSRScatterSources(self): lib().moc_oracle_compute_fsr_scatter_sources(self.h)
    def transportSweep(self): lib().moc_oracle_transport_sweep(self.h)
    def addSourceToScalarFlux(self): lib().moc_oracle_add_source_to_scalar_flux(self.h)
    def computeKeff(self): lib().moc_oracle_compute_keff(self.h)
    def computeResidual(self, res_type): return lib().moc_oracle_compute_residual(self.h, res_type)
    def computeStabilizingFlux(self): lib().moc_oracle_compute_stabilizing_flux(self.h)
    def stabilizeFlux(self): lib().moc_oracle_stabilize_flux(self.h)

    # --- drivers ---
    def computeEigenvalue(self, max_iters=1000, tol=1e-5, res_type=FISSION_SOURCE):
        self.num_iterations = lib().moc_oracle_compute_eigenvalue(self.h, max_iters, tol, res_type)
        return self.num_iterations

    def computeFlux(self, max_iters=1000, tol=1e-5, only_fixed_source=True):
        self.num_iterations = lib().moc_oracle_compute_flux(self.h, max_iters, tol, int(only_fixed_source))
        return self.num_iterations

    def computeSource(self, max_iters=1000, k_eff=1.0, tol=1e-5, res_type=TOTAL_SOURCE):
        self.num_iterations = lib().moc_oracle_compute_source(self.h, max_iters, k_eff, tol, res_type)
        return self.num_iterations

    # --- state ---
    def getKeff(self): return lib().moc_oracle_get_keff(self.h)
    def setKeff(self, k): lib().moc_oracle_set_keff(self.h, float(k))
    def getNumIterations(self): return self.num_iterations

    def getFluxes(self):
        out = np.empty(self.ft.n_fsrs * self.G); lib().moc_oracle_get_fluxes(self.h, _p(out)); return out

    def setFluxes(self, x):
        x = np.ascontiguousarray(x, dtype="f8"); lib().moc_oracle_set_fluxes(self.h, _p(x))

    def getSources(self):
        out = np.empty(self.ft.n_fsrs * self.G); lib().moc_oracle_get_sources(self.h, _p(out)); return out

    def setSources(self, x):
        x = np.ascontiguousarray(x, dtype="f8"); lib().moc_oracle_set_sources(self.h, _p(x))

    def getStartFluxes(self):
        out = np.empty(self.ft.n_tracks * 2 * self.F, dtype="f4")
        lib().moc_oracle_get_start_fluxes(self.h, _p(out)); return out

    def setStartFluxes(self, x):
        x = np.ascontiguousarray(x, dtype="f4"); lib().moc_oracle_set_start_fluxes(self.h, _p(x))

    def setFixedSourceByFSR(self, fsr_id, group, source):
        """group is 1-based like the reference (src/Solver.cpp:479-497)."""
        lib().moc_oracle_set_fixed_source(self.h, fsr_id, group - 1, float(source))

    def setFixedSourceMomentsByFSR(self, fsr_id, group, src_x, src_y, src_z):
        """CPULSSolver::setFixedSourceMomentByFSR, 1-based group."""
        lib().moc_oracle_set_fixed_source_moments(self.h, fsr_id, group - 1, float(src_x), float(src_y), float(src_z))

    def allowNegativeFluxes(self, allowed):
        lib().moc_oracle_allow_negative_fluxes(self.h, int(bool(allowed)))

    def stabilizeTransport(self, factor, stab_type=0):
        lib().moc_oracle_stabilize_transport(self.h, float(factor), stab_type)

    def computeFSRFissionRates(self, nu=False):
        out = np.empty(self.ft.n_fsrs); lib().moc_oracle_compute_fission_rates(self.h, _p(out), int(nu)); return out

    def setNumThreads(self, n): lib().moc_oracle_set_num_threads(self.h, int(n))
    def setKeffFromNeutronBalance(self): lib().moc_oracle_set_keff_from_neutron_balance(self.h, 1)
    def sweepSeconds(self, reset=False): return lib().moc_oracle_sweep_seconds(self.h, int(reset))


def format_harness_results(num_iters, keff, fluxes=None) -> str:
    """The string tests/testing_harness.py:158-207 builds (and hashes)."""
    s = "# Iterations: {0}\n".format(num_iters)
    s += "keff: {0:12.5E}\n".format(keff)
    if fluxes is not None:
        s += "fluxes:\n" + "\n".join("{0:12.6E}".format(f) for f in np.ravel(fluxes)) + "\n"
    return s
