"""ctypes binding of the C ABI in include/b200moc.h (libb200moc.so).

This is the same set of entry points the C++ ``B200Solver : public Solver``
plug-in calls (openmoc_b200/cpp/B200Solver.cpp); nothing here computes anything.
Loading fails loudly when the CUDA library has not been built - there is no
CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200moc.so")

SCALAR_FLUX, FISSION_SOURCE, TOTAL_SOURCE = 0, 1, 2
DIAGONAL, YAMAMOTO, GLOBAL = 0, 1, 2
PRECISION_DOUBLE, PRECISION_MIXED, PRECISION_TABLE = 0, 1, 2


class B200Error(RuntimeError):
    """Raised for any non-zero status of the C ABI (the reference raises
    RuntimeError from log_printf(ERROR), openmoc/swig/openmoc.i:106-112)."""


class Config(C.Structure):
    _fields_ = [("num_groups", C.c_int32), ("num_azim", C.c_int32), ("num_polar", C.c_int32),
                ("solve_3d", C.c_int32), ("n_tracks", C.c_int64), ("n_segments", C.c_int64),
                ("n_fsrs", C.c_int64), ("n_materials", C.c_int32), ("device", C.c_int32),
                ("precision", C.c_int32), ("deterministic", C.c_int32), ("n_fsrs_global", C.c_int64),
                ("linear_source", C.c_int32), ("reserved", C.c_int32)]


class CmfdConfig(C.Structure):
    """b200_cmfd_config (include/b200moc.h): the options of a reference Cmfd object."""
    _fields_ = [("num_x", C.c_int32), ("num_y", C.c_int32), ("num_z", C.c_int32), ("num_cmfd_groups", C.c_int32),
                ("boundaries", C.c_int32 * 6), ("linear_source", C.c_int32), ("flux_limiting", C.c_int32),
                ("centroid_update", C.c_int32), ("axial_interpolation", C.c_int32),
                ("num_unbounded_iterations", C.c_int32), ("num_azim_2", C.c_int32), ("num_polar_2", C.c_int32),
                ("sor_factor", C.c_double), ("relaxation_factor", C.c_double), ("linalg_tolerance", C.c_double)]


class CmfdStats(C.Structure):
    """b200_cmfd_stats: struct ConvergenceData of the reference (src/linalg.h:31-68) + device time."""
    _fields_ = [("pf", C.c_double), ("cmfd_res_1", C.c_double), ("cmfd_res_end", C.c_double),
                ("linear_res_1", C.c_double), ("linear_res_end", C.c_double), ("cmfd_iters", C.c_int32),
                ("linear_iters_1", C.c_int32), ("linear_iters_end", C.c_int32), ("linear_iters_total", C.c_int32),
                ("failed", C.c_int32), ("bad_tallies", C.c_int32), ("device_ms", C.c_double)]


_lib = None

# name -> argument ctypes (after the solver handle); all return int status
_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
SIGNATURES = {
    "b200_destroy": [],
    "b200_upload_tracks": [_vp] * 10,
    "b200_upload_quadrature": [_vp, _vp],
    "b200_upload_fsrs": [_vp, _vp],
    "b200_upload_materials": [_vp] * 7,
    "b200_upload_linear_source": [_vp, _vp, _vp, _vp],
    "b200_upload_cmfd_surfaces": [_vp, _vp],
    "b200_set_cmfd_groups": [_vp, _i32, _i64],
    "b200_get_cmfd_currents": [_vp, _i64],
    "b200_upload_otf_cmfd": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "b200_cmfd_configure": [C.POINTER(CmfdConfig)] + [_vp] * 9,
    "b200_cmfd_set_stencils": [_vp] * 5,
    "b200_cmfd_set_axial_interpolants": [_vp],
    "b200_cmfd_set_keff": [_dbl],
    "b200_cmfd_solve": [_i32, _dbl, C.POINTER(_dbl), C.POINTER(CmfdStats)],
    "b200_cmfd_set_in_loop": [_i32],
    "b200_get_flux_moments": [_vp, _i64],
    "b200_set_flux_moments": [_vp, _i64],
    "b200_upload_otf_geometry": [_i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp],
    "b200_otf_compute_volumes": [_i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "b200_upload_tracks_otf": [_vp] * 10 + [C.POINTER(_i64)],
    "b200_set_max_optical_length": [_dbl],
    "b200_get_num_segments": [C.POINTER(_i64)],
    "b200_get_segments": [_vp, _vp, _i64, _vp],
    "b200_get_volumes": [_vp, _i64],
    "b200_otf_count_segments": [_i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "b200_set_devices": [_i32, _vp],
    "b200_get_num_devices": [C.POINTER(_i32)],
    "b200_finalize": [],
    "b200_zero_track_fluxes": [],
    "b200_flatten_fsr_fluxes": [_dbl],
    "b200_flatten_fsr_fluxes_chi_spectrum": [_i32],
    "b200_store_fsr_fluxes": [],
    "b200_normalize_fluxes": [C.POINTER(_dbl)],
    "b200_compute_stabilizing_flux": [],
    "b200_stabilize_flux": [],
    "b200_compute_fsr_sources": [_i32],
    "b200_compute_fsr_fission_sources": [],
    "b200_compute_fsr_scatter_sources": [],
    "b200_compute_residual": [_i32, C.POINTER(_dbl)],
    "b200_compute_keff": [C.POINTER(_dbl)],
    "b200_add_source_to_scalar_flux": [],
    "b200_transport_sweep": [],
    "b200_get_fluxes": [_vp, _i64],
    "b200_set_fluxes": [_vp, _i64],
    "b200_get_fluxes_keff": [_vp, _i64, C.POINTER(_dbl)],
    "b200_set_fixed_source_by_fsr": [_i64, _i32, _dbl],
    "b200_set_fixed_source_moments_by_fsr": [_i64, _i32, _dbl, _dbl, _dbl],
    "b200_reset_fixed_sources": [],
    "b200_compute_fsr_fission_rates": [_vp, _i64, _i32],
    "b200_stabilize_transport": [_dbl, _i32],
    "b200_allow_negative_fluxes": [_i32],
    "b200_set_keff_from_neutron_balance": [_i32],
    "b200_get_keff": [C.POINTER(_dbl)],
    "b200_set_keff": [_dbl],
    "b200_get_fsr_sources": [_vp, _i64],
    "b200_set_fsr_sources": [_vp, _i64],
    "b200_get_start_fluxes": [_vp, _i64],
    "b200_set_start_fluxes": [_vp, _i64],
    "b200_compute_eigenvalue": [_i32, _dbl, _i32, C.POINTER(_i32)],
    "b200_compute_flux": [_i32, _dbl, _i32, C.POINTER(_i32)],
    "b200_compute_source": [_i32, _dbl, _dbl, _i32, C.POINTER(_i32)],
    "b200_iterate": [_i32, _i32, C.POINTER(_dbl), C.POINTER(_dbl)],
    "b200_eigen_loop_init": [_i32, _dbl],
    "b200_iteration_begin": [_i32],
    "b200_iteration_end": [_i32, _i32, _i32],
    "b200_eigen_loop_status": [_i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_dbl), C.POINTER(_dbl)],
    "b200_get_sweep_stats": [C.POINTER(_dbl), C.POINTER(_i64), C.POINTER(_i64)],
    "b200_reset_sweep_stats": [],
    "b200_synchronize": [],
    "b200_device_pointer": [C.c_char_p, C.POINTER(_vp), C.POINTER(_i64)],
    "b200_set_stream": [_vp],
    "b200_set_capturing": [_i32],
    "b200_defer_fixed_tally": [_i32],
    "b200_finish_fixed_tally": [],
}
#: every symbol include/b200moc.h declares (checked by tests/test_abi.py)
EXPORTS = sorted(list(SIGNATURES) + ["b200_last_error", "b200_version", "b200_device_count", "b200_create",
                                      "b200_eval_expF1", "b200_measure_ceilings", "b200_ls_prepass",
                                      "b200_cmfd_split_targets"])


def load():
    """Load libb200moc.so; raise B200Error with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} is missing: build the CUDA extension with "
            "`make -C openmoc_b200/csrc` (or __graft_entry__.build()). "
            "The B200 solver has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.b200_last_error.restype = C.c_char_p
    L.b200_last_error.argtypes = []
    L.b200_version.restype = C.c_int
    L.b200_device_count.restype = C.c_int
    L.b200_create.restype = C.c_int
    L.b200_create.argtypes = [C.POINTER(Config), C.POINTER(_vp)]
    L.b200_eval_expF1.restype = C.c_int
    L.b200_eval_expF1.argtypes = [_i32, _i32, _vp, _i64, _vp]
    L.b200_ls_prepass.restype = C.c_int
    L.b200_ls_prepass.argtypes = [_i32, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _i32] + [_vp] * 16 + [_vp, _vp, C.POINTER(_i32)]
    L.b200_cmfd_split_targets.restype = C.c_int
    L.b200_cmfd_split_targets.argtypes = [_i32, _i32, _i32, _vp, _i32, _i32, _vp, C.POINTER(_i32)]
    L.b200_measure_ceilings.restype = C.c_int
    L.b200_measure_ceilings.argtypes = [_i32, _i64, C.POINTER(_dbl), C.POINTER(_dbl)]
    for name, args in SIGNATURES.items():
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [_vp] + args
    _lib = L
    return L


def eval_expF1(x, precision: int = PRECISION_DOUBLE, device: int = 0):
    """Device evaluation of expF1_fractional for an array of optical lengths."""
    import numpy as np
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    check(load().b200_eval_expF1(device, precision, x.ctypes.data_as(_vp), x.size, out.ctypes.data_as(_vp)))
    return out


def measure_ceilings(device: int = 0, table_rows: int = 23869):
    """(FP64 instructions / s, RED.ADD.F64 / s) measured on the device (microbench.cuh)."""
    a, b = _dbl(), _dbl()
    check(load().b200_measure_ceilings(device, table_rows, C.byref(a), C.byref(b)))
    return a.value, b.value


def check(status: int) -> None:
    if status != 0:
        raise B200Error(load().b200_last_error().decode(errors="replace"))
