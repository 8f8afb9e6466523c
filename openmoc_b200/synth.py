"""Synthetic tracks of the named benchmark shapes, without OpenMOC.

``make_tracks(model, num_azim, spacing, ...)`` returns the same ``FlatTracks``
that OpenMOC's ``TrackGenerator`` + ``b200_flatten`` would hand to the solver
for the reference's decks

  pin-cell        tests/input_set.py:95-137
  simple-lattice  tests/input_set.py:310-417
  c5g7-2d         sample-input/benchmarks/c5g7/c5g7-2d.py (+ cells/lattices.py)

using the host-only generator in ``csrc/trackgen.cpp`` (cyclic track laydown,
links and quadrature restated from the reference; analytic ray tracing of pin
lattices).  It exists so that ``bench.py`` and the GPU box can build the
BASELINE.json workloads - up to ~1e8 segments for 2D C5G7 - in seconds, with
neither the reference nor a multi-GB track file.  By default FSR ids follow the
reference's discovery order, which makes seg_fsr bit-identical to tracks dumped
from the reference (tests/test_trackgen.py).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from typing import Dict, List, Tuple

import numpy as np

from .trackfile import FlatTracks, VACUUM, REFLECTIVE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200trackgen.so")

KIND_PIN, KIND_GRID = 0, 1
QUAD_TY, QUAD_EQUAL_ANGLE, QUAD_GAUSS_LEGENDRE, QUAD_EQUAL_WEIGHT = 0, 1, 2, 3


class CellType(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_rings", C.c_int32), ("n_sectors_fuel", C.c_int32),
                ("n_sectors_mod", C.c_int32), ("mat_fuel", C.c_int32), ("mat_mod", C.c_int32),
                ("subdiv", C.c_int32), ("pad", C.c_int32), ("fuel_radius", C.c_double)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `make -C openmoc_b200/csrc`")
        L = C.CDLL(LIB_PATH)
        L.b200_trackgen_create_2d.restype = C.c_void_p
        L.b200_trackgen_create_2d.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                              C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(C.c_int)]
        L.b200_trackgen_create_3d.restype = C.c_void_p
        L.b200_trackgen_create_3d.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                              C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                              C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                                              C.c_int, C.POINTER(C.c_int)]
        L.b200_trackgen_destroy.argtypes = [C.c_void_p]
        L.b200_trackgen_error.restype = C.c_char_p
        L.b200_trackgen_error.argtypes = [C.c_void_p]
        L.b200_trackgen_get.restype = C.c_int64
        L.b200_trackgen_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        _lib = L
    return _lib


# --------------------------------------------------------------------- materials
def c5g7_materials() -> Tuple[List[str], Dict[str, np.ndarray]]:
    """The seven C5G7 materials as solver tables (reference storage order)."""
    doc = json.load(open(os.path.join(HERE, "data", "c5g7_xs.json")))
    G = doc["num_groups"]
    names = list(doc["materials"])
    n = len(names)
    t = {k: np.zeros(n * G) for k in ("mat_sigma_t", "mat_sigma_a", "mat_sigma_f", "mat_nu_sigma_f", "mat_chi")}
    t["mat_sigma_s"] = np.zeros(n * G * G)
    t["mat_fiss_matrix"] = np.zeros(n * G * G)
    t["mat_fissionable"] = np.zeros(n, dtype=np.uint8)
    for m, name in enumerate(names):
        d = doc["materials"][name]
        for key, src in (("mat_sigma_t", "sigma_t"), ("mat_sigma_a", "sigma_a"), ("mat_sigma_f", "sigma_f"),
                         ("mat_nu_sigma_f", "nu_sigma_f"), ("mat_chi", "chi")):
            t[key][m * G:(m + 1) * G] = d[src]
        s_in = np.array(d["sigma_s"]).reshape(G, G)            # [origin][destination] as given to setSigmaS
        t["mat_sigma_s"][m * G * G:(m + 1) * G * G] = s_in.T.ravel()   # stored [dest*G+orig] (Material.cpp:728-731)
        chi, nsf = np.array(d["chi"]), np.array(d["nu_sigma_f"])
        chi_sum = 0.0
        for c in chi:                       # Material::setChi normalises chi (Material.cpp:900-909)
            chi_sum += c
        if abs(chi_sum) >= 1e-12:
            chi = chi / chi_sum
        t["mat_chi"][m * G:(m + 1) * G] = chi
        t["mat_fiss_matrix"][m * G * G:(m + 1) * G * G] = np.outer(chi, nsf).ravel()  # Material.cpp:975-978
        t["mat_fissionable"][m] = 1 if nsf.sum() > 0 else 0     # Material::setNuSigmaF marks fissionable
    return names, t


# ------------------------------------------------------------------------ models
def _model(name: str):
    """-> (nx, ny, pitch_x, pitch_y, xmin, ymin, cell_type[ny][nx] (row 0 = bottom), types, bcs, default quad)"""
    names, _ = c5g7_materials()
    M = {n: i for i, n in enumerate(names)}
    pin = lambda r, rings, sf, sm, fuel, mod="Water": CellType(KIND_PIN, rings, sf, sm, M[fuel], M[mod], 0, 0, r)
    grid = lambda k, mat="Water": CellType(KIND_GRID, 0, 0, 0, M[mat], M[mat], k, 0, 0.0)
    R = REFLECTIVE
    if name == "pin-cell":
        types = [pin(1.0, 0, 0, 0, "UO2")]
        return 1, 1, 4.0, 4.0, -2.0, -2.0, np.zeros((1, 1), "i4"), types, (R, R, R, R), QUAD_TY
    if name == "simple-lattice":
        types = [pin(0.4, 3, 8, 8, "UO2"), pin(0.3, 3, 8, 8, "UO2"), pin(0.2, 3, 8, 8, "UO2")]
        # 2x2 assembly [[pin1, pin2], [pin1, pin3]] (top row first) tiled 2x2
        top_down = np.array([[0, 1, 0, 1], [0, 2, 0, 2], [0, 1, 0, 1], [0, 2, 0, 2]], "i4")
        return 4, 4, 1.0, 1.0, -2.0, -2.0, top_down[::-1].copy(), types, (R, R, R, R), QUAD_TY
    if name == "c5g7-2d":
        # cells.py:6-8,73-91: fuel pins 4 sectors/no rings, tubes 5 rings x 4 sectors, moderator 8 sectors
        u, m, o, x = (pin(0.54, 0, 4, 8, f) for f in ("UO2", "MOX-4.3%", "MOX-7%", "MOX-8.7%"))
        g, f = pin(0.54, 5, 4, 8, "Guide Tube"), pin(0.54, 5, 4, 8, "Fission Chamber")
        a, r = grid(3), grid(1)
        types = [u, m, o, x, g, f, a, r]
        U, Mx, O, X, Gt, Fc, Ar, Rr = range(8)
        tubes = ["." * 17, "." * 17, ".....g..g..g.....", "...g.........g...", "." * 17,
                 "..g..g..g..g..g..", "." * 17, "." * 17, "..g..g..f..g..g..", "." * 17, "." * 17,
                 "..g..g..g..g..g..", "." * 17, "...g.........g...", ".....g..g..g.....", "." * 17, "." * 17]
        mox = ["mmmmmmmmmmmmmmmmm", "mooooooooooooooom", "mooooooooooooooom", "mooooxxxxxxxoooom",
               "moooxxxxxxxxxooom", "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom",
               "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom",
               "moooxxxxxxxxxooom", "mooooxxxxxxxoooom", "mooooooooooooooom", "mooooooooooooooom",
               "mmmmmmmmmmmmmmmmm"]
        tube = lambda ch: {"g": Gt, "f": Fc}.get(ch)
        uo2_a = np.array([[tube(tubes[j][i]) if tube(tubes[j][i]) is not None else U for i in range(17)]
                          for j in range(17)], "i4")
        mox_a = np.array([[tube(tubes[j][i]) if tube(tubes[j][i]) is not None else {"m": Mx, "o": O, "x": X}[mox[j][i]]
                           for i in range(17)] for j in range(17)], "i4")
        right = np.array([[Ar if i < 11 else Rr for i in range(17)] for j in range(17)], "i4")
        bottom = np.array([[Ar if j < 11 else Rr for i in range(17)] for j in range(17)], "i4")
        corner = np.array([[Ar if (i < 11 and j < 11) else Rr for i in range(17)] for j in range(17)], "i4")
        top_down = np.block([[uo2_a, mox_a, right], [mox_a, uo2_a, right], [bottom, bottom, corner]])
        # surfaces.py:24-27: x-min reflective, x-max vacuum, y-min vacuum, y-max reflective
        return (51, 51, 1.26, 1.26, -32.13, -32.13, top_down[::-1].copy().astype("i4"), types,
                (R, VACUUM, VACUUM, R), QUAD_TY)   # see _model docstring note below
    # NOTE c5g7-2d.py builds an EqualAnglePolarQuad but never calls setNumAzimAngles on
    # it, so TrackGenerator::generateTracks (src/TrackGenerator.cpp:802-806,887-896)
    # discards it and falls back to the default TY quadrature with 6 polar angles:
    # TY is what the reference actually runs for this deck.
    raise ValueError(f"unknown model {name!r} (pin-cell, simple-lattice, c5g7-2d)")


def _renumber_by_discovery(arrays: Dict[str, np.ndarray]) -> None:
    """Number FSRs in order of first appearance along the segment stream and drop
    regions no track crosses - the numbering OpenMOC's single-threaded ray tracer
    produces (Geometry::findFSRId appends a new FSR on first sight)."""
    seg = arrays["seg_fsr"]
    uniq, first = np.unique(seg, return_index=True)
    order = np.argsort(first, kind="stable")
    new_id = np.full(arrays["fsr_volume"].size, -1, dtype=np.int64)
    new_id[uniq[order]] = np.arange(uniq.size)
    arrays["seg_fsr"] = new_id[seg].astype("i4")
    arrays["fsr_volume"] = arrays["fsr_volume"][uniq[order]]
    arrays["fsr_mat"] = arrays["fsr_mat"][uniq[order]]
    if "fsr_cell" in arrays:
        arrays["fsr_cell"] = arrays["fsr_cell"][uniq[order]]


def materials_70g(names: List[str]) -> Dict[str, np.ndarray]:
    """The synthetic 70-group set of tests/test_forward_3D_lattice_70g/
    test_forward_3D_lattice_70g.py:43-61 (UO2 and Water; every other material gets the
    Water data), in solver storage order."""
    G = 70
    n = len(names)
    t = {k: np.zeros(n * G) for k in ("mat_sigma_t", "mat_sigma_a", "mat_sigma_f", "mat_nu_sigma_f", "mat_chi")}
    t["mat_sigma_s"] = np.zeros(n * G * G)
    t["mat_fiss_matrix"] = np.zeros(n * G * G)
    t["mat_fissionable"] = np.zeros(n, dtype=np.uint8)
    for m, name in enumerate(names):
        if name == "UO2":
            nsf, chi = np.linspace(0, 1, G) * 7, np.full(G, 1 / 70.)
            s_in, sig_t = (np.linspace(0, 1, G * G) / 1000).reshape(G, G), np.linspace(2, 3, G)
        else:
            nsf, chi = np.zeros(G), np.zeros(G)
            s_in, sig_t = (np.linspace(1, 2, G * G) / 1000).reshape(G, G), np.linspace(3, 4, G)
        if chi.sum() > 0:
            chi = chi / chi.sum()
        sl = slice(m * G, (m + 1) * G)
        t["mat_sigma_t"][sl], t["mat_nu_sigma_f"][sl], t["mat_chi"][sl] = sig_t, nsf, chi
        t["mat_sigma_a"][sl] = sig_t - s_in.sum(axis=1)
        t["mat_sigma_s"][m * G * G:(m + 1) * G * G] = s_in.T.ravel()
        t["mat_fiss_matrix"][m * G * G:(m + 1) * G * G] = np.outer(chi, nsf).ravel()
        t["mat_fissionable"][m] = 1 if nsf.sum() > 0 else 0
    return t


def make_tracks(model: str, num_azim: int = 4, spacing: float = 0.1, num_polar: int = 6,
                polar_quad: int = None, num_threads: int = 0,
                fsr_numbering: str = "discovery", groups70: bool = False,
                as_3d: bool = False, linear_source: bool = False) -> FlatTracks:
    """fsr_numbering: "discovery" (reference order, untouched regions dropped) or
    "lattice" (by lattice cell, every geometric region kept).
    groups70: swap the C5G7 data for the reference's synthetic 70-group set.
    linear_source: also produce the FSR centroids and centroid-relative segment starting points.
    as_3d: label the tracks as 3D tracks of polar index 0 (one angular flux per group per
    track, F = G) - a kernel-shape stand-in for 3D decks, not a physical 3D problem."""
    L = _load()
    nx, ny, px, py, xmin, ymin, cells, types, bcs, default_quad = _model(model)
    if polar_quad is None:
        polar_quad = default_quad
    cells = np.ascontiguousarray(cells, dtype="i4")
    tarr = (CellType * len(types))(*types)
    status = C.c_int()
    h = L.b200_trackgen_create_2d(nx, ny, px, py, xmin, ymin, cells.ctypes.data_as(C.c_void_p),
                                  C.cast(tarr, C.c_void_p), len(types), bcs[0], bcs[1], bcs[2], bcs[3],
                                  num_azim, float(spacing), num_polar, polar_quad, num_threads, C.byref(status))
    try:
        if status.value != 0:
            raise ValueError("track generation failed: " + L.b200_trackgen_error(h).decode())
        dtypes = {"seg_length": "f8", "seg_fsr": "i4", "seg_mat": "i4", "trk_seg_offset": "i8",
                  "trk_next_fwd": "i8", "trk_next_bwd": "i8", "trk_azim": "i4", "trk_polar": "i4",
                  "trk_xy": "i4", "trk_flags": "u1", "trk_bc_fwd": "u1", "trk_bc_bwd": "u1",
                  "trk_phi": "f8", "trk_theta": "f8", "trk_start": "f8", "quad_weight": "f8",
                  "quad_sin_theta": "f8", "fsr_volume": "f8", "fsr_mat": "i4",
                  "quad_azim_spacing": "f8", "quad_azim_weight": "f8", "quad_polar_spacing": "f8",
                  "quad_polar_weight": "f8", "seg_cmfd_fwd": "i4", "seg_cmfd_bwd": "i4", "fsr_cell": "i4"}
        arrays = {}
        for k, dt in dtypes.items():
            n = L.b200_trackgen_get(h, k.encode(), None)
            a = np.empty(n, dtype=dt)
            L.b200_trackgen_get(h, k.encode(), a.ctypes.data_as(C.c_void_p))
            arrays[k] = a
    finally:
        L.b200_trackgen_destroy(h)
    if fsr_numbering == "discovery":
        _renumber_by_discovery(arrays)
    names, mats = c5g7_materials()
    G = 7
    if groups70:
        mats, G = materials_70g(names), 70
    arrays.update(mats)
    ft = FlatTracks(num_groups=G, num_azim=num_azim, num_polar=num_polar, solve_3d=int(as_3d),
                    fluxes_per_track=G if as_3d else G * num_polar // 2, n_tracks=int(arrays["trk_azim"].size),
                    n_segments=int(arrays["seg_length"].size), n_fsrs=int(arrays["fsr_volume"].size),
                    n_materials=len(names), arrays=arrays)
    if linear_source:
        add_linear_source_data(ft)
    return ft


def add_linear_source_data(ft: FlatTracks) -> None:
    """What a linear-source solve needs on top of the flat data (2D decks): FSR centroids from the
    tracks (CentroidGenerator::onTrack, src/TrackTraversingAlgorithms.cpp:377-460) and the starting
    point of every segment relative to the centroid of its FSR (RecenterSegments, :1322-1332)."""
    if ft.solve_3d:
        raise ValueError("synthetic linear-source data exist for 2D decks only")
    a = ft.arrays
    off = a["trk_seg_offset"].astype(np.int64)
    nseg = np.diff(off)
    trk = np.repeat(np.arange(ft.n_tracks), nseg)
    length = a["seg_length"]
    before = np.cumsum(length) - length                       # arc length before each segment ...
    first = np.minimum(off[:-1], max(ft.n_segments - 1, 0))
    before -= np.repeat(np.where(nseg > 0, before[first], 0.0), nseg)    # ... counted from the track start
    phi = a["trk_phi"][trk]
    cos_phi, sin_phi = np.cos(phi), np.sin(phi)
    start = a["trk_start"].reshape(-1, 2)[trk]
    x = start[:, 0] + before * cos_phi
    y = start[:, 1] + before * sin_phi
    azim = a["trk_azim"][trk].astype(np.int64)
    wgt = a["quad_azim_spacing"][azim] * a["quad_azim_weight"][azim]
    fsr = a["seg_fsr"].astype(np.int64)
    vol = a["fsr_volume"][fsr]
    cx = np.bincount(fsr, weights=wgt * (x + cos_phi * length / 2.0) * length / vol, minlength=ft.n_fsrs)
    cy = np.bincount(fsr, weights=wgt * (y + sin_phi * length / 2.0) * length / vol, minlength=ft.n_fsrs)
    a["fsr_centroid"] = np.stack([cx, cy, np.zeros_like(cx)], axis=1).ravel()
    a["seg_start"] = np.stack([x - cx[fsr], y - cy[fsr], np.zeros_like(x)], axis=1).ravel()


# ----------------------------------------------------------------------- 3D decks
#: axial extent and z boundary conditions of the reference's 3D decks:
#:   pin-cell / simple-lattice  tests/input_set.py (PinCellInput / SimpleLatticeInput, num_dimensions=3)
#:   c5g7-3d                    profile/models/c5g7/c5g7-3d-cmfd.cpp (z in [-32.13, 32.13], reflective
#:                              bottom, vacuum top; the 2D C5G7 core extruded)
AXIAL = {"pin-cell": (-2.0, 2.0, REFLECTIVE, REFLECTIVE),
         "simple-lattice": (-5.0, 5.0, REFLECTIVE, VACUUM),
         "c5g7-2d": (-32.13, 32.13, REFLECTIVE, VACUUM)}

_DT3 = {"seg_length": "f8", "seg_fsr": "i4", "seg_mat": "i4", "trk_seg_offset": "i8",
        "trk_next_fwd": "i8", "trk_next_bwd": "i8", "trk_azim": "i4", "trk_polar": "i4",
        "trk_xy": "i4", "trk_2d": "i4", "trk_lz": "i4", "trk_flags": "u1", "trk_bc_fwd": "u1", "trk_bc_bwd": "u1",
        "trk_phi": "f8", "trk_theta": "f8", "trk_start": "f8", "trk_end": "f8", "trk_l0": "f8", "z_mesh": "f8",
        "quad_weight": "f8", "quad_sin_theta": "f8", "fsr_volume": "f8", "fsr_mat": "i4",
        "quad_azim_spacing": "f8", "quad_azim_weight": "f8", "quad_polar_spacing": "f8", "quad_polar_weight": "f8",
        "seg2d_length": "f8", "seg2d_fsr": "i4", "seg2d_mat": "i4", "trk2d_seg_offset": "i8",
        "trk2d_start": "f8", "trk2d_phi": "f8", "fsr2d_mat": "i4",
        "seg2d_surf_fwd": "i1", "seg2d_surf_bwd": "i1", "fsr2d_cell": "i4"}


def make_tracks_3d(model: str, num_azim: int = 4, spacing: float = 0.1, num_polar: int = 2,
                   z_spacing: float = 0.1, n_axial: int = 1, polar_quad: int = QUAD_GAUSS_LEGENDRE,
                   num_threads: int = 0, expand: bool = True, fsr_numbering: str = "discovery",
                   groups70: bool = False) -> FlatTracks:
    """3D tracks (z-stacks over the 2D tracks, TrackGenerator3D) of the named deck extruded in
    n_axial equal layers.  polar_quad defaults to Gauss-Legendre, the reference's 3D default
    (TrackGenerator3D::initializeDefaultQuadrature).

    expand=True: explicit 3D segments + FSR volumes from the host tracer (tests, small decks).
    expand=False: the arrays of the device tracer instead (2D segments `seg2d_*`, `trk2d_*`,
    the axial mesh `z_mesh`, per-track `trk_2d`, `trk_l0`, `trk_start`): the solver traces the
    z-stacks on the GPU (b200_upload_tracks_otf); FSR numbering is then always "lattice"."""
    L = _load()
    nx, ny, px, py, xmin, ymin, cells, types, bcs, _ = _model(model)
    zmin, zmax, bc_zmin, bc_zmax = AXIAL[model]
    cells = np.ascontiguousarray(cells, dtype="i4")
    tarr = (CellType * len(types))(*types)
    status = C.c_int()
    h = L.b200_trackgen_create_3d(nx, ny, px, py, xmin, ymin, cells.ctypes.data_as(C.c_void_p),
                                  C.cast(tarr, C.c_void_p), len(types), bcs[0], bcs[1], bcs[2], bcs[3],
                                  num_azim, float(spacing), num_polar, polar_quad, num_threads,
                                  zmin, zmax, bc_zmin, bc_zmax, int(n_axial), float(z_spacing),
                                  int(bool(expand)), C.byref(status))
    try:
        if status.value != 0:
            raise ValueError("3D track generation failed: " + L.b200_trackgen_error(h).decode())
        arrays = {}
        for k, dt in _DT3.items():
            n = L.b200_trackgen_get(h, k.encode(), None)
            if n < 0:
                raise RuntimeError("trackgen has no array %r" % k)
            a = np.empty(n, dtype=dt)
            L.b200_trackgen_get(h, k.encode(), a.ctypes.data_as(C.c_void_p))
            arrays[k] = a
    finally:
        L.b200_trackgen_destroy(h)
    if expand and fsr_numbering == "discovery":
        _renumber_by_discovery(arrays)
    names, mats = c5g7_materials()
    G = 7
    if groups70:
        mats, G = materials_70g(names), 70
    arrays.update(mats)
    ft = FlatTracks(num_groups=G, num_azim=num_azim, num_polar=num_polar, solve_3d=1, fluxes_per_track=G,
                    n_tracks=int(arrays["trk_azim"].size), n_segments=int(arrays["seg_length"].size),
                    n_fsrs=int(arrays["fsr_volume"].size), n_materials=len(names), arrays=arrays)
    ft.n_axial = int(n_axial)
    return ft


# ----------------------------------------------------------------------- CMFD mesh of a synthetic deck
def cmfd_mesh(ft: FlatTracks, model: str, num_z: int = 1, group_structure=None, **options):
    """A CMFD mesh laid over the pin lattice of the named deck (51 x 51 for C5G7, what
    sample-input/benchmarks/c5g7/c5g7-2d.py:51-56 and profile/models/c5g7/c5g7-3d-cmfd.cpp:537-556 use), `num_z`
    equal cells axially for the 3D decks (must divide the number of axial layers).  The generator has marked the
    surface of the lattice cell every segment ends on (`seg_cmfd_fwd/bwd`, `seg2d_surf_fwd/bwd`) and the cell of every
    FSR.  Returns the openmoc_b200.solver.CmfdMesh to hand to B200Solver(tracks, cmfd=...)."""
    from .solver import CmfdMesh
    nx, ny, px, py, xmin, ymin, _, _, bcs, _ = _model(model)
    a = ft.arrays
    if ft.solve_3d and "fsr2d_cell" in a:
        zmin, zmax, bc_zmin, bc_zmax = AXIAL[model]
        n_axial = int(getattr(ft, "n_axial", 1))
        if n_axial % num_z:
            raise ValueError("num_z must divide the number of axial layers")
        cell2d = a["fsr2d_cell"].astype(np.int64)
        layer = np.arange(n_axial, dtype=np.int64)
        # 3D FSR id = 2D FSR id * n_axial + layer (trackgen.cpp); CMFD cell = (z cell * ny + y) * nx + x
        fsr_cell = (cell2d[:, None] + (layer[None, :] * num_z // n_axial) * (nx * ny)).ravel()
        wz = np.full(num_z, (zmax - zmin) / num_z)
        z_planes = zmin + np.arange(num_z + 1) * (zmax - zmin) / num_z
        bz = (bc_zmin, bc_zmax)
    else:
        if num_z != 1:
            raise ValueError("a 2D deck has one axial CMFD cell")
        fsr_cell, wz, z_planes, bz = a["fsr_cell"], np.ones(1), None, (REFLECTIVE, REFLECTIVE)   # Cmfd.cpp:3860-3869
    # boundaries in face order X_MIN, Y_MIN, Z_MIN, X_MAX, Y_MAX, Z_MAX (src/constants.h:120-125); bcs = (xmin, xmax, ymin, ymax)
    boundaries = (bcs[0], bcs[2], bz[0], bcs[1], bcs[3], bz[1])
    return CmfdMesh(num_x=nx, num_y=ny, num_z=num_z, widths_x=np.full(nx, px), widths_y=np.full(ny, py), widths_z=wz,
                    boundaries=boundaries, fsr_cell=np.ascontiguousarray(fsr_cell, dtype="i4"),
                    group_structure=group_structure, z_planes=z_planes, **options)
