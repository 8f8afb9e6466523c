"""Pre-pass of the linear-source solver: the per-FSR linear expansion matrices and source constants the
sweep and closure kernels read (`b200_upload_linear_source`).  `linear_expansion_tables_device` runs it on
the GPU (csrc/ls_prepass.cuh) and is what the solver uses; `linear_expansion_tables` is the numpy
restatement the CPU tests check against the oracle and the GPU tests check the device version against.

Restates `LinearExpansionGenerator::onTrack` / `execute`
(src/TrackTraversingAlgorithms.cpp:536-831) on flattened tracks, vectorised over segments with
numpy.  The C++ plug-in does not need it (B200LSSolver inherits the reference's own pre-pass from
CPULSSolver); this module is what lets a track file alone - the Python `B200Solver` path - carry a
linear-source solve.  Needs the LS chunks of the track file: `seg_start` (centroid-relative),
`trk_phi`, `trk_theta`, `quad_azim_spacing/_weight`, `quad_polar_spacing/_weight`.
"""
from __future__ import annotations

import numpy as np

from .trackfile import FlatTracks

MIN_DET = 1e-10                       # src/constants.h:70


def expG2(x: np.ndarray) -> np.ndarray:
    """src/exponentials.h:293-323, the 5/5-order rational."""
    a1, a2, a3, a4, a5 = (-8.335775885589858e-2, -3.603942303847604e-3, 3.7673183263550827e-3,
                          1.124183494990467e-5, 1.6837426505799449e-4)
    b1, b2, b3, b4, b5 = (7.454048371823628e-1, 2.3794300531408347e-1, 5.367250964303789e-2,
                          6.125197988351906e-3, 1.0102514456857377e-3)
    num = a5 * x + a4
    num = num * x + a3
    num = num * x + a2
    num = num * x + a1
    num = num * x
    den = b5 * x + b4
    den = den * x + b3
    den = den * x + b2
    den = den * x + b1
    den = den * x + 1.0
    return num / den


def track_directions(ft: FlatTracks) -> np.ndarray:
    """[n_tracks, 3] unit vectors of the forward direction (TrackTraversingAlgorithms.cpp:913-926)."""
    phi = ft.arrays["trk_phi"]
    if ft.solve_3d:
        theta = ft.arrays["trk_theta"]
        st, ct = np.sin(theta), np.cos(theta)
    else:
        st, ct = np.ones_like(phi), np.zeros_like(phi)
    return np.stack([np.cos(phi) * st, np.sin(phi) * st, ct], axis=1)


def linear_expansion_tables_device(ft: FlatTracks, device: int = 0):
    """The same tables from the device kernels (csrc/ls_prepass.cuh through b200_ls_prepass): what
    B200Solver(linear_source=True) uses.  Returns (lin_exp, src_const, n_flat) like linear_expansion_tables."""
    import ctypes as C
    from . import capi
    a = ft.arrays
    G, P = ft.num_groups, ft.num_polar
    nc = 6 if ft.solve_3d else 3
    c = lambda k, dt: np.ascontiguousarray(a[k], dtype=dt)
    ins = [c("seg_length", "f8"), c("seg_fsr", "i4"), c("seg_start", "f8"), c("trk_seg_offset", "i8"), c("trk_azim", "i4"),
           c("trk_polar", "i4"), c("trk_phi", "f8"), c("trk_theta", "f8"), c("quad_azim_spacing", "f8"),
           c("quad_azim_weight", "f8"), c("quad_polar_spacing", "f8"), c("quad_polar_weight", "f8"),
           c("quad_sin_theta", "f8"), c("fsr_volume", "f8"), c("fsr_mat", "i4"), c("mat_sigma_t", "f8")]
    lin_exp = np.zeros(ft.n_fsrs * nc)
    src_const = np.zeros(ft.n_fsrs * nc * G)
    n_flat = C.c_int32()
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    capi.check(capi.load().b200_ls_prepass(device, G, ft.num_azim, P, int(ft.solve_3d), ft.n_tracks, ft.n_segments, ft.n_fsrs,
                                           ft.n_materials, *[p(x) for x in ins], p(lin_exp), p(src_const), C.byref(n_flat)))
    return lin_exp, src_const, int(n_flat.value)


def linear_expansion_tables(ft: FlatTracks):
    """Returns (lin_exp [n_fsrs*nc], src_const [n_fsrs*G*nc], n_flat): nc = 3 in 2D, 6 in 3D;
    layouts `[r][i]` and `[r][i][e]` as in CPULSSolver (src/CPULSSolver.h:36-50)."""
    a = ft.arrays
    G, P, is3d = ft.num_groups, ft.num_polar, bool(ft.solve_3d)
    nc = 6 if is3d else 3
    nseg = np.diff(a["trk_seg_offset"].astype(np.int64))
    trk = np.repeat(np.arange(ft.n_tracks), nseg)                    # track of every segment
    azim = a["trk_azim"][trk].astype(np.int64)
    phi = a["trk_phi"][trk]
    sin_phi, cos_phi = np.sin(phi), np.cos(phi)
    wgt = a["quad_azim_spacing"][azim] * a["quad_azim_weight"][azim]
    if is3d:
        polar = a["trk_polar"][trk].astype(np.int64)
        theta = a["trk_theta"][trk]
        sin_t, cos_t = np.sin(theta), np.cos(theta)
        wgt = wgt * a["quad_polar_spacing"][azim * P + polar] * a["quad_polar_weight"][azim * P + polar]
    else:
        sin_t, cos_t = np.ones_like(phi), np.zeros_like(phi)
    fsr = a["seg_fsr"].astype(np.int64)
    length = a["seg_length"]
    volume = a["fsr_volume"][fsr]
    start = a["seg_start"].reshape(-1, 3)
    xc = start[:, 0] + length * 0.5 * cos_phi * sin_t
    yc = start[:, 1] + length * 0.5 * sin_phi * sin_t
    zc = start[:, 2] + length * 0.5 * cos_t
    vol_impact = wgt * length / volume
    src_constant = vol_impact * length / 2.0
    sigma_t = a["mat_sigma_t"].reshape(-1, G)[a["fsr_mat"][fsr]]      # [n_seg, G]
    tau = length[:, None] * sigma_t

    geo = [xc * xc, yc * yc, xc * yc] + ([xc * zc, yc * zc, zc * zc] if is3d else [])
    tsc = np.stack([np.repeat((vol_impact * g)[:, None], G, axis=1) for g in geo], axis=1)   # [n_seg, nc, G]
    if not is3d:
        for p in range(P // 2):
            st = a["quad_sin_theta"][azim * P + p]
            g2 = (length[:, None] * expG2(tau / st[:, None]) * (src_constant * 2
                  * a["quad_polar_weight"][azim * P + p] * st)[:, None])
            tsc[:, 0] += (cos_phi * cos_phi)[:, None] * g2
            tsc[:, 1] += (sin_phi * sin_phi)[:, None] * g2
            tsc[:, 2] += (sin_phi * cos_phi)[:, None] * g2
    else:
        g2 = expG2(tau) * (length * src_constant)[:, None]
        tsc[:, 0] += (cos_phi * cos_phi * sin_t * sin_t)[:, None] * g2
        tsc[:, 1] += (sin_phi * sin_phi * sin_t * sin_t)[:, None] * g2
        tsc[:, 2] += (sin_phi * cos_phi * sin_t * sin_t)[:, None] * g2
        tsc[:, 3] += (cos_phi * cos_t * sin_t)[:, None] * g2
        tsc[:, 4] += (sin_phi * cos_t * sin_t)[:, None] * g2
        tsc[:, 5] += (cos_t * cos_t)[:, None] * g2
    src_const = np.zeros((ft.n_fsrs, nc, G))
    np.add.at(src_const, fsr, tsc)

    l2 = length * length
    terms = [xc * xc + (cos_phi * sin_t) ** 2 * l2 / 12.0,
             yc * yc + (sin_phi * sin_t) ** 2 * l2 / 12.0,
             xc * yc + sin_phi * cos_phi * sin_t ** 2 * l2 / 12.0]
    if is3d:
        terms += [xc * zc + cos_phi * cos_t * sin_t * l2 / 12.0,
                  yc * zc + sin_phi * cos_t * sin_t * l2 / 12.0,
                  zc * zc + cos_t ** 2 * l2 / 12.0]
    lem = np.zeros((ft.n_fsrs, nc))
    for i, t in enumerate(terms):
        lem[:, i] = np.bincount(fsr, weights=vol_impact * t, minlength=ft.n_fsrs)

    ilem = np.zeros_like(lem)
    m = lem.T
    if is3d:
        det = (m[0] * m[1] * m[5] + m[2] * m[4] * m[3] + m[3] * m[2] * m[4]
               - m[0] * m[4] * m[4] - m[3] * m[1] * m[3] - m[2] * m[2] * m[5])
        ok = ~((np.abs(det) < MIN_DET) | (a["fsr_volume"] < 1e-6))
        d = np.where(ok, det, 1.0)
        inv = [(m[1] * m[5] - m[4] * m[4]) / d, (m[0] * m[5] - m[3] * m[3]) / d, (m[3] * m[4] - m[2] * m[5]) / d,
               (m[2] * m[4] - m[3] * m[1]) / d, (m[3] * m[2] - m[0] * m[4]) / d, (m[0] * m[1] - m[2] * m[2]) / d]
    else:
        det = m[0] * m[1] - m[2] * m[2]
        ok = ~(np.abs(det) < MIN_DET)
        d = np.where(ok, det, 1.0)
        inv = [m[1] / d, m[0] / d, -m[2] / d]
    for i, v in enumerate(inv):
        ilem[:, i] = np.where(ok, v, 0.0)
    return ilem.ravel(), src_const.ravel(), int((~ok).sum())
