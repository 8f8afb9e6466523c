"""Axial on-the-fly ray tracing on the device (openmoc_b200/csrc/otf.cuh, b200_upload_tracks_otf)
against the host tracer of csrc/trackgen.cpp - itself pinned to the reference's TrackGenerator3D
dumps in tests/test_trackgen3d.py - and against the oracle on the expanded tracks.
Integer results bit-exact; segment lengths bit-exact (same IEEE operations on both sides)."""
import ctypes as C

import numpy as np
import pytest

from openmoc_b200 import capi
from openmoc_b200.capi import FISSION_SOURCE, Config, check
from openmoc_b200.synth import make_tracks_3d, QUAD_EQUAL_ANGLE, QUAD_GAUSS_LEGENDRE, QUAD_TY
from oracle.oracle_py import OracleSolver

pytestmark = pytest.mark.gpu

DECKS = [("simple-lattice", dict(num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9, n_axial=1)),
         ("simple-lattice", dict(num_azim=8, spacing=0.3, num_polar=4, z_spacing=0.7, n_axial=5, polar_quad=QUAD_TY)),
         ("c5g7-2d", dict(num_azim=4, spacing=1.0, num_polar=4, z_spacing=8.0, n_axial=3, polar_quad=QUAD_EQUAL_ANGLE)),
         ("pin-cell", dict(num_azim=16, spacing=0.2, num_polar=4, z_spacing=0.3, n_axial=2, polar_quad=QUAD_GAUSS_LEGENDRE))]


@pytest.mark.parametrize("model,kw", DECKS)
def test_device_expansion_equals_host_tracer(model, kw):
    from openmoc_b200.solver import B200Solver
    full = make_tracks_3d(model, fsr_numbering="lattice", **kw)
    lean = make_tracks_3d(model, expand=False, **kw)
    s = B200Solver(lean)
    length, fsr, off = s.getSegments()
    assert s.num_segments == full.n_segments
    assert np.array_equal(off, full.arrays["trk_seg_offset"])
    assert np.array_equal(fsr, full.arrays["seg_fsr"])
    assert np.array_equal(length, full.arrays["seg_length"])          # bit-exact
    np.testing.assert_allclose(s.getVolumes(), full.arrays["fsr_volume"], rtol=1e-12, atol=1e-14)
    assert s.integrationsPerSweep() == 2 * 7 * full.n_segments


def test_on_the_fly_solve_equals_explicit_solve_and_oracle():
    """the reference's test_forward_3D_lattice deck: same k_eff and iteration count whether the 3D
    segments are uploaded or traced on the device, and both equal to the oracle"""
    from openmoc_b200.solver import B200Solver
    kw = dict(num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9, n_axial=1)
    full = make_tracks_3d("simple-lattice", fsr_numbering="lattice", **kw)
    lean = make_tracks_3d("simple-lattice", expand=False, **kw)
    a, b = B200Solver(full), B200Solver(lean)
    for s in (a, b):
        s.setConvergenceThreshold(1e-5)
        s.computeEigenvalue(500, FISSION_SOURCE)
    o = OracleSolver(full)
    n = o.computeEigenvalue(500, 1e-5, FISSION_SOURCE)
    assert a.getNumIterations() == b.getNumIterations() == n
    assert abs(a.getKeff() - b.getKeff()) < 1e-11
    assert abs(b.getKeff() - o.getKeff()) * 1e5 < 1e-4                 # pcm; north star: 1 pcm
    used = full.arrays["fsr_volume"] > 0
    err = np.max(np.abs(b.getFluxes().reshape(-1, 7)[used] - o.getFluxes().reshape(-1, 7)[used])
                 / o.getFluxes().reshape(-1, 7)[used])
    assert err < 2e-9, err                                              # north star: 1e-4


def _py_trace(seg_len, seg_ext, ext_off, mesh, ext_fsr, l0, z0, cos_t, sin_t):
    """TraverseSegments::traceSegmentsOTF (src/TraverseSegments.cpp:304-505) in plain Python,
    per-FSR axial meshes."""
    out = []
    sign = 1 if cos_t > 0 else -1
    s, n = 0, len(seg_len)
    while s < n and l0 > seg_len[s]:
        l0 -= seg_len[s]; s += 1
    z = z0
    while s < n:
        e = seg_ext[s]
        m = mesh[ext_off[e] + e: ext_off[e + 1] + e + 1]
        nf = len(m) - 1
        lo, hi = 0, nf
        zi = None
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if z > m[mid]: lo = mid
            elif z < m[mid]: hi = mid
            else: zi = mid if sign > 0 else mid - 1; break
        if zi is None: zi = lo
        rem = seg_len[s] - l0; l0 = 0.0
        done = False
        while rem > 0:
            zd = ((m[zi + 1] if sign > 0 else m[zi]) - z) / cos_t
            sd = rem / sin_t
            if zd <= sd: d2, d3, mv = zd * sin_t, zd, sign
            else: d2, d3, mv = rem, sd, 0
            if d3 > 1e-8: out.append((d3, ext_fsr[ext_off[e] + zi]))
            z += d3 * cos_t; rem -= d2; zi += mv
            if zi < 0 or zi >= nf: done = True; break
        if done: break
        s += 1
    return out


def test_ragged_per_fsr_axial_meshes():
    """struct ExtrudedFSR with a different mesh per extruded region (src/Geometry.h:84-107): one 2D
    track over three extruded FSRs with 1, 4 and 2 axial cells, tracks going up and down from
    several start points - through the C ABI, against the plain-Python restatement"""
    lib = capi.load()
    rng = np.random.default_rng(7)
    seg_len = np.array([0.7, 1.1, 0.4, 0.9, 1.3], "f8")
    seg_ext = np.array([0, 1, 2, 1, 0], "i4")
    ext_off = np.array([0, 1, 5, 7], "i8")
    mesh = np.array([-1.0, 2.0,   -1.0, -0.5, 0.3, 1.1, 2.0,   -1.0, 0.75, 2.0], "f8")
    ext_fsr = np.array([6, 0, 1, 2, 3, 4, 5], "i4")
    A, P = 4, 4
    theta = np.array([[0.6, 1.2, np.pi - 1.2, np.pi - 0.6]] * 2).ravel()
    nt = 24
    trk_2d = np.zeros(nt, "i4")
    azim = np.zeros(nt, "i4")
    polar = rng.integers(0, P, nt).astype("i4")
    l0 = rng.uniform(0, 4.0, nt)
    z0 = np.where(polar < 2, rng.uniform(-1.0, 1.5, nt), rng.uniform(-0.5, 2.0, nt))
    l0[:4] = 0.0; z0[0] = -1.0; polar[0] = 0; z0[1] = 2.0; polar[1] = 3; z0[2] = 0.3; polar[2] = 1; z0[3] = 0.3; polar[3] = 2
    nxt = np.full(nt, -1, "i8"); flags = np.zeros(nt, "u1"); bc = np.zeros(nt, "u1")
    cfg = Config(num_groups=2, num_azim=A, num_polar=P, solve_3d=1, n_tracks=nt, n_segments=0, n_fsrs=7,
                 n_materials=1, device=0)
    h = C.c_void_p()
    check(lib.b200_create(C.byref(cfg), C.byref(h)))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    try:
        check(lib.b200_upload_otf_geometry(h, 1, 5, p(seg_len), p(seg_ext), p(np.array([0, 5], "i8")), 3,
                                           p(ext_off), p(mesh), p(ext_fsr), 0, p(theta)))
        ns = C.c_int64()
        check(lib.b200_upload_tracks_otf(h, p(trk_2d), p(l0), p(z0), p(azim), p(polar), p(nxt), p(nxt), p(flags),
                                         p(bc), p(bc), C.byref(ns)))
        length, fsr, off = np.empty(ns.value, "f8"), np.empty(ns.value, "i4"), np.empty(nt + 1, "i8")
        check(lib.b200_get_segments(h, p(length), p(fsr), ns.value, p(off)))
    finally:
        lib.b200_destroy(h)
    total = 0
    for t in range(nt):
        th = theta[polar[t]]
        want = _py_trace(seg_len, seg_ext, ext_off, mesh, ext_fsr, l0[t], z0[t], np.cos(th), np.sin(th))
        got = list(zip(length[off[t]:off[t + 1]], fsr[off[t]:off[t + 1]]))
        assert [f for _, f in got] == [f for _, f in want], t
        np.testing.assert_allclose([l for l, _ in got], [l for l, _ in want], rtol=1e-13)
        total += len(want)
    assert total == ns.value and total > nt
