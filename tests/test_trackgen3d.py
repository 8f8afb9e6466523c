"""The synthetic 3D track generator (z-stacks over the 2D tracks: csrc/trackgen.cpp, synth.make_tracks_3d)
against the reference's TrackGenerator3D: a full 3D track file dumped from the reference
(tests/golden/lattice3d_7g.b2trk) and the signatures of four larger dumps
(tests/golden/trackgen3d_signatures.json, made by make_trackgen3d_signatures.py)."""
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import load_case
from openmoc_b200.synth import (make_tracks_3d, QUAD_TY, QUAD_EQUAL_ANGLE, QUAD_GAUSS_LEGENDRE,
                                QUAD_EQUAL_WEIGHT)
from oracle.oracle_py import OracleSolver, FISSION_SOURCE

HERE = os.path.dirname(os.path.abspath(__file__))
QUADS = {"ty": QUAD_TY, "equal-angle": QUAD_EQUAL_ANGLE, "gl": QUAD_GAUSS_LEGENDRE, "equal-weight": QUAD_EQUAL_WEIGHT}
SIGS = json.load(open(os.path.join(HERE, "golden", "trackgen3d_signatures.json")))

_spec = importlib.util.spec_from_file_location("mk3d", os.path.join(HERE, "golden", "make_trackgen3d_signatures.py"))
_mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mk)


def test_matches_reference_3d_track_file():
    """test_forward_3D_lattice shape: every link, flag, BC and segment offset bit-exact, lengths to 1e-12,
    FSR ids equal up to the reference's (hash-map ordered) numbering."""
    ref, _ = load_case("lattice3d_7g")
    ft = make_tracks_3d("simple-lattice", num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9)
    ft.validate()
    a, b = ref.arrays, ft.arrays
    assert (ft.n_tracks, ft.n_segments, ft.n_fsrs) == (ref.n_tracks, ref.n_segments, ref.n_fsrs)
    for k in ("trk_seg_offset", "trk_azim", "trk_polar", "trk_xy", "trk_next_fwd", "trk_next_bwd", "trk_flags",
              "trk_bc_fwd", "trk_bc_bwd"):
        assert np.array_equal(a[k], b[k]), k
    np.testing.assert_allclose(b["seg_length"], a["seg_length"], rtol=0, atol=1e-12)
    pairs = np.unique(b["seg_fsr"].astype(np.int64) * (ref.n_fsrs + 1) + a["seg_fsr"])
    assert pairs.size == ref.n_fsrs                       # a bijection between the two numberings
    perm = np.empty(ref.n_fsrs, dtype=np.int64)
    perm[pairs // (ref.n_fsrs + 1)] = pairs % (ref.n_fsrs + 1)
    np.testing.assert_allclose(b["fsr_volume"], a["fsr_volume"][perm], rtol=1e-12)
    assert np.array_equal(b["fsr_mat"], a["fsr_mat"][perm])
    for k in ("quad_weight", "quad_sin_theta", "quad_polar_spacing", "quad_polar_weight", "trk_theta", "trk_phi"):
        np.testing.assert_allclose(b[k], a[k], rtol=1e-14)


@pytest.mark.parametrize("name", sorted(SIGS))
def test_matches_reference_signature(name):
    """vacuum sides, 1-3 axial layers, four polar quadratures, 4-16 azimuthal angles"""
    model, az, sp, pol, zs, nax, quad, _ = SIGS[name]["case"]
    want = SIGS[name]["sig"]
    ft = make_tracks_3d(model, num_azim=az, spacing=sp, num_polar=pol, z_spacing=zs, n_axial=nax,
                        polar_quad=QUADS[quad])
    got = _mk.signature(ft)
    for k in ("n_tracks", "n_segments", "n_fsrs", "fsr_mat", *_mk.INT_KEYS):
        assert got[k] == want[k], k
    assert abs(got["seg_length_sum"] - want["seg_length_sum"]) < 1e-9 * want["seg_length_sum"]
    assert abs(got["fsr_volume_sum"] - want["fsr_volume_sum"]) < 1e-9 * want["fsr_volume_sum"]
    np.testing.assert_allclose(got["seg_length_head"], want["seg_length_head"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(got["fsr_volume_head"], want["fsr_volume_head"], rtol=1e-8)
    np.testing.assert_allclose(got["quad_weight"], want["quad_weight"], rtol=1e-13)
    np.testing.assert_allclose(got["quad_sin_theta"], want["quad_sin_theta"], rtol=1e-14)


def test_hand_off_table_is_one_to_one():
    """every (track, direction) start slot is fed by at most one track end (b200_finalize requires it)"""
    ft = make_tracks_3d("c5g7-2d", num_azim=4, spacing=1.0, num_polar=4, z_spacing=8.0, n_axial=2,
                        polar_quad=QUAD_EQUAL_ANGLE)
    a = ft.arrays
    slots = []
    for nxt, bc, bit in ((a["trk_next_fwd"], a["trk_bc_fwd"], 1), (a["trk_next_bwd"], a["trk_bc_bwd"], 2)):
        linked = bc != 0
        fwd = (a["trk_flags"][linked] & bit) != 0
        slots.append(nxt[linked] * 2 + np.where(fwd, 0, 1))
    slots = np.concatenate(slots)
    assert slots.min() >= 0 and np.unique(slots).size == slots.size


def test_volume_is_conserved_3d():
    ft = make_tracks_3d("c5g7-2d", num_azim=4, spacing=0.8, num_polar=2, z_spacing=5.0, n_axial=4)
    vol = 64.26 ** 3
    assert abs(ft.arrays["fsr_volume"].sum() - vol) / vol < 1e-12


def test_device_tracer_inputs_describe_the_same_tracks():
    """expand=False hands over 2D segments + axial mesh + per-track start data instead of 3D segments"""
    kw = dict(num_azim=4, spacing=0.5, num_polar=2, z_spacing=1.5, n_axial=3)
    full = make_tracks_3d("simple-lattice", fsr_numbering="lattice", **kw)
    lean = make_tracks_3d("simple-lattice", expand=False, **kw)
    assert lean.n_segments == 0 and lean.n_tracks == full.n_tracks and lean.n_fsrs == full.n_fsrs
    for k in ("trk_next_fwd", "trk_next_bwd", "trk_flags", "trk_2d", "trk_l0", "trk_start", "z_mesh", "seg2d_length"):
        assert np.array_equal(lean.arrays[k], full.arrays[k]), k
    assert lean.arrays["z_mesh"].size == 4 and lean.arrays["trk_l0"].min() >= -1e-12


def test_synthetic_3d_lattice_reproduces_reference_eigenvalue():
    """tests/test_forward_3D_lattice/results_true.dat through the oracle on synthetic tracks"""
    _, res = load_case("lattice3d_7g")
    ft = make_tracks_3d("simple-lattice", num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9)
    s = OracleSolver(ft)
    n = s.computeEigenvalue(500, 1e-5, FISSION_SOURCE)
    assert n == res["iterations"] and abs(s.getKeff() - res["keff"]) < 1e-9
