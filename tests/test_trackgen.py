"""The synthetic track generator (openmoc_b200/synth.py, csrc/trackgen.cpp) against
tracks dumped from the unmodified reference (tests/golden/*.b2trk)."""
import numpy as np
import pytest

from conftest import load_case
from openmoc_b200.synth import make_tracks
from oracle.oracle_py import OracleSolver, FISSION_SOURCE

CASES = [("pin_cell", "pin-cell", dict(num_azim=4, spacing=0.1)),
         ("simple_lattice", "simple-lattice", dict(num_azim=4, spacing=0.12)),
         ("c5g7_2d_coarse", "c5g7-2d", dict(num_azim=4, spacing=0.5))]


@pytest.mark.parametrize("fixture,model,kw", CASES)
def test_matches_reference_tracks(fixture, model, kw):
    ref, _ = load_case(fixture)
    ft = make_tracks(model, **kw)
    ft.validate()
    a, b = ref.arrays, ft.arrays
    assert (ft.n_tracks, ft.n_segments, ft.n_fsrs) == (ref.n_tracks, ref.n_segments, ref.n_fsrs)
    for k in ("trk_seg_offset", "trk_azim", "trk_xy", "trk_next_fwd", "trk_next_bwd", "trk_flags",
              "trk_bc_fwd", "trk_bc_bwd", "seg_fsr"):
        assert np.array_equal(a[k], b[k]), k           # integer data: bit-exact
    np.testing.assert_allclose(b["seg_length"], a["seg_length"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(b["fsr_volume"], a["fsr_volume"], rtol=1e-8, atol=1e-10)  # reference ray tracer nudges points by TINY_MOVE
    np.testing.assert_allclose(b["quad_weight"], a["quad_weight"], rtol=1e-14)
    np.testing.assert_allclose(b["quad_sin_theta"], a["quad_sin_theta"], rtol=1e-15)
    np.testing.assert_allclose(b["trk_phi"], a["trk_phi"], rtol=1e-15)
    G = ref.num_groups
    for k in ("mat_sigma_t", "mat_nu_sigma_f", "mat_chi"):
        np.testing.assert_array_equal(a[k].reshape(-1, G)[a["fsr_mat"]], b[k].reshape(-1, G)[b["fsr_mat"]])
    sa = a["mat_sigma_s"].reshape(-1, G * G)[a["fsr_mat"]]
    sb = b["mat_sigma_s"].reshape(-1, G * G)[b["fsr_mat"]]
    np.testing.assert_array_equal(sa, sb)


def test_bench_shapes_have_the_survey_counts():
    # SURVEY.md section 8 / BASELINE.md: measured with the reference at (128 azim, 0.01 cm)
    ft = make_tracks("simple-lattice", num_azim=128, spacing=0.01)
    assert (ft.n_tracks, ft.n_segments, ft.n_fsrs) == (32656, 763256, 512)
    ft = make_tracks("pin-cell", num_azim=128, spacing=0.01)
    assert (ft.n_tracks, ft.n_segments, ft.n_fsrs) == (32656, 58320, 2)


def test_synthetic_pin_cell_reproduces_reference_eigenvalue():
    ft = make_tracks("pin-cell", num_azim=4, spacing=0.1)
    s = OracleSolver(ft)
    n = s.computeEigenvalue(500, 1e-5, FISSION_SOURCE)
    assert n == 261 and abs(s.getKeff() - 1.0466609855939686) < 1e-10


def test_volume_is_conserved():
    ft = make_tracks("c5g7-2d", num_azim=8, spacing=0.2)
    assert abs(ft.arrays["fsr_volume"].sum() - 64.26 ** 2) / 64.26 ** 2 < 1e-12


def test_lattice_numbering_keeps_every_region():
    a = make_tracks("c5g7-2d", num_azim=4, spacing=0.5, fsr_numbering="lattice")
    b = make_tracks("c5g7-2d", num_azim=4, spacing=0.5)
    assert a.n_fsrs >= b.n_fsrs and a.n_segments == b.n_segments
    np.testing.assert_allclose(np.sort(a.arrays["fsr_volume"])[a.n_fsrs - b.n_fsrs:], np.sort(b.arrays["fsr_volume"]), rtol=1e-12)


def test_bad_arguments():
    with pytest.raises(ValueError):
        make_tracks("no-such-model")
    with pytest.raises(ValueError):
        make_tracks("pin-cell", num_azim=6)
    with pytest.raises(ValueError):
        make_tracks("pin-cell", num_polar=8)      # TY has 2, 4, 6


def test_linear_source_data_match_a_reference_cpulssolver_dump():
    """FSR centroids, centroid-relative segment starting points and the quadrature factors of the
    synthetic generator against the track file dumped after a reference CPULSSolver run; the LS oracle
    (pinned to the reference's LS goldens) then gives the reference's k_eff on the synthetic tracks."""
    import numpy as np
    from oracle.oracle_py import OracleSolver
    ref, rj = load_case("simple_lattice_ls")
    ft = make_tracks("simple-lattice", num_azim=4, spacing=0.12, linear_source=True)
    assert np.array_equal(ft.arrays["seg_fsr"], ref.arrays["seg_fsr"])
    for k in ("quad_azim_spacing", "quad_azim_weight", "quad_polar_weight", "quad_polar_spacing"):
        np.testing.assert_allclose(ft.arrays[k], ref.arrays[k], rtol=1e-13, atol=0)
    np.testing.assert_allclose(ft.arrays["fsr_centroid"], ref.arrays["fsr_centroid"], atol=1e-9)
    np.testing.assert_allclose(ft.arrays["seg_start"], ref.arrays["seg_start"], atol=5e-8)
    o = OracleSolver(ft, linear_source=True)
    n = o.computeEigenvalue(500, 1e-5)
    assert n == rj["iterations"]
    assert abs(o.getKeff() - rj["keff"]) * 1e5 < 1e-3
