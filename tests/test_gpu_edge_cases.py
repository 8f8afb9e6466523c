"""Edge cases of the flat layout on the GPU, against the oracle: ragged random track sets
(tests/ragged.py) with empty tracks, group counts that do not fill the lane map, one to
three polar angles, self links, PERIODIC / REFLECTIVE / VACUUM mixes, FSRs no track crosses,
tiny and huge optical lengths.  The reference covers the same ground piecemeal
(tests/test_periodic, test_vacuum_bcs, test_1g_mgxs ...)."""
import numpy as np
import pytest

from ragged import make_ragged
from openmoc_b200.capi import PRECISION_DOUBLE, PRECISION_MIXED
from oracle.oracle_py import OracleSolver

pytestmark = pytest.mark.gpu

SHAPES = [  # G, NP, 3D
    (1, 1, False), (1, 3, False), (2, 2, False), (3, 1, False), (5, 3, False), (7, 3, False),
    (8, 2, False), (10, 3, False), (33, 2, False), (70, 3, False),
    (1, 1, True), (7, 1, True), (10, 1, True), (33, 1, True), (70, 1, True)]


def pair(ft, **kw):
    from openmoc_b200.solver import B200Solver
    return B200Solver(ft, **kw), OracleSolver(ft)


@pytest.mark.parametrize("G,NP,d3", SHAPES)
def test_ragged_sweep_and_iterations(G, NP, d3):
    ft = make_ragged(G=G, NP=NP, solve_3d=d3, seed=100 + G + NP)
    gpu, cpu = pair(ft)
    rng = np.random.default_rng(G)
    q = rng.uniform(0.0, 1.0, ft.n_fsrs * G)
    psi = rng.uniform(0.0, 1.0, ft.n_tracks * 2 * ft.fluxes_per_track).astype(np.float32)
    for s in (gpu, cpu):
        s.zeroTrackFluxes()
    gpu.setFSRSources(q); cpu.setSources(q)
    gpu.setStartFluxes(psi); cpu.setStartFluxes(psi)
    for _ in range(3):                       # three sweeps: both psi buffers and the hand-off
        gpu.transportSweep(); cpu.transportSweep()
        pg, pc = gpu.getFluxes(), cpu.getFluxes()
        np.testing.assert_allclose(pg, pc, rtol=1e-6, atol=1e-9 * np.abs(pc).max())
        np.testing.assert_allclose(gpu.getStartFluxes(), cpu.getStartFluxes(), rtol=1e-6, atol=1e-9)
    # 40 source iterations from the same start (tolerance 0: fixed count)
    gpu, cpu = pair(ft)
    gpu.computeEigenvalue(40, 1e-30); cpu.computeEigenvalue(40, 1e-30)
    assert abs(gpu.getKeff() - cpu.getKeff()) / cpu.getKeff() < 1e-8
    pc = cpu.getFluxes()
    np.testing.assert_allclose(gpu.getFluxes(), pc, rtol=1e-6, atol=1e-9 * pc.max())


def test_empty_tracks_pass_the_flux_through():
    ft = make_ragged(G=7, NP=3, seed=3, vacuum_fraction=0.0)
    gpu, _ = pair(ft)
    F = ft.fluxes_per_track
    nseg = np.diff(ft.trk_seg_offset)
    empty = np.nonzero(nseg == 0)[0]
    assert empty.size > 3
    psi = np.random.default_rng(0).uniform(0, 1, ft.n_tracks * 2 * F).astype(np.float32)
    gpu.zeroTrackFluxes(); gpu.setStartFluxes(psi); gpu.setFSRSources(np.zeros(ft.n_fsrs * 7))
    gpu.transportSweep()
    out = gpu.getStartFluxes().reshape(ft.n_tracks, 2, F)
    src = psi.reshape(ft.n_tracks, 2, F)
    for t in empty:
        for d, (nx, bit) in enumerate(((ft.trk_next_fwd[t], 1), (ft.trk_next_bwd[t], 2))):
            slot_dir = 0 if ft.trk_flags[t] & bit else 1
            np.testing.assert_array_equal(out[nx, slot_dir], src[t, d])


@pytest.mark.parametrize("vac", [0.0, 1.0])
def test_all_linked_and_all_vacuum(vac):
    ft = make_ragged(G=7, NP=3, seed=11, vacuum_fraction=vac)
    if vac == 1.0:      # slot 0 is kept reflective by the generator; cut it as well
        ft.arrays["trk_bc_fwd"][0] = 0
        ft.arrays["trk_next_fwd"][0] = -1
    gpu, cpu = pair(ft)
    gpu.computeEigenvalue(30, 1e-30); cpu.computeEigenvalue(30, 1e-30)
    assert abs(gpu.getKeff() - cpu.getKeff()) / cpu.getKeff() < 1e-8
    pc = cpu.getFluxes()
    np.testing.assert_allclose(gpu.getFluxes(), pc, rtol=1e-6, atol=1e-9 * pc.max())


@pytest.mark.parametrize("n_tracks,max_segments", [(1, 5), (2, 0), (3, 1), (225, 3)])
def test_tiny_problems(n_tracks, max_segments):
    ft = make_ragged(G=3, NP=2, seed=n_tracks, n_tracks=n_tracks, n_fsrs=4, max_segments=max_segments,
                     long_track=0, vacuum_fraction=0.2)
    gpu, cpu = pair(ft)
    gpu.computeEigenvalue(20, 1e-30); cpu.computeEigenvalue(20, 1e-30)
    assert abs(gpu.getKeff() - cpu.getKeff()) / cpu.getKeff() < 1e-8
    pc = cpu.getFluxes()
    np.testing.assert_allclose(gpu.getFluxes(), pc, rtol=1e-6, atol=1e-9 * pc.max())


@pytest.mark.parametrize("G,NP,d3", [(5, 3, False), (33, 2, False), (10, 1, True)])
def test_ragged_mixed_and_deterministic(G, NP, d3):
    ft = make_ragged(G=G, NP=NP, solve_3d=d3, seed=7)
    _, cpu = pair(ft)
    cpu.computeEigenvalue(40, 1e-30)
    pc = cpu.getFluxes()
    from openmoc_b200.solver import B200Solver
    mixed = B200Solver(ft, precision=PRECISION_MIXED)
    mixed.computeEigenvalue(40, 1e-30)
    assert abs(mixed.getKeff() - cpu.getKeff()) / cpu.getKeff() < 1e-5          # 1 pcm
    np.testing.assert_allclose(mixed.getFluxes(), pc, rtol=1e-4, atol=1e-5 * pc.max())
    runs = []
    for _ in range(2):
        det = B200Solver(ft, deterministic=True)
        det.computeEigenvalue(40, 1e-30)
        runs.append((det.getKeff(), det.getFluxes()))
    assert runs[0][0] == runs[1][0] and np.array_equal(runs[0][1], runs[1][1])
    assert abs(runs[0][0] - cpu.getKeff()) / cpu.getKeff() < 1e-7
    np.testing.assert_allclose(runs[0][1], pc, rtol=1e-5, atol=1e-7 * pc.max())


def test_shards_of_a_ragged_problem_add_up():
    """chain partition of a ragged link graph: per-shard tallies sum to the whole sweep"""
    from openmoc_b200.partition import partition_by_chain
    from openmoc_b200.solver import B200Solver
    ft = make_ragged(G=7, NP=3, seed=5, n_tracks=301)
    whole = B200Solver(ft)
    q = np.random.default_rng(1).uniform(0, 1, ft.n_fsrs * 7)
    whole.zeroTrackFluxes(); whole.setFSRSources(q); whole.flattenFSRFluxes(0.0)
    whole.transportSweep()
    total = np.zeros(ft.n_fsrs * 7)
    n_tracks = 0
    for sub in partition_by_chain(ft, 3):
        n_tracks += sub.n_tracks
        s = B200Solver(sub)
        s.zeroTrackFluxes(); s.setFSRSources(q); s.transportSweep()
        total += s.getFluxes()
    assert n_tracks == ft.n_tracks
    np.testing.assert_allclose(total, whole.getFluxes(), rtol=1e-11, atol=1e-13)


def test_two_ends_feeding_one_slot_is_rejected():
    """the sweep writes hand-offs concurrently: a link table that is not one-to-one must fail loudly"""
    from openmoc_b200.capi import B200Error
    from openmoc_b200.solver import B200Solver
    ft = make_ragged(G=2, NP=1, seed=1, n_tracks=20, vacuum_fraction=0.0)
    a = ft.arrays
    a["trk_next_fwd"][1] = a["trk_next_fwd"][2]
    a["trk_flags"][1] = (a["trk_flags"][1] & ~np.uint8(1)) | (a["trk_flags"][2] & np.uint8(1))
    with pytest.raises(B200Error, match="same slot"):
        B200Solver(ft)
