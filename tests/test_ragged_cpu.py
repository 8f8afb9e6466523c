"""CPU checks on the ragged track sets used by the GPU edge-case tests: the oracle's
behaviour on them and the host-side partitioning."""
import numpy as np

from ragged import make_ragged
from openmoc_b200.partition import partition_by_chain, track_components
from openmoc_b200.trackfile import REFLECTIVE, PERIODIC
from oracle.oracle_py import OracleSolver


def test_oracle_passes_flux_through_empty_tracks_and_zeroes_vacuum_targets():
    ft = make_ragged(G=5, NP=2, seed=3)
    o = OracleSolver(ft)
    F = ft.fluxes_per_track
    psi = np.random.default_rng(0).uniform(0.5, 1, ft.n_tracks * 2 * F).astype(np.float32)
    o.zeroTrackFluxes(); o.setStartFluxes(psi); o.setSources(np.zeros(ft.n_fsrs * 5))
    o.transportSweep()
    out = o.getStartFluxes().reshape(ft.n_tracks, 2, F)
    src = psi.reshape(ft.n_tracks, 2, F)
    nseg = np.diff(ft.trk_seg_offset)
    written = np.zeros((ft.n_tracks, 2), dtype=bool)
    for t in range(ft.n_tracks):
        for d, (nx, bit, bc) in enumerate(((ft.trk_next_fwd[t], 1, ft.trk_bc_fwd[t]),
                                           (ft.trk_next_bwd[t], 2, ft.trk_bc_bwd[t]))):
            if bc not in (REFLECTIVE, PERIODIC):
                continue
            sd = 0 if ft.trk_flags[t] & bit else 1
            written[nx, sd] = True
            if nseg[t] == 0:
                np.testing.assert_array_equal(out[nx, sd], src[t, d])
            else:
                assert np.all(out[nx, sd] <= src[t, d])        # q = 0: pure attenuation
    # slots no track hands off to keep their incoming flux (CPUSolver never rewrites them)
    np.testing.assert_array_equal(out[~written], src[~written])


def test_chain_partition_of_ragged_links_is_closed_and_complete():
    ft = make_ragged(G=2, NP=1, seed=9, n_tracks=400, vacuum_fraction=0.5)
    labels = track_components(ft)
    n_comp = labels.max() + 1
    assert n_comp >= 2
    world = min(4, n_comp)
    subs = partition_by_chain(ft, world)
    assert sum(s.n_tracks for s in subs) == ft.n_tracks
    assert sum(s.n_segments for s in subs) == ft.n_segments
    for s in subs:
        s.validate()
    # the shards' sweeps add up to the whole (oracle)
    q = np.random.default_rng(2).uniform(0, 1, ft.n_fsrs * 2)
    whole = OracleSolver(ft)
    whole.zeroTrackFluxes(); whole.setSources(q); whole.transportSweep()
    total = np.zeros_like(q)
    for s in subs:
        o = OracleSolver(s)
        o.zeroTrackFluxes(); o.setSources(q); o.transportSweep()
        total += o.getFluxes()
    np.testing.assert_allclose(total, whole.getFluxes(), rtol=1e-12, atol=1e-14)
