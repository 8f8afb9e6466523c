"""Property tests (hypothesis) of the track partitions on random ragged link tables: every
partition keeps all tracks and segments, stays self-consistent, and - simulated rank by rank
with the oracle - reproduces the unpartitioned sweep."""
import numpy as np
from hypothesis import given, settings, strategies as st

from ragged import make_ragged
from openmoc_b200.partition import (partition_by_chain, partition_by_track, track_components,
                                    assign_tracks)
from openmoc_b200.trackfile import REFLECTIVE, PERIODIC
from oracle.oracle_py import OracleSolver


def _sweep(ft, q, psi=None):
    o = OracleSolver(ft)
    o.zeroTrackFluxes(); o.setSources(q)
    if psi is not None:
        o.setStartFluxes(psi)
    o.transportSweep()
    return o.getFluxes(), o.getStartFluxes()


@settings(max_examples=12, deadline=None)
@given(seed=st.integers(0, 10_000), world=st.integers(2, 6), n_tracks=st.integers(8, 80),
       vac=st.sampled_from([0.0, 0.3, 0.8]))
def test_track_partition_reproduces_one_sweep(seed, world, n_tracks, vac):
    ft = make_ragged(G=2, NP=2, seed=seed, n_tracks=n_tracks, n_fsrs=7, vacuum_fraction=vac, long_track=40)
    F = ft.fluxes_per_track
    rng = np.random.default_rng(seed)
    q = rng.uniform(0, 1, ft.n_fsrs * 2)
    psi = rng.uniform(0, 1, ft.n_tracks * 2 * F).astype(np.float32)
    phi_ref, psi_ref = _sweep(ft, q, psi)
    owner = assign_tracks(ft, world)
    parts = partition_by_track(ft, world, owner=owner)
    assert sum(s.n_segments for s, _ in parts) == ft.n_segments
    phi = np.zeros_like(phi_ref)
    outs = []
    for r, (sub, plan) in enumerate(parts):
        sub.validate()
        ids = np.nonzero(owner == r)[0]
        local = np.zeros((sub.n_tracks, 2, F), dtype=np.float32)
        local[:ids.size] = psi.reshape(-1, 2, F)[ids]                 # ghosts start empty
        p, out = _sweep(sub, q, local.ravel())
        phi += p
        outs.append(out.reshape(-1, F))
    np.testing.assert_allclose(phi, phi_ref, rtol=1e-12, atol=1e-14)
    # hand the ghost slots over exactly as exchange_boundary_fluxes does and compare all start fluxes
    outbox = {}
    for r, (_, plan) in enumerate(parts):
        off = plan.ghost0
        for qk in range(world):
            outbox[(r, qk)] = outs[r][off:off + plan.send_counts[qk]].copy()
            off += plan.send_counts[qk]
    full = np.zeros((ft.n_tracks, 2, F), dtype=np.float32)
    for r, (sub, plan) in enumerate(parts):
        ids = np.nonzero(owner == r)[0]
        recv = np.concatenate([outbox[(s_, r)] for s_ in range(world)]) if plan.n_recv else np.zeros((0, F), np.float32)
        outs[r][plan.recv_slots] = recv
        full[ids] = outs[r][:2 * ids.size].reshape(-1, 2, F)
    np.testing.assert_array_equal(full.ravel(), psi_ref)


@settings(max_examples=12, deadline=None)
@given(seed=st.integers(0, 10_000), n_tracks=st.integers(10, 120), vac=st.sampled_from([0.2, 0.6, 1.0]))
def test_chain_partition_is_closed_and_additive(seed, n_tracks, vac):
    ft = make_ragged(G=1, NP=1, seed=seed, n_tracks=n_tracks, n_fsrs=5, vacuum_fraction=vac, long_track=0)
    labels = track_components(ft)
    n_comp = int(labels.max()) + 1
    # links never leave a component
    a = ft.arrays
    for d in ("fwd", "bwd"):
        linked = np.nonzero((a["trk_bc_" + d] == REFLECTIVE) | (a["trk_bc_" + d] == PERIODIC))[0]
        assert np.all(labels[linked] == labels[a["trk_next_" + d][linked]])
    world = min(3, n_comp)
    parts = partition_by_chain(ft, world)
    assert sum(p.n_tracks for p in parts) == ft.n_tracks
    q = np.random.default_rng(seed).uniform(0, 1, ft.n_fsrs)
    total = sum(_sweep(p, q)[0] for p in parts)
    np.testing.assert_allclose(total, _sweep(ft, q)[0], rtol=1e-12, atol=1e-14)
