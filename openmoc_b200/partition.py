"""Decomposition of flattened tracks across GPUs without angular-flux exchange.

Boundary hand-offs (`trk_next_fwd/bwd`) define a graph over the tracks; its connected
components - the cyclic track chains of the reference's cyclic tracking, or open paths
between two vacuum ends - never hand a flux to another component.  Sharding whole
components across ranks therefore needs no psi exchange at all; only the FSR tally is
summed (one all-reduce per sweep).  FSR, material and quadrature tables are replicated.

Every chain lives inside one azimuthal *pair* {a, A/2-1-a}: reflective links pair
azimuthal index a with A/2-1-a (src/TrackGenerator.cpp:1092,1169-1217), periodic links
stay inside a, and 3D tracks add the polar complement on the same 2D track
(src/TrackGenerator3D.cpp:2054-2057).  `partition_by_azim_pair` (the north-star
partition) keeps whole pairs together; `partition_by_chain` balances at chain
granularity, which also works when there are fewer pairs than GPUs.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .trackfile import FlatTracks, REFLECTIVE, PERIODIC


def assign_pairs(num_azim: int, seg_per_azim: np.ndarray, world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of the A/4 azimuthal pairs to
    `world` ranks, balanced by segment count.  Returns per-rank lists of azim indices."""
    a2 = num_azim // 2
    n_pairs = num_azim // 4
    if world > n_pairs:
        raise ValueError(f"{world} ranks but only {n_pairs} azimuthal pairs (num_azim={num_azim}); "
                         "use fewer GPUs or more azimuthal angles")
    load = [(int(seg_per_azim[a] + seg_per_azim[a2 - 1 - a]), a) for a in range(n_pairs)]
    load.sort(key=lambda x: (-x[0], x[1]))
    totals = [0] * world
    owned: List[List[int]] = [[] for _ in range(world)]
    for w, a in load:
        r = min(range(world), key=lambda i: (totals[i], i))
        totals[r] += w
        owned[r] += [a, a2 - 1 - a]
    return [sorted(o) for o in owned]


def track_components(ft: FlatTracks) -> np.ndarray:
    """Connected components of the boundary hand-off graph: label per track."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    a = ft.arrays
    n = ft.n_tracks
    src, dst = [], []
    for d in ("fwd", "bwd"):
        bc = a["trk_bc_" + d]
        linked = (bc == REFLECTIVE) | (bc == PERIODIC)
        ids = np.nonzero(linked)[0]
        src.append(ids)
        dst.append(a["trk_next_" + d][ids].astype(np.int64))
    src, dst = np.concatenate(src), np.concatenate(dst)
    g = coo_matrix((np.ones(src.size, dtype=np.int8), (src, dst)), shape=(n, n))
    _, labels = connected_components(g, directed=False)
    return labels


def partition_by_chain(ft: FlatTracks, world: int) -> List[FlatTracks]:
    """Shard whole track chains (connected components of the link graph) across `world`
    ranks, longest-processing-time first by segment count."""
    labels = track_components(ft)
    nseg = np.diff(ft.arrays["trk_seg_offset"].astype(np.int64))
    n_comp = int(labels.max()) + 1 if labels.size else 0
    if n_comp < world:
        raise ValueError(f"{world} ranks but only {n_comp} independent track chains")
    load = np.bincount(labels, weights=nseg + 1e-3, minlength=n_comp)
    order = np.argsort(-load, kind="stable")
    totals = np.zeros(world)
    owner = np.empty(n_comp, dtype=np.int64)
    for c in order:
        r = int(np.argmin(totals))
        owner[c] = r
        totals[r] += load[c]
    track_owner = owner[labels]
    return [_extract(ft, np.nonzero(track_owner == r)[0]) for r in range(world)]


def partition_by_azim_pair(ft: FlatTracks, world: int) -> List[FlatTracks]:
    a = ft.arrays
    nseg = np.diff(a["trk_seg_offset"].astype(np.int64))
    azim = a["trk_azim"].astype(np.int64)
    seg_per_azim = np.bincount(azim, weights=nseg, minlength=ft.num_azim // 2)
    owned = assign_pairs(ft.num_azim, seg_per_azim, world)
    return [_extract(ft, np.nonzero(np.isin(azim, owned[rank]))[0]) for rank in range(world)]


def _extract(ft: FlatTracks, ids: np.ndarray) -> FlatTracks:
    """The sub-problem made of tracks `ids` (closed under links), renumbered 0..n-1."""
    a = ft.arrays
    off = a["trk_seg_offset"].astype(np.int64)
    nseg = np.diff(off)
    per_track = ("trk_azim", "trk_polar", "trk_xy", "trk_flags", "trk_bc_fwd", "trk_bc_bwd",
                 "trk_phi", "trk_theta")
    per_seg = ("seg_length", "seg_fsr", "seg_mat", "seg_cmfd_fwd", "seg_cmfd_bwd")
    new_id = np.full(ft.n_tracks, -1, dtype=np.int64)
    new_id[ids] = np.arange(ids.size)
    sub = FlatTracks(num_groups=ft.num_groups, num_azim=ft.num_azim, num_polar=ft.num_polar,
                     solve_3d=ft.solve_3d, fluxes_per_track=ft.fluxes_per_track,
                     n_tracks=int(ids.size), n_segments=int(nseg[ids].sum()), n_fsrs=ft.n_fsrs,
                     n_materials=ft.n_materials)
    # segment gather indices, track by track, forward order preserved
    lens = nseg[ids]
    new_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    seg_idx = (np.repeat(off[ids] - new_off[:-1], lens) + np.arange(new_off[-1])).astype(np.int64)
    b = sub.arrays
    b["trk_seg_offset"] = new_off
    for k in per_track:
        if k in a:
            b[k] = a[k][ids]
    for k in per_seg:
        if k in a and a[k].size == ft.n_segments:
            b[k] = a[k][seg_idx]
    if "seg_start" in a and a["seg_start"].size == 3 * ft.n_segments:
        b["seg_start"] = a["seg_start"].reshape(-1, 3)[seg_idx].ravel()
    for d in ("fwd", "bwd"):
        nxt = a["trk_next_" + d][ids].astype(np.int64)
        bc = a["trk_bc_" + d][ids]
        linked = (bc == REFLECTIVE) | (bc == PERIODIC)
        mapped = np.where(linked, new_id[np.clip(nxt, 0, ft.n_tracks - 1)], -1)
        if linked.any() and mapped[linked].min() < 0:
            raise ValueError("track partition is not closed under boundary links")
        b["trk_next_" + d] = mapped
    for k, v in a.items():
        if k.startswith(("quad_", "fsr_", "mat_")):
            b[k] = v
    return sub
