"""Decomposition of flattened tracks across GPUs: two partitions without angular-flux
exchange (pairs, chains) and one with it (`partition_by_track`, at the end of the file).

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


def track_load(ft: FlatTracks) -> np.ndarray:
    """Work per track: its segment count, or - for an on-the-fly 3D track set, whose segments
    only exist on the device - its length (segments per cm are nearly uniform over a deck)."""
    a = ft.arrays
    if "trk_seg_offset" in a and a["trk_seg_offset"].size == ft.n_tracks + 1 and ft.n_segments > 0:
        return np.diff(a["trk_seg_offset"].astype(np.int64)).astype(np.float64)
    if "trk_end" in a and a["trk_end"].size == 3 * ft.n_tracks:
        d = a["trk_end"].reshape(-1, 3) - a["trk_start"].reshape(-1, 3)
        return np.sqrt((d * d).sum(axis=1))
    return np.ones(ft.n_tracks)


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


def partition_by_chain(ft: FlatTracks, world: int, only: int = None) -> List[FlatTracks]:
    """Shard whole track chains (connected components of the link graph) across `world`
    ranks, longest-processing-time first by segment count.  `only`: build that rank's shard
    alone (None for the others)."""
    labels = track_components(ft)
    nseg = track_load(ft)
    n_comp = int(labels.max()) + 1 if labels.size else 0
    if n_comp < world:
        raise ValueError(f"{world} ranks but only {n_comp} independent track chains")
    load = np.bincount(labels, weights=nseg + 1e-3, minlength=n_comp)
    order = np.argsort(-load, kind="stable")
    owner = np.empty(n_comp, dtype=np.int64)
    if n_comp <= 200000:
        totals = np.zeros(world)
        for c in order:                      # longest processing time first
            r = int(np.argmin(totals))
            owner[c] = r
            totals[r] += load[c]
    else:
        # millions of chains (3D decks with vacuum sides): deal them in a snake over the ranks in
        # order of decreasing load - the same balance to within one chain, without a Python loop
        pos = np.arange(n_comp) % (2 * world)
        owner[order] = np.where(pos < world, pos, 2 * world - 1 - pos)
    track_owner = owner[labels]
    return [_extract(ft, np.nonzero(track_owner == r)[0]) if only is None or r == only else None
            for r in range(world)]


def partition_by_azim_pair(ft: FlatTracks, world: int, only: int = None) -> List[FlatTracks]:
    a = ft.arrays
    nseg = track_load(ft)
    azim = a["trk_azim"].astype(np.int64)
    seg_per_azim = np.bincount(azim, weights=nseg, minlength=ft.num_azim // 2)
    owned = assign_pairs(ft.num_azim, seg_per_azim, world)
    return [_extract(ft, np.nonzero(np.isin(azim, owned[rank]))[0]) if only is None or rank == only else None
            for rank in range(world)]


def _extract(ft: FlatTracks, ids: np.ndarray, closed: bool = True) -> FlatTracks:
    """The sub-problem made of tracks `ids`, renumbered 0..n-1.  `closed`: the set must be
    closed under links; otherwise links that leave it are returned as -2."""
    a = ft.arrays
    explicit = "trk_seg_offset" in a and a["trk_seg_offset"].size == ft.n_tracks + 1
    off = a["trk_seg_offset"].astype(np.int64) if explicit else np.zeros(ft.n_tracks + 1, dtype=np.int64)
    nseg = np.diff(off)
    per_track = ("trk_azim", "trk_polar", "trk_xy", "trk_flags", "trk_bc_fwd", "trk_bc_bwd",
                 "trk_phi", "trk_theta", "trk_2d", "trk_l0", "trk_lz")
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
    for k in ("trk_start", "trk_end"):
        if k in a and ft.n_tracks and a[k].size % ft.n_tracks == 0:
            b[k] = a[k].reshape(ft.n_tracks, -1)[ids].ravel()
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
            if closed:
                raise ValueError("track partition is not closed under boundary links")
            mapped = np.where(linked & (mapped < 0), -2, mapped)
        b["trk_next_" + d] = mapped
    for k, v in a.items():
        if k.startswith(("quad_", "fsr_", "mat_", "seg2d_", "trk2d_", "fsr2d_")) or k == "z_mesh":
            b[k] = v
    return sub


# ---------------------------------------------------------------------------------------
# Partition with angular-flux exchange
# ---------------------------------------------------------------------------------------
class ExchangePlan:
    """What one rank sends and receives after every sweep (all sizes in start-flux slots of
    F floats; a slot is (track, direction), index track*2 + direction).

    ghost0        first ghost slot: outgoing fluxes bound for other ranks are written by the
                  sweep itself into slots ghost0 .. ghost0 + n_send - 1, grouped by destination
    send_counts   [world] slots sent to each rank (contiguous ranges of the ghost slots)
    recv_counts   [world] slots received from each rank
    recv_slots    [n_recv] local slot every received flux is stored to, in arrival order
                  (source rank major, then the sender's order)
    """
    def __init__(self, ghost0, send_counts, recv_counts, recv_slots):
        self.ghost0 = int(ghost0)
        self.send_counts = [int(x) for x in send_counts]
        self.recv_counts = [int(x) for x in recv_counts]
        self.recv_slots = np.asarray(recv_slots, dtype=np.int64)
        self.n_send = sum(self.send_counts)
        self.n_recv = sum(self.recv_counts)


def assign_tracks(ft: FlatTracks, world: int) -> np.ndarray:
    """Owner rank per track: tracks sorted by segment count and dealt in a snake
    (0..w-1, w-1..0, ...), so every rank gets the same number of segments to within one
    track and the same mix of long and short tracks - for any number of ranks."""
    nseg = track_load(ft)
    order = np.argsort(-nseg, kind="stable")
    pos = np.arange(ft.n_tracks) % (2 * world)
    owner = np.empty(ft.n_tracks, dtype=np.int64)
    owner[order] = np.where(pos < world, pos, 2 * world - 1 - pos)
    return owner


def assign_blocks(ft: FlatTracks, world: int) -> np.ndarray:
    """Owner rank per track: contiguous blocks of the Track uid order (azimuthal angle, 2D track,
    polar angle, position in the z-stack), cut where the cumulative load crosses k/world.  Every rank
    then sweeps whole neighbouring z-stacks one after the other, exactly like a single GPU does, so the
    FSR rows its resident tracks touch stay in L2 (a chain partition deals the chains of a 3D deck with
    vacuum sides - millions of them - all over the core: 8 GPUs, 7.9 ms per sweep instead of 6.1)."""
    load = track_load(ft) + 1e-9
    cum = np.cumsum(load)
    owner = np.minimum((cum - 0.5 * load) * world / cum[-1], world - 1).astype(np.int64)
    return owner


def partition_by_track(ft: FlatTracks, world: int, owner: np.ndarray = None, only: int = None):
    """Shard individual tracks across `world` ranks; hand-offs that cross ranks are exchanged
    after every sweep (the role of the reference's interface-flux exchange,
    CPUSolver::transferAllInterfaceFluxes, src/CPUSolver.cpp:1063-1211, here with
    torch.distributed send/recv over NCCL).  Unlike the reference's domain decomposition the
    exchanged flux is used by the very next sweep, exactly as in a single-process run, so the
    iteration (and its count) does not depend on the number of ranks.

    Returns [(FlatTracks, ExchangePlan)] per rank (`only`: build that rank's entry, None for
    the others).  A remote hand-off is expressed with plain
    track data: every rank gets ceil(n_send/2) *ghost tracks* with no segments and vacuum ends,
    appended after its own; the link of a track whose successor lives elsewhere points at a
    ghost (track, direction), so the sweep kernel packs the send buffer by itself."""
    a = ft.arrays
    nt = ft.n_tracks
    if owner is None:
        owner = assign_tracks(ft, world)
    owner = np.asarray(owner, dtype=np.int64)
    ids_of = [np.nonzero(owner == r)[0] for r in range(world)]
    new_id = np.empty(nt, dtype=np.int64)
    for r in range(world):
        new_id[ids_of[r]] = np.arange(ids_of[r].size)

    # every cross-rank hand-off: (source rank, destination rank, global target slot, source slot)
    src_slot, dst_slot = [], []
    for d, bit in (("fwd", 1), ("bwd", 2)):
        bc = a["trk_bc_" + d]
        linked = np.nonzero((bc == REFLECTIVE) | (bc == PERIODIC))[0]
        nx = a["trk_next_" + d][linked].astype(np.int64)
        to_fwd = (a["trk_flags"][linked] & bit) != 0
        remote = owner[nx] != owner[linked]
        src_slot.append(linked[remote] * 2 + (0 if d == "fwd" else 1))
        dst_slot.append(nx[remote] * 2 + np.where(to_fwd[remote], 0, 1))
    src_slot = np.concatenate(src_slot) if src_slot else np.zeros(0, np.int64)
    dst_slot = np.concatenate(dst_slot) if dst_slot else np.zeros(0, np.int64)
    src_rank, dst_rank = owner[src_slot // 2], owner[dst_slot // 2]

    out = []
    for r in range(world):
        if only is not None and r != only:
            out.append(None)
            continue
        sub = _extract(ft, ids_of[r], closed=False)
        n_loc = sub.n_tracks
        mine = np.nonzero(src_rank == r)[0]
        # send order: destination rank, then target slot (the receiver sorts the same way)
        mine = mine[np.lexsort((dst_slot[mine], dst_rank[mine]))]
        n_send = mine.size
        n_ghost = (n_send + 1) // 2
        b = sub.arrays
        nf, nb, fl = b["trk_next_fwd"].copy(), b["trk_next_bwd"].copy(), b["trk_flags"].copy()
        k = np.arange(n_send)
        ghost_trk, ghost_is_fwd = n_loc + k // 2, (k % 2 == 0)
        st = new_id[src_slot[mine] // 2]
        sd = src_slot[mine] % 2
        f = sd == 0
        nf[st[f]] = ghost_trk[f]
        fl[st[f]] = (fl[st[f]] & ~np.uint8(1)) | ghost_is_fwd[f].astype(np.uint8)
        nb[st[~f]] = ghost_trk[~f]
        fl[st[~f]] = (fl[st[~f]] & ~np.uint8(2)) | (ghost_is_fwd[~f].astype(np.uint8) << 1)
        assert not (nf == -2).any() and not (nb == -2).any()

        def pad(key, fill, dtype=None):
            v = b[key]
            b[key] = np.concatenate([v, np.full(n_ghost, fill, dtype=dtype or v.dtype)])
        b["trk_next_fwd"], b["trk_next_bwd"], b["trk_flags"] = nf, nb, fl
        for key, fill in (("trk_next_fwd", -1), ("trk_next_bwd", -1), ("trk_flags", 0), ("trk_bc_fwd", 0),
                          ("trk_bc_bwd", 0), ("trk_azim", 0), ("trk_polar", 0), ("trk_xy", 0),
                          ("trk_phi", 0.0), ("trk_theta", 0.0), ("trk_2d", 0), ("trk_lz", 0),
                          ("trk_l0", 1e300)):      # on-the-fly ghost: starts beyond its 2D track, no segments
            if key in b:
                pad(key, fill)
        for key in ("trk_start", "trk_end"):
            if key in b and n_loc and b[key].size % n_loc == 0:
                b[key] = np.concatenate([b[key], np.zeros(n_ghost * (b[key].size // n_loc), dtype=b[key].dtype)])
        b["trk_seg_offset"] = np.concatenate([b["trk_seg_offset"],
                                              np.full(n_ghost, b["trk_seg_offset"][-1], dtype=np.int64)])
        sub.n_tracks = n_loc + n_ghost

        theirs = np.nonzero(dst_rank == r)[0]
        theirs = theirs[np.lexsort((dst_slot[theirs], src_rank[theirs]))]
        recv_slots = new_id[dst_slot[theirs] // 2] * 2 + dst_slot[theirs] % 2
        plan = ExchangePlan(2 * n_loc, np.bincount(dst_rank[mine], minlength=world),
                            np.bincount(src_rank[theirs], minlength=world), recv_slots)
        out.append((sub, plan))
    return out


def exchange_boundary_fluxes(psi, plan: ExchangePlan, dist, group=None):
    """One hand-off round after a sweep.  `psi` is this rank's start-flux array as a torch
    tensor of shape [slots, F] (a zero-copy view of the device buffer on the GPU): the ghost
    slots are sent (and cleared, so that the ghost tracks carry nothing into the next sweep and
    its leakage tally), the received fluxes are stored to the slots the plan names.
    Point-to-point `isend/irecv` batched in one group: NCCL send/recv over NVLink on GPUs,
    gloo on CPUs."""
    import torch
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    # persistent staging buffers: allocating per call would make torch's caching allocator
    # fall back to cudaMalloc (a device-wide sync) whenever the host runs ahead of the GPU,
    # because blocks still in use on NCCL's stream cannot be recycled
    key = (psi.device, psi.dtype, psi.shape[1])
    if getattr(plan, "_key", None) != key:
        plan._key = key
        plan._send = torch.empty((plan.n_send, psi.shape[1]), dtype=psi.dtype, device=psi.device)
        plan._recv = torch.empty((plan.n_recv, psi.shape[1]), dtype=psi.dtype, device=psi.device)
        plan._recv_idx = torch.as_tensor(plan.recv_slots, device=psi.device)
    send, recv = plan._send, plan._recv
    ghosts = psi[plan.ghost0:plan.ghost0 + plan.n_send]
    send.copy_(ghosts)
    ghosts.zero_()
    ops, so, ro = [], 0, 0
    for q in range(world):
        ns, nr = plan.send_counts[q], plan.recv_counts[q]
        peer = q if group is None else dist.get_global_rank(group, q)
        if q != rank and ns:
            ops.append(dist.P2POp(dist.isend, send[so:so + ns], peer, group))
        if q != rank and nr:
            ops.append(dist.P2POp(dist.irecv, recv[ro:ro + nr], peer, group))
        so += ns
        ro += nr
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if plan.n_recv:
        psi.index_copy_(0, plan._recv_idx, recv)
