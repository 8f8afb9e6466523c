"""Spatial domain decomposition of flattened tracks (partition B).

The reference decomposes the geometry into nx x ny x nz boxes (`Geometry::setDomainDecomposition`,
src/Geometry.cpp:854): every MPI rank lays its own tracks over its box, a track that reaches a box face has the
boundary type INTERFACE and hands its outgoing angular flux to the track that continues it in the neighbouring box
(`CPUSolver::setupMPIBuffers / packBuffers / transferAllInterfaceFluxes`, src/CPUSolver.cpp:545-1211); the received
flux starts the neighbour's track in the NEXT sweep, so interface fluxes lag one iteration.

Here the same decomposition is made from the flattened tracks of the whole geometry: every track is cut at the
planes between the boxes (`split_tracks_2d`), a segment that straddles a plane is split there, the pieces of a track
are linked piece to piece like tracks that meet at a periodic boundary, and every piece belongs to the box that
contains it (`assign_domains`).  `partition_by_domain` then hands the pieces to the ranks with
`partition.partition_by_track(owner=...)`: the hand-offs between boxes become the NCCL send / recv of the exchange
plan, between the ranks that share a face (a track through an edge or a corner goes to the diagonal neighbour
directly).  As in the reference the flux crosses one box per sweep: the converged solution is that of the undivided
problem, the number of iterations grows slightly with the number of boxes.  FSR, material and quadrature data stay
replicated (the FSR tally is summed over the ranks like in every other partition of `partition.py`).

Explicit 2D and 3D track sets with `trk_start` / `trk_phi` (/ `trk_theta`): what `openmoc_b200.synth.make_tracks`,
`make_tracks_3d(expand=True)` and the track files carry; 3D boxes are nx x ny x nz.  Axially traced 3D track sets
(no explicit segments on the host) cannot be cut here: the device tracer walks a 3D track to the end of the geometry.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

from .trackfile import FlatTracks, REFLECTIVE, PERIODIC

PER_TRACK = ("trk_azim", "trk_polar", "trk_xy", "trk_phi", "trk_theta")


def track_geometry(ft: FlatTracks):
    """start points [nt, dim], unit directions [nt, dim] and lengths [nt] of explicit tracks (dim = 2 or 3)"""
    a = ft.arrays
    nt = ft.n_tracks
    dim = 3 if ft.solve_3d else 2
    explicit = "trk_seg_offset" in a and a["trk_seg_offset"].size == nt + 1 and ft.n_segments > 0
    if not explicit or "trk_start" not in a or a["trk_start"].size != dim * nt or "trk_phi" not in a \
            or (dim == 3 and "trk_theta" not in a):
        raise ValueError("the domain decomposition needs explicit tracks with trk_start and trk_phi (3D: trk_theta; "
                         "axially traced track sets have no explicit segments to cut)")
    off = a["trk_seg_offset"].astype(np.int64)
    cum = np.concatenate(([0.0], np.cumsum(a["seg_length"].astype(np.longdouble))))
    length = (cum[off[1:]] - cum[off[:-1]]).astype(np.float64)
    phi = a["trk_phi"].astype(np.float64)
    if dim == 2:
        direction = np.stack([np.cos(phi), np.sin(phi)], axis=1)
    else:
        theta = a["trk_theta"].astype(np.float64)
        direction = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], axis=1)
    return a["trk_start"].reshape(nt, dim).astype(np.float64), direction, length


track_geometry_2d = track_geometry


def bounding_box(ft: FlatTracks):
    """(xmin, xmax, ymin, ymax[, zmin, zmax]) of the geometry: every track starts and ends on its boundary"""
    start, direction, length = track_geometry(ft)
    pts = np.concatenate([start, start + direction * length[:, None]])
    return tuple(float(f(pts[:, i])) for i in range(pts.shape[1]) for f in (np.min, np.max))


def domain_planes(ft: FlatTracks, domains: Sequence[int], balance: bool = False):
    """Interior cut planes of nx x ny (x nz) boxes, one array per axis, and the bounding box.  Equal boxes, as
    Geometry::setDomainDecomposition makes them; `balance`: planes at the quantiles of the segment count along
    every axis instead, so that the boxes of a core with a reflector hold about the same number of segments (the
    2D C5G7 quarter core in 4 x 2 equal boxes: 0.59 - 1.36 of the mean)."""
    dim = 3 if ft.solve_3d else 2
    n = [int(x) for x in domains] + [1] * (dim - len(domains))
    if len(n) != dim or min(n) < 1:
        raise ValueError("the number of domains must be positive in each of the %d directions" % dim)
    box = bounding_box(ft)
    if balance:
        mid = segment_midpoints(ft)
        planes = []
        for i in range(dim):
            hist, edges = np.histogram(mid[:, i], bins=4096, range=(box[2 * i], box[2 * i + 1]))
            cdf = np.concatenate(([0.0], np.cumsum(hist))) / max(hist.sum(), 1)
            planes.append(np.interp(np.arange(1, n[i]) / n[i], cdf, edges))
        planes = tuple(planes)
    else:
        planes = tuple(box[2 * i] + (box[2 * i + 1] - box[2 * i]) * np.arange(1, n[i]) / n[i] for i in range(dim))
    return planes + (box,)


def segment_midpoints(ft: FlatTracks) -> np.ndarray:
    """[n_segments, dim] midpoint of every segment"""
    start, direction, _ = track_geometry(ft)
    a = ft.arrays
    off = a["trk_seg_offset"].astype(np.int64)
    seg_len = a["seg_length"].astype(np.float64)
    cum0 = np.concatenate(([np.longdouble(0)], np.cumsum(seg_len.astype(np.longdouble))))
    trk = np.repeat(np.arange(ft.n_tracks, dtype=np.int64), np.diff(off))
    along = (cum0[:-1] - cum0[off[trk]]).astype(np.float64) + 0.5 * seg_len
    return start[trk] + direction[trk] * along[:, None]


def split_tracks(ft: FlatTracks, x_planes, y_planes, z_planes=(), eps: float = None) -> FlatTracks:
    """Cut every track at the planes x = X, y = Y (and, 3D tracks, z = Z).  Returns a track set with one track per
    piece, in the order of the original tracks and along them; `piece_of` (original track of every piece) and
    `piece_d0` / `piece_d1` (distances from the original track's start) are added to its arrays.  A cut that falls on a segment boundary
    (the usual case: box faces are lattice-cell faces) splits no segment; otherwise the segment is split in two
    with the same FSR, like the reference does when it ray-traces every box on its own."""
    a = ft.arrays
    nt, ns = ft.n_tracks, ft.n_segments
    start, direction, tlen = track_geometry(ft)
    dim = start.shape[1]
    off = a["trk_seg_offset"].astype(np.int64)
    seg_len = a["seg_length"].astype(np.float64)
    if eps is None:
        eps = 1e-9 * max(1.0, float(tlen.max()) if nt else 1.0)

    # ---- cuts: (track, distance from its start), sorted along the tracks, one per crossing point
    cut_t, cut_d = [np.zeros(0, np.int64)], [np.zeros(0)]
    for axis, planes in ((0, x_planes), (1, y_planes), (2, z_planes)):
        if axis >= dim:
            if np.size(planes):
                raise ValueError("z planes need 3D tracks")
            continue
        for plane in np.asarray(planes, dtype=np.float64).ravel():
            with np.errstate(divide="ignore", invalid="ignore"):
                d = (plane - start[:, axis]) / direction[:, axis]
            ok = np.isfinite(d) & (d > eps) & (d < tlen - eps)
            cut_t.append(np.nonzero(ok)[0])
            cut_d.append(d[ok])
    cut_t, cut_d = np.concatenate(cut_t), np.concatenate(cut_d)
    order = np.lexsort((cut_d, cut_t))
    cut_t, cut_d = cut_t[order], cut_d[order]

    # ---- where every cut falls in the segment stream (extended precision: offsets of ~1e8 segments add up)
    cum = np.cumsum(seg_len.astype(np.longdouble))                 # end of every segment, whole stream
    cum0 = np.concatenate(([np.longdouble(0)], cum))               # begin of every segment
    pos = cum0[off[cut_t]] + cut_d.astype(np.longdouble)
    k = np.searchsorted(cum, pos, side="left").astype(np.int64)    # segment that contains the cut
    k = np.clip(k, off[cut_t], off[cut_t + 1] - 1)
    before = (pos - cum0[k]).astype(np.float64)                     # from the segment's begin to the cut
    after = (cum[k] - pos).astype(np.float64)
    at_begin, at_end = before < eps, after < eps
    # the boundary every cut makes, in units of "original segments + fraction": cuts on the same boundary merge
    bound_seg = np.where(at_end, k + 1, k)
    bound_frac = np.where(at_begin | at_end, 0.0, before)
    keep = np.ones(cut_t.size, dtype=bool)
    keep[1:] = (bound_seg[1:] != bound_seg[:-1]) | (np.abs(bound_frac[1:] - bound_frac[:-1]) >= eps)
    # a cut on the first / last boundary of its track cuts nothing off
    keep &= (bound_seg > off[cut_t]) | (bound_frac > 0)
    keep &= bound_seg < off[cut_t + 1]
    cut_t, cut_d, bound_seg, bound_frac = cut_t[keep], cut_d[keep], bound_seg[keep], bound_frac[keep]
    splits = bound_frac > 0
    n_cut = cut_t.size

    # ---- new segment stream: a segment with m interior cuts becomes m + 1 pieces
    split_seg, split_at = bound_seg[splits], bound_frac[splits]
    n_in_seg = np.bincount(split_seg, minlength=ns).astype(np.int64)
    first_split = np.concatenate(([0], np.cumsum(n_in_seg)))[:-1]            # first entry of split_* of a segment
    pieces = n_in_seg + 1
    first_new = np.concatenate(([0], np.cumsum(pieces)))                      # new index of a segment's first piece
    new_ns = int(first_new[-1])
    orig = np.repeat(np.arange(ns, dtype=np.int64), pieces)                   # original segment of every new one
    r = np.arange(new_ns, dtype=np.int64) - first_new[orig]                   # piece number inside its segment
    at = np.concatenate((split_at, [0.0]))                                    # padded: every index below is valid
    last = r == pieces[orig] - 1
    lo = np.where(r > 0, at[np.minimum(first_split[orig] + r - 1, split_at.size)], 0.0)
    hi = np.where(last, seg_len[orig], at[np.minimum(first_split[orig] + r, split_at.size)])
    out = FlatTracks(num_groups=ft.num_groups, num_azim=ft.num_azim, num_polar=ft.num_polar, solve_3d=ft.solve_3d,
                     fluxes_per_track=ft.fluxes_per_track, n_tracks=nt + n_cut, n_segments=new_ns, n_fsrs=ft.n_fsrs,
                     n_materials=ft.n_materials)
    b = out.arrays
    b["seg_length"] = hi - lo
    for key in ("seg_fsr", "seg_mat"):
        if key in a and a[key].size == ns:
            b[key] = a[key][orig]
    # the CMFD surface a segment ends on belongs to its last piece (forward) / first piece (backward)
    if "seg_cmfd_fwd" in a and a["seg_cmfd_fwd"].size == ns:
        b["seg_cmfd_fwd"] = np.where(last, a["seg_cmfd_fwd"][orig], -1).astype(a["seg_cmfd_fwd"].dtype)
        b["seg_cmfd_bwd"] = np.where(r == 0, a["seg_cmfd_bwd"][orig], -1).astype(a["seg_cmfd_bwd"].dtype)
    # linear source: starting points relative to the FSR centroid move along the track (CPULSSolver.cpp:736-738)
    if "seg_start" in a and a["seg_start"].size == 3 * ns:
        s0 = a["seg_start"].reshape(ns, 3)[orig].astype(np.float64)
        trk_of_seg = np.repeat(np.arange(nt, dtype=np.int64), np.diff(off))[orig]
        s0[:, :dim] += direction[trk_of_seg] * lo[:, None]
        b["seg_start"] = s0.ravel()

    # ---- new tracks: piece j of track t has the id t + (cuts in earlier tracks) + j
    cuts_in = np.bincount(cut_t, minlength=nt).astype(np.int64)
    n_pieces = cuts_in + 1
    base = np.arange(nt, dtype=np.int64) + np.concatenate(([0], np.cumsum(cuts_in)))[:-1]
    piece_of = np.repeat(np.arange(nt, dtype=np.int64), n_pieces)
    j = np.arange(nt + n_cut, dtype=np.int64) - base[piece_of]
    first_cut = np.concatenate(([0], np.cumsum(cuts_in)))[:-1]                # first entry of cut_* of a track
    # new segment index at which the piece after every cut begins: a cut on a segment boundary -> the first piece
    # of the next segment; the (q + 1)-th split of a segment -> its piece q + 1
    q = np.zeros(n_cut, dtype=np.int64)
    q[splits] = np.arange(split_seg.size, dtype=np.int64) - first_split[split_seg]
    cut_new_seg = np.concatenate((first_new[bound_seg] + np.where(splits, q + 1, 0), [0]))   # padded
    cut_dist = np.concatenate((cut_d, [0.0]))
    prev_cut = np.minimum(first_cut[piece_of] + j - 1, n_cut)                 # the cut a piece begins at (j > 0)
    next_cut = np.minimum(first_cut[piece_of] + j, n_cut)                     # the cut it ends at (not the last piece)
    piece_begin = np.where(j == 0, first_new[off[piece_of]], cut_new_seg[prev_cut])
    b["trk_seg_offset"] = np.concatenate((piece_begin, [new_ns])).astype(np.int64)
    is_last = j == n_pieces[piece_of] - 1
    d0 = np.where(j == 0, 0.0, cut_dist[prev_cut])
    d1 = np.where(is_last, tlen[piece_of], cut_dist[next_cut])
    b["piece_of"], b["piece_d0"], b["piece_d1"] = piece_of, d0, d1
    for key in PER_TRACK:
        if key in a and a[key].size == nt:
            b[key] = a[key][piece_of]
    b["trk_start"] = (start[piece_of] + direction[piece_of] * d0[:, None]).ravel()
    if dim == 3:
        b["trk_end"] = (start[piece_of] + direction[piece_of] * d1[:, None]).ravel()
        for key in ("trk_2d", "trk_lz"):
            if key in a and a[key].size == nt:
                b[key] = a[key][piece_of]
        if "trk_l0" in a and a["trk_l0"].size == nt:        # distance of the start point along the 2D track
            b["trk_l0"] = a["trk_l0"][piece_of] + d0 * np.hypot(direction[piece_of, 0], direction[piece_of, 1])

    # ---- links: inside a track piece to piece, at its two ends the original hand-offs, to the piece that holds the
    # entered end of the target track
    flags = a["trk_flags"][piece_of].astype(np.uint8)
    ids = np.arange(nt + n_cut, dtype=np.int64)
    for d, bit, inner, at_end_of_track in (("fwd", np.uint8(1), ids + 1, is_last), ("bwd", np.uint8(2), ids - 1, j == 0)):
        bc0 = a["trk_bc_" + d][piece_of]
        nxt0 = a["trk_next_" + d][piece_of].astype(np.int64)
        linked = (bc0 == REFLECTIVE) | (bc0 == PERIODIC)
        tgt = np.clip(nxt0, 0, nt - 1)
        enters_fwd = (flags & bit) != 0
        outer = np.where(linked, base[tgt] + np.where(enters_fwd, 0, n_pieces[tgt] - 1), nxt0)
        b["trk_next_" + d] = np.where(at_end_of_track, outer, inner)
        b["trk_bc_" + d] = np.where(at_end_of_track, bc0, PERIODIC).astype(a["trk_bc_" + d].dtype)
        # forward through a cut enters the next piece forward; backward through a cut enters the previous one backward
        inner_bit = bit if d == "fwd" else np.uint8(0)
        flags = np.where(at_end_of_track, flags, (flags & ~bit) | inner_bit).astype(np.uint8)
    b["trk_flags"] = flags
    for key, v in a.items():
        if key.startswith(("quad_", "fsr_", "mat_")) or key == "z_mesh":
            b[key] = v
    return out


def assign_domains(split: FlatTracks, planes) -> np.ndarray:
    """Box of every piece (by its midpoint; `planes`: the cut planes per axis), numbered x fastest, then y, then z:
    the owner rank of `partition_by_track`"""
    start, direction, _ = track_geometry(split)
    a = split.arrays
    mid = start + direction * (0.5 * (a["piece_d1"] - a["piece_d0"]))[:, None]
    owner, stride = np.zeros(split.n_tracks, dtype=np.int64), 1
    for i, cuts in enumerate(planes):
        cuts = np.asarray(cuts, dtype=np.float64)
        owner += stride * np.searchsorted(cuts, mid[:, i], side="right")
        stride *= cuts.size + 1
    return owner


def default_domains(world: int, dim: int = 2) -> Tuple[int, ...]:
    """nx x ny (x nz) = world, as square / cubic as it gets (2D: 8 -> 4 x 2; 3D: 8 -> 2 x 2 x 2)"""
    if dim == 3:
        nz = max(d for d in range(1, int(np.floor(world ** (1.0 / 3.0) + 1e-9)) + 1) if world % d == 0)
        return default_domains(world // nz, 2) + (nz,)
    ny = int(np.floor(np.sqrt(world)))
    while world % ny:
        ny -= 1
    return world // ny, ny


def partition_by_domain(ft: FlatTracks, world: int, domains: Sequence[int] = None, only: int = None,
                        balance: bool = False):
    """[(FlatTracks, ExchangePlan)] per rank: rank r sweeps the pieces of the tracks inside box r.  `domains`
    (nx, ny) / (nx, ny, nz) defaults to the squarest / most cubic factorisation of `world`; `balance`: see
    `domain_planes`."""
    from .partition import partition_by_track
    dim = 3 if ft.solve_3d else 2
    domains = default_domains(world, dim) if domains is None else tuple(int(x) for x in domains)
    if len(domains) > dim or int(np.prod(domains)) != world:
        raise ValueError("%s domains for %d ranks of a %dD problem" % (" x ".join(map(str, domains)), world, dim))
    *planes, _ = domain_planes(ft, domains, balance)
    split = split_tracks(ft, *planes)
    return partition_by_track(split, world, owner=assign_domains(split, planes), only=only)


split_tracks_2d = split_tracks
