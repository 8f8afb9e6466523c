"""Spatial domain decomposition of flattened tracks (openmoc_b200/domain.py; Geometry::setDomainDecomposition,
src/Geometry.cpp:854, with the interface-flux exchange of src/CPUSolver.cpp:1063-1211) - host logic on the CPU,
the oracle as the sweeping engine."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT
from openmoc_b200.domain import (assign_domains, bounding_box, default_domains, domain_planes, partition_by_domain,
                                 split_tracks_2d, track_geometry_2d)
from openmoc_b200.trackfile import PERIODIC, REFLECTIVE, VACUUM


def lattice(**kw):
    from openmoc_b200.synth import make_tracks
    return make_tracks("simple-lattice", num_azim=8, spacing=0.1, **kw)


def c5g7():
    from openmoc_b200.synth import make_tracks
    return make_tracks("c5g7-2d", num_azim=4, spacing=0.5)


def fsr_track_length(ft):
    return np.bincount(ft.arrays["seg_fsr"], weights=ft.arrays["seg_length"], minlength=ft.n_fsrs)


@pytest.mark.parametrize("domains", [(2, 1), (2, 2), (3, 2), (1, 4)])
@pytest.mark.parametrize("deck", [lattice, c5g7])
def test_split_keeps_every_track_and_cuts_it_at_the_planes(deck, domains):
    ft = deck()
    xs, ys, box = domain_planes(ft, domains)
    sp = split_tracks_2d(ft, xs, ys)
    sp.validate()
    a, b = ft.arrays, sp.arrays
    # the same track length in every FSR: volumes and the converged solution do not change
    np.testing.assert_allclose(fsr_track_length(sp), fsr_track_length(ft), rtol=1e-12)
    assert b["seg_length"].min() > 1e-10
    assert sp.n_segments >= ft.n_segments and sp.n_tracks > ft.n_tracks
    # walking the pieces of a track gives back its segments (split ones merged)
    piece_of, off, noff = b["piece_of"], a["trk_seg_offset"], b["trk_seg_offset"]
    assert np.array_equal(np.unique(piece_of), np.arange(ft.n_tracks))
    first = np.nonzero(np.concatenate(([True], piece_of[1:] != piece_of[:-1])))[0]
    for t in (0, ft.n_tracks // 3, ft.n_tracks - 1):
        p, fsrs, lens = first[t], [], []
        while True:
            sl = slice(noff[p], noff[p + 1])
            fsrs += list(b["seg_fsr"][sl]); lens += list(b["seg_length"][sl])
            if p + 1 >= sp.n_tracks or piece_of[p + 1] != t:
                break
            assert b["trk_next_fwd"][p] == p + 1 and b["trk_bc_fwd"][p] == PERIODIC and b["trk_flags"][p] & 1
            assert b["trk_next_bwd"][p + 1] == p and b["trk_bc_bwd"][p + 1] == PERIODIC and not b["trk_flags"][p + 1] & 2
            p += 1
        merged_f, merged_l = [fsrs[0]], [lens[0]]
        for f, l in zip(fsrs[1:], lens[1:]):
            if f == merged_f[-1]:
                merged_l[-1] += l
            else:
                merged_f.append(f); merged_l.append(l)
        of, ol = [a["seg_fsr"][off[t]]], [a["seg_length"][off[t]]]
        for f, l in zip(a["seg_fsr"][off[t] + 1:off[t + 1]], a["seg_length"][off[t] + 1:off[t + 1]]):
            if f == of[-1]:
                ol[-1] += l
            else:
                of.append(f); ol.append(l)
        assert merged_f == of
        np.testing.assert_allclose(merged_l, ol, rtol=1e-10)
        # the ends of the track keep their boundary conditions
        assert b["trk_bc_bwd"][first[t]] == a["trk_bc_bwd"][t] and b["trk_bc_fwd"][p] == a["trk_bc_fwd"][t]
    # every piece lies inside one box
    owner = assign_domains(sp, (xs, ys))
    assert set(np.unique(owner)) == set(range(domains[0] * domains[1]))
    start, direction, length = track_geometry_2d(sp)
    end = start + direction * length[:, None]
    nx, ny = domains
    wx, wy = (box[1] - box[0]) / nx, (box[3] - box[2]) / ny
    ix, iy = owner % nx, owner // nx
    for pts in (start, end):
        assert np.all(pts[:, 0] >= box[0] + ix * wx - 1e-8) and np.all(pts[:, 0] <= box[0] + (ix + 1) * wx + 1e-8)
        assert np.all(pts[:, 1] >= box[2] + iy * wy - 1e-8) and np.all(pts[:, 1] <= box[2] + (iy + 1) * wy + 1e-8)
    np.testing.assert_allclose(length, b["piece_d1"] - b["piece_d0"], atol=1e-9)


def test_links_between_tracks_enter_the_right_piece():
    """A hand-off at a reflective boundary enters the piece that holds the entered end of the target track, and the
    link graph stays one-to-one (every (track, direction) slot is fed by at most one hand-off)."""
    ft = lattice()
    xs, ys, _ = domain_planes(ft, (2, 2))
    sp = split_tracks_2d(ft, xs, ys)
    b = sp.arrays
    start, direction, length = track_geometry_2d(sp)
    end = start + direction * length[:, None]
    fed = np.zeros(2 * sp.n_tracks, dtype=np.int64)
    for d, bit, leave in (("fwd", 1, end), ("bwd", 2, start)):
        linked = np.nonzero((b["trk_bc_" + d] == REFLECTIVE) | (b["trk_bc_" + d] == PERIODIC))[0]
        nxt = b["trk_next_" + d][linked]
        to_fwd = (b["trk_flags"][linked] & bit) != 0
        np.add.at(fed, nxt * 2 + np.where(to_fwd, 0, 1), 1)
        # the flux leaves one piece where it enters the next (reflective: same point; this deck has no periodic side)
        enter = np.where(to_fwd[:, None], start[nxt], end[nxt])
        np.testing.assert_allclose(enter, leave[linked], atol=1e-8)
    assert fed.max() == 1


def test_split_without_planes_is_the_identity():
    ft = lattice()
    sp = split_tracks_2d(ft, [], [])
    for k in ("seg_length", "seg_fsr", "trk_seg_offset", "trk_next_fwd", "trk_next_bwd", "trk_flags", "trk_bc_fwd",
              "trk_bc_bwd", "trk_start"):
        assert np.array_equal(sp.arrays[k], ft.arrays[k]), k
    assert default_domains(8) == (4, 2) and default_domains(4) == (2, 2) and default_domains(2) == (2, 1)
    assert default_domains(7) == (7, 1)
    with pytest.raises(ValueError):
        partition_by_domain(ft, 4, domains=(3, 1))


def test_a_plane_inside_segments_splits_them():
    """planes that are no lattice-cell faces: segments are split, CMFD surfaces stay on the right pieces, the
    linear source's starting points move along the track"""
    ft = lattice(linear_source=True)
    box = bounding_box(ft)
    sp = split_tracks_2d(ft, [box[0] + 0.37 * (box[1] - box[0])], [box[2] + 0.61 * (box[3] - box[2])])
    sp.validate()
    assert sp.n_segments > ft.n_segments
    np.testing.assert_allclose(fsr_track_length(sp), fsr_track_length(ft), rtol=1e-12)
    a, b = ft.arrays, sp.arrays
    for k in ("seg_cmfd_fwd", "seg_cmfd_bwd"):
        assert np.array_equal(np.sort(b[k][b[k] >= 0]), np.sort(a[k][a[k] >= 0])), k
    # the two pieces of a split segment end / begin two pieces of one track: the second starts where the first ends
    s = b["seg_start"].reshape(-1, 3)
    trk = np.repeat(np.arange(sp.n_tracks), np.diff(b["trk_seg_offset"]))
    i = b["trk_seg_offset"][1:-1] - 1                                       # last segment of every piece but the last
    i = i[(b["piece_of"][trk[i]] == b["piece_of"][trk[i + 1]]) & (b["seg_fsr"][i] == b["seg_fsr"][i + 1])]
    assert i.size == sp.n_segments - ft.n_segments
    phi = b["trk_phi"][trk[i]]
    nxt = s[i, :2] + b["seg_length"][i, None] * np.stack([np.cos(phi), np.sin(phi)], axis=1)
    np.testing.assert_allclose(nxt, s[i + 1, :2], atol=1e-12)
    assert np.all(b["seg_cmfd_fwd"][i] < 0) and np.all(b["seg_cmfd_bwd"][i + 1] < 0)


# ------------------------------------------------------------------ physics: the oracle on the decomposed tracks
def simulate(ft, world, domains, max_iters, tol, balance=False):
    """All ranks of partition_by_domain in one process: sweep per box, sum of the FSR tallies, interface fluxes moved
    by hand with the plan's index lists (what exchange_boundary_fluxes does over NCCL)."""
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    parts = partition_by_domain(ft, world, domains, balance=balance)
    F = ft.fluxes_per_track
    solvers = [OracleSolver(sub) for sub, _ in parts]
    plans = [p for _, p in parts]
    for s in solvers:
        s.setKeff(1.0); s.zeroTrackFluxes()
        s.flattenFSRFluxes(0.0); s.storeFSRFluxes()
        s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
    k_prev, iters = 1.0, 0
    for i in range(max_iters):
        for s in solvers:
            s.computeFSRSources(i); s.transportSweep()
        phi = sum(s.getFluxes() for s in solvers)
        psi = [s.getStartFluxes().reshape(-1, F) for s in solvers]
        outbox = {}
        for r, p in enumerate(plans):
            o = p.ghost0
            for q in range(world):
                outbox[(r, q)] = psi[r][o:o + p.send_counts[q]].copy()
                o += p.send_counts[q]
            psi[r][p.ghost0:p.ghost0 + p.n_send] = 0.0
        for q, p in enumerate(plans):
            if p.n_recv:
                psi[q][p.recv_slots] = np.concatenate([outbox[(r, q)] for r in range(world)])
        res = None
        for s, ps in zip(solvers, psi):
            s.setStartFluxes(ps.ravel()); s.setFluxes(phi)
            s.addSourceToScalarFlux(); s.computeKeff(); s.normalizeFluxes()
            res = s.computeResidual(FISSION_SOURCE)
            s.storeFSRFluxes()
        k = solvers[0].getKeff()
        dk = int(1e5 * (k - k_prev)); k_prev = k
        iters += 1
        if res < tol and abs(dk) < 1:
            break
    return solvers[0].getKeff(), solvers[0].getFluxes(), iters, parts


@pytest.mark.parametrize("deck,world,domains", [(lattice, 4, (2, 2)), (lattice, 2, None)])
def test_decomposed_solve_converges_to_the_undivided_solution(deck, world, domains):
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    ft = deck()
    ref = OracleSolver(ft)
    n_ref = ref.computeEigenvalue(2000, 1e-9, FISSION_SOURCE)
    k, phi, iters, parts = simulate(ft, world, domains, 2000, 1e-9)
    # interface fluxes lag one sweep per box crossed (as in the reference): some more iterations, the same answer
    # (both runs stop at a residual of 1e-9: what is left of the difference is their distance from convergence)
    assert n_ref <= iters <= 1.3 * n_ref
    assert abs(k - ref.getKeff()) * 1e5 < 0.05
    assert np.max(np.abs(phi - ref.getFluxes()) / ref.getFluxes()) < 2e-6
    # only neighbouring boxes talk to each other
    nx = (domains or default_domains(world))[0]
    for r, (_, plan) in enumerate(parts):
        for q in range(world):
            if plan.send_counts[q]:
                assert q != r and abs(q % nx - r % nx) <= 1 and abs(q // nx - r // nx) <= 1
    assert sum(p.n_send for _, p in parts) == sum(p.n_recv for _, p in parts) > 0


def test_boxes_on_ranks_equal_the_cut_tracks_in_one_process():
    """C5G7 core (two vacuum sides), 3 x 2 boxes: distributing the pieces over ranks changes nothing - after 8
    iterations k_eff and the flux equal those of the cut track set swept by one solver (same lag, same hand-offs)."""
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    ft = c5g7()
    xs, ys, _ = domain_planes(ft, (3, 2))
    one = OracleSolver(split_tracks_2d(ft, xs, ys))
    one.computeEigenvalue(8, 1e-30, FISSION_SOURCE)
    k, phi, iters, parts = simulate(ft, 6, (3, 2), 8, 1e-30)
    assert iters == 8 and abs(k - one.getKeff()) < 1e-12
    np.testing.assert_allclose(phi, one.getFluxes(), rtol=1e-10, atol=1e-14)
    assert sum(sub.n_segments for sub, _ in parts) == split_tracks_2d(ft, xs, ys).n_segments


def test_balanced_boxes_hold_about_the_same_number_of_segments():
    """C5G7 quarter core (dense fuel block, sparse reflector): equal boxes are unbalanced, planes at the quantiles of
    the segment count are not; the decomposed iteration is the same whatever the planes"""
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    ft = c5g7()

    def spread(balance):
        parts = partition_by_domain(ft, 8, (4, 2), balance=balance)
        n = np.array([sub.n_segments for sub, _ in parts], dtype=float)
        return n.max() / n.mean(), parts
    equal, _ = spread(False)
    balanced, parts = spread(True)
    assert equal > 1.25 and balanced < 1.12
    *planes, _ = domain_planes(ft, (4, 2), balance=True)
    one = OracleSolver(split_tracks_2d(ft, *planes))
    one.computeEigenvalue(5, 1e-30, FISSION_SOURCE)
    k, phi, iters, _ = simulate(ft, 8, (4, 2), 5, 1e-30, balance=True)
    assert abs(k - one.getKeff()) < 1e-12
    np.testing.assert_allclose(phi, one.getFluxes(), rtol=1e-10, atol=1e-14)


def test_linear_source_on_cut_tracks():
    """CPULSSolver physics on the cut tracks (one process): box faces on lattice-cell faces split no segment - same
    pre-pass tables, same converged solution; a face inside FSRs splits segments, which refines the linear-source
    discretisation exactly like the reference's own per-box ray tracing (and its optical-length cuts) does: the
    geometric table is unchanged, the per-segment source constants and the solution move within the north-star
    tolerance (1 pcm, 1e-4)."""
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    ft = lattice(linear_source=True)
    ref = OracleSolver(ft, linear_source=True)
    ref.computeEigenvalue(3000, 1e-9, FISSION_SOURCE)
    xs, ys, box = domain_planes(ft, (2, 2))
    inside = ([box[0] + 0.37 * (box[1] - box[0])], [box[2] + 0.61 * (box[3] - box[2])])
    for planes, aligned in (((xs, ys), True), (inside, False)):
        sp = split_tracks_2d(ft, *planes)
        assert (sp.n_segments == ft.n_segments) == aligned
        s = OracleSolver(sp, linear_source=True)
        s.computeEigenvalue(3000, 1e-9, FISSION_SOURCE)
        (lin_a, con_a), (lin_b, con_b) = ref.getLinearSourceTables(), s.getLinearSourceTables()
        np.testing.assert_allclose(lin_b, lin_a, rtol=0, atol=1e-9)
        if aligned:
            assert np.array_equal(con_a, con_b)
        assert abs(s.getKeff() - ref.getKeff()) * 1e5 < (0.05 if aligned else 1.0)
        assert np.max(np.abs(s.getFluxes() - ref.getFluxes()) / ref.getFluxes()) < (2e-6 if aligned else 1e-4)


# ------------------------------------------------------------------ 3D: nx x ny x nz boxes of explicit 3D tracks
def lattice3d():
    from openmoc_b200.synth import make_tracks_3d
    return make_tracks_3d("simple-lattice", num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9, n_axial=2, expand=True)


def test_3d_tracks_are_cut_into_2x2x2_boxes():
    from openmoc_b200.domain import split_tracks
    ft = lattice3d()
    xs, ys, zs, box = domain_planes(ft, (2, 2, 2))
    assert len(box) == 6 and xs.size == ys.size == zs.size == 1
    sp = split_tracks(ft, xs, ys, zs)
    sp.validate()
    np.testing.assert_allclose(fsr_track_length(sp), fsr_track_length(ft), rtol=1e-12)
    owner = assign_domains(sp, (xs, ys, zs))
    assert set(np.unique(owner)) == set(range(8))
    start, direction, length = track_geometry_2d(sp)
    end = start + direction * length[:, None]
    np.testing.assert_allclose(end.ravel(), sp.arrays["trk_end"], atol=1e-9)
    idx = np.stack([owner % 2, owner // 2 % 2, owner // 4], axis=1)
    for i in range(3):
        lo, w = box[2 * i], (box[2 * i + 1] - box[2 * i]) / 2
        for pts in (start, end):
            assert np.all(pts[:, i] >= lo + idx[:, i] * w - 1e-8) and np.all(pts[:, i] <= lo + (idx[:, i] + 1) * w + 1e-8)
    assert default_domains(8, 3) == (2, 2, 2) and default_domains(4, 3) == (2, 2, 1) and default_domains(2, 3) == (2, 1, 1)
    with pytest.raises(ValueError):
        split_tracks(lattice(), [], [], [0.0])               # z planes need 3D tracks
    from openmoc_b200.synth import make_tracks_3d
    with pytest.raises(ValueError):                          # axially traced set: nothing to cut on the host
        partition_by_domain(make_tracks_3d("simple-lattice", num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9,
                                           n_axial=2, expand=False), 2)


def test_3d_decomposed_solve_converges_to_the_undivided_solution():
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    ft = lattice3d()
    ref = OracleSolver(ft)
    n_ref = ref.computeEigenvalue(2000, 1e-9, FISSION_SOURCE)
    k, phi, iters, parts = simulate(ft, 8, None, 2000, 1e-9)
    assert n_ref <= iters <= 1.3 * n_ref
    assert abs(k - ref.getKeff()) * 1e5 < 0.05
    assert np.max(np.abs(phi - ref.getFluxes()) / ref.getFluxes()) < 2e-6


# ------------------------------------------------------------------ tracks of the reference's own ray tracer
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def reference_tracks(tmp_path, name, args):
    import subprocess
    from openmoc_b200.trackfile import read_trackfile
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_driver not built")
    trk = os.path.join(tmp_path, name + ".b2trk")
    subprocess.run([DRIVER] + args + ["--quiet", "--max-iters", "1", "--dump-tracks", trk,
                                      "--json", os.path.join(tmp_path, name + ".json")],
                   check=True, capture_output=True, cwd=tmp_path)
    return read_trackfile(trk)


def test_reference_track_dumps_are_cut_into_boxes(tmp_path):
    """Track files dumped from the reference's TrackGenerator / TrackGenerator3D carry the start point of every track
    (trk_start, b200_flatten.cpp): the decomposition applies to real OpenMOC geometries, 2D and 3D."""
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    ft = reference_tracks(tmp_path, "sl", ["--model", "simple-lattice", "--azim", "4", "--spacing", "0.12"])
    assert ft.arrays["trk_start"].size == 2 * ft.n_tracks
    ref = OracleSolver(ft)
    n_ref = ref.computeEigenvalue(3000, 1e-9, FISSION_SOURCE)
    k, phi, iters, _ = simulate(ft, 4, None, 3000, 1e-9)
    assert n_ref <= iters <= 1.3 * n_ref and abs(k - ref.getKeff()) * 1e5 < 0.05
    assert np.max(np.abs(phi - ref.getFluxes()) / ref.getFluxes()) < 2e-6

    ft = reference_tracks(tmp_path, "l3", ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2",
                                           "--spacing", "0.24", "--zspacing", "0.9"])
    assert ft.solve_3d and ft.arrays["trk_start"].size == 3 * ft.n_tracks
    from openmoc_b200.domain import split_tracks
    *planes, box = domain_planes(ft, (2, 2, 2))
    np.testing.assert_allclose(box, (-2, 2, -2, 2, -5, 5), atol=1e-9)
    sp = split_tracks(ft, *planes)
    sp.validate()
    np.testing.assert_allclose(fsr_track_length(sp), fsr_track_length(ft), rtol=1e-12)
    one = OracleSolver(sp)
    one.computeEigenvalue(6, 1e-30, FISSION_SOURCE)
    k, phi, iters, parts = simulate(ft, 8, (2, 2, 2), 6, 1e-30)
    assert abs(k - one.getKeff()) < 1e-12
    np.testing.assert_allclose(phi, one.getFluxes(), rtol=1e-10, atol=1e-14)


def test_device_tracer_hand_over_carries_the_same_track_data(tmp_path):
    """b200_flatten(device_otf=true) - what the plug-in gives the device tracer for an OTF deck, no 3D segment on the
    host - and the host expansion agree on every per-track array, start points included"""
    args = ["--model", "simple-lattice", "--dims", "3", "--azim", "4", "--polar", "2", "--spacing", "0.24",
            "--zspacing", "0.9", "--formation", "otf-stacks", "--mode", "none"]
    host = reference_tracks(tmp_path, "host", args)
    dev = reference_tracks(tmp_path, "dev", args + ["--dump-device-otf"])
    assert dev.n_segments == 0 and host.n_segments > 0 and dev.n_tracks == host.n_tracks
    for k in ("trk_start", "trk_phi", "trk_theta", "trk_azim", "trk_polar", "trk_next_fwd", "trk_next_bwd", "trk_flags",
              "trk_bc_fwd", "trk_bc_bwd"):
        assert np.array_equal(dev.arrays[k], host.arrays[k]), k


# ------------------------------------------------------------------ the same over gloo, one process per box
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from openmoc_b200.partition import exchange_boundary_fluxes
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    part, plan = partition_by_domain(lattice(), world, only=rank)[rank]
    F = part.fluxes_per_track
    s = OracleSolver(part)
    s.setKeff(1.0); s.zeroTrackFluxes()
    s.flattenFSRFluxes(0.0); s.storeFSRFluxes()
    s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
    k_prev, iters = 1.0, 0
    for i in range(600):
        s.computeFSRSources(i)
        s.transportSweep()
        phi = torch.from_numpy(s.getFluxes())
        dist.all_reduce(phi, op=dist.ReduceOp.SUM)
        s.setFluxes(phi.numpy())
        psi = torch.from_numpy(s.getStartFluxes()).view(-1, F)
        exchange_boundary_fluxes(psi, plan, dist)
        s.setStartFluxes(psi.numpy().ravel())
        s.addSourceToScalarFlux()
        s.computeKeff(); k = s.getKeff()
        s.normalizeFluxes()
        res = s.computeResidual(FISSION_SOURCE)
        dk = int(1e5 * (k - k_prev)); k_prev = k
        s.storeFSRFluxes(); iters += 1
        if res < 1e-6 and abs(dk) < 1:
            break
    np.save(os.path.join(out_dir, f"phi{rank}.npy"), s.getFluxes())
    np.save(os.path.join(out_dir, f"k{rank}.npy"), np.array([s.getKeff(), iters]))
    dist.destroy_process_group()


def test_gloo_one_process_per_box(tmp_path):
    import torch.multiprocessing as mp
    world = 4
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    k_sim, phi_sim, iters_sim, _ = simulate(lattice(), world, None, 600, 1e-6)
    for r in range(world):
        k, iters = np.load(os.path.join(tmp_path, f"k{r}.npy"))
        assert int(iters) == iters_sim and abs(k - k_sim) < 1e-12
        np.testing.assert_allclose(np.load(os.path.join(tmp_path, f"phi{r}.npy")), phi_sim, rtol=1e-10)
