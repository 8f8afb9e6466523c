"""Random ragged track sets for edge-case tests: not a geometry, just the data the sweep
consumes (segments, links, quadrature, cross sections) with every irregularity the flat
layout allows.

* tracks with zero segments, with one segment, and one very long track
* runs of consecutive segments in the same FSR (tally flushes) and isolated ones
* optical lengths from 1e-9 to ~60 (the reference splits segments at tau = 10..100)
* a random one-to-one hand-off between (track, direction) slots mixing REFLECTIVE,
  PERIODIC and VACUUM ends, forward and backward targets, self links
* any number of groups (1, primes, 70) and 1..3 polar angles per 2D track, or 3D tracks
"""
import numpy as np

from openmoc_b200.trackfile import FlatTracks, VACUUM, REFLECTIVE, PERIODIC


def make_ragged(G=7, NP=3, solve_3d=False, n_tracks=97, n_fsrs=23, n_mats=3, seed=0,
                vacuum_fraction=0.3, max_segments=14, long_track=300, fissionable=True):
    rng = np.random.default_rng(seed)
    num_azim = 8
    A2 = num_azim // 2
    P = 2 if solve_3d else 2 * NP
    ft = FlatTracks(num_groups=G, num_azim=num_azim, num_polar=P, solve_3d=int(solve_3d),
                    fluxes_per_track=G * (1 if solve_3d else NP), n_tracks=n_tracks,
                    n_fsrs=n_fsrs, n_materials=n_mats)
    a = ft.arrays

    nseg = rng.integers(0, max_segments + 1, n_tracks)
    nseg[rng.choice(n_tracks, max(1, n_tracks // 8), replace=False)] = 0      # empty tracks
    nseg[rng.integers(0, n_tracks)] = 1
    if long_track:
        nseg[rng.integers(0, n_tracks)] = long_track
    off = np.concatenate(([0], np.cumsum(nseg))).astype(np.int64)
    ns = int(off[-1])
    ft.n_segments = ns
    # FSR ids in runs of 1..4 equal values
    fsr = np.empty(ns, dtype=np.int32)
    i = 0
    while i < ns:
        run = int(rng.integers(1, 5))
        fsr[i:i + run] = rng.integers(0, n_fsrs - 2)      # the last two FSRs are crossed by no track
        i += run
    length = rng.uniform(0.01, 2.0, ns)
    length[rng.random(ns) < 0.05] = 1e-9
    length[rng.random(ns) < 0.05] = rng.uniform(20.0, 40.0)
    fsr_mat = rng.integers(0, n_mats, n_fsrs).astype(np.int32)
    fsr_mat[-2:] = 0                                     # (non-fissionable: no isolated multiplying region)
    fsr_mat[0] = n_mats - 1                              # at least one fissionable FSR
    a["seg_length"] = length
    a["seg_fsr"] = fsr
    a["seg_mat"] = fsr_mat[fsr] if ns else np.zeros(0, np.int32)
    a["trk_seg_offset"] = off
    a["trk_azim"] = rng.integers(0, A2, n_tracks).astype(np.int32)
    a["trk_polar"] = (rng.integers(0, P, n_tracks) if solve_3d else np.zeros(n_tracks)).astype(np.int32)
    a["trk_xy"] = np.arange(n_tracks, dtype=np.int32)

    # one-to-one hand-off between slots
    perm = rng.permutation(2 * n_tracks)
    j = int(np.nonzero(perm == 0)[0][0])
    perm[0], perm[j] = perm[j], perm[0]                  # slot 0 hands off to itself
    bc = rng.choice([REFLECTIVE, PERIODIC, VACUUM], 2 * n_tracks,
                    p=[(1 - vacuum_fraction) / 2, (1 - vacuum_fraction) / 2, vacuum_fraction]).astype(np.uint8)
    bc[0] = REFLECTIVE
    nxt = perm // 2
    nxt_is_fwd = (perm % 2 == 0)
    a["trk_next_fwd"] = np.where(bc[0::2] == VACUUM, -1, nxt[0::2]).astype(np.int64)
    a["trk_next_bwd"] = np.where(bc[1::2] == VACUUM, -1, nxt[1::2]).astype(np.int64)
    a["trk_flags"] = (nxt_is_fwd[0::2].astype(np.uint8) | (nxt_is_fwd[1::2].astype(np.uint8) << 1))
    a["trk_bc_fwd"] = bc[0::2].copy()
    a["trk_bc_bwd"] = bc[1::2].copy()
    a["trk_phi"] = rng.uniform(0, np.pi, n_tracks)
    a["trk_theta"] = rng.uniform(0.2, np.pi - 0.2, n_tracks)

    w = rng.uniform(0.05, 0.3, A2 * P)
    a["quad_weight"] = w
    a["quad_sin_theta"] = rng.uniform(0.15, 1.0, A2 * P)
    # track-based volumes, V = sum over segments, both directions and polar angles of
    # w * l / sin(theta) / 4 pi, so that the tally conserves neutrons and iterations converge
    azim = a["trk_azim"][np.repeat(np.arange(n_tracks), nseg)]
    if solve_3d:
        pol = a["trk_polar"][np.repeat(np.arange(n_tracks), nseg)]
        wl = 2.0 * w[azim * P + pol] * length
    else:
        wl = np.zeros(ns)
        for p in range(NP):
            wl += 2.0 * w[azim * P + p] * length / a["quad_sin_theta"][azim * P + p]
    vol = np.bincount(fsr, weights=wl / (4 * np.pi), minlength=n_fsrs)
    vol[vol == 0.0] = 1.0
    a["fsr_volume"] = vol
    a["fsr_mat"] = fsr_mat
    a["fsr_centroid"] = np.zeros(3 * n_fsrs)

    st = rng.uniform(0.2, 1.5, (n_mats, G))
    ss = rng.uniform(0.0, 1.0, (n_mats, G, G))          # [mat][dest][orig] (Material.cpp:728-731)
    ss *= 0.6 * st[:, None, :] / ss.sum(axis=1, keepdims=True)   # out-scatter of a group = 0.6 sigma_t
    nusf = rng.uniform(0.0, 0.5, (n_mats, G))
    fiss = np.ones(n_mats, dtype=np.uint8)
    if n_mats > 1:
        nusf[0] = 0.0                                    # one non-fissionable material
        fiss[0] = 0
    if not fissionable:
        nusf[:] = 0.0
        fiss[:] = 0
    chi = rng.uniform(0.0, 1.0, (n_mats, G))
    chi /= chi.sum(axis=1, keepdims=True)
    chi[fiss == 0] = 0.0
    a["mat_sigma_t"] = st.ravel()
    a["mat_sigma_a"] = (0.4 * st).ravel()
    a["mat_sigma_f"] = (nusf / 2.4).ravel()
    a["mat_nu_sigma_f"] = nusf.ravel()
    a["mat_chi"] = chi.ravel()
    a["mat_sigma_s"] = ss.ravel()
    # fission matrix [mat][G dest][g src] = chi[G] * nu_sigma_f[g] (Material::buildFissionMatrix)
    a["mat_fiss_matrix"] = (chi[:, :, None] * nusf[:, None, :]).ravel()
    a["mat_fissionable"] = fiss
    ft.validate()
    return ft
