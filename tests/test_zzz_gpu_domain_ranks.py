"""Spatial domain decomposition (openmoc_b200/domain.py) on the GPU, on a single-GPU lease: the solver of every
box lives on cuda:0 of this process, the FSR tallies are summed and the interface fluxes moved by hand through the
host with the index lists of the exchange plan (what B200Solver does with NCCL all-reduce and send / recv).  The
checker is the oracle driven the same way (tests/test_domain.py::simulate): same cut tracks, same lag, so the two
agree iteration by iteration."""
import numpy as np
import pytest

from openmoc_b200.capi import FISSION_SOURCE

pytestmark = pytest.mark.gpu


def run_boxes(make_solver, parts, F, n_iter):
    """n_iter source iterations of Solver::computeEigenvalue over the boxes; returns k_eff and the flux"""
    world = len(parts)
    solvers = [make_solver(sub) for sub, _ in parts]
    plans = [p for _, p in parts]
    for s in solvers:
        s.setKeff(1.0); s.zeroTrackFluxes()
        s.flattenFSRFluxes(0.0); s.storeFSRFluxes()
        s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
    for i in range(n_iter):
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
        for s, ps in zip(solvers, psi):
            s.setStartFluxes(ps.ravel()); s.setFluxes(phi)
            s.addSourceToScalarFlux(); s.computeKeff(); s.normalizeFluxes()
            s.computeResidual(FISSION_SOURCE); s.storeFSRFluxes()
    return solvers[0].getKeff(), solvers[0].getFluxes()


@pytest.mark.parametrize("model,azim,spacing,world,domains", [("simple-lattice", 8, 0.1, 4, (2, 2)),
                                                              ("c5g7-2d", 4, 0.5, 6, (3, 2))])
def test_boxes_on_the_gpu_follow_the_oracle(model, azim, spacing, world, domains):
    from openmoc_b200.domain import partition_by_domain
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks
    from oracle.oracle_py import OracleSolver
    ft = make_tracks(model, num_azim=azim, spacing=spacing)
    parts = partition_by_domain(ft, world, domains)
    F = ft.fluxes_per_track
    k_gpu, phi_gpu = run_boxes(lambda sub: B200Solver(sub, global_tracks=ft), parts, F, 10)
    k_cpu, phi_cpu = run_boxes(OracleSolver, parts, F, 10)
    assert abs(k_gpu - k_cpu) * 1e5 < 1e-3                      # pcm
    np.testing.assert_allclose(phi_gpu, phi_cpu, rtol=1e-6, atol=1e-12)


def test_3d_boxes_on_the_gpu_follow_the_oracle():
    """2 x 2 x 2 boxes of an explicit 3D track set (the decomposition of configs[4], profile/models/c5g7/
    c5g7-3d-cmfd.cpp:557 `geometry.setDomainDecomposition(nx, ny, nz, ...)`, on the simple lattice)"""
    from openmoc_b200.domain import partition_by_domain
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks_3d
    from oracle.oracle_py import OracleSolver
    ft = make_tracks_3d("simple-lattice", num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9, n_axial=2, expand=True)
    parts = partition_by_domain(ft, 8)
    F = ft.fluxes_per_track
    k_gpu, phi_gpu = run_boxes(lambda sub: B200Solver(sub, global_tracks=ft), parts, F, 10)
    k_cpu, phi_cpu = run_boxes(OracleSolver, parts, F, 10)
    assert abs(k_gpu - k_cpu) * 1e5 < 1e-3
    np.testing.assert_allclose(phi_gpu, phi_cpu, rtol=1e-6, atol=1e-12)
