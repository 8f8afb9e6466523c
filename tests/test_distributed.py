"""Multi-process path: azimuthal-pair partition + all-reduce of the FSR tally.

CPU (gloo, world_size 2): the host-side logic - partition, per-rank sweep (through the
oracle, the checker), torch.distributed all-reduce, replicated FSR steps - reproduces the
single-process solution.  GPU (nccl, needs >= 2 devices): the same through B200Solver.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tracks():
    from openmoc_b200.synth import make_tracks
    return make_tracks("simple-lattice", num_azim=8, spacing=0.1)


# ------------------------------------------------------------------ gloo / CPU
def _cpu_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from openmoc_b200.partition import partition_by_azim_pair
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    ft = _tracks()
    part = partition_by_azim_pair(ft, world)[rank]
    s = OracleSolver(part)
    # Solver::computeEigenvalue step by step; every FSR step is replicated on all ranks
    s.setKeff(1.0); s.zeroTrackFluxes()
    s.flattenFSRFluxes(0.0); s.storeFSRFluxes()
    s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
    k_prev, iters = 1.0, 0
    for i in range(400):
        s.computeFSRSources(i)
        s.transportSweep()
        phi = torch.from_numpy(s.getFluxes())
        dist.all_reduce(phi, op=dist.ReduceOp.SUM)          # the one data-path collective
        s.setFluxes(phi.numpy())
        s.addSourceToScalarFlux()
        s.computeKeff(); k = s.getKeff()
        s.normalizeFluxes()
        res = s.computeResidual(FISSION_SOURCE)
        dk = int(1e5 * (k - k_prev)); k_prev = k
        s.storeFSRFluxes(); iters += 1
        if res < 1e-5 and abs(dk) < 1:
            break
    np.save(os.path.join(out_dir, f"phi{rank}.npy"), s.getFluxes())
    np.save(os.path.join(out_dir, f"k{rank}.npy"), np.array([s.getKeff(), iters]))
    dist.destroy_process_group()


def test_gloo_world2_matches_single_process(tmp_path):
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    port = _free_port()
    mp.spawn(_cpu_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ft = _tracks()
    ref = OracleSolver(ft)
    n = ref.computeEigenvalue(400, 1e-5, FISSION_SOURCE)
    for r in range(2):
        k, iters = np.load(os.path.join(tmp_path, f"k{r}.npy"))
        phi = np.load(os.path.join(tmp_path, f"phi{r}.npy"))
        assert int(iters) == n
        assert abs(k - ref.getKeff()) * 1e5 < 1e-4
        np.testing.assert_allclose(phi, ref.getFluxes(), rtol=1e-8)
    # ranks hold bit-identical replicated state
    assert np.array_equal(np.load(os.path.join(tmp_path, "phi0.npy")), np.load(os.path.join(tmp_path, "phi1.npy")))


def _cpu_worker_track(rank, world, port, out_dir):
    """partition_by_track + the isend/irecv exchange of boundary fluxes, over gloo"""
    import sys
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from openmoc_b200.partition import partition_by_track, exchange_boundary_fluxes
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    ft = _tracks()
    part, plan = partition_by_track(ft, world)[rank]
    F = part.fluxes_per_track
    s = OracleSolver(part)
    s.setKeff(1.0); s.zeroTrackFluxes()
    s.flattenFSRFluxes(0.0); s.storeFSRFluxes()
    s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
    k_prev, iters = 1.0, 0
    for i in range(400):
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
        if res < 1e-5 and abs(dk) < 1:
            break
    np.save(os.path.join(out_dir, f"phi{rank}.npy"), s.getFluxes())
    np.save(os.path.join(out_dir, f"k{rank}.npy"), np.array([s.getKeff(), iters]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_track_partition_with_flux_exchange(tmp_path, world):
    from oracle.oracle_py import OracleSolver, FISSION_SOURCE
    port = _free_port()
    mp.spawn(_cpu_worker_track, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ref = OracleSolver(_tracks())
    n = ref.computeEigenvalue(400, 1e-5, FISSION_SOURCE)
    for r in range(world):
        k, iters = np.load(os.path.join(tmp_path, f"k{r}.npy"))
        phi = np.load(os.path.join(tmp_path, f"phi{r}.npy"))
        assert int(iters) == n
        assert abs(k - ref.getKeff()) * 1e5 < 1e-4
        np.testing.assert_allclose(phi, ref.getFluxes(), rtol=1e-8)


class _ShardStepper:
    """The step methods of Solver over one rank's shard: the oracle sweeps the shard, the ranks sum their FSR
    tallies - what B200Solver.transportSweep does over NCCL.  Drives openmoc_b200.loops on the CPU."""
    def __init__(self, part, reduce_phi):
        from oracle.oracle_py import OracleSolver
        self.o, self.reduce_phi = OracleSolver(part), reduce_phi

    def __getattr__(self, name):
        return getattr(self.o, name)

    def transportSweep(self):
        self.o.transportSweep()
        self.o.setFluxes(self.reduce_phi(self.o.getFluxes()))


def _cpu_worker_fixed_source(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from openmoc_b200.loops import flux_loop, source_loop
    from openmoc_b200.partition import partition_by_azim_pair
    from oracle.oracle_py import TOTAL_SOURCE

    def reduce_phi(x):
        t = torch.from_numpy(x.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()
    part = partition_by_azim_pair(_tracks(), world)[rank]
    s = _ShardStepper(part, reduce_phi)
    s.setFixedSourceByFSR(3, 1, 1.0); s.setFixedSourceByFSR(100, 2, 0.5)
    n_flux = flux_loop(s, 300, 1e-6)
    phi_flux = s.getFluxes().copy()
    s = _ShardStepper(part, reduce_phi)
    s.setFixedSourceByFSR(3, 1, 1.0)
    n_src = source_loop(s, 600, 3.0, 1e-4, TOTAL_SOURCE)
    np.savez(os.path.join(out_dir, f"fixed{rank}.npz"), n_flux=n_flux, phi_flux=phi_flux, n_src=n_src,
             phi_src=s.getFluxes())
    dist.destroy_process_group()


def test_gloo_world2_fixed_source_loops(tmp_path):
    """computeFlux / computeSource with one process per rank (openmoc_b200.loops, the loops B200Solver runs at
    world > 1) against the oracle's own single-process drivers: same iteration counts, same fluxes."""
    from oracle.oracle_py import OracleSolver, TOTAL_SOURCE
    port = _free_port()
    mp.spawn(_cpu_worker_fixed_source, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = OracleSolver(_tracks())
    ref.setFixedSourceByFSR(3, 1, 1.0); ref.setFixedSourceByFSR(100, 2, 0.5)
    n_flux = ref.computeFlux(300, 1e-6)
    phi_flux = ref.getFluxes().copy()
    ref = OracleSolver(_tracks())
    ref.setFixedSourceByFSR(3, 1, 1.0)
    n_src = ref.computeSource(600, 3.0, 1e-4, TOTAL_SOURCE)
    assert 2 < n_flux < 300 and 2 < n_src < 600
    for r in range(2):
        d = np.load(os.path.join(tmp_path, f"fixed{r}.npz"))
        assert int(d["n_flux"]) == n_flux and int(d["n_src"]) == n_src
        np.testing.assert_allclose(d["phi_flux"], phi_flux, rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(d["phi_src"], ref.getFluxes(), rtol=1e-9, atol=1e-14)


# ------------------------------------------------------------------ nccl / GPU
def _gpu_worker(rank, world, port, out_dir, partition="pair"):
    import sys
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.capi import FISSION_SOURCE
    s = B200Solver(_tracks(), device=rank, process_group=dist.group.WORLD, partition=partition)
    s.setConvergenceThreshold(1e-5)
    s.computeEigenvalue(400, FISSION_SOURCE)
    np.save(os.path.join(out_dir, f"phi{rank}.npy"), s.getFluxes())
    np.save(os.path.join(out_dir, f"k{rank}.npy"), np.array([s.getKeff(), s.getNumIterations()]))
    s.close()          # the solver's CUDA graphs hold captured NCCL kernels: release them before the communicator
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(240)
@pytest.mark.parametrize("partition", ["pair", "chain", "track"])
def test_nccl_world2_matches_single_gpu(tmp_path, partition):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.capi import FISSION_SOURCE
    port = _free_port()
    mp.spawn(_gpu_worker, args=(2, port, str(tmp_path), partition), nprocs=2, join=True)
    one = B200Solver(_tracks())
    one.setConvergenceThreshold(1e-5)
    one.computeEigenvalue(400, FISSION_SOURCE)
    for r in range(2):
        k, iters = np.load(os.path.join(tmp_path, f"k{r}.npy"))
        phi = np.load(os.path.join(tmp_path, f"phi{r}.npy"))
        assert int(iters) == one.getNumIterations()
        assert abs(k - one.getKeff()) * 1e5 < 1e-3
        np.testing.assert_allclose(phi, one.getFluxes(), rtol=1e-7)


# ------------------------------------------- one process, several shards on ONE GPU
def _simulated_ranks(ft, world, max_iters=400, tol=1e-5, deterministic=False, linear=False):
    """The multi-GPU loop with every rank's solver living on cuda:0 of this process and the
    all-reduce done by hand (sum of the ranks' device tallies, written back to every rank): the
    same b200_iteration_begin -> reduce -> b200_iteration_end sequence B200Solver drives over NCCL,
    runnable on a single-GPU lease."""
    import ctypes as C
    from openmoc_b200.solver import B200Solver, _DeviceArray
    from openmoc_b200.capi import FISSION_SOURCE, check
    from openmoc_b200.partition import partition_by_chain
    parts = partition_by_chain(ft, world)
    solvers = [B200Solver(p, deterministic=deterministic, linear_source=linear,
                          global_tracks=ft) for p in parts]
    L = solvers[0]._lib
    views = []
    for s in solvers:
        s.useTorchStream()
        if deterministic:
            check(L.b200_defer_fixed_tally(s._h, 1))
        v = []
        names = [(b"scalar_flux_fixed", "<i8")] if deterministic else [(b"scalar_flux", "<f8")]
        if linear:
            names.append((b"scalar_flux_moments", "<f8"))
        for name, typ in names:
            p, n = C.c_void_p(), C.c_int64()
            check(L.b200_device_pointer(s._h, name, C.byref(p), C.byref(n)))
            v.append(torch.as_tensor(_DeviceArray(p.value, n.value, typ), device="cuda:0"))
        views.append(v)
        check(L.b200_eigen_loop_init(s._h, max_iters, tol))
    done, iters = C.c_int32(0), C.c_int32(0)
    i = 0
    while i < max_iters and not done.value:
        for s in solvers:
            check(L.b200_iteration_begin(s._h, i))
        for k in range(len(views[0])):
            total = views[0][k].clone()
            for r in range(1, world):
                total += views[r][k]
            for r in range(world):
                views[r][k].copy_(total)
        for s in solvers:
            if deterministic:
                check(L.b200_finish_fixed_tally(s._h))
            check(L.b200_iteration_end(s._h, i, FISSION_SOURCE, 1))
        i += 1
        for s in solvers:
            check(L.b200_eigen_loop_status(s._h, i, C.byref(done), C.byref(iters), None, None))
    return solvers, iters.value


@pytest.mark.gpu
def test_simulated_ranks_deterministic_tally_is_bitwise_equal_to_one_gpu():
    """The fixed-point scale derives from replicated quantities only, so 1, 2 and 3 ranks sum the
    same integers: fluxes and k_eff bit-identical (the scale used to follow each rank's
    own max |psi|)."""
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.capi import FISSION_SOURCE
    from openmoc_b200.synth import make_tracks
    # the fully reflective lattice has two chains; the C5G7 core (two vacuum sides) has many
    for ft, world, max_iters in ((_tracks(), 2, 400), (make_tracks("c5g7-2d", num_azim=4, spacing=0.5), 3, 30),
                                 (make_tracks("c5g7-2d", num_azim=4, spacing=0.5), 5, 30)):
        one = B200Solver(ft, deterministic=True)
        one.setConvergenceThreshold(1e-5)
        one.computeEigenvalue(max_iters, FISSION_SOURCE)
        solvers, iters = _simulated_ranks(ft, world, max_iters=max_iters, deterministic=True)
        assert iters == one.getNumIterations()
        for s in solvers:
            assert s.getKeff() == one.getKeff()
            assert np.array_equal(s.getFluxes(), one.getFluxes())


@pytest.mark.gpu
def test_simulated_ranks_linear_source_matches_one_gpu():
    """linear source across ranks: the moment tallies are reduced like the scalar flux"""
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.capi import FISSION_SOURCE
    from openmoc_b200.synth import make_tracks
    ft = make_tracks("simple-lattice", num_azim=8, spacing=0.1, linear_source=True)
    one = B200Solver(ft, linear_source=True)
    one.setConvergenceThreshold(1e-5)
    one.computeEigenvalue(400, FISSION_SOURCE)
    solvers, iters = _simulated_ranks(ft, 2, linear=True)
    assert iters == one.getNumIterations()
    for s in solvers:
        assert abs(s.getKeff() - one.getKeff()) * 1e5 < 1e-4
        np.testing.assert_allclose(s.getFluxes(), one.getFluxes(), rtol=1e-8)
        np.testing.assert_allclose(s.getFluxMoments(), one.getFluxMoments(), rtol=1e-6, atol=1e-10)


@pytest.mark.gpu
def test_simulated_ranks_3d_on_the_fly_tracks():
    """3D deck, device-side axial tracing, chains sharded over 3 ranks: volumes from all tracks,
    segments from each rank's shard"""
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.capi import FISSION_SOURCE
    from openmoc_b200.synth import make_tracks_3d
    ft = make_tracks_3d("simple-lattice", num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9, n_axial=2, expand=False)
    one = B200Solver(ft)
    one.setConvergenceThreshold(1e-5)
    one.computeEigenvalue(500, FISSION_SOURCE)
    solvers, iters = _simulated_ranks(ft, 3, max_iters=500)
    assert iters == one.getNumIterations()
    assert sum(s.num_segments for s in solvers) == one.num_segments
    for s in solvers:
        assert abs(s.getKeff() - one.getKeff()) * 1e5 < 1e-4
        np.testing.assert_allclose(s.getFluxes(), one.getFluxes(), rtol=1e-8)
