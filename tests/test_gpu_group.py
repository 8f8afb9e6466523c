"""Several devices behind one solver handle (b200_set_devices; csrc/group.cuh): the library shards the
tracks by chain, runs replicated FSR steps and sums the shards' tallies with its own two-shot
all-reduce over peer memory.  On a single-GPU lease the same device is listed two or three times
(several shards on one GPU: every code path but the NVLink hop); with more GPUs the real ones too.
Checked against the single-device solver, the oracle and the reference's goldens."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_case
from openmoc_b200.capi import FISSION_SOURCE, B200Error
from oracle.oracle_py import OracleSolver, format_harness_results

pytestmark = pytest.mark.gpu

GOLDENS = json.load(open(os.path.join(GOLDEN, "ref_goldens.json")))


def device_lists():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 1
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], list(range(min(n, 8)))]
    return lists


@pytest.mark.parametrize("devices", device_lists())
def test_group_eigenvalue_matches_single_device_and_oracle(devices):
    from openmoc_b200.solver import B200Solver
    ft, ref = load_case("simple_lattice")
    one, grp = B200Solver(ft), B200Solver(ft, devices=devices)
    for s in (one, grp):
        s.setConvergenceThreshold(1e-5)
        s.computeEigenvalue(500, FISSION_SOURCE)
    o = OracleSolver(ft)
    n = o.computeEigenvalue(500, 1e-5, FISSION_SOURCE)
    assert grp.getNumIterations() == one.getNumIterations() == n == ref["iterations"]
    assert abs(grp.getKeff() - o.getKeff()) * 1e5 < 1e-4                 # north star: 1 pcm
    np.testing.assert_allclose(grp.getFluxes(), o.getFluxes(), rtol=2e-9)   # north star: 1e-4
    np.testing.assert_allclose(grp.getFluxes(), one.getFluxes(), rtol=1e-10)
    # the reference's golden file, from the group
    import hashlib
    text = format_harness_results(grp.getNumIterations(), grp.getKeff(), grp.getFluxes())
    assert hashlib.sha512(text.encode()).hexdigest() == GOLDENS["test_forward_simple_lattice"].strip()
    assert grp.integrationsPerSweep() == one.integrationsPerSweep()


def test_group_deterministic_tally_is_bitwise_equal_to_one_device():
    from openmoc_b200.solver import B200Solver
    ft, _ = load_case("simple_lattice")
    one = B200Solver(ft, deterministic=True)
    one.setConvergenceThreshold(1e-5)
    one.computeEigenvalue(500, FISSION_SOURCE)
    for devices in device_lists():
        grp = B200Solver(ft, devices=devices, deterministic=True)
        grp.setConvergenceThreshold(1e-5)
        grp.computeEigenvalue(500, FISSION_SOURCE)
        assert grp.getNumIterations() == one.getNumIterations()
        assert grp.getKeff() == one.getKeff()
        assert np.array_equal(grp.getFluxes(), one.getFluxes())


def test_group_step_by_step_api_and_start_fluxes():
    """the Solver virtuals one by one (what the plug-in's base-class loop calls) + getStartFluxes /
    setStartFluxes gathered from / scattered to the shards by global track id"""
    from openmoc_b200.solver import B200Solver
    ft, _ = load_case("pin_cell")
    one, grp = B200Solver(ft), B200Solver(ft, devices=[0, 0])
    for s in (one, grp):
        s.zeroTrackFluxes(); s.flattenFSRFluxes(1.0); s.normalizeFluxes(); s.storeFSRFluxes()
        for i in range(5):
            s.computeFSRSources(i); s.transportSweep(); s.addSourceToScalarFlux()
            s.computeKeff(); s.normalizeFluxes(); s.computeResidual(FISSION_SOURCE); s.storeFSRFluxes()
    assert abs(one.getKeff() - grp.getKeff()) < 1e-13
    np.testing.assert_allclose(grp.getFluxes(), one.getFluxes(), rtol=1e-12)
    psi1, psig = one.getStartFluxes(), grp.getStartFluxes()
    np.testing.assert_allclose(psig, psi1, rtol=1e-5, atol=1e-12)
    grp.setStartFluxes(psi1 * 2)
    np.testing.assert_array_equal(grp.getStartFluxes(), (psi1 * 2).astype(np.float32))


def test_group_fixed_source_flux():
    """tests/test_compute_flux golden (water box, fixed source) from a two-shard group"""
    from openmoc_b200.solver import B200Solver
    ft, ref = load_case("water_box")
    one, grp = B200Solver(ft), B200Solver(ft, devices=[0, 0])
    for s in (one, grp):
        for fsr in ref["source_fsrs"]:
            for g, v in ((1, 1.0), (2, 0.5), (3, 0.25)):
                s.setFixedSourceByFSR(fsr, g, v)
        s.setConvergenceThreshold(1e-5)
        s.computeFlux(100, only_fixed_source=True)
    assert grp.getNumIterations() == one.getNumIterations()
    np.testing.assert_allclose(grp.getFluxes(), one.getFluxes(), rtol=1e-10)


def test_group_linear_source():
    from openmoc_b200.solver import B200Solver
    ft, ref = load_case("simple_lattice_ls")
    one, grp = B200Solver(ft, linear_source=True), B200Solver(ft, linear_source=True, devices=[0, 0, 0])
    for s in (one, grp):
        s.setConvergenceThreshold(1e-5)
        s.computeEigenvalue(500, FISSION_SOURCE)
    assert grp.getNumIterations() == one.getNumIterations() == ref["iterations"]
    assert abs(grp.getKeff() - ref["keff"]) * 1e5 < 1e-3
    np.testing.assert_allclose(grp.getFluxes(), one.getFluxes(), rtol=1e-9)
    np.testing.assert_allclose(grp.getFluxMoments(), one.getFluxMoments(), rtol=1e-6, atol=1e-11)


def test_group_3d_on_the_fly_tracks():
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks_3d
    ft = make_tracks_3d("simple-lattice", num_azim=4, spacing=0.24, num_polar=2, z_spacing=0.9, n_axial=2, expand=False)
    one, grp = B200Solver(ft), B200Solver(ft, devices=[0, 0, 0])
    assert grp.num_segments == one.num_segments > 0
    for s in (one, grp):
        s.setConvergenceThreshold(1e-5)
        s.computeEigenvalue(500, FISSION_SOURCE)
    assert grp.getNumIterations() == one.getNumIterations()
    assert abs(grp.getKeff() - one.getKeff()) * 1e5 < 1e-4
    np.testing.assert_allclose(grp.getFluxes(), one.getFluxes(), rtol=1e-8)
    np.testing.assert_allclose(grp.getVolumes(), one.getVolumes(), rtol=1e-12)


def test_group_refuses_what_it_cannot_do():
    from openmoc_b200.solver import B200Solver
    ft, _ = load_case("pin_cell")
    with pytest.raises(B200Error, match="between 1 and 16"):
        B200Solver(ft, devices=[0] * 17)
    grp = B200Solver(ft, devices=[0, 0])
    with pytest.raises(B200Error, match="neutron balance"):
        grp.setKeffFromNeutronBalance()


@pytest.mark.parametrize("devices", device_lists())
def test_group_track_partition_hands_fluxes_over_through_peer_memory(devices, monkeypatch):
    """Tracks dealt one by one (forced here; automatic when a deck has few chains, like the fully
    reflective lattice below): a hand-off whose next track lives on another shard is stored by the sweep
    kernel straight into that shard's start-flux buffer.  Same iteration count and fluxes as one device -
    unlike the reference's domain decomposition, the exchanged fluxes do not lag an iteration."""
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks
    monkeypatch.setenv("B200_GROUP_PARTITION", "track")
    for ft, iters in ((load_case("simple_lattice")[0], 500), (make_tracks("c5g7-2d", num_azim=4, spacing=0.5), 40)):
        one, grp = B200Solver(ft), B200Solver(ft, devices=devices)
        for s in (one, grp):
            s.setConvergenceThreshold(1e-5)
            s.computeEigenvalue(iters, FISSION_SOURCE)
        assert grp.getNumIterations() == one.getNumIterations()
        assert abs(grp.getKeff() - one.getKeff()) * 1e5 < 1e-4
        np.testing.assert_allclose(grp.getFluxes(), one.getFluxes(), rtol=1e-9)
        np.testing.assert_allclose(grp.getStartFluxes(), one.getStartFluxes(), rtol=1e-5, atol=1e-12)


def test_group_more_shards_than_chains_falls_back_to_track_partition():
    from openmoc_b200.solver import B200Solver
    from openmoc_b200.synth import make_tracks
    ft = make_tracks("simple-lattice", num_azim=8, spacing=0.1)          # fully reflective: two chains
    one, grp = B200Solver(ft), B200Solver(ft, devices=[0] * 5)
    for s in (one, grp):
        s.setConvergenceThreshold(1e-5)
        s.computeEigenvalue(400, FISSION_SOURCE)
    assert grp.getNumIterations() == one.getNumIterations()
    np.testing.assert_allclose(grp.getFluxes(), one.getFluxes(), rtol=1e-9)
