"""computeFlux / computeSource with the tracks sharded over several ranks (openmoc_b200.loops, what B200Solver runs
with one process per GPU), on a single-GPU lease: every rank's solver lives on cuda:0 of this process and the
all-reduce of the FSR tally is done by hand through the host.  Checked against the one-GPU drivers inside the library
(b200_compute_flux / b200_compute_source), which tests/test_gpu_parity.py checks against the oracle."""
import numpy as np
import pytest

from openmoc_b200.capi import TOTAL_SOURCE

pytestmark = pytest.mark.gpu


class SimulatedRanks:
    """The step methods of Solver, each applied to every rank's shard; transportSweep sums the ranks' tallies."""
    STEPS = ("setKeff", "zeroTrackFluxes", "flattenFSRFluxes", "storeFSRFluxes", "computeFSRSources",
             "addSourceToScalarFlux", "setFixedSourceByFSR", "setConvergenceThreshold")

    def __init__(self, ft, world):
        from openmoc_b200.partition import partition_by_chain
        from openmoc_b200.solver import B200Solver
        self.solvers = [B200Solver(p, global_tracks=ft) for p in partition_by_chain(ft, world)]
        assert sum(s.num_segments for s in self.solvers) == ft.n_segments

    def __getattr__(self, name):
        if name not in self.STEPS:
            raise AttributeError(name)
        return lambda *a: [getattr(s, name)(*a) for s in self.solvers][0]

    def transportSweep(self):
        for s in self.solvers:
            s.transportSweep()
        phi = sum(s.getFluxes() for s in self.solvers)
        for s in self.solvers:
            s.setFluxes(phi)

    def computeResidual(self, res_type):
        r = [s.computeResidual(res_type) for s in self.solvers]
        assert all(abs(x - r[0]) <= 1e-12 * abs(r[0]) for x in r)    # replicated FSR state: all ranks stop together
        return r[0]


def decks():
    from openmoc_b200.synth import make_tracks
    # (tracks, ranks, max_iters of the source loop, its k_eff and tolerance)
    yield make_tracks("simple-lattice", num_azim=8, spacing=0.1), 2, 600, 3.0, 1e-4     # converges (418 iterations)
    yield make_tracks("c5g7-2d", num_azim=4, spacing=0.5), 3, 30, 1.5, 1e-5              # stopped at max_iters


def test_fixed_source_loops_over_ranks_match_one_gpu():
    from openmoc_b200.loops import flux_loop, source_loop
    from openmoc_b200.solver import B200Solver
    for ft, world, src_iters, k_eff, src_tol in decks():
        one, ranks = B200Solver(ft), SimulatedRanks(ft, world)
        for s in (one, ranks):
            s.setFixedSourceByFSR(3, 1, 1.0)
            s.setFixedSourceByFSR(100, 2, 0.5)
        one.setConvergenceThreshold(1e-6)
        one.computeFlux(300)
        n = flux_loop(ranks, 300, 1e-6)
        assert n == one.getNumIterations() and 2 < n < 300
        for s in ranks.solvers:
            np.testing.assert_allclose(s.getFluxes(), one.getFluxes(), rtol=1e-9, atol=1e-14)

        one.setConvergenceThreshold(src_tol)
        one.computeSource(src_iters, k_eff=k_eff, res_type=TOTAL_SOURCE)
        n = source_loop(ranks, src_iters, k_eff, src_tol, TOTAL_SOURCE)
        assert n == one.getNumIterations()
        for s in ranks.solvers:
            np.testing.assert_allclose(s.getFluxes(), one.getFluxes(), rtol=1e-8, atol=1e-14)
