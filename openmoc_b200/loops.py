"""Host-driven source-iteration loops over the Solver virtuals.

With one GPU the whole loop runs inside the library (b200_compute_flux / b200_compute_source, device-side
stopping rule).  With one process per GPU the transport sweep has a collective in its middle (the all-reduce of
the FSR tallies, the hand-over of boundary fluxes between ranks: B200Solver.transportSweep), so the fixed-source
loops are driven from the host, step by step through the same virtuals the reference's base class calls, on
every rank alike: the FSR-side state is replicated, every rank sees the same residual and stops at the same
iteration without a further exchange.

The `solver` argument is anything with the step methods of `Solver` (src/Solver.h:334-431): B200Solver on the
GPUs; the CPU tests drive the same functions over gloo with a stand-in (tests/test_distributed.py).
"""
from .capi import SCALAR_FLUX, TOTAL_SOURCE


def flux_loop(solver, max_iters: int, tol: float, only_fixed_source: bool = True) -> int:
    """Solver::computeFlux (src/Solver.cpp:1352-1420): the sources are computed once, from the fixed source (and,
    unless only_fixed_source, the flux the solver holds); returns the reference's _num_iterations."""
    solver.setKeff(1.0)
    if only_fixed_source:                       # :1381-1387
        solver.zeroTrackFluxes()
        solver.flattenFSRFluxes(0.0)
        solver.storeFSRFluxes()
    solver.computeFSRSources(0)                 # :1390
    for i in range(int(max_iters)):             # :1397-1413
        solver.transportSweep()
        solver.addSourceToScalarFlux()
        residual = solver.computeResidual(SCALAR_FLUX)
        solver.storeFSRFluxes()
        if i > 1 and residual < tol:
            return i
    return int(max_iters)


def source_loop(solver, max_iters: int, k_eff: float, tol: float, res_type: int = TOTAL_SOURCE) -> int:
    """Solver::computeSource (src/Solver.cpp:1459-1516): source iteration at a fixed k_eff."""
    if k_eff <= 0.0:                            # :1464-1466
        raise ValueError("The Solver is unable to compute the source with keff = %f since it is not a positive value"
                         % k_eff)
    solver.setKeff(float(k_eff))
    solver.zeroTrackFluxes()                    # computeInitialFluxGuess(true)
    solver.flattenFSRFluxes(1.0)
    solver.storeFSRFluxes()
    for i in range(int(max_iters)):             # :1492-1508
        solver.computeFSRSources(i)
        solver.transportSweep()
        solver.addSourceToScalarFlux()
        residual = solver.computeResidual(res_type)
        solver.storeFSRFluxes()
        if i > 1 and residual < tol:
            return i
    return int(max_iters)
