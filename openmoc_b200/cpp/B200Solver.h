/**
 * @file B200Solver.h
 * @brief `B200Solver : public Solver` - the drop-in plug-in that runs OpenMOC's flat-source
 *        transport sweep and the per-FSR steps around it on an NVIDIA B200 through the C ABI
 *        of include/b200moc.h.
 *
 * It is used exactly like `CPUSolver` / `GPUSolver` (src/CPUSolver.h,
 * src/accel/cuda/GPUSolver.h:78-179):
 * @code
 *   TrackGenerator tg(&geometry, num_azim, spacing);  tg.generateTracks();
 *   B200Solver solver(&tg);
 *   solver.setConvergenceThreshold(1e-5);
 *   solver.computeEigenvalue(1000);          // Solver.cpp:1542, unchanged base-class loop
 *   solver.getFluxes(fluxes, n);  solver.getKeff();
 * @endcode
 * Every pure virtual of src/Solver.h:334-431 is a one-line call into libb200moc.so
 * (B200SolverT.h).  Host mirrors of the FSR arrays (_scalar_flux, _old_scalar_flux,
 * _reduced_sources) stay allocated so that the base class's non-virtual helpers
 * (getFluxesArray, dumpFSRFluxes, getFSRSource ...) keep working; syncHostMirrors()
 * refreshes them.
 *
 * Not supported in this build (log_printf(ERROR) like GPUSolver.cu:1156-1161):
 * CMFD flux update and domain decomposition.
 */
#ifndef B200SOLVER_H_
#define B200SOLVER_H_

#include "B200SolverT.h"

class B200Solver : public B200SolverT<Solver> {
public:
  B200Solver(TrackGenerator* track_generator = NULL, int device = 0, int precision = 0)
      : B200SolverT<Solver>(track_generator, device, precision) {}
};

#endif /* B200SOLVER_H_ */
