/**
 * @file B200LSSolver.h
 * @brief `B200LSSolver : public CPULSSolver` - linear-source drop-in.  The reference's own
 *        host pre-pass (CPULSSolver::initializeFSRs -> LinearExpansionGenerator,
 *        src/CPULSSolver.cpp:174-186, src/TrackTraversingAlgorithms.cpp:470-831) still
 *        computes the FSR expansion matrices and source constants; every per-iteration step
 *        (moment sources, the linear-source sweep with its four tallies, closure,
 *        normalisation) runs on the B200.
 */
#ifndef B200LSSOLVER_H_
#define B200LSSOLVER_H_

#include "CPULSSolver.h"
#include "B200SolverT.h"

class B200LSSolver : public B200SolverT<CPULSSolver> {
protected:
  bool isLinearSource() { return true; }
  void uploadExtras();
  void syncExtraMirrors();
  void pushExtraHostFlux();
  void pushExtraFixedSources();
  void allocateHostFluxMirrors();
  void allocateHostSourceMirrors();
public:
  B200LSSolver(TrackGenerator* track_generator = NULL, int device = 0)
      : B200SolverT<CPULSSolver>(track_generator, device, 0) {}
  /** flux moments, reference layout [r*3G + c*G + e] */
  void getFluxMoments(FP_PRECISION* out, long n);
};

#endif /* B200LSSOLVER_H_ */
