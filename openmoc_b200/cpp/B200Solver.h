/**
 * @file B200Solver.h
 * @brief `B200Solver : public Solver` - the drop-in plug-in that runs OpenMOC's
 *        transport sweep and the per-FSR steps around it on an NVIDIA B200
 *        through the C ABI of include/b200moc.h.
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
 * Every pure virtual of src/Solver.h:334-431 is a one-line call into libb200moc.so.
 * Host mirrors of the FSR arrays (_scalar_flux, _old_scalar_flux, _reduced_sources)
 * are kept allocated so that the base class's non-virtual helpers
 * (getFluxesArray, dumpFSRFluxes, getFSRSource ...) keep working; they are refreshed
 * by syncHostMirrors(), which getFluxes()/getFlux()/storeFSRFluxes() call.
 *
 * Not supported in this build (log_printf(ERROR) like GPUSolver.cu:1156-1161):
 * CMFD flux update, k_eff from neutron balance, linear source, domain decomposition.
 */
#ifndef B200SOLVER_H_
#define B200SOLVER_H_

#include "Solver.h"
#include "b200_flatten.h"

struct b200_solver;

class B200Solver : public Solver {

private:
  b200_solver* _h;
  B200FlatTracks _flat;
  long _flattened_segments;
  bool _materials_dirty, _fixed_dirty, _mirror_stale;
  int _device, _precision;
  double _device_keff;

  void check(int status, const char* what);
  void ensureDevice();
  void pushMaterialsIfDirty();
  void pushFixedSourcesIfDirty();
  void pushKeff();

protected:
  /* Solver pure virtuals, src/Solver.h:334-431 */
  void initializeFluxArrays();
  void initializeSourceArrays();
  void zeroTrackFluxes();
  void flattenFSRFluxes(FP_PRECISION value);
  void flattenFSRFluxesChiSpectrum();
  void storeFSRFluxes();
  double normalizeFluxes();
  void computeStabilizingFlux();
  void stabilizeFlux();
  void computeFSRSources(int iteration);
  void computeFSRFissionSources();
  void computeFSRScatterSources();
  double computeResidual(residualType res_type);
  void computeKeff();
  void addSourceToScalarFlux();
  void transportSweep();

  /* hooks (virtual in the base) */
  void initializeExpEvaluators();
  void initializeMaterials(solverMode mode);
  void initializeCmfd();

public:
  B200Solver(TrackGenerator* track_generator = NULL, int device = 0, int precision = 0);
  virtual ~B200Solver();

  void getFluxes(FP_PRECISION* out_fluxes, int num_fluxes);
  void setFluxes(FP_PRECISION* in_fluxes, int num_fluxes);
  double getFlux(long fsr_id, int group);
  double getFSRSource(long fsr_id, int group);
  void setFixedSourceByFSR(long fsr_id, int group, double source);
  void resetFixedSources();
  void initializeFixedSources();
  void computeFSRFissionRates(double* fission_rates, long num_FSRs, bool nu = false);

  /** Copy phi, old phi and q from the device into the base-class host arrays. */
  void syncHostMirrors();
  /** Fused device-side source iteration (b200_compute_eigenvalue): same results as
   *  computeEigenvalue() without a host round trip per step. */
  void computeEigenvalueFused(int max_iters = 1000, residualType res_type = FISSION_SOURCE);
  /** Accumulated device time of the sweep kernel (ms) and number of sweeps. */
  void getSweepStats(double* ms, long* sweeps);
};

#endif /* B200SOLVER_H_ */
