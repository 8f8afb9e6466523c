/**
 * @file B200Solver.cpp
 * @brief See B200Solver.h.  Plug-in side of the boundary: compiled against the
 *        reference's headers, links libb200moc.so.  No numerics here.
 */
#include "B200Solver.h"

#include <cstring>

#include "TrackGenerator3D.h"
#include "Cmfd.h"
#include "../../include/b200moc.h"

B200Solver::B200Solver(TrackGenerator* track_generator, int device, int precision)
    : Solver(track_generator) {
  _h = NULL;
  _flattened_segments = -1;
  _materials_dirty = false;
  _fixed_dirty = false;
  _mirror_stale = false;
  _device = device;
  _precision = precision;
  _device_keff = -1.;
  _gpu_solver = true;   /* switches the wording of the timer report, Solver.cpp:1908-1929 */
}

B200Solver::~B200Solver() {
  if (_h != NULL) b200_destroy(_h);
}

/* CUDA / library errors follow the reference's convention: log_printf(ERROR) throws
 * std::logic_error, which SWIG turns into a Python RuntimeError (src/log.cpp:535-599). */
void B200Solver::check(int status, const char* what) {
  if (status != 0)
    log_printf(ERROR, "B200Solver::%s failed: %s", what, b200_last_error());
}

/* (Re)flatten the tracks and upload everything; runs from initializeExpEvaluators(),
 * i.e. after FSR centroids are final and over-long segments have been split
 * (Solver.cpp:716-741) - the one point every compute* entry passes. */
void B200Solver::ensureDevice() {
  long n_seg = _track_generator->getNumSegments();
  if (_h != NULL && n_seg == _flattened_segments) return;
  if (_h != NULL) { b200_destroy(_h); _h = NULL; }

  b200_flatten(_track_generator, &_flat, false);
  b200_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.num_groups = _flat.num_groups;
  cfg.num_azim = _flat.num_azim;
  cfg.num_polar = _flat.num_polar;
  cfg.solve_3d = _flat.solve_3d;
  cfg.n_tracks = _flat.n_tracks;
  cfg.n_segments = _flat.n_segments;
  cfg.n_fsrs = _flat.n_fsrs;
  cfg.n_materials = _flat.n_materials;
  cfg.device = _device;
  cfg.precision = _precision;
  check(b200_create(&cfg, &_h), "b200_create");
  check(b200_upload_tracks(_h, _flat.seg_length.data(), _flat.seg_fsr.data(), _flat.trk_seg_offset.data(),
                           _flat.trk_azim.data(), _flat.trk_polar.data(), _flat.trk_next_fwd.data(),
                           _flat.trk_next_bwd.data(), _flat.trk_flags.data(), _flat.trk_bc_fwd.data(),
                           _flat.trk_bc_bwd.data()), "b200_upload_tracks");
  check(b200_upload_quadrature(_h, _flat.quad_weight.data(), _flat.quad_sin_theta.data()), "b200_upload_quadrature");
  check(b200_upload_fsrs(_h, _flat.fsr_volume.data(), _flat.fsr_mat.data()), "b200_upload_fsrs");
  check(b200_upload_materials(_h, _flat.mat_sigma_t.data(), _flat.mat_sigma_s.data(), _flat.mat_fiss_matrix.data(),
                              _flat.mat_nu_sigma_f.data(), _flat.mat_sigma_f.data(), _flat.mat_chi.data(),
                              _flat.mat_fissionable.data()), "b200_upload_materials");
  check(b200_finalize(_h), "b200_finalize");
  if (_stabilize_transport)
    check(b200_stabilize_transport(_h, _stabilization_factor, (int)_stabilization_type), "b200_stabilize_transport");
  check(b200_allow_negative_fluxes(_h, _negative_fluxes_allowed), "b200_allow_negative_fluxes");
  /* the segment stream only lives on the device from here on */
  std::vector<double>().swap(_flat.seg_length);
  std::vector<int32_t>().swap(_flat.seg_fsr);
  std::vector<int32_t>().swap(_flat.seg_mat);
  _flattened_segments = n_seg;
  _materials_dirty = false;
  _fixed_dirty = true;
  _device_keff = -1.;
}

void B200Solver::pushMaterialsIfDirty() {
  if (!_materials_dirty || _h == NULL) return;
  B200FlatTracks tmp;
  /* cheap: re-read the material tables only */
  Geometry* geometry = _track_generator->getGeometry();
  std::map<int, Material*> mats = geometry->getAllMaterials();
  int G = _num_groups, m = 0;
  size_t n = mats.size();
  tmp.mat_sigma_t.assign(n * G, 0.); tmp.mat_nu_sigma_f = tmp.mat_sigma_f = tmp.mat_chi = tmp.mat_sigma_t;
  tmp.mat_sigma_s.assign(n * G * G, 0.); tmp.mat_fiss_matrix = tmp.mat_sigma_s;
  tmp.mat_fissionable.assign(n, 0);
  for (std::map<int, Material*>::iterator it = mats.begin(); it != mats.end(); ++it, ++m) {
    Material* mat = it->second;
    tmp.mat_fissionable[m] = mat->isFissionable();
    for (int g = 0; g < G; g++) {
      tmp.mat_sigma_t[m * G + g] = mat->getSigmaT()[g];
      tmp.mat_nu_sigma_f[m * G + g] = mat->getNuSigmaF()[g];
      tmp.mat_chi[m * G + g] = mat->getChi()[g];
    }
    for (int i = 0; i < G * G; i++) {
      tmp.mat_sigma_s[(size_t)m * G * G + i] = mat->getSigmaS()[i];
      tmp.mat_fiss_matrix[(size_t)m * G * G + i] = mat->getFissionMatrix()[i];
    }
  }
  check(b200_upload_materials(_h, tmp.mat_sigma_t.data(), tmp.mat_sigma_s.data(), tmp.mat_fiss_matrix.data(),
                              tmp.mat_nu_sigma_f.data(), NULL, tmp.mat_chi.data(), tmp.mat_fissionable.data()),
        "b200_upload_materials");
  /* on a finalized solver b200_upload_materials refreshes the derived tables itself */
  _materials_dirty = false;
}

void B200Solver::pushFixedSourcesIfDirty() {
  if (!_fixed_dirty || _h == NULL) return;
  check(b200_reset_fixed_sources(_h), "b200_reset_fixed_sources");
  if (_fixed_sources_on) {
    std::map< std::pair<int, int>, FP_PRECISION >::iterator it;
    for (it = _fix_src_FSR_map.begin(); it != _fix_src_FSR_map.end(); ++it)
      check(b200_set_fixed_source_by_fsr(_h, it->first.first, it->first.second, it->second),
            "b200_set_fixed_source_by_fsr");
  }
  _fixed_dirty = false;
}

/* the base-class loops assign _k_eff directly (Solver.cpp:1372,1473,1566) */
void B200Solver::pushKeff() {
  if (_k_eff != _device_keff) {
    check(b200_set_keff(_h, _k_eff), "b200_set_keff");
    _device_keff = _k_eff;
  }
}

/* ------------------------------ hooks ------------------------------------ */
void B200Solver::initializeExpEvaluators() {
  Solver::initializeExpEvaluators();
  ensureDevice();
  check(b200_set_keff_from_neutron_balance(_h, !_keff_from_fission_rates), "b200_set_keff_from_neutron_balance");
}

void B200Solver::initializeMaterials(solverMode mode) {
  Solver::initializeMaterials(mode);
  /* adjoint mode transposes the production matrices in place (Solver.cpp:806-807) */
  if (_h != NULL) _materials_dirty = true;
}

void B200Solver::initializeCmfd() {
  Cmfd* cmfd = _geometry->getCmfd();
  if (cmfd != NULL && cmfd->isFluxUpdateOn())
    log_printf(ERROR, "CMFD acceleration is not supported by the B200Solver in this build");
  _cmfd = NULL;
}

/* host mirrors only; device arrays are (re)zeroed, which is what a fresh
 * CPUSolver::initializeFluxArrays (CPUSolver.cpp:281-370) gives */
void B200Solver::initializeFluxArrays() {
  long size = _num_FSRs * _num_groups;
  if (_scalar_flux != NULL && !_user_fluxes) delete [] _scalar_flux;
  if (_old_scalar_flux != NULL) delete [] _old_scalar_flux;
  _scalar_flux = new FP_PRECISION[size]();
  _old_scalar_flux = new FP_PRECISION[size]();
  _user_fluxes = false;
  pushMaterialsIfDirty();
  check(b200_zero_track_fluxes(_h), "b200_zero_track_fluxes");
  check(b200_flatten_fsr_fluxes(_h, 0.), "b200_flatten_fsr_fluxes");
  check(b200_store_fsr_fluxes(_h), "b200_store_fsr_fluxes");
}

void B200Solver::initializeSourceArrays() {
  long size = _num_FSRs * _num_groups;
  if (_reduced_sources != NULL) delete [] _reduced_sources;
  _reduced_sources = new FP_PRECISION[size]();
  if (_fixed_sources_on && !_fixed_sources_initialized) initializeFixedSources();
}

void B200Solver::initializeFixedSources() {
  Solver::initializeFixedSources();     /* cell / material maps -> FSR map */
  _fixed_sources_initialized = true;
  _fixed_dirty = true;
}

/* ------------------------- Solver pure virtuals -------------------------- */
void B200Solver::zeroTrackFluxes() { check(b200_zero_track_fluxes(_h), "zeroTrackFluxes"); }

void B200Solver::flattenFSRFluxes(FP_PRECISION value) {
  check(b200_flatten_fsr_fluxes(_h, value), "flattenFSRFluxes");
  _mirror_stale = true;
}

void B200Solver::flattenFSRFluxesChiSpectrum() {
  if (_chi_spectrum_material == NULL)
    log_printf(ERROR, "A flattening of the FSR fluxes for a chi spectrum was "
               "requested but no chi spectrum material was set.");
  std::map<int, Material*> mats = _geometry->getAllMaterials();
  int m = 0;
  for (std::map<int, Material*>::iterator it = mats.begin(); it != mats.end(); ++it, ++m)
    if (it->second == _chi_spectrum_material) break;
  check(b200_flatten_fsr_fluxes_chi_spectrum(_h, m), "flattenFSRFluxesChiSpectrum");
  _mirror_stale = true;
}

void B200Solver::storeFSRFluxes() {
  check(b200_store_fsr_fluxes(_h), "storeFSRFluxes");
}

double B200Solver::normalizeFluxes() {
  double norm = 0.;
  check(b200_normalize_fluxes(_h, &norm), "normalizeFluxes");
  _mirror_stale = true;
  return norm;
}

void B200Solver::computeStabilizingFlux() { check(b200_compute_stabilizing_flux(_h), "computeStabilizingFlux"); }
void B200Solver::stabilizeFlux() { check(b200_stabilize_flux(_h), "stabilizeFlux"); _mirror_stale = true; }

void B200Solver::computeFSRSources(int iteration) {
  pushFixedSourcesIfDirty();
  pushKeff();
  check(b200_compute_fsr_sources(_h, iteration), "computeFSRSources");
}
void B200Solver::computeFSRFissionSources() { check(b200_compute_fsr_fission_sources(_h), "computeFSRFissionSources"); }
void B200Solver::computeFSRScatterSources() { check(b200_compute_fsr_scatter_sources(_h), "computeFSRScatterSources"); }

double B200Solver::computeResidual(residualType res_type) {
  double residual = 0.;
  pushKeff();
  check(b200_compute_residual(_h, (int)res_type, &residual), "computeResidual");
  return residual;
}

void B200Solver::computeKeff() {
  pushKeff();
  check(b200_compute_keff(_h, &_k_eff), "computeKeff");
  _device_keff = _k_eff;
}

void B200Solver::addSourceToScalarFlux() {
  check(b200_add_source_to_scalar_flux(_h), "addSourceToScalarFlux");
  _mirror_stale = true;
}

/* The "Transport Sweep" timer split keeps meaning what the reference's report expects
 * (Solver.cpp:1901-1929): wall time of the sweep, the stream is drained before stopping. */
void B200Solver::transportSweep() {
  _timer->startTimer();
  check(b200_transport_sweep(_h), "transportSweep");
  check(b200_synchronize(_h), "transportSweep");
  _timer->stopTimer();
  _timer->recordSplit("Transport Sweep");
  _mirror_stale = true;
}

/* ------------------------------ public API ------------------------------- */
void B200Solver::syncHostMirrors() {
  if (_h == NULL || _scalar_flux == NULL) return;
  long n = _num_FSRs * _num_groups;
  check(b200_get_fluxes(_h, _scalar_flux, n), "syncHostMirrors");
  if (_reduced_sources != NULL) check(b200_get_fsr_sources(_h, _reduced_sources, n), "syncHostMirrors");
  _mirror_stale = false;
}

void B200Solver::getFluxes(FP_PRECISION* out_fluxes, int num_fluxes) {
  if (num_fluxes != _num_groups * _num_FSRs)
    log_printf(ERROR, "Unable to get FSR scalar fluxes since there are "
               "%d groups and %d FSRs which does not match the requested "
               "%d flux values", _num_groups, _num_FSRs, num_fluxes);
  if (_h == NULL)
    log_printf(ERROR, "Unable to get FSR scalar fluxes since they have not yet been allocated");
  check(b200_get_fluxes(_h, out_fluxes, num_fluxes), "getFluxes");
}

/* CPUSolver aliases the caller's buffer (CPUSolver.cpp:190); like GPUSolver::setFluxes
 * (GPUSolver.cu:1088) the values are copied to the device instead. */
void B200Solver::setFluxes(FP_PRECISION* in_fluxes, int num_fluxes) {
  if (num_fluxes != _num_groups * _num_FSRs)
    log_printf(ERROR, "Unable to set an array with %d flux values for %d "
               " groups and %d FSRs", num_fluxes, _num_groups, _num_FSRs);
  if (_h == NULL)
    log_printf(ERROR, "Unable to set FSR scalar fluxes before the solver is initialized "
               "(call initializeSolver first)");
  check(b200_set_fluxes(_h, in_fluxes, num_fluxes), "setFluxes");
  _mirror_stale = true;
}

double B200Solver::getFlux(long fsr_id, int group) {
  if (_mirror_stale) syncHostMirrors();
  return Solver::getFlux(fsr_id, group);
}

double B200Solver::getFSRSource(long fsr_id, int group) {
  syncHostMirrors();
  return Solver::getFSRSource(fsr_id, group);
}

void B200Solver::setFixedSourceByFSR(long fsr_id, int group, double source) {
  Solver::setFixedSourceByFSR(fsr_id, group, source);
  _fixed_dirty = true;
}

void B200Solver::resetFixedSources() {
  _fix_src_FSR_map.clear();
  _fix_src_cell_map.clear();
  _fix_src_material_map.clear();
  _fixed_dirty = true;
}

void B200Solver::computeFSRFissionRates(double* fission_rates, long num_FSRs, bool nu) {
  if (_h == NULL)
    log_printf(ERROR, "Unable to compute FSR fission rates since the "
               "source distribution has not been calculated");
  check(b200_compute_fsr_fission_rates(_h, fission_rates, num_FSRs, nu), "computeFSRFissionRates");
}

void B200Solver::computeEigenvalueFused(int max_iters, residualType res_type) {
  clearTimerSplits();
  _timer->startTimer();
  initializeMaterials(_solver_mode);
  initializeFSRs();
  countFissionableFSRs();
  initializeExpEvaluators();
  initializeFluxArrays();
  initializeSourceArrays();
  initializeCmfd();
  pushFixedSourcesIfDirty();
  int iters = 0;
  check(b200_compute_eigenvalue(_h, max_iters, _converge_thresh, (int)res_type, &iters), "computeEigenvalueFused");
  _num_iterations = iters;
  check(b200_get_keff(_h, &_k_eff), "computeEigenvalueFused");
  _device_keff = _k_eff;
  syncHostMirrors();
  _timer->stopTimer();
  _timer->recordSplit("Total time");
}

void B200Solver::getSweepStats(double* ms, long* sweeps) {
  int64_t n = 0, launches = 0;
  check(b200_get_sweep_stats(_h, ms, &n, &launches), "getSweepStats");
  if (sweeps) *sweeps = n;
}
