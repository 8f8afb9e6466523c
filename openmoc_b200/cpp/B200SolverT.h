/**
 * @file B200SolverT.h
 * @brief Implementation of the B200 plug-in as a class template over its OpenMOC base:
 *        B200Solver   = B200SolverT<Solver>        flat source   (see B200Solver.h)
 *        B200LSSolver = B200SolverT<CPULSSolver>+  linear source (see B200LSSolver.h)
 *        Everything is a thin forwarder into the C ABI of include/b200moc.h.
 */
#ifndef B200SOLVERT_H_
#define B200SOLVERT_H_

#include <cstring>
#include <map>
#include <vector>
#include <omp.h>

#include "Solver.h"
#include "TrackGenerator3D.h"
#include "Cmfd.h"
#include "b200_flatten.h"
#include "b200_cmfd_view.h"
#include "../../include/b200moc.h"

template <class Base>
class B200SolverT : public Base {

protected:
  /* members of Solver used below (dependent names) */
  using Base::_track_generator; using Base::_geometry; using Base::_num_groups; using Base::_num_FSRs;
  using Base::_scalar_flux; using Base::_old_scalar_flux; using Base::_reduced_sources;
  using Base::_user_fluxes; using Base::_fixed_sources_on; using Base::_fixed_sources_initialized;
  using Base::_fix_src_FSR_map; using Base::_fix_src_cell_map; using Base::_fix_src_material_map;
  using Base::_k_eff; using Base::_keff_from_fission_rates; using Base::_stabilize_transport;
  using Base::_stabilization_factor; using Base::_stabilization_type; using Base::_negative_fluxes_allowed;
  using Base::_chi_spectrum_material; using Base::_timer; using Base::_cmfd; using Base::_gpu_solver;
  using Base::_FSR_volumes;
  using Base::_converge_thresh; using Base::_SOLVE_3D; using Base::_num_iterations; using Base::_solver_mode;

  b200_solver* _h;
  B200FlatTracks _flat;
  long _flattened_segments;
  double _flattened_key;           /* fingerprint of the tracks the device image was built from */
  bool _materials_dirty, _fixed_dirty, _mirror_stale;
  bool _cmfd_active, _host_flux_newer;
  bool _cmfd_on_device;            /* user switch (setCmfdOnDevice); B200_HOST_CMFD=1 forces the host Cmfd */
  bool _cmfd_device_active;        /* this solve: collapse, diffusion solve and prolongation run on the device */
  double _cmfd_device_ms; long _cmfd_linear_iters;   /* accumulated over the solve */
  double _cmfd_device_keff;        /* the CMFD k_eff the device holds */
  Cmfd* _cmfd_suspended;           /* Cmfd whose flux update is switched off while the device does its work */
  std::vector<double> _cmfd_currents;
  int _device, _precision;
  std::vector<int> _devices;       /* more than one entry: one handle drives them all (b200_set_devices) */
  double _device_keff;

  void check(int status, const char* what);
  void ensureDevice();
  void pushMaterialsIfDirty();
  void pushFixedSourcesIfDirty();
  void pushKeff();
  void handCurrentsToCmfd();
  void configureDeviceCmfd();
  void restoreCmfdFluxUpdate();
  void tallyStartingCurrents();
  std::vector<float> _start_flux_host;
  void pushHostFluxIfNewer();

  /* customisation points of the linear-source subclass */
  virtual bool isLinearSource() { return false; }
  virtual void uploadExtras() {}
  virtual void syncExtraMirrors() {}
  virtual void pushExtraHostFlux() {}
  virtual void pushExtraFixedSources() {}
  virtual void allocateHostFluxMirrors() {
    long size = _num_FSRs * _num_groups;
    if (_scalar_flux != NULL && !_user_fluxes) delete [] _scalar_flux;
    if (_old_scalar_flux != NULL) delete [] _old_scalar_flux;
    _scalar_flux = new FP_PRECISION[size]();
    _old_scalar_flux = new FP_PRECISION[size]();
    _user_fluxes = false;
  }
  virtual void allocateHostSourceMirrors() {
    long size = _num_FSRs * _num_groups;
    if (_reduced_sources != NULL) delete [] _reduced_sources;
    _reduced_sources = new FP_PRECISION[size]();
  }

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
  B200SolverT(TrackGenerator* track_generator = NULL, int device = 0, int precision = 0);
  virtual ~B200SolverT();

  void getFluxes(FP_PRECISION* out_fluxes, int num_fluxes);
  void setFluxes(FP_PRECISION* in_fluxes, int num_fluxes);
  double getFlux(long fsr_id, int group);
  double getFSRSource(long fsr_id, int group);
  void setFixedSourceByFSR(long fsr_id, int group, double source);
  void resetFixedSources();
  void initializeFixedSources();
  void computeFSRFissionRates(double* fission_rates, long num_FSRs, bool nu = false);

  /** Host threads for what stays on the CPU (track flattening, the reference Cmfd, the
   *  linear-source pre-pass); same meaning as CPUSolver::setNumThreads (CPUSolver.cpp:132-159). */
  void setNumThreads(int num_threads) {
    if (num_threads <= 0)
      log_printf(ERROR, "Unable to set the number of threads to %d since it is less than or equal to 0", num_threads);
    omp_set_num_threads(num_threads);
  }
  /** Several GPUs behind this solver: the tracks are sharded by chain inside the library, FSR steps run
   *  replicated, the tallies (scalar flux, LS moments, CMFD currents) are summed over NVLink peer memory
   *  by the library's own all-reduce kernels (include/b200moc.h: b200_set_devices).  The role of the
   *  reference's MPI decomposition (Geometry::setDomainDecomposition + CPUSolver.cpp:545-1211) on one
   *  multi-GPU node; CMFD and the linear source work unchanged.  Call before the first compute*(). */
  void setNumDevices(int n) {
    if (n < 1) log_printf(ERROR, "Unable to use %d devices", n);
    std::vector<int> d(n);
    for (int i = 0; i < n; i++) d[i] = _device + i;
    setDevices(d);
  }
  /** Explicit device list (a device may repeat: several shards on one GPU). */
  void setDevices(const std::vector<int>& devices) {
    _devices = devices;
    if (_h != NULL) { b200_destroy(_h); _h = NULL; }      /* the device image is rebuilt on the next solve */
  }
  /** CMFD collapse / diffusion eigenvalue solve / prolongation on the device (default) or by the reference's
   *  host Cmfd fed with the device's fluxes and currents.  The host path also serves what the device path does
   *  not reproduce: the sigma-t rebalance and the neutron-balance check. */
  void setCmfdOnDevice(bool on) { _cmfd_on_device = on; }
  bool isCmfdOnDevice() { return _cmfd_device_active; }
  /** Device time (ms) of the CMFD solves of the last compute*() and their SOR iterations. */
  void getCmfdStats(double* ms, long* linear_iterations) { *ms = _cmfd_device_ms; *linear_iterations = _cmfd_linear_iters; }
  /** Solver::computeEigenvalue is not virtual; this overload only restores Cmfd::isFluxUpdateOn afterwards. */
  void computeEigenvalue(int max_iters = 1000, residualType res_type = FISSION_SOURCE) {
    Base::computeEigenvalue(max_iters, res_type);
    restoreCmfdFluxUpdate();
  }
  /** The summary of the base class with the Cmfd's flux update shown as the user set it (it is only switched
   *  off on the host object because the device does that work). */
  void printInputParamsSummary() {
    if (_cmfd_suspended != NULL) _cmfd_suspended->setFluxUpdateOn(true);
    Base::printInputParamsSummary();
    if (_cmfd_suspended != NULL) {
      _cmfd_suspended->setFluxUpdateOn(false);
      log_printf(NORMAL, "CMFD collapse, diffusion solve and prolongation: on the B200 device");
    }
  }
  /** Copy phi, old phi and q from the device into the base-class host arrays. */
  void syncHostMirrors();
  /** Fused device-side source iteration (b200_compute_eigenvalue): same results as
   *  computeEigenvalue() without a host round trip per step. */
  void computeEigenvalueFused(int max_iters = 1000, residualType res_type = FISSION_SOURCE);
  /** Accumulated device time of the sweep kernel (ms) and number of sweeps. */
  void getSweepStats(double* ms, long* sweeps);
};

/* ------------------------------------------------------------------------------------ */
template <class Base>
B200SolverT<Base>::B200SolverT(TrackGenerator* track_generator, int device, int precision)
    : Base(track_generator) {
  _h = NULL;
  _flattened_segments = -1;
  _flattened_key = -1.;
  _materials_dirty = false;
  _fixed_dirty = false;
  _mirror_stale = false;
  _cmfd_active = false;
  _cmfd_on_device = true;
  _cmfd_device_active = false;
  _cmfd_suspended = NULL;
  _cmfd_device_keff = 1.;
  _cmfd_device_ms = 0.; _cmfd_linear_iters = 0;
  _host_flux_newer = false;
  _device = device;
  _precision = precision;
  _device_keff = -1.;
  _gpu_solver = true;   /* switches the wording of the timer report, Solver.cpp:1908-1929 */
}

template <class Base>
B200SolverT<Base>::~B200SolverT() {
  restoreCmfdFluxUpdate();
  if (_h != NULL) b200_destroy(_h);
}

/* CUDA / library errors follow the reference's convention: log_printf(ERROR) throws
 * std::logic_error, which SWIG turns into a Python RuntimeError (src/log.cpp:535-599). */
template <class Base>
void B200SolverT<Base>::check(int status, const char* what) {
  if (status != 0)
    log_printf(ERROR, "B200Solver::%s failed: %s", what, b200_last_error());
}

/* (Re)flatten the tracks and upload everything; runs from initializeExpEvaluators(),
 * i.e. after FSR centroids are final and over-long segments have been split
 * (Solver.cpp:716-741) - the one point every compute* entry passes. */
template <class Base>
void B200SolverT<Base>::ensureDevice() {
  long n_seg = _track_generator->getNumSegments();
  /* The device image is reused only for the very same tracks: a re-traced geometry can keep its
   * segment count while volumes, quadrature or FSR numbering change, so the key also covers the
   * track counts, the FSR volumes and the quadrature weights. */
  double key = (double)n_seg * 1e-3 + (double)_track_generator->getNumTracks() + 7. * (double)_num_FSRs
             + 1e6 * (double)_num_groups;     /* a Material given another group structure between two solves */
  if (_FSR_volumes != NULL)
    for (long r = 0; r < _num_FSRs; r++) key += (double)_FSR_volumes[r] * (1. + 1e-3 * (double)(r % 977));
  /* ... and the identity of the FSR materials: the device's FSR -> material indices follow the order of
   * Geometry::getAllMaterials() at flatten time, which changes when cells are refilled with other (e.g. cloned)
   * Material objects between two solves (tests/test_multisim_materials) */
  if (this->_FSR_materials != NULL)
    for (long r = 0; r < _num_FSRs; r++)
      if (this->_FSR_materials[r] != NULL) key += 1e-3 * (double)this->_FSR_materials[r]->getId() * (double)(1 + r % 31);
  {
    Quadrature* q = _track_generator->getQuadrature();
    for (int a = 0; a < q->getNumAzimAngles() / 2; a++)
      for (int p = 0; p < q->getNumPolarAngles(); p++) key += 13. * q->getWeightInline(a, p) * (1 + a) * (3 + p);
  }
  if (_h != NULL && n_seg == _flattened_segments && key == _flattened_key) return;
  if (_h != NULL) { b200_destroy(_h); _h = NULL; }
  _flattened_key = key;

  /* On-the-fly 3D formations go to the device tracer (no 3D segment is made on the host), CMFD surfaces
   * included, unless a per-segment datum only the host traversal produces is needed: linear-source starting
   * points.  B200_HOST_OTF=1 forces the host expansion. */
  bool device_otf = b200_can_trace_on_device(_track_generator) && !isLinearSource() && getenv("B200_HOST_OTF") == NULL;
  b200_flatten(_track_generator, &_flat, isLinearSource(), device_otf);
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
  cfg.linear_source = isLinearSource() ? 1 : 0;
  check(b200_create(&cfg, &_h), "b200_create");
  if (_devices.size() > 1 || (_devices.size() == 1 && _devices[0] != _device))
    check(b200_set_devices(_h, (int)_devices.size(), _devices.data()), "b200_set_devices");
  if (_flat.device_otf) {
    check(b200_upload_otf_geometry(_h, _flat.n_tracks_2d, (int64_t)_flat.seg2d_length.size(), _flat.seg2d_length.data(),
                                   _flat.seg2d_ext.data(), _flat.trk2d_seg_offset.data(), _flat.n_extruded,
                                   _flat.ext_offset.data(), _flat.ext_mesh.data(), _flat.ext_fsr.data(), 0,
                                   _flat.otf_theta.data()), "b200_upload_otf_geometry");
    if (_flat.otf_cmfd)
      check(b200_upload_otf_cmfd(_h, _flat.seg2d_surf_fwd.data(), _flat.seg2d_surf_bwd.data(), _flat.fsr_cmfd_cell.data(),
                                 _flat.cmfd_nx, _flat.cmfd_ny, _flat.cmfd_nz, _flat.cmfd_z_planes.data()),
            "b200_upload_otf_cmfd");
    check(b200_upload_tracks_otf(_h, _flat.trk_2d.data(), _flat.trk_l0.data(), _flat.trk_z0.data(), _flat.trk_azim.data(),
                                 _flat.trk_polar.data(), _flat.trk_next_fwd.data(), _flat.trk_next_bwd.data(),
                                 _flat.trk_flags.data(), _flat.trk_bc_fwd.data(), _flat.trk_bc_bwd.data(), NULL),
          "b200_upload_tracks_otf");
  } else
  check(b200_upload_tracks(_h, _flat.seg_length.data(), _flat.seg_fsr.data(), _flat.trk_seg_offset.data(),
                           _flat.trk_azim.data(), _flat.trk_polar.data(), _flat.trk_next_fwd.data(),
                           _flat.trk_next_bwd.data(), _flat.trk_flags.data(), _flat.trk_bc_fwd.data(),
                           _flat.trk_bc_bwd.data()), "b200_upload_tracks");
  check(b200_upload_quadrature(_h, _flat.quad_weight.data(), _flat.quad_sin_theta.data()), "b200_upload_quadrature");
  check(b200_upload_fsrs(_h, _flat.fsr_volume.data(), _flat.fsr_mat.data()), "b200_upload_fsrs");
  check(b200_upload_materials(_h, _flat.mat_sigma_t.data(), _flat.mat_sigma_s.data(), _flat.mat_fiss_matrix.data(),
                              _flat.mat_nu_sigma_f.data(), _flat.mat_sigma_f.data(), _flat.mat_chi.data(),
                              _flat.mat_fissionable.data()), "b200_upload_materials");
  if (_flat.device_otf)
    check(b200_set_max_optical_length(_h, _track_generator->retrieveMaxOpticalLength()), "b200_set_max_optical_length");
  uploadExtras();
  {
    Cmfd* cmfd = _geometry->getCmfd();
    if (cmfd != NULL && cmfd->isFluxUpdateOn() && !_flat.device_otf)
      check(b200_upload_cmfd_surfaces(_h, _flat.seg_cmfd_fwd.data(), _flat.seg_cmfd_bwd.data()),
            "b200_upload_cmfd_surfaces");
  }
  check(b200_finalize(_h), "b200_finalize");
  /* the segment stream only lives on the device from here on */
  std::vector<double>().swap(_flat.seg_length);
  std::vector<int32_t>().swap(_flat.seg_fsr);
  std::vector<int32_t>().swap(_flat.seg_mat);
  std::vector<double>().swap(_flat.seg_start);
  std::vector<int32_t>().swap(_flat.seg_cmfd_fwd);
  std::vector<int32_t>().swap(_flat.seg_cmfd_bwd);
  _flattened_segments = n_seg;
  _materials_dirty = false;
  _fixed_dirty = true;
  _device_keff = -1.;
}

template <class Base>
void B200SolverT<Base>::pushMaterialsIfDirty() {
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

template <class Base>
void B200SolverT<Base>::pushFixedSourcesIfDirty() {
  if (!_fixed_dirty || _h == NULL) return;
  check(b200_reset_fixed_sources(_h), "b200_reset_fixed_sources");
  if (_fixed_sources_on) {
    std::map< std::pair<int, int>, FP_PRECISION >::iterator it;
    for (it = _fix_src_FSR_map.begin(); it != _fix_src_FSR_map.end(); ++it)
      check(b200_set_fixed_source_by_fsr(_h, it->first.first, it->first.second, it->second),
            "b200_set_fixed_source_by_fsr");
    pushExtraFixedSources();
  }
  _fixed_dirty = false;
}

/* the base-class loops assign _k_eff directly (Solver.cpp:1372,1473,1566) */
template <class Base>
void B200SolverT<Base>::pushKeff() {
  if (_k_eff != _device_keff) {
    check(b200_set_keff(_h, _k_eff), "b200_set_keff");
    _device_keff = _k_eff;
  }
}

/* ------------------------------ hooks ------------------------------------ */
template <class Base>
void B200SolverT<Base>::initializeExpEvaluators() {
  Base::initializeExpEvaluators();
  ensureDevice();
  /* solver options may change between two solves on the same tracks: pushed every time */
  check(b200_set_keff_from_neutron_balance(_h, !_keff_from_fission_rates), "b200_set_keff_from_neutron_balance");
  if (_stabilize_transport)
    check(b200_stabilize_transport(_h, _stabilization_factor, (int)_stabilization_type), "b200_stabilize_transport");
  check(b200_allow_negative_fluxes(_h, _negative_fluxes_allowed), "b200_allow_negative_fluxes");
}

template <class Base>
void B200SolverT<Base>::initializeMaterials(solverMode mode) {
  Base::initializeMaterials(mode);
  /* adjoint mode transposes the production matrices in place (Solver.cpp:806-807) */
  if (_h != NULL) _materials_dirty = true;
}

template <class Base>
void B200SolverT<Base>::initializeCmfd() {
  /* Solver::initializeCmfd (src/Solver.cpp:1145-1181) hands Cmfd the HOST arrays (_scalar_flux,
   * _reduced_sources, volumes, materials): the CMFD solve stays the reference's host code,
   * fed every iteration with the device's fluxes and surface currents (addSourceToScalarFlux)
   * and read back before the next device step (pushHostFluxIfNewer). */
  restoreCmfdFluxUpdate();
  Base::initializeCmfd();
  _cmfd_active = (_cmfd != NULL && _cmfd->isFluxUpdateOn());
  _cmfd_device_active = false;
  if (!_cmfd_active) {
    check(b200_set_cmfd_groups(_h, NULL, 0, 0), "b200_set_cmfd_groups");
    return;
  }
  if (_cmfd->isSigmaTRebalanceOn() && dynamic_cast<TrackGenerator3D*>(_track_generator) == NULL)
    log_printf(ERROR, "Starting currents not implemented yet for 2D MOC");     /* CPUSolver.cpp:532, the reference's own limit */
  std::vector<int32_t> map(_num_groups);
  for (int e = 0; e < _num_groups; e++) map[e] = _cmfd->getCmfdGroup(e);
  check(b200_set_cmfd_groups(_h, map.data(), _cmfd->getNumCmfdGroups(), _cmfd->getNumCells()),
        "b200_set_cmfd_groups");
  configureDeviceCmfd();
  /* the host copy of the currents is only needed when the reference's host Cmfd does the work */
  if (_cmfd_device_active) std::vector<double>().swap(_cmfd_currents);
  else _cmfd_currents.assign((size_t)_cmfd->getNumCells() * NUM_SURFACES * _cmfd->getNumCmfdGroups(), 0.);
}

/* Hands the mesh, the group structure, the FSR lists, the options and the k-nearest stencils of the Cmfd object
 * to the device (b200_cmfd_configure) and takes the per-iteration work away from it: with the flux update
 * switched off the base-class loop calls the virtual computeKeff() instead of Cmfd::computeKeff
 * (Solver.cpp:1627-1630), which is where the device solve runs. */
template <class Base>
void B200SolverT<Base>::configureDeviceCmfd() {
  const char* env = getenv("B200_HOST_CMFD");
  if (!_cmfd_on_device || (env != NULL && atoi(env) != 0)) return;
  B200CmfdView v;
  b200_read_cmfd(_cmfd, _num_FSRs, &v);
  if (v.balance_sigma_t || v.check_neutron_balance || _geometry->isDomainDecomposed()) return;   /* host Cmfd */
  b200_cmfd_config c;
  memset(&c, 0, sizeof c);
  c.num_x = v.num_x; c.num_y = v.num_y; c.num_z = v.num_z;
  c.num_cmfd_groups = v.num_cmfd_groups;
  for (int s = 0; s < NUM_FACES; s++) c.boundaries[s] = v.boundaries[s];
  c.linear_source = v.linear_source;
  c.flux_limiting = v.flux_limiting;
  c.centroid_update = v.centroid_update;
  c.axial_interpolation = v.num_z >= 3 ? v.use_axial_interpolation : 0;
  c.num_unbounded_iterations = v.num_unbounded_iterations;
  c.sor_factor = v.sor_factor;
  c.relaxation_factor = v.relaxation_factor;
  c.linalg_tolerance = MIN_LINALG_TOLERANCE;
  Quadrature* quad = _track_generator->getQuadrature();
  c.num_azim_2 = quad->getNumAzimAngles() / 2;
  c.num_polar_2 = quad->getNumPolarAngles() / 2;
  std::vector<double> wa(c.num_azim_2), st((size_t)c.num_azim_2 * c.num_polar_2), wp(st.size());
  for (int a = 0; a < c.num_azim_2; a++) {
    wa[a] = quad->getAzimWeight(a);
    for (int p = 0; p < c.num_polar_2; p++) {
      st[a * c.num_polar_2 + p] = quad->getSinTheta(a, p);
      wp[a * c.num_polar_2 + p] = quad->getPolarWeight(a, p);
    }
  }
  check(b200_cmfd_configure(_h, &c, v.widths_x.data(), v.widths_y.data(), v.widths_z.data(), v.group_indices.data(),
                            v.cell_fsr_offset.data(), v.cell_fsrs.data(), wa.data(), st.data(), wp.data()),
        "b200_cmfd_configure");
  if (v.centroid_update)
    check(b200_cmfd_set_stencils(_h, v.st_offset.data(), v.st_cell.data(), v.st_weight.data(), v.st_own.data(),
                                 v.st_size.data()), "b200_cmfd_set_stencils");
  if (c.axial_interpolation)
    check(b200_cmfd_set_axial_interpolants(_h, v.axial_interpolants.data()), "b200_cmfd_set_axial_interpolants");
  check(b200_cmfd_set_keff(_h, v.k_eff), "b200_cmfd_set_keff");
  _cmfd_device_keff = v.k_eff;
  _cmfd_device_ms = 0.; _cmfd_linear_iters = 0;
  _cmfd->setFluxUpdateOn(false);
  _cmfd_suspended = _cmfd;
  _cmfd_device_active = true;
}

template <class Base>
void B200SolverT<Base>::restoreCmfdFluxUpdate() {
  if (_cmfd_suspended != NULL) _cmfd_suspended->setFluxUpdateOn(true);
  _cmfd_suspended = NULL;
}

/* Device tallies -> Cmfd.  Faces go straight into the public current Vector
 * (Cmfd::getLocalCurrents, src/Cmfd.h:442, layout [cell][surface*ncg + g]); edge and corner
 * currents live in a private map (Cmfd.h:233) and are replayed through the public
 * Cmfd::tallyCurrent with a synthetic one-surface segment whose float "track flux" carries the
 * value as a hi + lo pair (so that the double is not rounded to float). */
template <class Base>
void B200SolverT<Base>::handCurrentsToCmfd() {
  const int ncg = _cmfd->getNumCmfdGroups();
  const long n_cells = _cmfd->getNumCells();
  check(b200_get_cmfd_currents(_h, _cmfd_currents.data(), (long)_cmfd_currents.size()), "b200_get_cmfd_currents");
  Vector* faces = _cmfd->getLocalCurrents();
  /* first MOC group of every CMFD group, for the replay */
  std::vector<int> first_moc(ncg, -1);
  for (int e = _num_groups - 1; e >= 0; e--) first_moc[_cmfd->getCmfdGroup(e)] = e;
  Quadrature* quad = _track_generator->getQuadrature();
  const bool solve3d = _SOLVE_3D;
  const int np = solve3d ? 1 : quad->getNumPolarAngles() / 2;
  std::vector<float> flux((size_t)np * _num_groups + _num_groups, 0.f);
  const double w0 = quad->getWeightInline(0, 0);
  const double w1 = (!solve3d && np > 1) ? quad->getWeightInline(0, 1) : w0;
  for (long cell = 0; cell < n_cells; cell++) {
    for (int surf = 0; surf < NUM_SURFACES; surf++) {
      const double* v = &_cmfd_currents[((size_t)cell * NUM_SURFACES + surf) * ncg];
      if (surf < NUM_FACES) {
        for (int g = 0; g < ncg; g++)
          if (v[g] != 0.) faces->incrementValue(cell, surf * ncg + g, v[g]);
        continue;
      }
      bool any = false;
      for (int g = 0; g < ncg; g++) any |= (v[g] != 0.);
      if (!any) continue;
      segment seg;
      seg._cmfd_surface_fwd = cell * NUM_SURFACES + surf;
      /* 2D: hi part in polar 0, lo part in polar 1 of the same call; 3D: two calls */
      std::fill(flux.begin(), flux.end(), 0.f);
      std::vector<float> lo(flux.size(), 0.f);
      bool need_lo = false;
      for (int g = 0; g < ncg; g++) {
        const int e = first_moc[g];
        const float hi = (float)(v[g] / w0);
        flux[e] = hi;
        const double rest = v[g] - (double)hi * w0;
        if (!solve3d && np > 1) flux[_num_groups + e] = (float)(rest / w1);
        else { lo[e] = (float)(rest / w0); need_lo |= (lo[e] != 0.f); }
      }
      _cmfd->tallyCurrent(&seg, flux.data(), 0, 0, true);
      if (need_lo) _cmfd->tallyCurrent(&seg, lo.data(), 0, 0, true);
    }
  }
}

/* Cmfd::computeKeff rescales the host fluxes (updateMOCFlux, src/Cmfd.cpp:1509-1560); the next
 * device step must see them. */
template <class Base>
void B200SolverT<Base>::pushHostFluxIfNewer() {
  if (!_host_flux_newer) return;
  check(b200_set_fluxes(_h, _scalar_flux, (long)_num_FSRs * _num_groups), "pushHostFluxIfNewer");
  pushExtraHostFlux();
  _host_flux_newer = false;
}

/* host mirrors only; device arrays are (re)zeroed, which is what a fresh
 * CPUSolver::initializeFluxArrays (CPUSolver.cpp:281-370) gives */
template <class Base>
void B200SolverT<Base>::initializeFluxArrays() {
  allocateHostFluxMirrors();
  pushMaterialsIfDirty();
  check(b200_zero_track_fluxes(_h), "b200_zero_track_fluxes");
  check(b200_flatten_fsr_fluxes(_h, 0.), "b200_flatten_fsr_fluxes");
  check(b200_store_fsr_fluxes(_h), "b200_store_fsr_fluxes");
}

template <class Base>
void B200SolverT<Base>::initializeSourceArrays() {
  allocateHostSourceMirrors();
  if (_fixed_sources_on && !_fixed_sources_initialized) initializeFixedSources();
}

template <class Base>
void B200SolverT<Base>::initializeFixedSources() {
  Base::initializeFixedSources();     /* cell / material maps -> FSR map */
  _fixed_sources_initialized = true;
  _fixed_dirty = true;
}

/* ------------------------- Solver pure virtuals -------------------------- */
template <class Base>
void B200SolverT<Base>::zeroTrackFluxes() { check(b200_zero_track_fluxes(_h), "zeroTrackFluxes"); }

template <class Base>
void B200SolverT<Base>::flattenFSRFluxes(FP_PRECISION value) {
  check(b200_flatten_fsr_fluxes(_h, value), "flattenFSRFluxes");
  _mirror_stale = true;
}

template <class Base>
void B200SolverT<Base>::flattenFSRFluxesChiSpectrum() {
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

template <class Base>
void B200SolverT<Base>::storeFSRFluxes() {
  pushHostFluxIfNewer();
  check(b200_store_fsr_fluxes(_h), "storeFSRFluxes");
}

template <class Base>
double B200SolverT<Base>::normalizeFluxes() {
  double norm = 0.;
  pushHostFluxIfNewer();
  check(b200_normalize_fluxes(_h, &norm), "normalizeFluxes");
  _mirror_stale = true;
  return norm;
}

template <class Base>
void B200SolverT<Base>::computeStabilizingFlux() { check(b200_compute_stabilizing_flux(_h), "computeStabilizingFlux"); }
template <class Base>
void B200SolverT<Base>::stabilizeFlux() { pushHostFluxIfNewer(); check(b200_stabilize_flux(_h), "stabilizeFlux"); _mirror_stale = true; }

template <class Base>
void B200SolverT<Base>::computeFSRSources(int iteration) {
  pushFixedSourcesIfDirty();
  pushHostFluxIfNewer();
  pushKeff();
  check(b200_compute_fsr_sources(_h, iteration), "computeFSRSources");
}
template <class Base>
void B200SolverT<Base>::computeFSRFissionSources() { check(b200_compute_fsr_fission_sources(_h), "computeFSRFissionSources"); }
template <class Base>
void B200SolverT<Base>::computeFSRScatterSources() { check(b200_compute_fsr_scatter_sources(_h), "computeFSRScatterSources"); }

template <class Base>
double B200SolverT<Base>::computeResidual(residualType res_type) {
  double residual = 0.;
  pushHostFluxIfNewer();
  pushKeff();
  check(b200_compute_residual(_h, (int)res_type, &residual), "computeResidual");
  return residual;
}

template <class Base>
void B200SolverT<Base>::computeKeff() {
  if (_cmfd_device_active) {
    /* Cmfd::computeKeff(_num_iterations) on the device; the threshold is the one the base-class loop keeps
     * setting on the Cmfd object (Solver.cpp:1159, 1674) */
    b200_cmfd_stats st;
    const double host_k = b200_cmfd_keff(_cmfd);          /* Cmfd::setKeff since the last solve (Solver.cpp:1249) */
    if (host_k != _cmfd_device_keff) check(b200_cmfd_set_keff(_h, host_k), "b200_cmfd_set_keff");
    _timer->startTimer();
    check(b200_cmfd_solve(_h, _num_iterations, b200_cmfd_source_threshold(_cmfd), &_k_eff, &st), "b200_cmfd_solve");
    _timer->stopTimer();
    _timer->recordSplit("Total CMFD time");           /* the split Cmfd::computeKeff records (Cmfd.cpp:1286) */
    _cmfd_device_ms += st.device_ms;
    _cmfd_linear_iters += st.linear_iters_total;
    if (st.failed)
      log_printf(WARNING, "The CMFD solve on the device did not converge in MOC iteration %d: k_eff and fluxes "
                 "are left as they are", _num_iterations);
    if (st.bad_tallies > 0)
      log_printf(WARNING_ONCE, "Negative or zero reaction tally calculated in %d CMFD cell-groups", st.bad_tallies);
    ConvergenceData* cd = b200_cmfd_convergence_data(_cmfd);
    if (cd != NULL) {
      cd->pf = st.pf; cd->cmfd_res_1 = st.cmfd_res_1; cd->cmfd_res_end = st.cmfd_res_end;
      cd->linear_res_1 = st.linear_res_1; cd->linear_res_end = st.linear_res_end;
      cd->cmfd_iters = st.cmfd_iters; cd->linear_iters_1 = st.linear_iters_1; cd->linear_iters_end = st.linear_iters_end;
    }
    _cmfd->setKeff(_k_eff);
    _cmfd_device_keff = _k_eff;
    _device_keff = _k_eff;
    _mirror_stale = true;
    return;
  }
  pushKeff();
  check(b200_compute_keff(_h, &_k_eff), "computeKeff");
  _device_keff = _k_eff;
}

template <class Base>
void B200SolverT<Base>::addSourceToScalarFlux() {
  check(b200_add_source_to_scalar_flux(_h), "addSourceToScalarFlux");
  _mirror_stale = true;
  if (_cmfd_active && !_cmfd_device_active) {
    /* the base-class loop calls _cmfd->computeKeff() right after this step (Solver.cpp:1628-1629) */
    syncHostMirrors();
    handCurrentsToCmfd();
    _host_flux_newer = true;
  }
}

/* The "Transport Sweep" timer split keeps meaning what the reference's report expects
 * (Solver.cpp:1901-1929): wall time of the sweep, the stream is drained before stopping. */
template <class Base>
void B200SolverT<Base>::transportSweep() {
  pushHostFluxIfNewer();
  if (_cmfd_active && !_cmfd_device_active) _cmfd->zeroCurrents();       /* CPUSolver.cpp:2343-2344 */
  if (_cmfd_active && !_cmfd_device_active && _cmfd->isSigmaTRebalanceOn()) tallyStartingCurrents();   /* CPUSolver.cpp:2356-2357 */
  _timer->startTimer();
  check(b200_transport_sweep(_h), "transportSweep");
  check(b200_synchronize(_h), "transportSweep");
  _timer->stopTimer();
  _timer->recordSplit("Transport Sweep");
  _mirror_stale = true;
}

/* CMFD sigma-t rebalance: the currents the starting angular fluxes carry into the boundary CMFD cells
 * (CPUSolver::tallyStartingCurrents, src/CPUSolver.cpp:498-537).  The start fluxes come back from the
 * device once per sweep (n_tracks * 2 * G floats); Cmfd::tallyStartingCurrent (src/Cmfd.cpp:5412) locates the
 * cell and tallies, exactly as for the reference's own solver. */
template <class Base>
void B200SolverT<Base>::tallyStartingCurrents() {
  TrackGenerator3D* tg3 = dynamic_cast<TrackGenerator3D*>(_track_generator);
  if (tg3 == NULL) log_printf(ERROR, "Starting currents not implemented yet for 2D MOC");
  const long nt = _flat.n_tracks;
  const int F = _num_groups;
  _start_flux_host.resize((size_t)nt * 2 * F);
  check(b200_get_start_fluxes(_h, _start_flux_host.data(), (long)_start_flux_host.size()), "b200_get_start_fluxes");
  Quadrature* quad = _track_generator->getQuadrature();
#pragma omp parallel for schedule(static)
  for (long t = 0; t < nt; t++) {
    TrackStackIndexes tsi;
    Track3D track;
    tg3->getTSIByIndex(t, &tsi);
    tg3->getTrackOTF(&track, &tsi);
    const double azim = track.getPhi(), polar = track.getTheta();
    const double dx = cos(azim) * sin(polar) * TINY_MOVE, dy = sin(azim) * sin(polar) * TINY_MOVE, dz = cos(polar) * TINY_MOVE;
    const double weight = quad->getWeightInline(track.getAzimIndex(), track.getPolarIndex());
    _cmfd->tallyStartingCurrent(track.getStart(), dx, dy, dz, &_start_flux_host[(size_t)(t * 2) * F], weight);
    _cmfd->tallyStartingCurrent(track.getEnd(), -dx, -dy, -dz, &_start_flux_host[(size_t)(t * 2 + 1) * F], weight);
  }
}

/* ------------------------------ public API ------------------------------- */
template <class Base>
void B200SolverT<Base>::syncHostMirrors() {
  if (_h == NULL || _scalar_flux == NULL) return;
  long n = _num_FSRs * _num_groups;
  check(b200_get_fluxes(_h, _scalar_flux, n), "syncHostMirrors");
  if (_reduced_sources != NULL) check(b200_get_fsr_sources(_h, _reduced_sources, n), "syncHostMirrors");
  syncExtraMirrors();
  _mirror_stale = false;
}

template <class Base>
void B200SolverT<Base>::getFluxes(FP_PRECISION* out_fluxes, int num_fluxes) {
  if (num_fluxes != _num_groups * _num_FSRs)
    log_printf(ERROR, "Unable to get FSR scalar fluxes since there are "
               "%d groups and %d FSRs which does not match the requested "
               "%d flux values", _num_groups, _num_FSRs, num_fluxes);
  if (_h == NULL)
    log_printf(ERROR, "Unable to get FSR scalar fluxes since they have not yet been allocated");
  pushHostFluxIfNewer();
  check(b200_get_fluxes(_h, out_fluxes, num_fluxes), "getFluxes");
}

/* CPUSolver aliases the caller's buffer (CPUSolver.cpp:190); like GPUSolver::setFluxes
 * (GPUSolver.cu:1088) the values are copied to the device instead. */
template <class Base>
void B200SolverT<Base>::setFluxes(FP_PRECISION* in_fluxes, int num_fluxes) {
  if (num_fluxes != _num_groups * _num_FSRs)
    log_printf(ERROR, "Unable to set an array with %d flux values for %d "
               " groups and %d FSRs", num_fluxes, _num_groups, _num_FSRs);
  if (_h == NULL)
    log_printf(ERROR, "Unable to set FSR scalar fluxes before the solver is initialized "
               "(call initializeSolver first)");
  check(b200_set_fluxes(_h, in_fluxes, num_fluxes), "setFluxes");
  _mirror_stale = true;
}

template <class Base>
double B200SolverT<Base>::getFlux(long fsr_id, int group) {
  if (_mirror_stale) syncHostMirrors();
  return Base::getFlux(fsr_id, group);
}

template <class Base>
double B200SolverT<Base>::getFSRSource(long fsr_id, int group) {
  syncHostMirrors();
  return Base::getFSRSource(fsr_id, group);
}

template <class Base>
void B200SolverT<Base>::setFixedSourceByFSR(long fsr_id, int group, double source) {
  Base::setFixedSourceByFSR(fsr_id, group, source);
  _fixed_dirty = true;
}

template <class Base>
void B200SolverT<Base>::resetFixedSources() {
  _fix_src_FSR_map.clear();
  _fix_src_cell_map.clear();
  _fix_src_material_map.clear();
  _fixed_dirty = true;
}

template <class Base>
void B200SolverT<Base>::computeFSRFissionRates(double* fission_rates, long num_FSRs, bool nu) {
  if (_h == NULL)
    log_printf(ERROR, "Unable to compute FSR fission rates since the "
               "source distribution has not been calculated");
  check(b200_compute_fsr_fission_rates(_h, fission_rates, num_FSRs, nu), "computeFSRFissionRates");
}

template <class Base>
void B200SolverT<Base>::computeEigenvalueFused(int max_iters, residualType res_type) {
  this->clearTimerSplits();
  _timer->startTimer();
  initializeMaterials(_solver_mode);
  this->initializeFSRs();
  this->countFissionableFSRs();
  initializeExpEvaluators();
  initializeFluxArrays();
  initializeSourceArrays();
  initializeCmfd();
  if (_cmfd_active && !_cmfd_device_active)
    log_printf(ERROR, "computeEigenvalueFused runs the whole source iteration on the device and cannot "
               "hand currents to the host Cmfd every iteration: use computeEigenvalue() with this Cmfd");
  pushFixedSourcesIfDirty();
  int iters = 0;
  check(b200_compute_eigenvalue(_h, max_iters, _converge_thresh, (int)res_type, &iters), "computeEigenvalueFused");
  _num_iterations = iters;
  restoreCmfdFluxUpdate();
  check(b200_get_keff(_h, &_k_eff), "computeEigenvalueFused");
  _device_keff = _k_eff;
  syncHostMirrors();
  _timer->stopTimer();
  _timer->recordSplit("Total time");
}

template <class Base>
void B200SolverT<Base>::getSweepStats(double* ms, long* sweeps) {
  int64_t n = 0, launches = 0;
  check(b200_get_sweep_stats(_h, ms, &n, &launches), "getSweepStats");
  if (sweeps) *sweeps = n;
}

#endif /* B200SOLVERT_H_ */
