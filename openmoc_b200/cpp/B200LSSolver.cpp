/**
 * @file B200LSSolver.cpp
 * @brief Linear-source specifics of the plug-in (see B200LSSolver.h).
 */
#include "B200LSSolver.h"

#include <cmath>

template class B200SolverT<CPULSSolver>;

/* host arrays: let the reference allocate everything it frees in its destructors */
void B200LSSolver::allocateHostFluxMirrors() { CPULSSolver::initializeFluxArrays(); }
void B200LSSolver::allocateHostSourceMirrors() { CPULSSolver::initializeSourceArrays(); }

/* segment starting points come with the flatten; track directions are those of
 * TransportSweep::onTrack (src/TrackTraversingAlgorithms.cpp:913-925) */
void B200LSSolver::uploadExtras() {
  std::vector<double> dir((size_t)_flat.n_tracks * 3);
  for (int64_t t = 0; t < _flat.n_tracks; t++) {
    double phi = _flat.trk_phi[t];
    double cos_theta = 0.0, sin_theta = 1.0;
    if (_flat.solve_3d) {
      double theta = _flat.trk_theta[t];
      cos_theta = cos(theta);
      sin_theta = sin(theta);
    }
    dir[3 * t] = cos(phi) * sin_theta;
    dir[3 * t + 1] = sin(phi) * sin_theta;
    dir[3 * t + 2] = cos_theta;
  }
  check(b200_upload_linear_source(_h, _flat.seg_start.data(), dir.data(), _FSR_lin_exp_matrix,
                                  _FSR_source_constants), "b200_upload_linear_source");
}

void B200LSSolver::syncExtraMirrors() {
  if (_scalar_flux_xyz != NULL)
    check(b200_get_flux_moments(_h, _scalar_flux_xyz, (long)_num_FSRs * _num_groups * 3), "syncHostMirrors");
}

/* Cmfd::updateMOCFlux rescales the moments too (src/Cmfd.cpp:1552-1559, setFluxMoments) */
void B200LSSolver::pushExtraHostFlux() {
  if (_scalar_flux_xyz != NULL)
    check(b200_set_flux_moments(_h, _scalar_flux_xyz, (long)_num_FSRs * _num_groups * 3), "pushHostFluxIfNewer");
}

void B200LSSolver::getFluxMoments(FP_PRECISION* out, long n) {
  check(b200_get_flux_moments(_h, out, n), "getFluxMoments");
}

/* fixed source moments (CPULSSolver::initializeFixedSources, src/CPULSSolver.cpp:154-205) follow the flat
 * fixed sources to the device */
void B200LSSolver::pushExtraFixedSources() {
  if (!_fixed_source_moments_on || _fixed_sources_xyz.empty()) return;
  for (long r = 0; r < _num_FSRs; r++)
    for (int g = 0; g < _num_groups; g++) {
      const std::vector<double>& v = _fixed_sources_xyz[r * _num_groups + g];
      if (v.size() == 3 && (v[0] != 0. || v[1] != 0. || v[2] != 0.))
        check(b200_set_fixed_source_moments_by_fsr(_h, r, g + 1, v[0], v[1], v[2]), "b200_set_fixed_source_moments_by_fsr");
    }
}
