/**
 * @file b200_cmfd_view.h
 * @brief What the device CMFD (include/b200moc.h: b200_cmfd_*) needs to know about a reference
 *        Cmfd object.  The Cmfd of the host stays the description of the mesh and of the user's
 *        options; most of those have setters but no getters (src/Cmfd.h:430-500), so
 *        b200_cmfd_view.cpp reads the private members through explicit-instantiation accessors
 *        (no change to the reference sources).
 */
#ifndef B200_CMFD_VIEW_H_
#define B200_CMFD_VIEW_H_

#include <cstdint>
#include <vector>

class Cmfd;
struct ConvergenceData;

struct B200CmfdView {
  int num_x, num_y, num_z, num_cmfd_groups;
  int boundaries[6];
  bool linear_source, flux_limiting, centroid_update, check_neutron_balance, balance_sigma_t;
  int use_axial_interpolation, num_unbounded_iterations, k_nearest;
  double sor_factor, relaxation_factor, k_eff;
  std::vector<double> widths_x, widths_y, widths_z;
  std::vector<int32_t> group_indices;                 /* num_cmfd_groups + 1 */
  std::vector<int64_t> cell_fsr_offset;               /* cells + 1 */
  std::vector<int32_t> cell_fsrs;
  /* k-nearest stencils in the form b200_cmfd_set_stencils takes */
  std::vector<int64_t> st_offset;
  std::vector<int32_t> st_cell, st_size;
  std::vector<double> st_weight, st_own;
  std::vector<double> axial_interpolants;             /* FSRs x 3, empty when unused */
};

/** Reads everything above; call after Cmfd::initialize (Solver::initializeCmfd). */
void b200_read_cmfd(Cmfd* cmfd, long num_fsrs, B200CmfdView* out);
/** Cmfd::_source_convergence_threshold (set by the base-class loop every iteration, Solver.cpp:1674). */
double b200_cmfd_source_threshold(Cmfd* cmfd);
/** Cmfd::_k_eff: the starting guess of the next diffusion solve (Cmfd::setKeff, Solver.cpp:1249). */
double b200_cmfd_keff(Cmfd* cmfd);
/** Cmfd::_convergence_data (the iteration report of a verbose solve), or NULL. */
ConvergenceData* b200_cmfd_convergence_data(Cmfd* cmfd);

#endif
