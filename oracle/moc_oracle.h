/*
 * moc_oracle.h - CPU restatement (plain C) of the reference's MOC source
 * iteration on flattened SoA tracks.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library; the product path
 * (openmoc_b200/, libb200moc.so) never does.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 * against the reference's own golden files through track fixtures dumped from
 * the unmodified reference (oracle/_ref/ref_driver):
 *   tests/test_forward_pin_cell/results_true.dat           (261 it, k, 14 fluxes)
 *   tests/test_forward_simple_lattice/results_true.dat     (SHA-512, 3584 fluxes)
 *   tests/test_forward_3D_lattice_70g/results_true.dat     (258 it, k)
 *   tests/test_forward_3D_lattice/results_true.dat, test_forward_hom_inf_medium
 *   tests/test_forward_3D_lattice_linear/results_true.dat      (linear source: 156 it, k, 480 FSRs)
 *   tests/test_forward_3D_lattice_linear_70g/results_true.dat  (linear source: 186 it, k)
 *   tests/test_compute_flux, tests/test_compute_source         (fixed-source drivers: 2 / 130 it)
 *   tests/test_adjoint_{pin_cell,simple_lattice,hom_inf_medium} (transposed production matrices)
 *   tests/unit_tests/test_exponentials.py:74-77            (expF1 known answers)
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef MOC_ORACLE_H_
#define MOC_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct moc_oracle moc_oracle;

/* residualType, src/Solver.h:75-85 */
enum { MOC_RES_SCALAR_FLUX = 0, MOC_RES_FISSION_SOURCE = 1, MOC_RES_TOTAL_SOURCE = 2 };

/* Arrays are copied.  Layouts as in include/b200moc.h / the B2TRK track file. */
moc_oracle* moc_oracle_create(
    int num_groups, int num_azim, int num_polar, int solve_3d,
    int64_t n_tracks, int64_t n_segments, int64_t n_fsrs, int n_materials,
    const double* seg_length, const int32_t* seg_fsr,
    const int64_t* trk_seg_offset, const int32_t* trk_azim, const int32_t* trk_polar,
    const int64_t* trk_next_fwd, const int64_t* trk_next_bwd,
    const uint8_t* trk_flags, const uint8_t* trk_bc_fwd, const uint8_t* trk_bc_bwd,
    const double* quad_weight, const double* quad_sin_theta,
    const double* fsr_volume, const int32_t* fsr_mat,
    const double* mat_sigma_t, const double* mat_sigma_s, const double* mat_fiss_matrix,
    const double* mat_nu_sigma_f, const double* mat_sigma_f, const double* mat_chi,
    const uint8_t* mat_fissionable);
void moc_oracle_destroy(moc_oracle* o);

/* one-to-one with the Solver virtuals (src/Solver.h:334-431) */
void   moc_oracle_zero_track_fluxes(moc_oracle* o);
void   moc_oracle_flatten_fsr_fluxes(moc_oracle* o, double value);
void   moc_oracle_store_fsr_fluxes(moc_oracle* o);
double moc_oracle_normalize_fluxes(moc_oracle* o);
void   moc_oracle_compute_fsr_sources(moc_oracle* o, int iteration);
void   moc_oracle_compute_fsr_fission_sources(moc_oracle* o);
void   moc_oracle_compute_fsr_scatter_sources(moc_oracle* o);
void   moc_oracle_transport_sweep(moc_oracle* o);
void   moc_oracle_add_source_to_scalar_flux(moc_oracle* o);
void   moc_oracle_compute_keff(moc_oracle* o);
double moc_oracle_compute_residual(moc_oracle* o, int res_type);
void   moc_oracle_compute_stabilizing_flux(moc_oracle* o);
void   moc_oracle_stabilize_flux(moc_oracle* o);

/* drivers (src/Solver.cpp:1542-1689, 1352-1420, 1459-1516); return #iterations */
int moc_oracle_compute_eigenvalue(moc_oracle* o, int max_iters, double tol, int res_type);
int moc_oracle_compute_flux(moc_oracle* o, int max_iters, double tol, int only_fixed_source);
int moc_oracle_compute_source(moc_oracle* o, int max_iters, double k_eff, double tol, int res_type);

/* state access */
double moc_oracle_get_keff(moc_oracle* o);
void   moc_oracle_set_keff(moc_oracle* o, double k);
void   moc_oracle_get_fluxes(moc_oracle* o, double* out);        /* [n_fsrs*G] */
void   moc_oracle_set_fluxes(moc_oracle* o, const double* in);
void   moc_oracle_get_sources(moc_oracle* o, double* out);       /* reduced sources q */
void   moc_oracle_set_sources(moc_oracle* o, const double* in);
void   moc_oracle_get_start_fluxes(moc_oracle* o, float* out);   /* [n_tracks*2*F] */
void   moc_oracle_set_start_fluxes(moc_oracle* o, const float* in);
void   moc_oracle_set_fixed_source(moc_oracle* o, int64_t fsr, int group0, double value);
void   moc_oracle_set_fixed_source_moments(moc_oracle* o, int64_t fsr, int group0, double sx, double sy, double sz);
void   moc_oracle_allow_negative_fluxes(moc_oracle* o, int allowed);
void   moc_oracle_stabilize_transport(moc_oracle* o, double factor, int type);
void   moc_oracle_compute_fission_rates(moc_oracle* o, double* out, int nu);
void   moc_oracle_set_num_threads(moc_oracle* o, int n);
/* Solver::setKeffFromNeutronBalance (src/Solver.cpp:2047): k = fission/(absorption+leakage) */
void   moc_oracle_set_keff_from_neutron_balance(moc_oracle* o, int on);
/* seconds spent inside transport_sweep since creation / last reset */
double moc_oracle_sweep_seconds(moc_oracle* o, int reset);

/* Linear source (CPULSSolver).  Call once after create: runs the LinearExpansionGenerator
 * pre-pass (src/TrackTraversingAlgorithms.cpp:470-831); afterwards every step function above
 * follows src/CPULSSolver.cpp.  seg_start [n_seg*3] is relative to the FSR centroids, the
 * quadrature factors are the chunks quad_azim_spacing/_weight [A/2], quad_polar_spacing/_weight
 * [A/2*P] of the track file.  Returns the number of FSRs that fall back to a flat source. */
int    moc_oracle_enable_linear_source(moc_oracle* o, const double* seg_start, const double* trk_phi,
                                       const double* trk_theta, const double* azim_spacing,
                                       const double* azim_weight, const double* polar_spacing,
                                       const double* polar_weight);
void   moc_oracle_get_flux_moments(moc_oracle* o, double* out);  /* [n_fsrs][3][G] */
void   moc_oracle_get_linear_source_tables(moc_oracle* o, double* lin_exp, double* src_const);

/* scalar exponential, src/exponentials.h:156-192 */
double moc_oracle_expF1(double x);

#ifdef __cplusplus
}
#endif
#endif
