/*
 * b200moc.h - C ABI of the B200-native MOC transport-sweep engine.
 *
 * This is the drop-in boundary: a `B200Solver : public Solver` on the OpenMOC
 * side (openmoc_b200/cpp/B200Solver.cpp) implements every pure virtual of
 * src/Solver.h:334-431 as a one-line call into this library, exactly like the
 * reference's GPUSolver (src/accel/cuda/GPUSolver.h:78-179) does with its own
 * kernels.  Only POD pointers and sizes cross the boundary; no OpenMOC, torch
 * or CUDA types.  Paths below are relative to the reference repository.
 *
 * Conventions (all "CPU convention"):
 *   scalar flux / reduced source  double  [r*G + e]          src/Solver.h:34-46
 *   track angular flux            float   [(t*2+dir)*F + p*G + e]  src/Solver.h:49-54
 *        dir 0 = forward, 1 = reverse;  F = G*P/2 (2D) or G (3D)   src/Solver.cpp:432-449
 *   sigma_s[dest*G + orig], fiss_matrix[dest*G + orig]       src/Material.cpp:728,977
 *   q = Q / 4pi (not divided by sigma_t)                     src/CPUSolver.cpp:1974
 *   fission source normalised to N_FSR                       src/CPUSolver.cpp:1910,2327
 *
 * Every function returns 0 on success, non-zero on failure with the message
 * available from b200_last_error() (the plug-in turns it into
 * log_printf(ERROR, ...) => std::logic_error => Python RuntimeError, the
 * reference's own convention, src/log.cpp:535-599).
 *
 * Requirements on the uploaded tracks: the hand-off table must be one-to-one (no two track
 * ends feed the same (track, direction) start slot - cyclic tracking guarantees it;
 * b200_finalize checks it, because the sweep performs all hand-offs concurrently).
 *
 * Tuning knobs read from the environment (defaults are what bench.py measures):
 *   B200_ORDER=natural|sorted Track uid order / longest track first; default: sorted for 2D decks,
 *                             natural for 3D decks (FSR locality in L2)
 *   B200_PHI_REPLICAS=R       copies of the FSR tally (power of two); default: enough for
 *                             n_fsrs*R >= 16 Ki rows, 1 for large decks
 *   B200_GRAPH=0|1            CUDA-graph replay of the fused iteration off / on; default: on
 *                             for decks with fewer than 1e8 integrations per sweep
 *   B200_GPL, B200_IPC, B200_CTA   lane map / CTA size of the sweep kernel (experiments)
 *   B200_CMFD_MODE=0|1|2      launch shape of the CMFD eigenvalue solve: one CTA / cooperative grid / one thread-block
 *                             cluster; default by mesh size (profiles/r02_cmfd.md).  B200_CMFD_CLUSTER (CTAs),
 *                             B200_CMFD_CLUSTER_THREADS, B200_CMFD_THREADS, B200_CMFD_BLOCKS refine it (experiments)
 *   (plug-in) B200_HOST_CMFD=1 keeps the reference's host Cmfd, B200_HOST_OTF=1 the host expansion of OTF tracks
 */
#ifndef B200MOC_H_
#define B200MOC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_solver b200_solver;

/* residualType, src/Solver.h:75-85 */
enum { B200_RES_SCALAR_FLUX = 0, B200_RES_FISSION_SOURCE = 1, B200_RES_TOTAL_SOURCE = 2 };
/* stabilizationType, src/Solver.h:103-113 */
enum { B200_STAB_DIAGONAL = 0, B200_STAB_YAMAMOTO = 1, B200_STAB_GLOBAL = 2 };
/* boundaryType, src/boundary_type.h:14-29 */
enum { B200_BC_VACUUM = 0, B200_BC_REFLECTIVE = 1, B200_BC_PERIODIC = 2, B200_BC_INTERFACE = 3 };
/* arithmetic of the per-segment attenuation */
enum { B200_PRECISION_DOUBLE = 0,   /* FP_PRECISION=double build of the reference: default */
       B200_PRECISION_MIXED = 1,    /* fp32 exponential + fp32 delta-psi, fp64 tally */
       B200_PRECISION_TABLE = 2 };  /* double arithmetic, F1 from a shared-memory quadratic interpolation table
                                     * (the role of ExpEvaluator's table, src/ExpEvaluator.cpp:190-330); held to the
                                     * north-star tolerance only: k_eff within 1 pcm, fluxes within 1e-4 */

typedef struct b200_config {
  int32_t num_groups;      /* G */
  int32_t num_azim;        /* A: azimuthal angles in (0, 2pi) */
  int32_t num_polar;       /* P: polar angles in (0, pi) */
  int32_t solve_3d;        /* 0: 2D tracks carry P/2 polar angles; 1: one per track */
  int64_t n_tracks, n_segments, n_fsrs;
  int32_t n_materials;
  int32_t device;          /* CUDA device ordinal */
  int32_t precision;       /* B200_PRECISION_* */
  int32_t deterministic;   /* 1: order-independent 64-bit fixed-point FSR tally: bitwise reproducible runs */
  int64_t n_fsrs_global;   /* normalisation count when FSRs are sharded; 0 => n_fsrs */
  int32_t linear_source;   /* 1: CPULSSolver physics (flux moments, linear source); needs b200_upload_linear_source */
  int32_t reserved;
} b200_config;

const char* b200_last_error(void);
int b200_version(void);
/* number of visible CUDA devices, or -1 with an error message */
int b200_device_count(void);

/* replaces GPUSolver::GPUSolver / ~GPUSolver (GPUSolver.cu:771-850) */
int b200_create(const b200_config* cfg, b200_solver** out);
int b200_destroy(b200_solver* s);

/* ---- several GPUs behind one handle ----
 * Call right after b200_create (cfg.device is then only the device the host talks to first).  The
 * uploads are parked on the host; b200_finalize shards the tracks by whole chains (connected components
 * of the boundary hand-off graph: no angular flux ever crosses a shard), balanced by segment count, and
 * builds one solver per entry of `devices` (a device may appear twice: two shards on one GPU) with
 * replicated FSR / material / quadrature data.  Every other entry point then acts on the group: FSR
 * steps run replicated and stay bit-identical, the transport sweep runs on every shard and the shards'
 * tallies (scalar flux, flux moments, CMFD currents, fixed-point tally) are summed by the library's own
 * two-shot all-reduce kernels over peer memory (NVLink P2P loads, CUDA events between the shards'
 * streams; no NCCL, no host copies).  Replaces the MPI reductions and interface exchange of
 * src/CPUSolver.cpp:545-1211, 1900, 2224, 2317 for tracks that are decomposed by chain.
 * Not available on a group: k_eff from the neutron balance, b200_set_stream, b200_get_segments. */
int b200_set_devices(b200_solver* s, int32_t n_devices, const int32_t* devices);
int b200_get_num_devices(b200_solver* s, int32_t* n_devices);

/* replaces GPUSolver::initializeTracks + clone_track (GPUSolver.cu:1312-1353,
 * clone.cu:86-126): one SoA upload instead of one cudaMalloc per track. */
int b200_upload_tracks(b200_solver* s,
                       const double* seg_length, const int32_t* seg_fsr,
                       const int64_t* trk_seg_offset,           /* n_tracks+1 */
                       const int32_t* trk_azim, const int32_t* trk_polar,
                       const int64_t* trk_next_fwd, const int64_t* trk_next_bwd,
                       const uint8_t* trk_flags,  /* bit0 next_fwd_is_fwd, bit1 next_bwd_is_fwd */
                       const uint8_t* trk_bc_fwd, const uint8_t* trk_bc_bwd);
/* replaces GPUSolver::copyQuadrature (GPUSolver.cu:1102-1143); [A/2][P] tables of
 * Quadrature::getWeightInline / getSinThetaInline (src/Quadrature.h:301-336) */
int b200_upload_quadrature(b200_solver* s, const double* weight, const double* sin_theta);
/* replaces GPUSolver::initializeFSRs (GPUSolver.cu:1165-1215) */
int b200_upload_fsrs(b200_solver* s, const double* volume, const int32_t* fsr_material);
/* replaces GPUSolver::initializeMaterials + clone_material (GPUSolver.cu:1224-1303) */
int b200_upload_materials(b200_solver* s, const double* sigma_t, const double* sigma_s,
                          const double* fiss_matrix, const double* nu_sigma_f,
                          const double* sigma_f, const double* chi,
                          const uint8_t* fissionable);
/* Linear source only (cfg.linear_source = 1), before b200_finalize.  Replaces the data the
 * reference keeps in struct segment::_starting_position (src/Track.h:50, relative to the FSR
 * centroid), the per-track direction of TransportSweep::onTrack
 * (src/TrackTraversingAlgorithms.cpp:913-925) and the two tables LinearExpansionGenerator
 * fills (src/CPULSSolver.h:36-40, src/TrackTraversingAlgorithms.cpp:470-831):
 *   seg_start[n_segments][3], trk_direction[n_tracks][3],
 *   lin_exp_matrix[n_fsrs][nc], source_constants[n_fsrs][nc][G]   (nc = 3 in 2D, 6 in 3D) */
int b200_upload_linear_source(b200_solver* s, const double* seg_start, const double* trk_direction,
                              const double* lin_exp_matrix, const double* source_constants);
/* The linear-source pre-pass itself on the device: LinearExpansionGenerator (src/TrackTraversingAlgorithms.cpp:
 * 470-831) from flattened tracks to the two tables b200_upload_linear_source takes.  Stand-alone (no solver
 * handle): host arrays in, host arrays out.  azim_spacing / azim_weight are [A/2], polar_spacing / polar_weight /
 * sin_theta [A/2][P] (Quadrature.cpp:674-750); n_flat_fsrs counts the FSRs whose moment matrix is singular and
 * which therefore keep a flat source.  The plug-in may keep using the reference's own host pre-pass instead. */
int b200_ls_prepass(int32_t device, int32_t num_groups, int32_t num_azim, int32_t num_polar, int32_t solve_3d,
                    int64_t n_tracks, int64_t n_segments, int64_t n_fsrs, int32_t n_materials,
                    const double* seg_length, const int32_t* seg_fsr, const double* seg_start,
                    const int64_t* trk_seg_offset, const int32_t* trk_azim, const int32_t* trk_polar,
                    const double* trk_phi, const double* trk_theta,
                    const double* azim_spacing, const double* azim_weight, const double* polar_spacing,
                    const double* polar_weight, const double* sin_theta,
                    const double* volume, const int32_t* fsr_material, const double* sigma_t,
                    double* lin_exp_matrix, double* source_constants, int32_t* n_flat_fsrs);
/* CMFD surface-current tally inside the sweep (Cmfd::tallyCurrent, src/Cmfd.h:572-670; the
 * reference calls it per segment from TransportSweep::onTrack).  seg_cmfd_fwd/bwd are
 * segment::_cmfd_surface_fwd/_bwd (src/Track.h:42-46: cell*26 + surface, or -1).  The tally
 * lands in a dense array [(cell*26 + surface)*ncg + g] the host hands to Cmfd. */
int b200_upload_cmfd_surfaces(b200_solver* s, const int32_t* seg_cmfd_fwd, const int32_t* seg_cmfd_bwd);
int b200_set_cmfd_groups(b200_solver* s, const int32_t* moc_to_cmfd_group, int32_t num_cmfd_groups,
                         int64_t num_cmfd_cells);   /* num_cmfd_groups <= 0 switches the tally off */
int b200_get_cmfd_currents(b200_solver* s, double* out, int64_t n);

/* CMFD on axially traced tracks (call between b200_upload_otf_geometry and b200_finalize): the device tracer then
 * also produces segment::_cmfd_surface_fwd/_bwd of every 3D segment, the way traceSegmentsOTF / traceStackOTF do
 * (src/TraverseSegments.cpp:429-457, 689-765; Cmfd::findCmfdSurfaceOTF -> Lattice::getLatticeSurfaceOTF,
 * src/Universe.cpp:2241-2326): seg2d_surface_fwd/bwd[n_segments_2d] = surface (0..9: x / y faces and z-parallel edges,
 * src/constants.h:120-129) the radial segment crosses at its forward / backward end, or -1;
 * fsr_cmfd_cell[n_fsrs] = Geometry::getCmfdCell; z_planes[num_z+1] = the z planes of the CMFD mesh. */
int b200_upload_otf_cmfd(b200_solver* s, const int8_t* seg2d_surface_fwd, const int8_t* seg2d_surface_bwd,
                         const int32_t* fsr_cmfd_cell, int32_t num_x, int32_t num_y, int32_t num_z,
                         const double* z_planes);

/* ---- CMFD acceleration on the device ----
 * Replaces what Cmfd::computeKeff (src/Cmfd.cpp:1192-1295) does between two sweeps: splitVertexCurrents /
 * splitEdgeCurrents (:2126-2331), collapseXS (:720-1007), constructMatrices (:1353-1500, with
 * getSurfaceDiffusionCoefficient :1048-1177 and computeLarsensEDCFactor :1625), eigenvalueSolve + linearSolve
 * (src/linalg.cpp:25-163, 179-395), rescaleFlux (:1304) and updateMOCFlux (:1509-1580, getUpdateRatio /
 * getFluxRatio :3178-3290) - on the scalar flux and the surface currents the sweep left in HBM.  The Cmfd object
 * of the host keeps its role as the description of the mesh: the plug-in reads the lattice, the group structure,
 * the FSR lists of the cells and the k-nearest stencils from it and hands them over once.
 * Not reproduced (the plug-in then keeps the host Cmfd): domain decomposition, the sigma-t rebalance, the
 * neutron-balance check, the few-group backup solver (a diverged solve keeps the last CMFD k_eff and skips the
 * flux update, like the reference when the backup fails too). */
typedef struct b200_cmfd_config {
  int32_t num_x, num_y, num_z;          /* Cmfd::setLatticeStructure; num_z = 1 in 2D (Cmfd.cpp:3860-3869) */
  int32_t num_cmfd_groups;              /* must equal b200_set_cmfd_groups' */
  int32_t boundaries[6];                /* B200_BC_* per face SURFACE_X_MIN .. SURFACE_Z_MAX (src/constants.h:120-125) */
  int32_t linear_source;                /* Cmfd::setFluxMoments was called: no Larsen factor, moments are scaled too */
  int32_t flux_limiting;                /* Cmfd::useFluxLimiting, default on */
  int32_t centroid_update;              /* Cmfd::setCentroidUpdateOn + setKNearest: needs b200_cmfd_set_stencils */
  int32_t axial_interpolation;          /* Cmfd::useAxialInterpolation 0 / 1 / 2: needs b200_cmfd_set_axial_interpolants */
  int32_t num_unbounded_iterations;     /* Cmfd::setNumUnboundedIterations */
  int32_t num_azim_2, num_polar_2;      /* quadrature of the Larsen factor */
  double sor_factor;                    /* Cmfd::setSORRelaxationFactor */
  double relaxation_factor;             /* Cmfd::setCMFDRelaxationFactor, default 0.7 */
  double linalg_tolerance;              /* MIN_LINALG_TOLERANCE = LINALG_TOL of the build (src/constants.h:77) */
} b200_cmfd_config;
/* widths_{x,y,z}: cell widths (Cmfd::_cell_widths_*); group_indices[ncg+1]: first MOC group of every CMFD group
 * (Cmfd::_group_indices); cell_fsr_offset[n_cells+1] / cell_fsrs: Cmfd::getCellFSRs in its own order;
 * azim_weight[num_azim_2], sin_theta / polar_weight[num_azim_2*num_polar_2]: Quadrature::getAzimWeight /
 * getSinTheta / getPolarWeight.  Call after b200_set_cmfd_groups. */
int b200_cmfd_configure(b200_solver* s, const b200_cmfd_config* cfg, const double* widths_x, const double* widths_y,
                        const double* widths_z, const int32_t* group_indices, const int64_t* cell_fsr_offset,
                        const int32_t* cell_fsrs, const double* azim_weight, const double* sin_theta,
                        const double* polar_weight);
/* k-nearest stencils (Cmfd::generateKNearestStencils, Cmfd.cpp:2887-2960) as getUpdateRatio reads them: for FSR r
 * the entries [offset[r], offset[r+1]) are the stencil cells other than the FSR's own (already resolved with
 * getCellByStencil) with their weights; own_weight[r] is the weight of the first entry of the stencil and
 * stencil_size[r] its length (Cmfd.cpp:3178-3210). */
int b200_cmfd_set_stencils(b200_solver* s, const int64_t* offset, const int32_t* cell, const double* weight,
                           const double* own_weight, const int32_t* stencil_size);
int b200_cmfd_set_axial_interpolants(b200_solver* s, const double* interpolants /* [n_fsrs*3] */);
int b200_cmfd_set_keff(b200_solver* s, double k_eff);       /* Cmfd::setKeff */
typedef struct b200_cmfd_stats {      /* struct ConvergenceData, src/linalg.h:31-68 */
  double pf, cmfd_res_1, cmfd_res_end, linear_res_1, linear_res_end;
  int32_t cmfd_iters, linear_iters_1, linear_iters_end, linear_iters_total, failed, bad_tallies;
  double device_ms;                   /* CUDA-event time of the seven kernels of this solve */
} b200_cmfd_stats;
/* One CMFD solve + prolongation (Cmfd::computeKeff(moc_iteration)); k_eff of the solver becomes the CMFD one
 * (Solver.cpp:1628).  source_threshold: Cmfd::setSourceConvergenceThreshold's value; < 0: the value the device
 * keeps (0.01 x the last residual, Solver.cpp:1671-1675).  k_eff / stats may be NULL (no host synchronisation). */
int b200_cmfd_solve(b200_solver* s, int32_t moc_iteration, double source_threshold, double* k_eff,
                    b200_cmfd_stats* stats);
/* The fused loops (b200_compute_eigenvalue, b200_iterate, b200_iteration_end) run the CMFD solve after the closure of
 * every iteration once b200_cmfd_configure has been called - the whole CMFD-accelerated source iteration then
 * runs without a host round trip per step; on = 0 takes it out again (the host then calls b200_cmfd_solve). */
int b200_cmfd_set_in_loop(b200_solver* s, int32_t on);
/* Cmfd::getVertexSplitSurfaces / getEdgeSplitSurfaces (Cmfd.cpp:2348-2480) as this library restates them: the
 * (cell*26 + surface) slots an edge or vertex current of `cell` is split onto.  Host-only; used by the tests. */
int b200_cmfd_split_targets(int32_t num_x, int32_t num_y, int32_t num_z, const int32_t* boundaries, int32_t cell,
                            int32_t surface, int32_t* targets /* [6] */, int32_t* num_targets);
/* flux moments in the reference layout [r*3G + c*G + e] (src/CPULSSolver.h:22) */
int b200_get_flux_moments(b200_solver* s, double* out, int64_t n);
int b200_set_flux_moments(b200_solver* s, const double* in, int64_t n);
/* ---- axial on-the-fly ray tracing on the device (3D solvers) ----
 * Replaces TraverseSegments::traceSegmentsOTF / traceStackOTF + SegmentationKernel
 * (src/TraverseSegments.cpp:304-505, 523-911; src/MOCKernel.cpp:216-268) and the per-sweep
 * TrackGenerator3D::getTrackOTF (src/TrackGenerator3D.cpp:1697): the host never materialises a 3D
 * segment.  It hands over the 2D segments of the radial tracks (length + extruded FSR id,
 * struct segment of the 2D Track), the axial meshes of the extruded FSRs (struct ExtrudedFSR,
 * src/Geometry.h:84-107: extruded FSR e owns 3D FSRs ext_fsr_ids[ext_offset[e]..ext_offset[e+1])
 * bottom-up and the planes ext_mesh[ext_offset[e]+e .. ext_offset[e+1]+e]), or - with
 * n_extruded_fsrs = 0 - one global mesh of n_axial_global layers in ext_mesh[0..n_axial_global]
 * (TrackGenerator3D::useGlobalZMesh; 3D FSR id = extruded id * n_axial_global + layer), and the
 * corrected polar angles theta[A/2][P] (Quadrature::getTheta).  Call with cfg.n_segments = 0. */
int b200_upload_otf_geometry(b200_solver* s, int64_t n_tracks_2d, int64_t n_segments_2d,
                             const double* seg2d_length, const int32_t* seg2d_extruded_fsr,
                             const int64_t* trk2d_seg_offset, int64_t n_extruded_fsrs,
                             const int64_t* ext_offset, const double* ext_mesh,
                             const int32_t* ext_fsr_ids, int32_t n_axial_global, const double* theta);
/* FSR volumes from the tracks (VolumeKernel, src/MOCKernel.cpp:80-162): every 3D track given here
 * (2D track id, distance of its start point from the start of the 2D track, start height, angle
 * indices) is traced and class_weight[azim*P+polar] * length tallied (class weight = azimuthal
 * spacing * azimuthal weight * polar spacing * polar weight).  A multi-GPU host passes ALL tracks
 * of the problem here and only its shard to b200_upload_tracks_otf.  Afterwards
 * b200_upload_fsrs accepts volume = NULL. */
int b200_otf_compute_volumes(b200_solver* s, int64_t n_tracks, const int32_t* trk_2d, const double* trk_l0,
                             const double* trk_z0, const int32_t* trk_azim, const int32_t* trk_polar,
                             const double* class_weight);
/* the solver's own 3D tracks (cfg.n_tracks of them): start data as above, links as in
 * b200_upload_tracks.  The device counts the 3D segments of every track, then writes its
 * segment stream in place; the total is returned and available from b200_get_num_segments. */
int b200_upload_tracks_otf(b200_solver* s, const int32_t* trk_2d, const double* trk_l0, const double* trk_z0,
                           const int32_t* trk_azim, const int32_t* trk_polar,
                           const int64_t* trk_next_fwd, const int64_t* trk_next_bwd,
                           const uint8_t* trk_flags, const uint8_t* trk_bc_fwd, const uint8_t* trk_bc_bwd,
                           int64_t* n_segments);
/* Solver::setMaxOpticalLength (src/Solver.cpp:550) / TrackGenerator::retrieveMaxOpticalLength: segments the
 * device traces are cut at this optical length exactly like the reference's on-the-fly kernels cut them
 * (src/MOCKernel.cpp:216-268, 353-410); default 100 (MAX_OPTICAL_LENGTH).  Takes effect at b200_finalize. */
int b200_set_max_optical_length(b200_solver* s, double max_tau);
int b200_get_num_segments(b200_solver* s, int64_t* n_segments);
/* number of 3D segments of arbitrary tracks over the uploaded geometry: the work estimate a host
 * needs to balance a partition of on-the-fly tracks */
int b200_otf_count_segments(b200_solver* s, int64_t n_tracks, const int32_t* trk_2d, const double* trk_l0,
                            const double* trk_z0, const int32_t* trk_azim, const int32_t* trk_polar,
                            int32_t* counts);
/* host copies of the device segment stream (tests, track dumps); trk_seg_offset may be NULL */
int b200_get_segments(b200_solver* s, double* seg_length, int32_t* seg_fsr, int64_t n_segments,
                      int64_t* trk_seg_offset);
int b200_get_volumes(b200_solver* s, double* volume, int64_t n_fsrs);
/* builds device-side derived tables; call after the four uploads (re-callable) */
int b200_finalize(b200_solver* s);

/* ---- one entry per Solver pure virtual (src/Solver.h:334-431) ---- */
int b200_zero_track_fluxes(b200_solver* s);
int b200_flatten_fsr_fluxes(b200_solver* s, double value);
int b200_flatten_fsr_fluxes_chi_spectrum(b200_solver* s, int32_t material);
int b200_store_fsr_fluxes(b200_solver* s);
int b200_normalize_fluxes(b200_solver* s, double* norm_factor);
int b200_compute_stabilizing_flux(b200_solver* s);
int b200_stabilize_flux(b200_solver* s);
int b200_compute_fsr_sources(b200_solver* s, int32_t iteration);
int b200_compute_fsr_fission_sources(b200_solver* s);
int b200_compute_fsr_scatter_sources(b200_solver* s);
int b200_compute_residual(b200_solver* s, int32_t res_type, double* residual);
int b200_compute_keff(b200_solver* s, double* k_eff);
int b200_add_source_to_scalar_flux(b200_solver* s);
/* zero phi, sweep every track both ways, hand boundary fluxes over.  On return
 * scalar_flux holds the raw tally sum(w*delta_psi); a multi-GPU host sums it
 * across ranks (b200_device_pointer) before b200_add_source_to_scalar_flux. */
int b200_transport_sweep(b200_solver* s);

/* ---- public Solver API (src/Solver.h:440-585) ---- */
int b200_get_fluxes(b200_solver* s, double* out_fluxes, int64_t num_fluxes);
int b200_set_fluxes(b200_solver* s, const double* in_fluxes, int64_t num_fluxes);
/* getFluxes and getKeff behind one host synchronisation */
int b200_get_fluxes_keff(b200_solver* s, double* out_fluxes, int64_t num_fluxes, double* k_eff);
int b200_set_fixed_source_by_fsr(b200_solver* s, int64_t fsr_id, int32_t group /*1-based*/, double source);
/* CPULSSolver::setFixedSourceMomentByFSR (src/CPULSSolver.cpp:154-205): volume-averaged x, y, z moments
 * of the fixed source in one FSR and group (1-based), linear-source solvers only */
int b200_set_fixed_source_moments_by_fsr(b200_solver* s, int64_t fsr_id, int32_t group, double src_x,
                                         double src_y, double src_z);
int b200_reset_fixed_sources(b200_solver* s);
int b200_compute_fsr_fission_rates(b200_solver* s, double* fission_rates, int64_t num_fsrs, int32_t nu);
int b200_stabilize_transport(b200_solver* s, double factor, int32_t stabilization_type);
int b200_allow_negative_fluxes(b200_solver* s, int32_t allowed);
/* Solver::setKeffFromNeutronBalance (src/Solver.cpp:2047): k = fission / (absorption + leakage),
 * with the vacuum leakage tallied in the sweep (src/CPUSolver.cpp:2264-2325, 2592-2600) */
int b200_set_keff_from_neutron_balance(b200_solver* s, int32_t on);
int b200_get_keff(b200_solver* s, double* k_eff);
int b200_set_keff(b200_solver* s, double k_eff);
int b200_get_fsr_sources(b200_solver* s, double* out, int64_t n);       /* reduced sources q */
int b200_set_fsr_sources(b200_solver* s, const double* in, int64_t n);
int b200_get_start_fluxes(b200_solver* s, float* out, int64_t n);       /* n = n_tracks*2*F */
int b200_set_start_fluxes(b200_solver* s, const float* in, int64_t n);

/* ---- fused drivers: the loops of Solver::computeEigenvalue / computeFlux /
 *      computeSource (src/Solver.cpp:1542-1689, 1352-1420, 1459-1516) run
 *      device-side, convergence flag evaluated on the GPU ---- */
int b200_compute_eigenvalue(b200_solver* s, int32_t max_iters, double tolerance,
                            int32_t res_type, int32_t* num_iterations);
int b200_compute_flux(b200_solver* s, int32_t max_iters, double tolerance,
                      int32_t only_fixed_source, int32_t* num_iterations);
int b200_compute_source(b200_solver* s, int32_t max_iters, double k_eff, double tolerance,
                        int32_t res_type, int32_t* num_iterations);
/* the same loop in pieces, for hosts that reduce the FSR tally across GPUs inside every
 * iteration: init; then per iteration begin (sources + sweep) -> [all-reduce scalar_flux] ->
 * end (closure, k_eff, normalisation, residual, store, device-side stopping rule); status
 * fetches the convergence flag (poll it every few iterations; once done, later begin/end
 * calls are no-ops on the device). */
int b200_eigen_loop_init(b200_solver* s, int32_t max_iters, double tolerance);
int b200_iteration_begin(b200_solver* s, int32_t iteration);
int b200_iteration_end(b200_solver* s, int32_t iteration, int32_t res_type, int32_t check_convergence);
int b200_eigen_loop_status(b200_solver* s, int32_t enqueued, int32_t* done, int32_t* iterations,
                           double* k_eff, double* residual);
/* run exactly n source iterations (sources..store) without convergence test;
 * benchmark hook.  residual/k of the last iteration are returned if non-NULL */
int b200_iterate(b200_solver* s, int32_t n, int32_t res_type, double* k_eff, double* residual);

/* the device's (1-exp(-x))/x, i.e. expF1_fractional (src/exponentials.h:156-192), evaluated
 * for n host values - the hook for the reference's known-answer vectors
 * (tests/unit_tests/test_exponentials.py:74-77) */
int b200_eval_expF1(int32_t device, int32_t precision, const double* x, int64_t n, double* out);

/* the two machine ceilings that bound the sweep besides HBM, measured on the spot: sustained FP64
 * instruction rate (DFMA, the sweep's instruction mix and occupancy) and RED.ADD.F64 rate into an
 * L2-resident table of table_rows x 7 doubles (the tally pattern of the sweep).  bench.py reports the
 * roofline fraction of the binding one next to the HBM fraction. */
int b200_measure_ceilings(int32_t device, int64_t table_rows, double* fp64_instr_per_s, double* red_f64_per_s);

/* ---- instrumentation (the "Transport Sweep" timer split, src/CPUSolver.cpp:2365-2378) ---- */
int b200_get_sweep_stats(b200_solver* s, double* sweep_ms_total, int64_t* num_sweeps,
                         int64_t* kernel_launches);
int b200_reset_sweep_stats(b200_solver* s);
int b200_synchronize(b200_solver* s);

/* ---- plumbing for multi-GPU hosts: raw device pointers of solver arrays so a
 *      host framework (torch.distributed / NCCL) can reduce them in place.
 *      name in {"scalar_flux","old_scalar_flux","reduced_sources","start_flux"} */
int b200_device_pointer(b200_solver* s, const char* name, void** ptr, int64_t* num_elements);
/* deterministic mode across GPUs: defer=1 makes b200_transport_sweep leave the tally in
 * the int64 fixed-point buffer ("scalar_flux_fixed") so the host can sum it exactly across
 * ranks (integer all-reduce); b200_finish_fixed_tally then converts it into scalar_flux. */
int b200_defer_fixed_tally(b200_solver* s, int32_t defer);
int b200_finish_fixed_tally(b200_solver* s);
/* bracket a stream capture made by the host (CUDA graph of begin -> all-reduce -> end): while set,
 * no timing events are recorded and nothing synchronises with the host */
int b200_set_capturing(b200_solver* s, int32_t capturing);
/* use an externally owned cudaStream_t (passed as void*) for all launches */
int b200_set_stream(b200_solver* s, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MOC_H_ */
