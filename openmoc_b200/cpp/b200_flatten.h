/**
 * @file b200_flatten.h
 * @brief Flattens an OpenMOC TrackGenerator(3D) into the SoA arrays that the
 *        B200 sweep kernels consume (the arguments of b200_upload_* in
 *        include/b200moc.h) and reads/writes them as a "B2TRK" track file.
 *
 * This file is PLUG-IN code: it is compiled against the reference's headers
 * (src/TraverseSegments.h, src/TrackGenerator3D.h ...) and lives on the
 * OpenMOC side of the C-ABI.  It replaces the per-track clone_track() /
 * cudaMalloc loop of the reference GPUSolver (src/accel/cuda/clone.cu:86-126)
 * by a single pass over TraverseSegments::loopOverTracks (the same traversal
 * TransportSweep uses, src/TrackTraversingAlgorithms.cpp:866-879), so all four
 * segmentation modes (EXPLICIT_2D/3D, OTF_TRACKS, OTF_STACKS) flatten the same
 * way.
 */
#ifndef B200_FLATTEN_H_
#define B200_FLATTEN_H_

#include <cstdint>
#include <string>
#include <vector>

class TrackGenerator;
class Geometry;

/** Host-side SoA image of everything the sweep path reads. */
struct B200FlatTracks {
  /* problem shape */
  int num_groups = 0;
  int num_azim = 0;        /* full number of azimuthal angles (A) */
  int num_polar = 0;       /* full number of polar angles (P) */
  int solve_3d = 0;
  int fluxes_per_track = 0; /* F: G*P/2 in 2D, G in 3D (Solver.cpp:432-449) */
  int64_t n_tracks = 0, n_segments = 0, n_fsrs = 0;
  int n_materials = 0;

  /* segment stream, contiguous per track, forward order */
  std::vector<double> seg_length;
  std::vector<int32_t> seg_fsr;
  std::vector<int32_t> seg_mat;
  std::vector<int32_t> seg_cmfd_fwd, seg_cmfd_bwd;
  std::vector<double> seg_start;  /* xyz relative to FSR centroid (LS) */

  /* per track, indexed by Track uid */
  std::vector<int64_t> trk_seg_offset;  /* n_tracks + 1 */
  std::vector<int32_t> trk_azim, trk_polar, trk_xy;
  std::vector<int64_t> trk_next_fwd, trk_next_bwd;
  std::vector<uint8_t> trk_flags;   /* bit0 next_fwd_is_fwd, bit1 next_bwd_is_fwd */
  std::vector<uint8_t> trk_bc_fwd, trk_bc_bwd;  /* boundaryType enum values */
  std::vector<double> trk_phi, trk_theta;
  std::vector<double> trk_start;   /* start point of every track, xy (2D) or xyz (3D): what a spatial domain
                                    * decomposition cuts the tracks with (openmoc_b200/domain.py) */

  /* quadrature: [A/2][P] total weights, [A/2][P] sin(theta) */
  std::vector<double> quad_weight, quad_sin_theta;
  /* the factors the total weight is made of (Quadrature.cpp:674-750): [A/2] azimuthal spacing and
   * weight, [A/2][P] polar spacing (3D only, else 0) and weight - the linear-source pre-pass
   * (LinearExpansionGenerator, TrackTraversingAlgorithms.cpp:670-692) weights tracks with them */
  std::vector<double> quad_azim_spacing, quad_azim_weight, quad_polar_spacing, quad_polar_weight;

  /* FSR data */
  std::vector<double> fsr_volume;
  std::vector<int32_t> fsr_mat;
  std::vector<double> fsr_centroid; /* xyz */

  /* material tables [n_materials][...] in the reference's storage order */
  std::vector<double> mat_sigma_t, mat_sigma_a, mat_sigma_f, mat_nu_sigma_f, mat_chi;
  std::vector<double> mat_sigma_s;      /* [dest*G+orig]  (Material.cpp:728-731) */
  std::vector<double> mat_fiss_matrix;  /* [G_dest*G+g_orig] (Material.cpp:975-978) */
  std::vector<uint8_t> mat_fissionable;

  /* Axial on-the-fly tracks for the device tracer (b200_upload_otf_geometry / b200_upload_tracks_otf):
   * filled by b200_flatten(..., device_otf = true) INSTEAD of the 3D segment stream.  2D segments of
   * the radial tracks, the axial mesh and 3D FSR ids of every ExtrudedFSR (src/Geometry.h:84-107), the
   * corrected polar angles, and per 3D track its 2D track, the distance of its start point from the
   * start of that 2D track and its start height (what TraverseSegments::traceSegmentsOTF starts from,
   * src/TraverseSegments.cpp:304-340). */
  bool device_otf = false;
  int64_t n_tracks_2d = 0, n_extruded = 0;
  std::vector<double> seg2d_length, ext_mesh, otf_theta, trk_l0, trk_z0;
  std::vector<int32_t> seg2d_ext, ext_fsr, trk_2d;
  std::vector<int64_t> trk2d_seg_offset, ext_offset;
  /* CMFD with the device tracer (b200_upload_otf_cmfd): the surface the 2D segments cross at their ends
   * (segment::_cmfd_surface_fwd/_bwd % NUM_SURFACES of the radial tracks, or -1), Geometry::getCmfdCell of every
   * 3D FSR and the z planes of the CMFD lattice; filled when the Geometry has a Cmfd */
  bool otf_cmfd = false;
  std::vector<int8_t> seg2d_surf_fwd, seg2d_surf_bwd;
  std::vector<int32_t> fsr_cmfd_cell;
  std::vector<double> cmfd_z_planes;
  int cmfd_nx = 0, cmfd_ny = 0, cmfd_nz = 0;
};

/**
 * Flatten the tracks of a TrackGenerator whose segments are final, i.e. after
 * Solver::initializeFSRs() and Solver::initializeExpEvaluators() have run
 * (centroid re-centring and tau>max splitting mutate segments).
 */
void b200_flatten(TrackGenerator* track_generator, B200FlatTracks* out,
                  bool with_ls_data = true, bool device_otf = false);

/** True when the tracks can go to the device tracer: an on-the-fly 3D formation (OTF_TRACKS or
 *  OTF_STACKS).  The caller still falls back to the host expansion when it needs per-segment data
 *  the tracer does not produce (linear-source starting points). */
bool b200_can_trace_on_device(TrackGenerator* track_generator);

/** Write / read the chunked binary track file (see openmoc_b200/trackfile.py). */
struct B200CmfdView;
/** cmfd != NULL (b200_read_cmfd of the Geometry's initialised Cmfd): the file also carries the CMFD mesh - chunks
 *  cmfd_dims (num_x, num_y, num_z, num_cmfd_groups), cmfd_widths_x/y/z, cmfd_boundaries, cmfd_group_indices,
 *  cmfd_options (SOR factor, relaxation factor, flux limiting) and fsr_cmfd_cell - next to the surfaces the
 *  segments cross (seg_cmfd_fwd / seg_cmfd_bwd): openmoc_b200.solver.CmfdMesh.from_tracks rebuilds the mesh. */
void b200_write_trackfile(const B200FlatTracks& ft, const std::string& path, const B200CmfdView* cmfd = NULL);

#endif /* B200_FLATTEN_H_ */
