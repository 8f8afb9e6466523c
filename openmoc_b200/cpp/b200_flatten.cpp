/**
 * @file b200_flatten.cpp
 * @brief TrackGenerator -> SoA flattening for the B200 sweep (plug-in side).
 *
 * Compiled against the reference's headers; see b200_flatten.h.
 */
#include "b200_flatten.h"
#include "b200_cmfd_view.h"
#include "Cmfd.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>

#include "TraverseSegments.h"
#include "TrackGenerator3D.h"
#include "MOCKernel.h"
#include "Quadrature.h"
#include "Geometry.h"
#include "Material.h"

namespace {

/** One visit of every track through the reference's own traversal, copying
 *  segments and links out.  Runs on the calling thread (the orphaned
 *  `omp for` in loopOverTracks binds to a team of one). */
class FlattenPass : public TraverseSegments {
 public:
  struct Seg { double len; int32_t fsr, mat, cf, cb; double x, y, z; };

  /* Explicit formations keep their segments in the Track objects, so their number is known before the
   * traversal: with `offsets` given (trk_seg_offset, n_tracks + 1) the segments are written straight into the
   * flat arrays of `out` and no per-track copy (56 bytes per segment) is ever made.  On-the-fly formations
   * are traced as they go: their segments are collected per track first. */
  FlattenPass(TrackGenerator* tg, std::map<Material*, int>* mat_index,
              B200FlatTracks* out, const std::vector<int64_t>* offsets = NULL, bool with_ls_data = true)
      : TraverseSegments(tg), _mat_index(mat_index), _out(out), _offsets(offsets), _with_ls(with_ls_data) {
    if (_offsets == NULL) _per_track.resize(out->n_tracks);
    _seen.assign(out->n_tracks, 0);
  }
  bool explicitFormation() { return _segment_formation == EXPLICIT_2D || _segment_formation == EXPLICIT_3D; }

  void execute() {
    if (_segment_formation != EXPLICIT_2D && _segment_formation != EXPLICIT_3D) {
      /* on-the-fly formations re-trace into the TrackGenerator's per-thread buffers: keep
       * the calling thread only (a team of one, as the orphaned `omp for` then binds) */
      MOCKernel* kernel = getKernel<SegmentationKernel>();
      loopOverTracks(kernel);
    } else {
      /* explicit segments are only read: give the reference's orphaned `omp for`
       * (TraverseSegments.cpp:75) a team; every track writes its own slots */
#pragma omp parallel
      loopOverTracks(NULL);
    }
  }

  void onTrack(Track* track, segment* segments) {
    long uid = track->getUid();
    int azim = track->getAzimIndex();
    int xy = track->getXYIndex();
    int polar = 0;
    Track3D* t3 = dynamic_cast<Track3D*>(track);
    if (t3 != NULL) polar = t3->getPolarIndex();

    /* The z-stack case hands over several tracks at once
     * (TrackTraversingAlgorithms.cpp:928-934). */
    Track* single[1] = {track};
    Track** tracks_array = single;
    int n_in_stack = 1;
    if (_segment_formation == OTF_STACKS) {
      int*** tps = _track_generator_3D->getTracksPerStack();
      n_in_stack = tps[azim][xy][polar];
      tracks_array = _track_generator_3D->getTemporaryTracksArray(0);
    }

    for (int i = 0; i < n_in_stack; i++) {
      Track* t = tracks_array[i];
      long id = uid + i;
      if (id < 0 || id >= _out->n_tracks)
        log_printf(ERROR, "b200_flatten: track uid %ld out of range", id);
      _seen[id] = 1;
      _out->trk_azim[id] = t->getAzimIndex();
      _out->trk_xy[id] = t->getXYIndex();
      Track3D* tt3 = dynamic_cast<Track3D*>(t);
      _out->trk_polar[id] = tt3 ? tt3->getPolarIndex() : 0;
      _out->trk_phi[id] = t->getPhi();
      _out->trk_theta[id] = tt3 ? tt3->getTheta() : M_PI_2;
      {
        const size_t dim = _out->solve_3d ? 3 : 2;
        Point* p0 = t->getStart();
        _out->trk_start[dim * id] = p0->getX();
        _out->trk_start[dim * id + 1] = p0->getY();
        if (dim == 3) _out->trk_start[dim * id + 2] = p0->getZ();
      }
      _out->trk_next_fwd[id] = t->getTrackNextFwd();
      _out->trk_next_bwd[id] = t->getTrackNextBwd();
      _out->trk_flags[id] = (t->getNextFwdFwd() ? 1 : 0) | (t->getNextBwdFwd() ? 2 : 0);
      _out->trk_bc_fwd[id] = (uint8_t)t->getBCFwd();
      _out->trk_bc_bwd[id] = (uint8_t)t->getBCBwd();
    }

    int n = track->getNumSegments();
    Material* last_mat = NULL;
    int last_idx = -1;
    if (_offsets != NULL) {
      /* explicit tracks: one track per call, straight into the flat arrays */
      size_t o = (size_t)(*_offsets)[uid];
      if ((int64_t)n != (*_offsets)[uid + 1] - (*_offsets)[uid])
        log_printf(ERROR, "b200_flatten: track %ld changed its segment count during the traversal", uid);
      for (int s = 0; s < n; s++, o++) {
        const segment& sg = segments[s];
        if (sg._material != last_mat) {
          std::map<Material*, int>::iterator it = _mat_index->find(sg._material);
          last_mat = sg._material;
          last_idx = (it == _mat_index->end()) ? -1 : it->second;
        }
        _out->seg_length[o] = sg._length;
        _out->seg_fsr[o] = sg._region_id;
        _out->seg_mat[o] = last_idx;
        _out->seg_cmfd_fwd[o] = sg._cmfd_surface_fwd;
        _out->seg_cmfd_bwd[o] = sg._cmfd_surface_bwd;
        if (_with_ls) {
          _out->seg_start[3 * o] = sg._starting_position[0];
          _out->seg_start[3 * o + 1] = sg._starting_position[1];
          _out->seg_start[3 * o + 2] = sg._starting_position[2];
        }
      }
      return;
    }
    if (n_in_stack == 1) _per_track[uid].reserve(n);
    for (int s = 0; s < n; s++) {
      const segment& sg = segments[s];
      long id = uid + sg._track_idx;
      Seg o;
      o.len = sg._length;
      o.fsr = sg._region_id;
      if (sg._material != last_mat) {       /* consecutive segments mostly share a material */
        std::map<Material*, int>::iterator it = _mat_index->find(sg._material);
        last_mat = sg._material;
        last_idx = (it == _mat_index->end()) ? -1 : it->second;
      }
      o.mat = last_idx;
      o.cf = sg._cmfd_surface_fwd;
      o.cb = sg._cmfd_surface_bwd;
      o.x = sg._starting_position[0];
      o.y = sg._starting_position[1];
      o.z = sg._starting_position[2];
      _per_track[id].push_back(o);
    }
  }

  std::vector<std::vector<Seg> > _per_track;
  std::vector<char> _seen;

 private:
  std::map<Material*, int>* _mat_index;
  B200FlatTracks* _out;
  const std::vector<int64_t>* _offsets;
  bool _with_ls;
};

/** Segment count of every explicit track (Track::getNumSegments), by uid. */
class CountPass : public TraverseSegments {
 public:
  CountPass(TrackGenerator* tg, std::vector<int64_t>* counts) : TraverseSegments(tg), _counts(counts) {}
  void execute() {
#pragma omp parallel
    loopOverTracks(NULL);
  }
  void onTrack(Track* track, segment* segments) { (*_counts)[track->getUid()] = track->getNumSegments(); }
 private:
  std::vector<int64_t>* _counts;
};

}  // namespace


bool b200_can_trace_on_device(TrackGenerator* tg) {
  TrackGenerator3D* tg3 = dynamic_cast<TrackGenerator3D*>(tg);
  if (tg3 == NULL) return false;
  segmentationType f = tg3->getSegmentFormation();
  return f == OTF_TRACKS || f == OTF_STACKS;
}

namespace {

/* The z-stacks handed to the device tracer: no 3D segment is ever made on the host.  Replaces the
 * host-side expansion through TraverseSegments::loopOverTracksByTrackOTF / ByStackOTF
 * (src/TraverseSegments.cpp:151-260), whose 56-byte struct segment per 3D segment is what stops
 * production-size 3D decks long before the GPU's memory does. */
void flatten_for_device_tracer(TrackGenerator3D* tg3, B200FlatTracks* ft) {
  Geometry* geometry = tg3->getGeometry();
  Quadrature* quad = tg3->getQuadrature();
  const int A2 = ft->num_azim / 2, P = ft->num_polar;
  ft->device_otf = true;

  /* 2D tracks and their segments (segment::_region_id is the extruded FSR id in the OTF formations) */
  Track** t2d = tg3->get2DTracksArray();
  const long n2 = tg3->getNum2DTracks();
  ft->n_tracks_2d = n2;
  ft->trk2d_seg_offset.assign(n2 + 1, 0);
  for (long t = 0; t < n2; t++) ft->trk2d_seg_offset[t + 1] = ft->trk2d_seg_offset[t] + t2d[t]->getNumSegments();
  const int64_t ns2 = ft->trk2d_seg_offset[n2];
  ft->seg2d_length.resize(ns2); ft->seg2d_ext.resize(ns2);
  Cmfd* cmfd = geometry->getCmfd();
  ft->otf_cmfd = (cmfd != NULL && cmfd->isFluxUpdateOn());
  if (ft->otf_cmfd) { ft->seg2d_surf_fwd.assign(ns2, -1); ft->seg2d_surf_bwd.assign(ns2, -1); }
  for (long t = 0; t < n2; t++) {
    segment* segs = t2d[t]->getSegments();
    int64_t o = ft->trk2d_seg_offset[t];
    for (int s = 0; s < t2d[t]->getNumSegments(); s++, o++) {
      ft->seg2d_length[o] = segs[s]._length;
      ft->seg2d_ext[o] = segs[s]._region_id;
      if (ft->otf_cmfd) {
        /* only the surface part matters: the 3D cell comes from the 3D FSR (TraverseSegments.cpp:447-456) */
        if (segs[s]._cmfd_surface_fwd != -1) ft->seg2d_surf_fwd[o] = (int8_t)(segs[s]._cmfd_surface_fwd % NUM_SURFACES);
        if (segs[s]._cmfd_surface_bwd != -1) ft->seg2d_surf_bwd[o] = (int8_t)(segs[s]._cmfd_surface_bwd % NUM_SURFACES);
      }
    }
  }
  if (ft->otf_cmfd) {
    ft->fsr_cmfd_cell.resize(ft->n_fsrs);
    for (int64_t r = 0; r < ft->n_fsrs; r++) ft->fsr_cmfd_cell[r] = geometry->getCmfdCell(r);
    Lattice* lat = cmfd->getLattice();
    ft->cmfd_nx = cmfd->getNumX(); ft->cmfd_ny = cmfd->getNumY(); ft->cmfd_nz = cmfd->getNumZ();
    const std::vector<double>& acc = lat->getAccumulateZ();
    ft->cmfd_z_planes.resize(ft->cmfd_nz + 1);
    for (int k = 0; k <= ft->cmfd_nz; k++) ft->cmfd_z_planes[k] = acc[k] + lat->getMinZ();
  }

  /* extruded FSRs: axial mesh (their own, or the global one) and 3D FSR ids, bottom-up */
  double* global_mesh = NULL;
  int global_n = 0;
  tg3->retrieveGlobalZMesh(global_mesh, global_n);
  const long n_ext = (long)geometry->getExtrudedFSRLookup().size();
  ft->n_extruded = n_ext;
  ft->ext_offset.assign(n_ext + 1, 0);
  for (long e = 0; e < n_ext; e++)
    ft->ext_offset[e + 1] = ft->ext_offset[e] + (global_mesh != NULL ? (long)global_n : (long)geometry->getExtrudedFSR(e)->_num_fsrs);
  ft->ext_fsr.resize(ft->ext_offset[n_ext]);
  ft->ext_mesh.resize(ft->ext_offset[n_ext] + n_ext);
  for (long e = 0; e < n_ext; e++) {
    ExtrudedFSR* ef = geometry->getExtrudedFSR(e);
    const long n = ft->ext_offset[e + 1] - ft->ext_offset[e];
    const double* mesh = global_mesh != NULL ? global_mesh : ef->_mesh;
    for (long k = 0; k < n; k++) ft->ext_fsr[ft->ext_offset[e] + k] = (int32_t)ef->_fsr_ids[k];
    for (long k = 0; k <= n; k++) ft->ext_mesh[ft->ext_offset[e] + e + k] = mesh[k];
  }

  /* corrected polar angles */
  ft->otf_theta.resize((size_t)A2 * P);
  for (int a = 0; a < A2; a++)
    for (int p = 0; p < P; p++) ft->otf_theta[a * P + p] = quad->getTheta(a, p);

  /* 3D tracks: TrackGenerator3D::getTrackOTF (src/TrackGenerator3D.cpp:1697-1745) once per track */
  const size_t nt = ft->n_tracks;
  ft->trk_2d.assign(nt, 0); ft->trk_l0.assign(nt, 0.); ft->trk_z0.assign(nt, 0.);
  int*** tps = tg3->getTracksPerStack();
  Track** rows = tg3->get2DTracks();
  std::vector<char> seen(nt, 0);
  for (int a = 0; a < A2; a++) {
    const int nxy = tg3->getNumX(a) + tg3->getNumY(a);
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < nxy; i++) {
      Track* flat2d = &rows[a][i];
      const double cos_phi = cos(flat2d->getPhi());
      for (int p = 0; p < P; p++)
        for (int z = 0; z < tps[a][i][p]; z++) {
          TrackStackIndexes tsi;
          tsi._azim = a; tsi._xy = i; tsi._polar = p; tsi._z = z;
          Track3D t;
          tg3->getTrackOTF(&t, &tsi);
          const long id = t.getUid();
          if (id < 0 || id >= (long)nt) log_printf(ERROR, "b200_flatten: 3D track uid %ld out of range", id);
          seen[id] = 1;
          ft->trk_azim[id] = a; ft->trk_xy[id] = i; ft->trk_polar[id] = p;
          ft->trk_phi[id] = t.getPhi(); ft->trk_theta[id] = t.getTheta();
          ft->trk_next_fwd[id] = t.getTrackNextFwd(); ft->trk_next_bwd[id] = t.getTrackNextBwd();
          ft->trk_flags[id] = (t.getNextFwdFwd() ? 1 : 0) | (t.getNextBwdFwd() ? 2 : 0);
          ft->trk_bc_fwd[id] = (uint8_t)t.getBCFwd(); ft->trk_bc_bwd[id] = (uint8_t)t.getBCBwd();
          ft->trk_2d[id] = (int32_t)flat2d->getUid();
          ft->trk_l0[id] = (t.getStart()->getX() - flat2d->getStart()->getX()) / cos_phi;
          ft->trk_z0[id] = t.getStart()->getZ();
          ft->trk_start[3 * id] = t.getStart()->getX();
          ft->trk_start[3 * id + 1] = t.getStart()->getY();
          ft->trk_start[3 * id + 2] = t.getStart()->getZ();
        }
    }
  }
  for (size_t t = 0; t < nt; t++)
    if (!seen[t]) log_printf(ERROR, "b200_flatten: 3D track %ld was never visited", (long)t);
  ft->trk_seg_offset.assign(nt + 1, 0);
  ft->n_segments = 0;
}

}  // namespace

void b200_flatten(TrackGenerator* tg, B200FlatTracks* ft, bool with_ls_data, bool device_otf) {

  Geometry* geometry = tg->getGeometry();
  Quadrature* quad = tg->getQuadrature();
  TrackGenerator3D* tg3 = dynamic_cast<TrackGenerator3D*>(tg);

  ft->num_groups = geometry->getNumEnergyGroups();
  ft->num_azim = tg->getNumAzim();
  ft->num_polar = quad->getNumPolarAngles();
  ft->solve_3d = (tg3 != NULL);
  ft->fluxes_per_track = ft->solve_3d ? ft->num_groups
                                      : ft->num_groups * ft->num_polar / 2;
  ft->n_tracks = tg->getNumTracks();
  ft->n_fsrs = geometry->getNumFSRs();
  const int G = ft->num_groups;

  /* ---- materials, in the (ordered) id map of the Geometry ---- */
  std::map<int, Material*> mats = geometry->getAllMaterials();
  std::map<Material*, int> mat_index;
  ft->n_materials = mats.size();
  ft->mat_sigma_t.assign((size_t)ft->n_materials * G, 0.);
  ft->mat_sigma_a = ft->mat_sigma_f = ft->mat_nu_sigma_f = ft->mat_chi = ft->mat_sigma_t;
  ft->mat_sigma_s.assign((size_t)ft->n_materials * G * G, 0.);
  ft->mat_fiss_matrix = ft->mat_sigma_s;
  ft->mat_fissionable.assign(ft->n_materials, 0);
  int m = 0;
  for (std::map<int, Material*>::iterator it = mats.begin(); it != mats.end(); ++it, ++m) {
    Material* mat = it->second;
    mat_index[mat] = m;
    if (mat->getNumEnergyGroups() != G)
      log_printf(ERROR, "b200_flatten: material %d has %d groups, expected %d",
                 mat->getId(), mat->getNumEnergyGroups(), G);
    /* getFissionMatrix() builds chi x nu_sigma_f on first use and otherwise
     * returns the matrix Solver::initializeMaterials prepared (transposed in
     * adjoint mode, Solver.cpp:794-812) - do not rebuild it here. */
    ft->mat_fissionable[m] = mat->isFissionable();
    FP_PRECISION* sigma_f = NULL;  /* optional: log_printf(ERROR) throws if unset */
    try { sigma_f = mat->getSigmaF(); } catch (std::exception&) { sigma_f = NULL; }
    for (int g = 0; g < G; g++) {
      ft->mat_sigma_t[m * G + g] = mat->getSigmaT()[g];
      ft->mat_sigma_a[m * G + g] = mat->getSigmaA()[g];
      ft->mat_sigma_f[m * G + g] = sigma_f ? sigma_f[g] : 0.;
      ft->mat_nu_sigma_f[m * G + g] = mat->getNuSigmaF()[g];
      ft->mat_chi[m * G + g] = mat->getChi()[g];
    }
    for (int i = 0; i < G * G; i++) {
      ft->mat_sigma_s[(size_t)m * G * G + i] = mat->getSigmaS()[i];
      ft->mat_fiss_matrix[(size_t)m * G * G + i] = mat->getFissionMatrix()[i];
    }
  }

  /* ---- quadrature tables ---- */
  int A2 = ft->num_azim / 2, P = ft->num_polar;
  ft->quad_weight.resize((size_t)A2 * P);
  ft->quad_sin_theta.resize((size_t)A2 * P);
  ft->quad_azim_spacing.resize(A2); ft->quad_azim_weight.resize(A2);
  ft->quad_polar_spacing.assign((size_t)A2 * P, 0.); ft->quad_polar_weight.resize((size_t)A2 * P);
  for (int a = 0; a < A2; a++) {
    ft->quad_azim_spacing[a] = quad->getAzimSpacing(a);
    ft->quad_azim_weight[a] = quad->getAzimWeight(a);
    for (int p = 0; p < P; p++) {
      ft->quad_weight[a * P + p] = quad->getWeightInline(a, p);
      ft->quad_sin_theta[a * P + p] = quad->getSinThetaInline(a, p);
      ft->quad_polar_weight[a * P + p] = quad->getPolarWeight(a, p);
      if (ft->solve_3d) ft->quad_polar_spacing[a * P + p] = quad->getPolarSpacing(a, p);
    }
  }

  /* ---- FSRs ---- */
  FP_PRECISION* vols = tg->getFSRVolumesBuffer();
  if (vols == NULL)
    log_printf(ERROR, "b200_flatten called before Solver::initializeFSRs()");
  ft->fsr_volume.resize(ft->n_fsrs);
  ft->fsr_mat.resize(ft->n_fsrs);
  ft->fsr_centroid.assign((size_t)ft->n_fsrs * 3, 0.);
  for (long r = 0; r < ft->n_fsrs; r++) {
    ft->fsr_volume[r] = vols[r];
    ft->fsr_mat[r] = mat_index[geometry->findFSRMaterial(r)];
    Point* c = geometry->getFSRCentroid(r);
    if (c != NULL) {
      ft->fsr_centroid[3 * r] = c->getX();
      ft->fsr_centroid[3 * r + 1] = c->getY();
      ft->fsr_centroid[3 * r + 2] = c->getZ();
    }
  }

  /* ---- tracks and segments ---- */
  size_t nt = ft->n_tracks;
  ft->trk_azim.assign(nt, 0); ft->trk_polar.assign(nt, 0); ft->trk_xy.assign(nt, 0);
  ft->trk_next_fwd.assign(nt, -1); ft->trk_next_bwd.assign(nt, -1);
  ft->trk_flags.assign(nt, 0); ft->trk_bc_fwd.assign(nt, 0); ft->trk_bc_bwd.assign(nt, 0);
  ft->trk_phi.assign(nt, 0.); ft->trk_theta.assign(nt, 0.);
  ft->trk_start.assign(nt * (ft->solve_3d ? 3 : 2), 0.);

  ft->device_otf = false;
  if (device_otf && b200_can_trace_on_device(tg)) {
    flatten_for_device_tracer(tg3, ft);
    return;
  }
  {
    FlattenPass probe(tg, &mat_index, ft);
    if (probe.explicitFormation()) {
      /* two traversals, no per-track copies: count, then write in place */
      std::vector<int64_t> counts(nt, 0);
      CountPass count(tg, &counts);
      count.execute();
      ft->trk_seg_offset.assign(nt + 1, 0);
      for (size_t t = 0; t < nt; t++) ft->trk_seg_offset[t + 1] = ft->trk_seg_offset[t] + counts[t];
      ft->n_segments = ft->trk_seg_offset[nt];
      const size_t ns = ft->n_segments;
      ft->seg_length.resize(ns); ft->seg_fsr.resize(ns); ft->seg_mat.resize(ns);
      ft->seg_cmfd_fwd.resize(ns); ft->seg_cmfd_bwd.resize(ns);
      if (with_ls_data) ft->seg_start.resize(3 * ns);
      FlattenPass fill(tg, &mat_index, ft, &ft->trk_seg_offset, with_ls_data);
      fill.execute();
      for (size_t t = 0; t < nt; t++)
        if (!fill._seen[t]) log_printf(ERROR, "b200_flatten: track %ld was never visited", (long)t);
      return;
    }
  }
  FlattenPass pass(tg, &mat_index, ft);
  pass.execute();

  ft->trk_seg_offset.assign(nt + 1, 0);
  for (size_t t = 0; t < nt; t++) {
    if (!pass._seen[t])
      log_printf(ERROR, "b200_flatten: track %ld was never visited", (long)t);
    ft->trk_seg_offset[t + 1] = ft->trk_seg_offset[t] + (int64_t)pass._per_track[t].size();
  }
  ft->n_segments = ft->trk_seg_offset[nt];
  size_t ns = ft->n_segments;
  ft->seg_length.resize(ns); ft->seg_fsr.resize(ns); ft->seg_mat.resize(ns);
  ft->seg_cmfd_fwd.resize(ns); ft->seg_cmfd_bwd.resize(ns);
  if (with_ls_data) ft->seg_start.resize(3 * ns);
#pragma omp parallel for schedule(static, 256)
  for (size_t t = 0; t < nt; t++) {
    size_t o = ft->trk_seg_offset[t];
    std::vector<FlattenPass::Seg>& v = pass._per_track[t];
    for (size_t s = 0; s < v.size(); s++) {
      ft->seg_length[o + s] = v[s].len;
      ft->seg_fsr[o + s] = v[s].fsr;
      ft->seg_mat[o + s] = v[s].mat;
      ft->seg_cmfd_fwd[o + s] = v[s].cf;
      ft->seg_cmfd_bwd[o + s] = v[s].cb;
      if (with_ls_data) {
        ft->seg_start[3 * (o + s)] = v[s].x;
        ft->seg_start[3 * (o + s) + 1] = v[s].y;
        ft->seg_start[3 * (o + s) + 2] = v[s].z;
      }
    }
    std::vector<FlattenPass::Seg>().swap(v);
  }
}


/* ------------------------------------------------------------------------- */
/* Track file: magic "B2TRK001", int64 n_chunks, then per chunk               */
/*   char name[24]; char dtype[8]; int64 count; data padded to 8 bytes.       */
/* ------------------------------------------------------------------------- */
namespace {
struct ChunkWriter {
  FILE* f; int64_t n;
  void put(const char* name, const char* dtype, const void* data, int64_t count, size_t elt) {
    char nm[24]; char dt[8];
    memset(nm, 0, sizeof nm); memset(dt, 0, sizeof dt);
    strncpy(nm, name, 23); strncpy(dt, dtype, 7);
    fwrite(nm, 1, 24, f); fwrite(dt, 1, 8, f); fwrite(&count, 8, 1, f);
    size_t bytes = (size_t)count * elt;
    if (bytes) fwrite(data, 1, bytes, f);
    static const char pad[8] = {0};
    if (bytes % 8) fwrite(pad, 1, 8 - bytes % 8, f);
    n++;
  }
  void scalar(const char* name, int64_t v) { put(name, "i8", &v, 1, 8); }
  void v(const char* name, const std::vector<double>& a) { put(name, "f8", a.data(), a.size(), 8); }
  void v(const char* name, const std::vector<int32_t>& a) { put(name, "i4", a.data(), a.size(), 4); }
  void v(const char* name, const std::vector<int64_t>& a) { put(name, "i8", a.data(), a.size(), 8); }
  void v(const char* name, const std::vector<uint8_t>& a) { put(name, "u1", a.data(), a.size(), 1); }
};
}  // namespace

void b200_write_trackfile(const B200FlatTracks& ft, const std::string& path, const B200CmfdView* cmfd) {
  FILE* f = fopen(path.c_str(), "wb");
  if (f == NULL) log_printf(ERROR, "b200_write_trackfile: cannot open %s", path.c_str());
  fwrite("B2TRK001", 1, 8, f);
  int64_t zero = 0;
  fwrite(&zero, 8, 1, f);  /* patched below */
  ChunkWriter w = {f, 0};
  w.scalar("num_groups", ft.num_groups);
  w.scalar("num_azim", ft.num_azim);
  w.scalar("num_polar", ft.num_polar);
  w.scalar("solve_3d", ft.solve_3d);
  w.scalar("fluxes_per_track", ft.fluxes_per_track);
  w.scalar("n_tracks", ft.n_tracks);
  w.scalar("n_segments", ft.n_segments);
  w.scalar("n_fsrs", ft.n_fsrs);
  w.scalar("n_materials", ft.n_materials);
  w.v("seg_length", ft.seg_length); w.v("seg_fsr", ft.seg_fsr); w.v("seg_mat", ft.seg_mat);
  w.v("seg_cmfd_fwd", ft.seg_cmfd_fwd); w.v("seg_cmfd_bwd", ft.seg_cmfd_bwd);
  w.v("seg_start", ft.seg_start);
  w.v("trk_seg_offset", ft.trk_seg_offset);
  w.v("trk_azim", ft.trk_azim); w.v("trk_polar", ft.trk_polar); w.v("trk_xy", ft.trk_xy);
  w.v("trk_next_fwd", ft.trk_next_fwd); w.v("trk_next_bwd", ft.trk_next_bwd);
  w.v("trk_flags", ft.trk_flags); w.v("trk_bc_fwd", ft.trk_bc_fwd); w.v("trk_bc_bwd", ft.trk_bc_bwd);
  w.v("trk_phi", ft.trk_phi); w.v("trk_theta", ft.trk_theta); w.v("trk_start", ft.trk_start);
  w.v("quad_weight", ft.quad_weight); w.v("quad_sin_theta", ft.quad_sin_theta);
  w.v("quad_azim_spacing", ft.quad_azim_spacing); w.v("quad_azim_weight", ft.quad_azim_weight);
  w.v("quad_polar_spacing", ft.quad_polar_spacing); w.v("quad_polar_weight", ft.quad_polar_weight);
  w.v("fsr_volume", ft.fsr_volume); w.v("fsr_mat", ft.fsr_mat); w.v("fsr_centroid", ft.fsr_centroid);
  w.v("mat_sigma_t", ft.mat_sigma_t); w.v("mat_sigma_a", ft.mat_sigma_a);
  w.v("mat_sigma_f", ft.mat_sigma_f); w.v("mat_nu_sigma_f", ft.mat_nu_sigma_f);
  w.v("mat_chi", ft.mat_chi); w.v("mat_sigma_s", ft.mat_sigma_s);
  w.v("mat_fiss_matrix", ft.mat_fiss_matrix); w.v("mat_fissionable", ft.mat_fissionable);
  if (cmfd != NULL) {
    std::vector<int32_t> dims = {cmfd->num_x, cmfd->num_y, cmfd->num_z, cmfd->num_cmfd_groups};
    std::vector<int32_t> bcs(cmfd->boundaries, cmfd->boundaries + 6);
    std::vector<double> options = {cmfd->sor_factor, cmfd->relaxation_factor, cmfd->flux_limiting ? 1. : 0.};
    std::vector<int32_t> fsr_cell(ft.n_fsrs, -1);
    for (size_t i = 0; i + 1 < cmfd->cell_fsr_offset.size(); i++)
      for (int64_t j = cmfd->cell_fsr_offset[i]; j < cmfd->cell_fsr_offset[i + 1]; j++)
        if (cmfd->cell_fsrs[j] >= 0 && cmfd->cell_fsrs[j] < ft.n_fsrs) fsr_cell[cmfd->cell_fsrs[j]] = (int32_t)i;
    w.v("cmfd_dims", dims); w.v("cmfd_boundaries", bcs); w.v("cmfd_options", options);
    w.v("cmfd_widths_x", cmfd->widths_x); w.v("cmfd_widths_y", cmfd->widths_y); w.v("cmfd_widths_z", cmfd->widths_z);
    w.v("cmfd_group_indices", cmfd->group_indices); w.v("fsr_cmfd_cell", fsr_cell);
  }
  fseek(f, 8, SEEK_SET);
  fwrite(&w.n, 8, 1, f);
  fclose(f);
}
