/*
 * otf.cuh - axial on-the-fly ray tracing on the device.
 *
 * Replaces TraverseSegments::traceSegmentsOTF / traceStackOTF + SegmentationKernel
 * (src/TraverseSegments.cpp:304-505, 523-911, src/MOCKernel.cpp:216-268): the 3D segments of a
 * track are never stored on the host; they follow from the 2D segments of its radial ("flattened")
 * track, the axial mesh of every extruded FSR (struct ExtrudedFSR, src/Geometry.h:84-107) and
 * the track's starting point.  The host hands those over once (b200_upload_tracks_otf); the
 * kernels here either expand them into the device segment stream (count + fill, no host copy
 * of a 3D segment ever exists) or trace them inside the sweep itself (sweep_otf.cuh).
 *
 * The walk restates traceSegmentsOTF: advance along the 2D segments, cut every 2D segment at the
 * axial mesh planes the track crosses inside it, drop pieces shorter than TINY_MOVE, stop at
 * the top / bottom of the mesh.  Arithmetic uses the non-contracting intrinsics so that the
 * pieces are bit-identical to a host evaluation of the same formulas (count pass == fill pass
 * == csrc/trackgen.cpp's host tracer).
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sweep.cuh"

namespace b200 {

constexpr double OTF_TINY_MOVE = 1e-8;   /* src/constants.h:41 */

struct OtfGeom {
  /* 2D ("flattened") tracks */
  const double* __restrict__ seg2d_len;
  const int32_t* __restrict__ seg2d_ext;       /* extruded FSR id of the 2D segment */
  const int64_t* __restrict__ trk2d_off;       /* n_trk2d + 1 */
  /* extruded FSRs: extruded FSR e owns 3D FSRs ext_fsr[ext_off[e] .. ext_off[e+1]) bottom-up and the
   * mesh planes ext_mesh[ext_off[e] + e .. ext_off[e+1] + e] (one more plane than FSRs).
   * ext_off == NULL: one global mesh (TrackGenerator3D::useGlobalZMesh) of n_axial layers in
   * ext_mesh[0 .. n_axial], 3D FSR id = e * n_axial + layer */
  const int64_t* __restrict__ ext_off;
  const double* __restrict__ ext_mesh;
  const int32_t* __restrict__ ext_fsr;
  int n_axial;
  /* 3D tracks */
  const int32_t* __restrict__ trk_2d;
  const double* __restrict__ trk_l0;           /* distance of the start point from the start of the 2D track */
  const double* __restrict__ trk_z0;
  const int32_t* __restrict__ trk_class;       /* azim * P + polar */
  const double* __restrict__ cls_cos_theta;    /* signed: negative for downward tracks */
  const double* __restrict__ cls_sin_theta;
  int64_t n_trk;
  /* maximum optical length of a segment (MOCKernel::_max_tau): longer pieces are cut the way
   * SegmentationKernel / TransportKernel cut them (src/MOCKernel.cpp:216-268, 353-410); NULL: no cuts */
  const double* __restrict__ fsr_max_sigma_t;
  double max_tau;
  /* CMFD surfaces of the 3D segments (TraverseSegments.cpp:429-457 + Lattice::getLatticeSurfaceOTF,
   * src/Universe.cpp:2241-2326): the surface (0..9, or -1) the 2D segment crosses at either end, the CMFD cell of
   * every 3D FSR (Geometry::getCmfdCell) and the z planes of the CMFD mesh; seg2d_surf_fwd == NULL: no CMFD */
  const int8_t* __restrict__ seg2d_surf_fwd;
  const int8_t* __restrict__ seg2d_surf_bwd;
  const int32_t* __restrict__ fsr_cmfd_cell;
  const double* __restrict__ cmfd_z;
  int cmfd_nxy;
};

/* Lattice::getLatticeSurfaceOTF: combines the 2D surface with a crossing of the cell's bottom / top plane */
__device__ __forceinline__ int otf_cmfd_surface(const OtfGeom& g, int cell, double z, int surface_2d) {
  const int lat_z = cell / g.cmfd_nxy;
  int at = -1;
  if (fabs(g.cmfd_z[lat_z] - z) < OTF_TINY_MOVE) at = 0;
  else if (fabs(g.cmfd_z[lat_z + 1] - z) < OTF_TINY_MOVE) at = 1;
  if (at < 0) return surface_2d < 0 ? -1 : cell * 26 + surface_2d;
  int surface;
  switch (surface_2d) {                         /* src/constants.h:120-145 */
    case 0: surface = 10 + 2 * at; break;      /* X_MIN -> X_MIN_Z_MIN / X_MIN_Z_MAX */
    case 3: surface = 11 + 2 * at; break;      /* X_MAX */
    case 1: surface = 14 + 2 * at; break;      /* Y_MIN */
    case 4: surface = 15 + 2 * at; break;      /* Y_MAX */
    case 6: surface = 18 + at; break;          /* X_MIN_Y_MIN */
    case 8: surface = 20 + at; break;          /* X_MIN_Y_MAX */
    case 7: surface = 22 + at; break;          /* X_MAX_Y_MIN */
    case 9: surface = 24 + at; break;          /* X_MAX_Y_MAX */
    default: surface = 2 + 3 * at;             /* Z_MIN / Z_MAX */
  }
  return cell * 26 + surface;
}

/* TraverseSegments::findMeshIndex (src/TraverseSegments.cpp:926-956) */
__device__ __forceinline__ int otf_mesh_index(const double* __restrict__ v, int size, double val, int sign) {
  int imin = 0, imax = size - 1;
  while (imax - imin > 1) {
    const int imid = (imin + imax) / 2;
    const double m = v[imid];
    if (val > m) imin = imid;
    else if (val < m) imax = imid;
    else return sign > 0 ? imid : imid - 1;
  }
  return imin;
}

/* Walks 3D track t forward and calls emit(length_3d, fsr_3d, cmfd_surface_fwd, cmfd_surface_bwd) for every 3D
 * segment (the surfaces are -1 without CMFD data). */
template <typename Emit>
__device__ __forceinline__ void otf_trace(const OtfGeom& g, int64_t t, Emit emit) {
  const int32_t t2 = g.trk_2d[t];
  const int cls = g.trk_class[t];
  const double cos_theta = g.cls_cos_theta[cls], sin_theta = g.cls_sin_theta[cls];
  const int sign = (cos_theta > 0) - (cos_theta < 0);
  double z = g.trk_z0[t];
  double start_dist = g.trk_l0[t];
  int64_t s = g.trk2d_off[t2];
  const int64_t s1 = g.trk2d_off[t2 + 1];
  for (; s < s1; s++) {
    const double l = g.seg2d_len[s];
    if (start_dist > l) start_dist = __dsub_rn(start_dist, l);
    else break;
  }
  if (s == s1) return;
  const bool global = g.ext_off == nullptr;
  const double* mesh = g.ext_mesh;
  int nf = g.n_axial;
  int64_t fsr0 = 0;
  int zi = 0;
  bool first = true;
  for (; s < s1; s++) {
    const int32_t e = g.seg2d_ext[s];
    if (first || !global) {
      if (global) {
        fsr0 = 0;
      } else {
        const int64_t o = g.ext_off[e];
        nf = (int)(g.ext_off[e + 1] - o);
        mesh = g.ext_mesh + o + e;
        fsr0 = o;
      }
      zi = otf_mesh_index(mesh, nf + 1, z, sign);
      first = false;
    }
    double remaining = __dsub_rn(g.seg2d_len[s], start_dist);
    start_dist = 0.0;
    bool complete = false;
    while (remaining > 0) {
      const double z_dist = sign > 0 ? __ddiv_rn(__dsub_rn(mesh[zi + 1], z), cos_theta)
                                     : __ddiv_rn(__dsub_rn(mesh[zi], z), cos_theta);
      const double seg_dist = __ddiv_rn(remaining, sin_theta);
      double d2, d3;
      int zmove;
      if (z_dist <= seg_dist) { d2 = __dmul_rn(z_dist, sin_theta); d3 = z_dist; zmove = sign; }
      else { d2 = remaining; d3 = seg_dist; zmove = 0; }
      if (d3 > OTF_TINY_MOVE) {
        const int32_t fsr = global ? (int32_t)((int64_t)e * nf + zi) : g.ext_fsr[fsr0 + zi];
        int cf = -1, cb = -1;
        if (g.seg2d_surf_fwd != nullptr) {
          int s2b = -1, s2f = -1;
          if (__dsub_rn(g.seg2d_len[s], remaining) <= OTF_TINY_MOVE) s2b = g.seg2d_surf_bwd[s];
          const double next_dist = __ddiv_rn(__dsub_rn(remaining, d2), sin_theta);
          if (zmove == 0 || next_dist <= OTF_TINY_MOVE) s2f = g.seg2d_surf_fwd[s];
          const int cell = g.fsr_cmfd_cell[fsr];
          cb = otf_cmfd_surface(g, cell, z, s2b);
          cf = otf_cmfd_surface(g, cell, __dadd_rn(z, __dmul_rn(d3, cos_theta)), s2f);
        }
        double len = d3;
        if (g.fsr_max_sigma_t != nullptr) {
          /* num_cuts = length * max_sigma_t * sin(theta) / max_tau + 1 pieces, all but the last of length
           * max_tau / (max_sigma_t * sin(theta)) (the reference's rule, sin(theta) included) */
          const double ms = __dmul_rn(g.fsr_max_sigma_t[fsr], sin_theta);
          const double t = __dmul_rn(len, ms);
          if (t > g.max_tau) {
            const int cuts = (int)__ddiv_rn(t, g.max_tau) + 1;
            const double piece = __ddiv_rn(g.max_tau, ms);
            /* SegmentationKernel::execute (MOCKernel.cpp:237-266): the backward surface stays with the first
             * piece, the forward surface with the last */
            for (int c = 0; c < cuts - 1; c++) { emit(piece, fsr, -1, c == 0 ? cb : -1); len = __dsub_rn(len, piece); }
            if (cuts > 1) cb = -1;
          }
        }
        emit(len, fsr, cf, cb);
      }
      z = __dadd_rn(z, __dmul_rn(d3, cos_theta));
      remaining = __dsub_rn(remaining, d2);
      zi += zmove;
      if (zi < 0 || zi >= nf) { zi = zi < 0 ? 0 : nf - 1; complete = true; break; }
    }
    if (complete) break;
  }
}

/* pass 1: number of 3D segments of every track */
__global__ void otf_count_kernel(const OtfGeom g, int32_t* __restrict__ count) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < g.n_trk; t += (int64_t)gridDim.x * blockDim.x) {
    int32_t n = 0;
    otf_trace(g, t, [&](double, int32_t, int, int) { n++; });
    count[t] = n;
  }
}

/* pass 2: writes the padded device segment stream of sweep.cuh directly ({length, FSR id * G} records,
 * seg already skips the front padding) and, when cls_vol_weight is given, tallies the FSR volumes
 * (VolumeKernel, src/MOCKernel.cpp:80-162: azimuthal x polar spacing and weight times length) */
__global__ void otf_fill_kernel(const OtfGeom g, const int64_t* __restrict__ trk_off, SegRec* __restrict__ seg,
                                int G, const double* __restrict__ cls_vol_weight, double* __restrict__ volume,
                                int2* __restrict__ seg_cmfd) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < g.n_trk; t += (int64_t)gridDim.x * blockDim.x) {
    SegRec* out = seg != nullptr ? seg + trk_off[t] : nullptr;
    int2* outc = (seg != nullptr && seg_cmfd != nullptr) ? seg_cmfd + trk_off[t] : nullptr;
    const double w = cls_vol_weight != nullptr ? cls_vol_weight[g.trk_class[t]] : 0.0;
    otf_trace(g, t, [&](double len, int32_t fsr, int cf, int cb) {
      if (out != nullptr) {
        SegRec r;
        r.len = len; r.base = (uint32_t)fsr * (uint32_t)G; r.spare = 0u;
        *out++ = r;
      }
      if (outc != nullptr) *outc++ = make_int2(cf, cb);
      if (volume != nullptr) atomicAdd(&volume[fsr], w * len);
    });
  }
}

/* the sentinel records before and after the stream */
__global__ void otf_pad_kernel(SegRec* __restrict__ padded, int64_t n_seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * SEG_PAD) return;
  SegRec r;
  r.len = 0.0; r.base = 0u; r.spare = 0u;
  padded[i < SEG_PAD ? i : n_seg + i] = r;
}

/* explicit copies of the stream for hosts that want to look at it (tests) */
__global__ void otf_unpack_kernel(const SegRec* __restrict__ seg, int64_t n_seg, int G, double* __restrict__ len,
                                  int32_t* __restrict__ fsr) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_seg; i += (int64_t)gridDim.x * blockDim.x) {
    len[i] = seg[i].len;
    fsr[i] = (int32_t)(seg[i].base / (uint32_t)G);
  }
}

}  // namespace b200
