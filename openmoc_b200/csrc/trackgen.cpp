/*
 * trackgen.cpp - synthetic 2D track generator for pin-lattice geometries.
 *
 * Benchmark / test infrastructure of the B200 solver: produces, without OpenMOC,
 * the flattened SoA tracks (the arguments of b200_upload_* in include/b200moc.h)
 * that OpenMOC's TrackGenerator + b200_flatten would produce for the named
 * shapes (pin cell, simple lattice, 2D C5G7), so that bench.py and the GPU box
 * need neither the reference nor a multi-GB track file.  It is host-only C++
 * (no CUDA) and is NOT on the sweep path.
 *
 * It restates, from scratch, three published algorithms of the reference:
 *   cyclic track laydown and angle correction   src/TrackGenerator.cpp:950-1070
 *   reflective / periodic track linking         src/TrackGenerator.cpp:1086-1222
 *   quadrature weights (TY, equal-angle)        src/Quadrature.cpp:674-745, 822-894, 1483-1553
 * and replaces the reference's CSG ray tracer (Geometry::segmentize) by an
 * analytic tracer specialised to rectangular lattices of pin cells with
 * equal-area rings and azimuthal sectors (src/Cell.cpp:1325-1389, 1397-1570).
 * tests/test_trackgen.py checks it against tracks dumped from the reference.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <omp.h>

namespace {

constexpr int KIND_PIN = 0, KIND_GRID = 1;
constexpr int BC_VACUUM = 0, BC_REFLECTIVE = 1, BC_PERIODIC = 2;

struct CellType {
  int32_t kind;            /* KIND_PIN / KIND_GRID */
  int32_t n_rings;         /* PIN: equal-area rings inside fuel_radius (0 or 1: none) */
  int32_t n_sectors_fuel;  /* PIN: azimuthal sectors inside fuel_radius (0/1: none) */
  int32_t n_sectors_mod;   /* PIN: azimuthal sectors outside */
  int32_t mat_fuel;        /* PIN: material inside; GRID: the material */
  int32_t mat_mod;         /* PIN: material outside */
  int32_t subdiv;          /* GRID: k x k plain sub-cells (k >= 1) */
  int32_t pad;
  double fuel_radius;      /* PIN; <= 0: no cylinder at all */
};

struct TypeInfo {
  CellType t;
  std::vector<double> radii;   /* descending, radii[0] = fuel radius */
  int n_regions;
  int n_fuel_regions;
};

struct Gen {
  /* inputs */
  int nx, ny, num_azim, num_polar;
  double px, py, xmin, ymin, spacing;
  int bc[4]; /* xmin, xmax, ymin, ymax */
  std::vector<int32_t> cell_type;
  std::vector<TypeInfo> types;
  /* outputs */
  std::vector<double> seg_length;
  std::vector<int32_t> seg_fsr, seg_mat;
  std::vector<int64_t> trk_seg_offset, trk_next_fwd, trk_next_bwd;
  std::vector<int32_t> trk_azim, trk_polar, trk_xy;
  std::vector<uint8_t> trk_flags, trk_bc_fwd, trk_bc_bwd;
  std::vector<double> trk_phi, trk_theta, trk_start;
  std::vector<double> quad_weight, quad_sin_theta;
  /* factors of the total weight, for the linear-source pre-pass (track file chunks of the same name) */
  std::vector<double> quad_azim_spacing, quad_azim_weight, quad_polar_spacing, quad_polar_weight;
  std::vector<double> fsr_volume;
  std::vector<int32_t> fsr_mat;
  std::vector<int64_t> fsr_base;
  int64_t n_fsrs = 0;
  std::string error;
};

inline int sector_of(double x, double y, int ns) {
  if (ns <= 1) return 0;
  const double delta = 2.0 * M_PI / ns;
  double th = atan2(y, x) + M_PI / 4.0;     /* sector i spans (-pi/4 + i*delta, -pi/4 + (i+1)*delta) */
  th = fmod(th, 2.0 * M_PI);
  if (th < 0) th += 2.0 * M_PI;
  int s = (int)(th / delta);
  return s >= ns ? ns - 1 : s;
}

/* local region of point (x, y) relative to the cell centre */
inline int region_of(const TypeInfo& ti, double x, double y, double w, double h, int* mat) {
  const CellType& t = ti.t;
  if (t.kind == KIND_GRID) {
    const int k = t.subdiv;
    int i = (int)floor((x + 0.5 * w) / (w / k));
    int j = (int)floor((y + 0.5 * h) / (h / k));
    i = std::min(std::max(i, 0), k - 1);
    j = std::min(std::max(j, 0), k - 1);
    *mat = t.mat_fuel;
    return j * k + i;
  }
  const double r2 = x * x + y * y;
  if (!ti.radii.empty() && r2 < ti.radii[0] * ti.radii[0]) {
    int ring = 0;
    for (size_t k = 1; k < ti.radii.size(); k++)
      if (r2 < ti.radii[k] * ti.radii[k]) ring = (int)k;
    *mat = t.mat_fuel;
    const int ns = std::max(t.n_sectors_fuel, 1);
    return ring * ns + sector_of(x, y, ns);
  }
  *mat = t.mat_mod;
  return ti.n_fuel_regions + sector_of(x, y, std::max(t.n_sectors_mod, 1));
}

struct Piece { double len; int32_t fsr, mat; };

/* Trace one track from (x0, y0) along (dx, dy) for total length L. */
void trace(const Gen& g, double x0, double y0, double dx, double dy, double L,
           std::vector<Piece>& out, std::vector<double>& brk, std::vector<double>& cuts) {
  out.clear();
  brk.clear();
  brk.push_back(0.0);
  brk.push_back(L);
  /* crossings with the main lattice lines */
  for (int i = 1; i < g.nx; i++) {
    double t = (g.xmin + i * g.px - x0) / dx;
    if (t > 0.0 && t < L) brk.push_back(t);
  }
  for (int j = 1; j < g.ny; j++) {
    double t = (g.ymin + j * g.py - y0) / dy;
    if (t > 0.0 && t < L) brk.push_back(t);
  }
  std::sort(brk.begin(), brk.end());

  const double tiny = 1e-12;
  for (size_t b = 0; b + 1 < brk.size(); b++) {
    const double ta = brk[b], tb = brk[b + 1];
    if (tb - ta <= tiny) continue;
    const double tm = 0.5 * (ta + tb);
    int ci = (int)floor((x0 + tm * dx - g.xmin) / g.px);
    int cj = (int)floor((y0 + tm * dy - g.ymin) / g.py);
    ci = std::min(std::max(ci, 0), g.nx - 1);
    cj = std::min(std::max(cj, 0), g.ny - 1);
    const int64_t cell = (int64_t)cj * g.nx + ci;
    const TypeInfo& ti = g.types[g.cell_type[cell]];
    const double cx = g.xmin + (ci + 0.5) * g.px, cy = g.ymin + (cj + 0.5) * g.py;
    /* ray in cell-centred coordinates: p(t) = (ox, oy) + t (dx, dy) */
    const double ox = x0 - cx, oy = y0 - cy;

    cuts.clear();
    cuts.push_back(ta);
    cuts.push_back(tb);
    if (ti.t.kind == KIND_GRID) {
      const int k = ti.t.subdiv;
      for (int i = 1; i < k; i++) {
        double t = (-0.5 * g.px + i * g.px / k - ox) / dx;
        if (t > ta && t < tb) cuts.push_back(t);
        t = (-0.5 * g.py + i * g.py / k - oy) / dy;
        if (t > ta && t < tb) cuts.push_back(t);
      }
    } else {
      /* circles: |o + t d|^2 = r^2, |d| = 1 */
      const double bq = ox * dx + oy * dy;
      const double cq0 = ox * ox + oy * oy;
      for (double r : ti.radii) {
        const double disc = bq * bq - (cq0 - r * r);
        if (disc <= 0.0) continue;
        const double sq = sqrt(disc);
        const double t1 = -bq - sq, t2 = -bq + sq;
        if (t1 > ta && t1 < tb) cuts.push_back(t1);
        if (t2 > ta && t2 < tb) cuts.push_back(t2);
      }
      /* sector planes A x + B y = 0 with (A, B) = (cos az, sin az), az = pi/4 + i*delta */
      for (int which = 0; which < 2; which++) {
        const int ns = which == 0 ? ti.t.n_sectors_fuel : ti.t.n_sectors_mod;
        if (ns < 2) continue;
        if (which == 0 && ti.radii.empty()) continue;
        const double delta = 2.0 * M_PI / ns;
        const int nlines = (ns % 2 == 0) ? ns / 2 : ns;
        for (int i = 0; i < nlines; i++) {
          const double az = M_PI / 4.0 + i * delta;
          const double A = cos(az), B = sin(az);
          const double nd = A * dx + B * dy;
          if (fabs(nd) < 1e-14) continue;
          const double t = -(A * ox + B * oy) / nd;
          if (t > ta && t < tb) cuts.push_back(t);
        }
      }
    }
    std::sort(cuts.begin(), cuts.end());
    const int64_t base = g.fsr_base[cell];
    for (size_t c = 0; c + 1 < cuts.size(); c++) {
      const double sa = cuts[c], sb = cuts[c + 1];
      if (sb - sa <= tiny) continue;
      const double sm = 0.5 * (sa + sb);
      int mat = 0;
      const int reg = region_of(ti, ox + sm * dx, oy + sm * dy, g.px, g.py, &mat);
      const int32_t fsr = (int32_t)(base + reg);
      if (!out.empty() && out.back().fsr == fsr && ti.t.kind == KIND_PIN && c > 0) {
        out.back().len += sb - sa;       /* a candidate cut that was no region boundary */
      } else {
        out.push_back({sb - sa, fsr, mat});
      }
    }
  }
}

int build(Gen& g, int polar_quad) {
  const int A = g.num_azim, A2 = A / 2, A4 = A / 4, P = g.num_polar, P2 = P / 2;
  const double width = g.nx * g.px, height = g.ny * g.py;
  if (A < 4 || A % 4) { g.error = "num_azim must be a positive multiple of 4"; return 1; }
  if (P < 2 || P % 2) { g.error = "num_polar must be even"; return 1; }
  if (g.spacing <= 0) { g.error = "spacing must be positive"; return 1; }

  /* ---- cell types and FSR numbering ---- */
  for (TypeInfo& ti : g.types) {
    CellType& t = ti.t;
    ti.radii.clear();
    if (t.kind == KIND_GRID) {
      if (t.subdiv < 1) t.subdiv = 1;
      ti.n_fuel_regions = 0;
      ti.n_regions = t.subdiv * t.subdiv;
    } else {
      if (t.fuel_radius > 0) {
        const int nr = std::max(t.n_rings, 1);
        /* equal-area rings (src/Cell.cpp:1497-1518) */
        double r1 = t.fuel_radius;
        const double increment = M_PI * r1 * r1 / nr;
        for (int i = 0; i < nr; i++) {
          ti.radii.push_back(r1);
          r1 = sqrt(std::max(r1 * r1 - increment / M_PI, 0.0));
        }
        ti.n_fuel_regions = nr * std::max(t.n_sectors_fuel, 1);
      } else {
        ti.n_fuel_regions = 0;
      }
      ti.n_regions = ti.n_fuel_regions + std::max(t.n_sectors_mod, 1);
    }
  }
  const int64_t n_cells = (int64_t)g.nx * g.ny;
  g.fsr_base.resize(n_cells + 1);
  g.fsr_base[0] = 0;
  for (int64_t c = 0; c < n_cells; c++) {
    if (g.cell_type[c] < 0 || g.cell_type[c] >= (int)g.types.size()) { g.error = "cell type index out of range"; return 1; }
    g.fsr_base[c + 1] = g.fsr_base[c] + g.types[g.cell_type[c]].n_regions;
  }
  g.n_fsrs = g.fsr_base[n_cells];
  if (g.n_fsrs > INT32_MAX) { g.error = "too many FSRs for 32-bit ids"; return 1; }

  /* ---- azimuthal angles, cyclic correction (src/TrackGenerator.cpp:978-1010) ---- */
  std::vector<int> num_x(A2), num_y(A2);
  std::vector<double> phi(A2), dx_eff(A2), dy_eff(A2), azim_spacing(A2), azim_weight(A2);
  for (int a = 0; a < A4; a++) {
    const double want = 2.0 * M_PI / A * (0.5 + a);
    num_x[a] = (int)(fabs(width / g.spacing * sin(want))) + 1;
    num_y[a] = (int)(fabs(height / g.spacing * cos(want))) + 1;
    num_x[A2 - a - 1] = num_x[a];
    num_y[A2 - a - 1] = num_y[a];
    const double p = atan((height * num_x[a]) / (width * num_y[a]));
    phi[a] = p;
    phi[A2 - a - 1] = M_PI - p;
    dx_eff[a] = dx_eff[A2 - a - 1] = width / num_x[a];
    dy_eff[a] = dy_eff[A2 - a - 1] = height / num_y[a];
    azim_spacing[a] = azim_spacing[A2 - a - 1] = dx_eff[a] * sin(p);
  }
  /* azimuthal weights (src/Quadrature.cpp:704-720) */
  for (int a = 0; a < A4; a++) {
    double x1 = (a < A4 - 1) ? 0.5 * (phi[a + 1] - phi[a]) : M_PI_2 - phi[a];
    double x2 = (a >= 1) ? 0.5 * (phi[a] - phi[a - 1]) : phi[a];
    azim_weight[a] = azim_weight[A2 - a - 1] = (x1 + x2) / M_PI;
  }
  /* polar quadrature */
  std::vector<double> theta(P2), pw(P2);
  if (polar_quad == 0) {            /* Tabuchi-Yamamoto (src/Quadrature.cpp:822-894) */
    if (P == 2) { theta[0] = asin(0.798184); pw[0] = 0.5; }
    else if (P == 4) { theta[0] = asin(0.363900); theta[1] = asin(0.899900); pw[0] = 0.212854 / 2.0; pw[1] = 0.787146 / 2.0; }
    else if (P == 6) { theta[0] = asin(0.166648); theta[1] = asin(0.537707); theta[2] = asin(0.932954);
                       pw[0] = 0.046233 / 2.0; pw[1] = 0.283619 / 2.0; pw[2] = 0.670148 / 2.0; }
    else { g.error = "TY quadrature supports 2, 4 or 6 polar angles"; return 1; }
  } else if (polar_quad == 1) {     /* equal angle (src/Quadrature.cpp:1483-1553) */
    const double dth = M_PI / P;
    double ta = 0.;
    for (int p = 0; p < P2; p++) {
      double tb = ta + dth;
      theta[p] = acos(0.5 * (cos(ta) + cos(tb)));
      ta = tb;
    }
    for (int p = 0; p < P2; p++) {
      double y1 = (p < P2 - 1) ? 0.5 * (cos(theta[p]) - cos(theta[p + 1])) : cos(theta[p]);
      double y2 = (p >= 1) ? 0.5 * (cos(theta[p - 1]) - cos(theta[p])) : 1.0 - cos(theta[p]);
      pw[p] = (y1 + y2) / 2.0;
    }
  } else { g.error = "unknown polar quadrature"; return 1; }
  /* total weights, 2D form (src/Quadrature.cpp:727-741), mirrored over polar halves */
  g.quad_weight.assign((size_t)A2 * P, 0.);
  g.quad_sin_theta.assign((size_t)A2 * P, 0.);
  g.quad_azim_spacing.assign(azim_spacing.begin(), azim_spacing.end());
  g.quad_azim_weight.assign(azim_weight.begin(), azim_weight.end());
  g.quad_polar_spacing.assign((size_t)A2 * P, 0.);          /* 2D: no polar spacing */
  g.quad_polar_weight.assign((size_t)A2 * P, 0.);
  for (int a = 0; a < A2; a++)
    for (int p = 0; p < P2; p++) {
      const double st = sin(theta[p]);
      const double w = 2.0 * M_PI * azim_weight[a] * azim_spacing[a] * pw[p] * 2.0 * st;
      g.quad_weight[a * P + p] = g.quad_weight[a * P + (P - 1 - p)] = w;
      g.quad_sin_theta[a * P + p] = g.quad_sin_theta[a * P + (P - 1 - p)] = st;
      g.quad_polar_weight[a * P + p] = g.quad_polar_weight[a * P + (P - 1 - p)] = pw[p];
    }

  /* ---- tracks: start / end points (src/TrackGenerator.cpp:1013-1062) ---- */
  std::vector<int64_t> first(A2 + 1, 0);
  for (int a = 0; a < A2; a++) first[a + 1] = first[a] + num_x[a] + num_y[a];
  const int64_t nt = first[A2];
  auto id = [&](int a, int i) { return first[a] + i; };
  g.trk_azim.resize(nt); g.trk_polar.assign(nt, 0); g.trk_xy.resize(nt);
  g.trk_next_fwd.resize(nt); g.trk_next_bwd.resize(nt); g.trk_flags.resize(nt);
  g.trk_bc_fwd.resize(nt); g.trk_bc_bwd.resize(nt); g.trk_phi.resize(nt);
  g.trk_theta.assign(nt, M_PI_2); g.trk_start.resize(2 * nt);
  std::vector<double> tx1(nt), ty1(nt);
  const int bxmin = g.bc[0], bxmax = g.bc[1], bymin = g.bc[2], bymax = g.bc[3];
  for (int a = 0; a < A2; a++) {
    const int ac = A2 - a - 1, nxa = num_x[a], nya = num_y[a];
    for (int i = 0; i < nxa + nya; i++) {
      const int64_t t = id(a, i);
      double sx, sy, ex, ey;
      if (a < A4) {
        if (i < nxa) { sx = g.xmin + width - dx_eff[a] * (i + 0.5); sy = g.ymin; }
        else { sx = g.xmin; sy = g.ymin + dy_eff[a] * (i - nxa + 0.5); }
        if (i < nya) { ex = g.xmin + width; ey = g.ymin + dy_eff[a] * (i + 0.5); }
        else { ex = g.xmin + width - dx_eff[a] * ((i - nya) + 0.5); ey = g.ymin + height; }
      } else {
        if (i < nxa) { sx = g.xmin + dx_eff[a] * (i + 0.5); sy = g.ymin; }
        else { sx = g.xmin + width; sy = g.ymin + dy_eff[a] * (i - nxa + 0.5); }
        if (i < nya) { ex = g.xmin; ey = g.ymin + dy_eff[a] * (i + 0.5); }
        else { ex = g.xmin + dx_eff[a] * (i - nya + 0.5); ey = g.ymin + height; }
      }
      g.trk_start[2 * t] = sx; g.trk_start[2 * t + 1] = sy;
      tx1[t] = ex; ty1[t] = ey;
      g.trk_azim[t] = a; g.trk_xy[t] = i; g.trk_phi[t] = phi[a];

      /* boundary conditions and links (src/TrackGenerator.cpp:1091-1222) */
      int bc_fwd, bc_bwd;
      if (a < A4) {
        bc_fwd = (i < nya) ? bxmax : bymax;
        bc_bwd = (i < nxa) ? bymin : bxmin;
      } else {
        bc_fwd = (i < nya) ? bxmin : bymax;
        bc_bwd = (i < nxa) ? bymin : bxmax;
      }
      g.trk_bc_fwd[t] = (uint8_t)bc_fwd; g.trk_bc_bwd[t] = (uint8_t)bc_bwd;
      bool next_fwd_fwd, next_bwd_fwd;
      int64_t next_fwd, next_bwd;
      if (i < nya) {
        next_fwd_fwd = true;
        next_fwd = (bc_fwd == BC_PERIODIC) ? id(a, i + nxa) : id(ac, i + nxa);
      } else {
        if (bymax == BC_PERIODIC) { next_fwd_fwd = true; next_fwd = id(a, i - nya); }
        else { next_fwd_fwd = false; next_fwd = id(ac, (nxa + nya) - (i - nya) - 1); }
      }
      if (i < nxa) {
        if (bymin == BC_PERIODIC) { next_bwd_fwd = false; next_bwd = id(a, i + nya); }
        else { next_bwd_fwd = true; next_bwd = id(ac, nxa - i - 1); }
      } else {
        next_bwd_fwd = false;
        next_bwd = (bc_bwd == BC_PERIODIC) ? id(a, i - nxa) : id(ac, i - nxa);
      }
      g.trk_next_fwd[t] = next_fwd; g.trk_next_bwd[t] = next_bwd;
      g.trk_flags[t] = (next_fwd_fwd ? 1 : 0) | (next_bwd_fwd ? 2 : 0);
    }
  }

  /* ---- ray tracing, parallel over tracks ---- */
  std::vector<std::vector<Piece>> per_track(nt);
#pragma omp parallel
  {
    std::vector<Piece> out;
    std::vector<double> brk, cuts;
#pragma omp for schedule(dynamic, 64)
    for (int64_t t = 0; t < nt; t++) {
      const double sx = g.trk_start[2 * t], sy = g.trk_start[2 * t + 1];
      const double ex = tx1[t] - sx, ey = ty1[t] - sy;
      const double L = sqrt(ex * ex + ey * ey);
      trace(g, sx, sy, ex / L, ey / L, L, out, brk, cuts);
      per_track[t] = out;
    }
  }
  g.trk_seg_offset.assign(nt + 1, 0);
  for (int64_t t = 0; t < nt; t++) g.trk_seg_offset[t + 1] = g.trk_seg_offset[t] + (int64_t)per_track[t].size();
  const int64_t ns = g.trk_seg_offset[nt];
  g.seg_length.resize(ns); g.seg_fsr.resize(ns); g.seg_mat.resize(ns);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t t = 0; t < nt; t++) {
    int64_t o = g.trk_seg_offset[t];
    for (const Piece& p : per_track[t]) {
      g.seg_length[o] = p.len; g.seg_fsr[o] = p.fsr; g.seg_mat[o] = p.mat; o++;
    }
    std::vector<Piece>().swap(per_track[t]);
  }

  /* ---- FSR volumes (VolumeKernel, src/MOCKernel.cpp:145-162: w_a * spacing_a * L) and materials ---- */
  g.fsr_volume.assign(g.n_fsrs, 0.);
  g.fsr_mat.assign(g.n_fsrs, -1);
  for (int64_t t = 0; t < nt; t++) {
    const int a = g.trk_azim[t];
    const double w = azim_weight[a] * azim_spacing[a];
    for (int64_t s = g.trk_seg_offset[t]; s < g.trk_seg_offset[t + 1]; s++) {
      g.fsr_volume[g.seg_fsr[s]] += w * g.seg_length[s];
      g.fsr_mat[g.seg_fsr[s]] = g.seg_mat[s];
    }
  }
  /* FSRs no track crossed still need a material: classify a representative point */
  for (int64_t c = 0; c < n_cells; c++) {
    const TypeInfo& ti = g.types[g.cell_type[c]];
    for (int r = 0; r < ti.n_regions; r++) {
      int64_t f = g.fsr_base[c] + r;
      if (g.fsr_mat[f] >= 0) continue;
      g.fsr_mat[f] = (ti.t.kind == KIND_PIN && r < ti.n_fuel_regions) ? ti.t.mat_fuel
                     : (ti.t.kind == KIND_PIN ? ti.t.mat_mod : ti.t.mat_fuel);
    }
  }
  return 0;
}

}  // namespace

/* ------------------------------- C ABI ------------------------------------ */
extern "C" {

struct b200_trackgen { Gen g; };

const char* b200_trackgen_error(b200_trackgen* h) { return h ? h->g.error.c_str() : "null handle"; }

b200_trackgen* b200_trackgen_create_2d(int nx, int ny, double pitch_x, double pitch_y, double xmin,
                                       double ymin, const int32_t* cell_type /* [ny][nx], row 0 = bottom */,
                                       const void* types /* CellType[n_types] */, int n_types,
                                       int bc_xmin, int bc_xmax, int bc_ymin, int bc_ymax, int num_azim,
                                       double spacing, int num_polar, int polar_quad, int num_threads,
                                       int* status) {
  b200_trackgen* h = new b200_trackgen();
  Gen& g = h->g;
  g.nx = nx; g.ny = ny; g.px = pitch_x; g.py = pitch_y; g.xmin = xmin; g.ymin = ymin;
  g.num_azim = num_azim; g.num_polar = num_polar; g.spacing = spacing;
  g.bc[0] = bc_xmin; g.bc[1] = bc_xmax; g.bc[2] = bc_ymin; g.bc[3] = bc_ymax;
  g.cell_type.assign(cell_type, cell_type + (size_t)nx * ny);
  const CellType* ct = (const CellType*)types;
  g.types.resize(n_types);
  for (int i = 0; i < n_types; i++) g.types[i].t = ct[i];
  if (num_threads > 0) omp_set_num_threads(num_threads);
  int rc = build(g, polar_quad);
  if (status) *status = rc;
  return h;
}

void b200_trackgen_destroy(b200_trackgen* h) { delete h; }

int64_t b200_trackgen_size(b200_trackgen* h, const char* name) {
  Gen& g = h->g;
  std::string n(name);
  if (n == "n_tracks") return (int64_t)g.trk_azim.size();
  if (n == "n_segments") return (int64_t)g.seg_length.size();
  if (n == "n_fsrs") return g.n_fsrs;
  return -1;
}

/* copy an output array into caller memory; returns element count or -1 */
int64_t b200_trackgen_get(b200_trackgen* h, const char* name, void* dst) {
  Gen& g = h->g;
  std::string n(name);
#define OUT(field)                                                             \
  if (n == #field) {                                                           \
    if (dst && !g.field.empty()) memcpy(dst, g.field.data(), g.field.size() * sizeof(g.field[0])); \
    return (int64_t)g.field.size();                                            \
  }
  OUT(seg_length) OUT(seg_fsr) OUT(seg_mat) OUT(trk_seg_offset) OUT(trk_next_fwd) OUT(trk_next_bwd)
  OUT(trk_azim) OUT(trk_polar) OUT(trk_xy) OUT(trk_flags) OUT(trk_bc_fwd) OUT(trk_bc_bwd) OUT(trk_phi)
  OUT(trk_theta) OUT(trk_start) OUT(quad_weight) OUT(quad_sin_theta) OUT(fsr_volume) OUT(fsr_mat)
  OUT(quad_azim_spacing) OUT(quad_azim_weight) OUT(quad_polar_spacing) OUT(quad_polar_weight)
#undef OUT
  return -1;
}

}  // extern "C"
