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
  /* CMFD mesh = the pin lattice: surface codes of the segment ends, segment::_cmfd_surface_fwd/_bwd, FSR -> cell */
  std::vector<int8_t> seg_surf_fwd, seg_surf_bwd;
  std::vector<int32_t> seg_cmfd_fwd, seg_cmfd_bwd, fsr_cell;
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
  /* laydown tables kept for the 3D stage (TrackGenerator3D works on top of the 2D tracks) */
  std::vector<int> num_x, num_y;
  std::vector<int64_t> first;                 /* first 2D track uid of azimuthal angle a */
  std::vector<double> phi, azim_spacing, azim_weight, trk_end;
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

/* cf / cb: surface of the lattice cell (= CMFD cell of a CMFD mesh laid over the pin lattice) the piece ends on in
 * the forward / backward direction, 0..9 as in src/constants.h:120-129, or -1 (Lattice::getLatticeSurface,
 * src/Universe.cpp:2100-2230, restricted to x / y) */
struct Piece { double len; int32_t fsr, mat; int32_t cell; int8_t cf, cb; };

inline int8_t lattice_surface_2d(double x, double y, double px, double py) {
  const double tol = 1e-10;
  const bool min_x = fabs(x + 0.5 * px) < tol, max_x = fabs(x - 0.5 * px) < tol;
  const bool min_y = fabs(y + 0.5 * py) < tol, max_y = fabs(y - 0.5 * py) < tol;
  if (min_x) return min_y ? 6 : (max_y ? 8 : 0);
  if (max_x) return min_y ? 7 : (max_y ? 9 : 3);
  if (min_y) return 1;
  if (max_y) return 4;
  return -1;
}

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
    const size_t first_piece = out.size();
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
        out.push_back({sb - sa, fsr, mat, (int32_t)cell, -1, -1});
      }
    }
    if (out.size() > first_piece) {
      out[first_piece].cb = lattice_surface_2d(ox + ta * dx, oy + ta * dy, g.px, g.py);
      out.back().cf = lattice_surface_2d(ox + tb * dx, oy + tb * dy, g.px, g.py);
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
  std::vector<int>& num_x = g.num_x; std::vector<int>& num_y = g.num_y;
  std::vector<double>& phi = g.phi; std::vector<double>& azim_spacing = g.azim_spacing;
  std::vector<double>& azim_weight = g.azim_weight;
  num_x.assign(A2, 0); num_y.assign(A2, 0); phi.assign(A2, 0.); azim_spacing.assign(A2, 0.); azim_weight.assign(A2, 0.);
  std::vector<double> dx_eff(A2), dy_eff(A2);
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
  std::vector<int64_t>& first = g.first;
  first.assign(A2 + 1, 0);
  for (int a = 0; a < A2; a++) first[a + 1] = first[a] + num_x[a] + num_y[a];
  const int64_t nt = first[A2];
  auto id = [&](int a, int i) { return first[a] + i; };
  g.trk_azim.resize(nt); g.trk_polar.assign(nt, 0); g.trk_xy.resize(nt);
  g.trk_next_fwd.resize(nt); g.trk_next_bwd.resize(nt); g.trk_flags.resize(nt);
  g.trk_bc_fwd.resize(nt); g.trk_bc_bwd.resize(nt); g.trk_phi.resize(nt);
  g.trk_theta.assign(nt, M_PI_2); g.trk_start.resize(2 * nt); g.trk_end.resize(2 * nt);
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
      g.trk_end[2 * t] = ex; g.trk_end[2 * t + 1] = ey;
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
  g.seg_surf_fwd.resize(ns); g.seg_surf_bwd.resize(ns); g.seg_cmfd_fwd.resize(ns); g.seg_cmfd_bwd.resize(ns);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t t = 0; t < nt; t++) {
    int64_t o = g.trk_seg_offset[t];
    for (const Piece& p : per_track[t]) {
      g.seg_length[o] = p.len; g.seg_fsr[o] = p.fsr; g.seg_mat[o] = p.mat;
      g.seg_surf_fwd[o] = p.cf; g.seg_surf_bwd[o] = p.cb;
      g.seg_cmfd_fwd[o] = p.cf < 0 ? -1 : p.cell * 26 + p.cf;      /* segment::_cmfd_surface_fwd, src/Track.h:42-46 */
      g.seg_cmfd_bwd[o] = p.cb < 0 ? -1 : p.cell * 26 + p.cb;
      o++;
    }
    std::vector<Piece>().swap(per_track[t]);
  }

  /* ---- FSR volumes (VolumeKernel, src/MOCKernel.cpp:145-162: w_a * spacing_a * L) and materials ---- */
  g.fsr_volume.assign(g.n_fsrs, 0.);
  g.fsr_mat.assign(g.n_fsrs, -1);
  g.fsr_cell.resize(g.n_fsrs);
  for (int64_t c = 0; c < n_cells; c++)
    for (int64_t f = g.fsr_base[c]; f < g.fsr_base[c + 1]; f++) g.fsr_cell[f] = (int32_t)c;
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


/* ========================================================================= */
/* 3D: z-stacks of tracks over the 2D tracks of an axially extruded geometry   */
/* ========================================================================= */
/*
 * Restates, in index arithmetic on flat tables, the published 3D cyclic laydown of the
 * reference (modular ray tracing in the (l, z) plane of every 2D track chain):
 *   polar-angle correction and (n_l, n_z) per angle   src/TrackGenerator3D.cpp:750-809
 *   chain tracks -> z-stacks of 3D tracks              src/TrackGenerator3D.cpp:975-1245
 *   3D reflective / periodic / vacuum links            src/TrackGenerator3D.cpp:1762-2374
 *   3D quadrature weights                              src/Quadrature.cpp:674-745, 1051-1076, 1425-1457, 1520-1553
 *   axial on-the-fly segmentation of one 3D track      src/TraverseSegments.cpp:304-505, 926-956
 * The geometry is the 2D pin lattice extruded between z_min and z_max and cut into n_axial
 * equal layers (what the reference's 3D C5G7 deck gets from its lattice planes and CMFD mesh,
 * profile/models/c5g7/c5g7-3d-cmfd.cpp:537-556); 3D FSR id = 2D FSR id * n_axial + layer.
 */
constexpr double TINY_MOVE = 1e-8;      /* src/constants.h:41 */
constexpr double FLT_EPS_REF = 1e-12;   /* src/constants.h:12 (the reference redefines FLT_EPSILON) */
constexpr int BC_INTERFACE = 3;

struct Chain { int a, x, p, lz, link; };   /* TrackChainIndexes, src/Track3D.h */
struct Stack { int a, xy, p, z; };         /* TrackStackIndexes */

struct Gen3 {
  Gen* g2 = nullptr;
  int P = 0, n_axial = 1, bc_zmin = BC_REFLECTIVE, bc_zmax = BC_REFLECTIVE;
  double zmin = 0, zmax = 0, z_spacing = 1, xmin = 0, xmax = 0, ymin = 0, ymax = 0;
  std::vector<int> nl, nz;                 /* [A2*P] */
  std::vector<double> dl, dz, theta;       /* [A2*P] */
  std::vector<int32_t> per_stack, first_lz;    /* [n_trk2d*P] */
  std::vector<int64_t> cum;                    /* [n_trk2d*P] first uid of the stack */
  /* outputs, one entry per 3D track in uid order (azim, xy, polar, z) */
  std::vector<int32_t> trk_azim, trk_polar, trk_xy, trk_2d, trk_lz;
  std::vector<int64_t> trk_next_fwd, trk_next_bwd;
  std::vector<uint8_t> trk_flags, trk_bc_fwd, trk_bc_bwd;
  std::vector<double> trk_phi, trk_theta, trk_start, trk_end, trk_l0;
  std::vector<double> z_mesh;
  /* explicit 3D segments (optional) */
  std::vector<double> seg_length;
  std::vector<int32_t> seg_fsr, seg_mat;
  std::vector<int64_t> trk_seg_offset;
  std::vector<double> fsr_volume;
  std::vector<int32_t> fsr_mat;
  std::vector<double> quad_weight, quad_sin_theta, quad_polar_spacing, quad_polar_weight;

  int A2() const { return g2->num_azim / 2; }
  int nx(int a) const { return g2->num_x[a]; }
  int ny(int a) const { return g2->num_y[a]; }
  int64_t id2(int a, int xy) const { return g2->first[a] + xy; }
  size_t ap(int a, int p) const { return (size_t)a * P + p; }
  size_t sp(int a, int xy, int p) const { return (size_t)id2(a, xy) * P + p; }
  /* link k of chain (a, x) is the 2D track xy = x + k * num_x (periodic-forward neighbour,
   * src/TrackGenerator.cpp:1171), the chain ends with the first track leaving through y_max */
  int chain_xy(int a, int x, int link) const { return x + link * nx(a); }
  int chain_len(int a, int x) const { int k = 0; while (chain_xy(a, x, k) < ny(a)) k++; return k + 1; }

  /* getFirst2DTrackLinkIndex, src/TrackGenerator3D.cpp:975-1083 */
  int first_link(const Chain& c, double* x1o, double* y1o, double* z1o, double* z2o) const {
    const int64_t t0 = id2(c.a, c.x);
    const double phi = g2->phi[c.a], cos_phi = cos(phi), sin_phi = sin(phi);
    const double x_start = g2->trk_start[2 * t0], y_start = g2->trk_start[2 * t0 + 1];
    const int n_l = nl[ap(c.a, c.p)], n_z = nz[ap(c.a, c.p)];
    const double d_z = dz[ap(c.a, c.p)], d_l = dl[ap(c.a, c.p)];
    const double width_x = xmax - xmin, width_y = ymax - ymin;
    double l_start = 0.0;
    if (c.p < P / 2 && c.lz < n_l) l_start = width_y / sin_phi - (c.lz + 0.5) * d_l;
    else if (c.p >= P / 2 && c.lz >= n_z) l_start = d_l * (c.lz - n_z + 0.5);
    double x_ext = x_start - xmin + l_start * cos_phi;
    double y_ext = y_start - ymin + l_start * sin_phi;
    bool nudged = false;
    if (fabs(l_start) > FLT_EPS_REF) {
      if (fabs(round(x_ext / width_x) * width_x - x_ext) < TINY_MOVE ||
          fabs(round(y_ext / width_y) * width_y - y_ext) < TINY_MOVE) {
        l_start += 10 * TINY_MOVE;
        x_ext = x_start - xmin + l_start * cos_phi;
        y_ext = y_start - ymin + l_start * sin_phi;
        nudged = true;
      }
    }
    const int link_index = abs((int)floor(x_ext / width_x));
    double x1 = (x_ext < 0.0) ? fmod(x_ext, width_x) + xmax : fmod(x_ext, width_x) + xmin;
    double y1 = fmod(y_ext, width_y) + ymin;
    double z1, z2;
    if (c.p < P / 2) {
      z1 = zmin + std::max(0., (c.lz - n_l + 0.5)) * d_z;
      z2 = zmax + std::min(0., (c.lz - n_z + 0.5)) * d_z;
    } else {
      z1 = zmax + std::min(0., (c.lz - n_z + 0.5)) * d_z;
      z2 = zmin + std::max(0., (c.lz - n_l + 0.5)) * d_z;
    }
    if (nudged) { x1 -= 10 * TINY_MOVE * cos_phi; y1 -= 10 * TINY_MOVE * sin_phi; }
    if (x1o) { *x1o = x1; *y1o = y1; *z1o = z1; *z2o = z2; }
    return link_index;
  }

  /* set3DTrackData, src/TrackGenerator3D.cpp:1099-1245.  count = true: registers every piece of
   * the chain track in its z-stack; otherwise stops at piece c.link and returns its end points.
   * Returns the number of pieces walked. */
  int walk(const Chain& c, bool count, double* se /* x1 y1 z1 x2 y2 z2 */) {
    const double th = theta[ap(c.a, c.p)];
    double x1 = 0, y1 = 0, z1 = 0, x2, y2, z2, z_end;
    int link = first_link(c, &x2, &y2, &z2, &z_end);
    const int first = link;
    const int K = chain_len(c.a, c.x);
    bool end_of_chain = false;
    while (!end_of_chain) {
      if (link >= K) break;                  /* defensive: the reference would read past the chain */
      const int xy = chain_xy(c.a, c.x, link);
      const int64_t t2 = id2(c.a, xy);
      const double phi = g2->phi[c.a];
      x1 = x2; y1 = y2; z1 = z2;
      double dl_xy;
      if (link == first) {
        const double dx = g2->trk_end[2 * t2] - x1, dy = g2->trk_end[2 * t2 + 1] - y1;
        dl_xy = sqrt(dx * dx + dy * dy);
      } else {
        const double dx = g2->trk_end[2 * t2] - g2->trk_start[2 * t2], dy = g2->trk_end[2 * t2 + 1] - g2->trk_start[2 * t2 + 1];
        dl_xy = sqrt(dx * dx + dy * dy);     /* Track::getLength */
      }
      double dl_z;
      if (c.p < P / 2) dl_z = (z_end - z1) * tan(th);
      else dl_z = (z1 - z_end) / tan(th - M_PI_2);
      const double d = std::min(dl_z, dl_xy);
      x2 = x1 + d * cos(phi);
      y2 = y1 + d * sin(phi);
      if (c.p < P / 2) z2 = z1 + d / tan(th);
      else z2 = z1 - d * tan(th - M_PI_2);
      if (fabs(x2 - x1) < TINY_MOVE || fabs(y2 - y1) < TINY_MOVE || fabs(z2 - z1) < TINY_MOVE) break;
      if (dl_z < dl_xy || xy >= ny(c.a) || c.link == link - first) end_of_chain = true;
      if (count) {
        const size_t s = sp(c.a, xy, c.p);
        if (per_stack[s] == 0) first_lz[s] = c.lz;
        per_stack[s]++;
      }
      link++;
      if (!end_of_chain) x2 = (c.a < g2->num_azim / 4) ? xmin : xmax;
    }
    if (se) {
      auto clampd = [](double v, double lo, double hi) { return std::max(lo, std::min(hi, v)); };
      se[0] = clampd(x1, xmin, xmax); se[1] = clampd(y1, ymin, ymax); se[2] = clampd(z1, zmin, zmax);
      se[3] = clampd(x2, xmin, xmax); se[4] = clampd(y2, ymin, ymax); se[5] = clampd(z2, zmin, zmax);
    }
    return link - first;
  }

  /* getNum3DTrackChainLinks, src/TrackGenerator3D.cpp:1821-1853 */
  int num_links(const Chain& c) const {
    const int first = first_link(c, nullptr, nullptr, nullptr, nullptr);
    int link = first;
    const int K = chain_len(c.a, c.x);
    while (true) {
      if (link >= K) break;
      const int xy = chain_xy(c.a, c.x, link);
      const size_t s = sp(c.a, xy, c.p);
      const int min_lz = first_lz[s], max_lz = per_stack[s] + min_lz - 1;
      if (c.p < P / 2 && c.lz > max_lz) break;
      else if (c.p >= P / 2 && c.lz < min_lz) break;
      link++;
      if (xy >= ny(c.a)) break;
    }
    return link - first;
  }

  /* convertTSItoTCI / convertTCItoTSI, src/TrackGenerator3D.cpp:1762-1813 */
  Chain to_chain(const Stack& s) const {
    Chain c;
    c.a = s.a; c.x = s.xy % nx(s.a); c.p = s.p;
    c.lz = first_lz[sp(s.a, s.xy, s.p)] + s.z;
    c.link = 0;
    c.link = s.xy / nx(s.a) - first_link(c, nullptr, nullptr, nullptr, nullptr);   /* Track::getLinkIndex - first link */
    return c;
  }
  /* false when the link falls off the chain (the reference would index past its array) */
  bool to_stack(const Chain& c, Stack* s) const {
    const int link = first_link(c, nullptr, nullptr, nullptr, nullptr) + c.link;
    if (link < 0 || link >= chain_len(c.a, c.x)) return false;
    s->a = c.a; s->xy = chain_xy(c.a, c.x, link); s->p = c.p;
    s->z = c.lz - first_lz[sp(s->a, s->xy, s->p)];
    return true;
  }
  bool in_stack(const Stack& s) const { return s.z >= 0 && s.z < per_stack[sp(s.a, s.xy, s.p)]; }
  int64_t uid(const Stack& s) const { return cum[sp(s.a, s.xy, s.p)] + s.z; }

  /* 2D neighbours of track (a, xy) (src/TrackGenerator.cpp:1163-1217), as xy indices in angle a / ac */
  int prdc_fwd_xy(int a, int i) const { return i < ny(a) ? i + nx(a) : i - ny(a); }
  int refl_fwd_xy(int a, int i) const { return i < ny(a) ? i + nx(a) : (nx(a) + ny(a)) - (i - ny(a)) - 1; }
  int prdc_bwd_xy(int a, int i) const { return i < nx(a) ? i + ny(a) : i - nx(a); }
  int refl_bwd_xy(int a, int i) const { return i < nx(a) ? nx(a) - i - 1 : i - nx(a); }

  /* setLinkingTracks, src/TrackGenerator3D.cpp:1864-2374: the track the outgoing (or incoming)
   * end of (s, c) connects to, whether it is entered in its forward direction, and the BC */
  bool link_of(const Stack& s, const Chain& c, bool outgoing, int64_t* next_uid, bool* next_is_fwd, int* bc_out) const {
    const Gen& g = *g2;
    const int a = c.a, A = g.num_azim;
    const int64_t t2 = id2(s.a, s.xy);
    Chain next = {c.a, c.x, c.p, c.lz, 0}, prdc = next;
    const int n_z = nz[ap(a, c.p)], n_l = nl[ap(a, c.p)], lz = c.lz;
    const int ac = A / 2 - a - 1, pc = P - c.p - 1;
    const int links = num_links(c);
    bool next_fwd = outgoing;
    int bc = outgoing ? g.trk_bc_fwd[t2] : g.trk_bc_bwd[t2];
    const int bc_xy_fwd = g.trk_bc_fwd[t2], bc_xy_bwd = g.trk_bc_bwd[t2];
    const int bymin = g.bc[2], bymax = g.bc[3];
    auto periodic_like = [](int b) { return b == BC_PERIODIC || b == BC_INTERFACE; };
    auto last = [&](Chain& q) { q.link = num_links(q) - 1; };
    /* a track through a z boundary that also sits on an x boundary of its 2D track ("double reflection") */
    auto double_reflection = [&](int bc_xy) {
      Stack sp_;
      if (!to_stack(prdc, &sp_)) return;
      if (sp_.xy != s.xy) {
        if (!periodic_like(bc_xy)) next.a = ac;
        if (bc_xy == BC_INTERFACE && bc != BC_VACUUM) bc = BC_INTERFACE;
        else if (bc_xy == BC_VACUUM) bc = BC_VACUUM;
      }
    };
    const bool up = c.p < P / 2;
    /* which face the end sits on: the last piece of the chain track going out / the first coming in */
    const bool at_last = c.link == links - 1 && outgoing, at_first = c.link == 0 && !outgoing;
    /* for upward tracks the chain track leaves through z_max when lz >= nz, else through y_max, and
     * enters through z_min when lz < nl, else through y_min; downward tracks mirror this in z */
    const bool z_out = up ? lz >= n_z : lz < n_l;
    const bool z_in = up ? lz < n_l : lz >= n_z;
    if (at_last && z_out) {                              /* SURFACE_Z_MAX (up) / SURFACE_Z_MIN (down) */
      bc = up ? bc_zmax : bc_zmin;
      const int lz_prdc = up ? lz - n_z : lz + n_z;
      const int lz_refl = up ? n_l + 2 * n_z - lz - 1 : n_l - lz - 1;
      prdc.lz = lz_prdc;
      if (periodic_like(bc)) next.lz = lz_prdc;
      else { next.p = pc; next.lz = lz_refl; }
      double_reflection(bc_xy_fwd);
    } else if (at_first && z_in) {                       /* SURFACE_Z_MIN (up) / SURFACE_Z_MAX (down) */
      bc = up ? bc_zmin : bc_zmax;
      const int lz_prdc = up ? lz + n_z : lz - n_z;
      const int lz_refl = up ? n_l - lz - 1 : n_l + 2 * n_z - lz - 1;
      prdc.lz = lz_prdc; last(prdc);
      if (periodic_like(bc)) { next.lz = lz_prdc; last(next); }
      else { next.p = pc; next.lz = lz_refl; last(next); }
      double_reflection(bc_xy_bwd);
    } else if (at_first) {                               /* SURFACE_Y_MIN */
      const int lz2 = up ? lz - n_l : lz + n_l;
      prdc.lz = lz2; prdc.x = prdc_bwd_xy(s.a, s.xy) % nx(a); last(prdc);
      next.lz = lz2;
      if (periodic_like(bymin)) { next.x = prdc_bwd_xy(s.a, s.xy) % nx(a); last(next); }
      else { next.a = ac; next.x = refl_bwd_xy(s.a, s.xy) % nx(a); next.p = pc; next_fwd = true; }
    } else if (at_last) {                                /* SURFACE_Y_MAX */
      const int lz2 = up ? n_l + lz : lz - n_l;
      prdc.lz = lz2; prdc.x = prdc_fwd_xy(s.a, s.xy) % nx(a);
      next.lz = lz2;
      if (periodic_like(bymax)) next.x = prdc_fwd_xy(s.a, s.xy) % nx(a);
      else { next.a = ac; next.x = refl_fwd_xy(s.a, s.xy) % nx(a); next.p = pc; last(next); next_fwd = false; }
    } else if (outgoing) {                               /* an x face, going on along the chain */
      next.link = c.link + 1;
      if (!periodic_like(bc_xy_fwd)) next.a = ac;
    } else {
      next.link = c.link - 1;
      if (!periodic_like(bc_xy_bwd)) next.a = ac;
    }
    *next_is_fwd = next_fwd;
    *bc_out = bc;
    Stack sn;
    if (!to_stack(next, &sn) || !in_stack(sn)) { *next_uid = -1; return false; }
    *next_uid = uid(sn);
    return true;
  }
};

/* TraverseSegments::findMeshIndex, src/TraverseSegments.cpp:926-956 */
inline int find_mesh_index(const double* v, int size, double val, int sign) {
  int imin = 0, imax = size - 1;
  while (imax - imin > 1) {
    const int imid = (imin + imax) / 2;
    if (val > v[imid]) imin = imid;
    else if (val < v[imid]) imax = imid;
    else return sign > 0 ? imid : imid - 1;
  }
  return imin;
}

/* TraverseSegments::traceSegmentsOTF (src/TraverseSegments.cpp:304-505) for one 3D track over the
 * segments of its 2D track and the global axial mesh; calls emit(length, fsr3d, mat) */
template <typename Emit>
void trace_otf(const Gen& g, const Gen3& h, int64_t t2, double x0, double z0, double theta, Emit emit) {
  const double phi = g.trk_phi[t2], cos_phi = cos(phi);
  const double cos_theta = cos(theta), sin_theta = sin(theta);
  const int sign = (cos_theta > 0) - (cos_theta < 0);
  double z = z0;
  double start_dist = (x0 - g.trk_start[2 * t2]) / cos_phi;
  const int64_t s0 = g.trk_seg_offset[t2], s1 = g.trk_seg_offset[t2 + 1];
  int64_t s = s0;
  for (; s < s1; s++) {
    if (start_dist > g.seg_length[s]) start_dist -= g.seg_length[s];
    else break;
  }
  if (s == s1) return;
  const double* mesh = h.z_mesh.data();
  const int nf = h.n_axial;
  int zi = find_mesh_index(mesh, nf + 1, z, sign);
  for (; s < s1; s++) {
    double remaining = g.seg_length[s] - start_dist;
    start_dist = 0;
    bool complete = false;
    while (remaining > 0) {
      const double z_dist = sign > 0 ? (mesh[zi + 1] - z) / cos_theta : (mesh[zi] - z) / cos_theta;
      const double seg_dist = remaining / sin_theta;
      double d2, d3; int zmove;
      if (z_dist <= seg_dist) { d2 = z_dist * sin_theta; d3 = z_dist; zmove = sign; }
      else { d2 = remaining; d3 = seg_dist; zmove = 0; }
      if (d3 > TINY_MOVE) emit(d3, (int32_t)((int64_t)g.seg_fsr[s] * nf + zi), g.seg_mat[s]);
      z += d3 * cos_theta;
      remaining -= d2;
      zi += zmove;
      if (zi < 0 || zi >= nf) { zi = zi < 0 ? 0 : nf - 1; complete = true; break; }
    }
    if (complete) break;
  }
}

/* polar angles of one octant before the correction: 0 TY, 1 equal angle, 2 Gauss-Legendre, 3 equal weight */
int polar_angles(int kind, int P, std::vector<double>& th, std::string& err) {
  const int P2 = P / 2;
  th.assign(P2, 0.);
  if (kind == 0) {
    if (P == 2) th[0] = asin(0.798184);
    else if (P == 4) { th[0] = asin(0.363900); th[1] = asin(0.899900); }
    else if (P == 6) { th[0] = asin(0.166648); th[1] = asin(0.537707); th[2] = asin(0.932954); }
    else { err = "TY quadrature supports 2, 4 or 6 polar angles"; return 1; }
  } else if (kind == 1) {
    const double dth = M_PI / P;
    double ta = 0.;
    for (int p = 0; p < P2; p++) { const double tb = ta + dth; th[p] = acos(0.5 * (cos(ta) + cos(tb))); ta = tb; }
  } else if (kind == 2) {
    /* positive roots of the Legendre polynomial of degree P, ascending (GLPolarQuad::getLegendreRoots,
     * src/Quadrature.cpp:1134-1240, solved here by Newton's method on the three-term recurrence) */
    std::vector<double> roots;
    for (int i = 0; i < P2; i++) {
      double x = cos(M_PI * (i + 0.75) / (P + 0.5));
      for (int it = 0; it < 100; it++) {
        double p0 = 1., p1 = x;
        for (int k = 2; k <= P; k++) { const double pk = ((2. * k - 1.) * x * p1 - (k - 1.) * p0) / k; p0 = p1; p1 = pk; }
        const double dp = P * (x * p1 - p0) / (x * x - 1.);
        const double dx = p1 / dp;
        x -= dx;
        if (fabs(dx) < 1e-16) break;
      }
      roots.push_back(x);
    }
    std::sort(roots.begin(), roots.end());
    for (int p = 0; p < P2; p++) th[p] = acos(roots[p]);
  } else if (kind == 3) {
    double ca = 1.;
    for (int p = 0; p < P2; p++) { const double cb = ca - 1. / P2; th[p] = acos(0.5 * (ca + cb)); ca = cb; }
  } else { err = "unknown polar quadrature"; return 1; }
  return 0;
}

/* polar weights of azimuthal angle a from its (corrected) polar angles */
void polar_weights(int kind, int P, const double* th_uncorrected, const double* th, double* w) {
  const int P2 = P / 2;
  if (kind == 0) {
    if (P == 2) w[0] = 0.5;
    else if (P == 4) { w[0] = 0.212854 / 2.0; w[1] = 0.787146 / 2.0; }
    else { w[0] = 0.046233 / 2.0; w[1] = 0.283619 / 2.0; w[2] = 0.670148 / 2.0; }
  } else if (kind == 2) {
    for (int p = 0; p < P2; p++) {      /* GLPolarQuad::getGLWeights on the uncorrected roots, halved (:1051-1076, 1247-1258) */
      const double x = cos(th_uncorrected[p]);
      double p0 = 1., p1 = x;
      for (int k = 2; k <= P - 1; k++) { const double pk = ((2. * k - 1.) * x * p1 - (k - 1.) * p0) / k; p0 = p1; p1 = pk; }
      const double pm1 = P - 1 == 0 ? 1. : p1;
      w[p] = -(2 * x * x - 2) / ((double)P * P * pm1 * pm1) / 2.0;
    }
  } else {                               /* equal angle / equal weight: from the current angles (:1425-1457, 1520-1553) */
    for (int p = 0; p < P2; p++) {
      const double y1 = (p < P2 - 1) ? 0.5 * (cos(th[p]) - cos(th[p + 1])) : cos(th[p]);
      const double y2 = (p >= 1) ? 0.5 * (cos(th[p - 1]) - cos(th[p])) : 1.0 - cos(th[p]);
      w[p] = (y1 + y2) / 2.0;
    }
  }
}

int build3d(Gen& g, Gen3& h, int polar_quad, int expand) {
  h.g2 = &g;
  const int A = g.num_azim, A2 = A / 2, A4 = A / 4, P = g.num_polar, P2 = P / 2;
  h.P = P;
  h.xmin = g.xmin; h.xmax = g.xmin + g.nx * g.px; h.ymin = g.ymin; h.ymax = g.ymin + g.ny * g.py;
  const double width_y = h.ymax - h.ymin, width_z = h.zmax - h.zmin;
  if (!(width_z > 0)) { g.error = "z_max must exceed z_min"; return 1; }
  if (!(h.z_spacing > 0)) { g.error = "z spacing must be positive"; return 1; }
  if (h.n_axial < 1) { g.error = "n_axial must be at least 1"; return 1; }
  if ((double)g.n_fsrs * h.n_axial > (double)INT32_MAX) { g.error = "too many 3D FSRs for 32-bit ids"; return 1; }

  /* ---- polar angles, their correction to the cyclic laydown, spacings (:750-809) ---- */
  std::vector<double> th0;
  if (polar_angles(polar_quad, P, th0, g.error)) return 1;
  h.nl.assign((size_t)A2 * P, 0); h.nz.assign((size_t)A2 * P, 0);
  h.dl.assign((size_t)A2 * P, 0.); h.dz.assign((size_t)A2 * P, 0.); h.theta.assign((size_t)A2 * P, 0.);
  std::vector<double> polar_spacing((size_t)A2 * P, 0.), polar_weight((size_t)A2 * P, 0.);
  for (int a = 0; a < A4; a++) {
    const double phi = g.phi[a];
    std::vector<double> thc(P2), w(P2);
    for (int p = 0; p < P2; p++) {
      double theta = th0[p];
      const double length = width_y / sin(phi);
      int n_l = (int)ceil(length * tan(M_PI_2 - theta) / h.z_spacing);
      int n_z = (int)ceil(width_z * n_l * tan(theta) / length);
      const double d_l = width_y / (sin(phi) * n_l), d_z = width_z / n_z;
      theta = atan(d_l / d_z);
      thc[p] = theta;
      const int as[2] = {a, A2 - a - 1}, ps[2] = {p, P - p - 1};
      for (int ia = 0; ia < 2; ia++)
        for (int ip = 0; ip < 2; ip++) {
          const size_t k = h.ap(as[ia], ps[ip]);
          h.nl[k] = n_l; h.nz[k] = n_z; h.dl[k] = d_l; h.dz[k] = d_z;
          h.theta[k] = ip == 0 ? theta : M_PI - theta;
          polar_spacing[k] = d_z * sin(theta);
        }
    }
    polar_weights(polar_quad, P, th0.data(), thc.data(), w.data());
    for (int p = 0; p < P2; p++) {
      const int as[2] = {a, A2 - a - 1}, ps[2] = {p, P - p - 1};
      for (int ia = 0; ia < 2; ia++)
        for (int ip = 0; ip < 2; ip++) polar_weight[h.ap(as[ia], ps[ip])] = w[p];
    }
  }
  /* total weights, 3D form (Quadrature::precomputeWeights, :727-741) */
  h.quad_weight.assign((size_t)A2 * P, 0.); h.quad_sin_theta.assign((size_t)A2 * P, 0.);
  for (int a = 0; a < A2; a++)
    for (int p = 0; p < P; p++) {
      const size_t k = h.ap(a, p);
      const int pp = p < P2 ? p : P - 1 - p;
      h.quad_sin_theta[k] = sin(h.theta[h.ap(a, pp)]);
      h.quad_weight[k] = 2.0 * M_PI * g.azim_weight[a] * g.azim_spacing[a] * polar_weight[k] * polar_spacing[k];
    }
  h.quad_polar_spacing = polar_spacing;
  h.quad_polar_weight = polar_weight;

  /* ---- z-stacks: walk every chain track once (getCycleTrackData, :934-964) ---- */
  const int64_t nt2 = g.first[A2];
  h.per_stack.assign((size_t)nt2 * P, 0); h.first_lz.assign((size_t)nt2 * P, -1); h.cum.assign((size_t)nt2 * P, 0);
  for (int a = 0; a < A2; a++)
#pragma omp parallel for schedule(dynamic)
    for (int x = 0; x < g.num_x[a]; x++)
      for (int p = 0; p < P; p++) {
        const int n = h.nl[h.ap(a, p)] + h.nz[h.ap(a, p)];
        for (int lz = 0; lz < n; lz++) {
          Chain c = {a, x, p, lz, -1};
          h.walk(c, true, nullptr);
        }
      }
  int64_t n3 = 0;
  for (int a = 0; a < A2; a++)
    for (int i = 0; i < g.num_x[a] + g.num_y[a]; i++)
      for (int p = 0; p < P; p++) { h.cum[h.sp(a, i, p)] = n3; n3 += h.per_stack[h.sp(a, i, p)]; }
  if (n3 > (int64_t)INT32_MAX) { g.error = "too many 3D tracks"; return 1; }

  /* ---- per-track data (getTrackOTF, :1697-1745) ---- */
  h.trk_azim.resize(n3); h.trk_polar.resize(n3); h.trk_xy.resize(n3); h.trk_2d.resize(n3); h.trk_lz.resize(n3);
  h.trk_next_fwd.resize(n3); h.trk_next_bwd.resize(n3); h.trk_flags.resize(n3);
  h.trk_bc_fwd.resize(n3); h.trk_bc_bwd.resize(n3); h.trk_phi.resize(n3); h.trk_theta.resize(n3);
  h.trk_start.resize(3 * n3); h.trk_end.resize(3 * n3); h.trk_l0.resize(n3);
  int bad = 0;
  for (int a = 0; a < A2; a++) {
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < g.num_x[a] + g.num_y[a]; i++)
      for (int p = 0; p < P; p++) {
        const int n = h.per_stack[h.sp(a, i, p)];
        for (int z = 0; z < n; z++) {
          const Stack s = {a, i, p, z};
          const int64_t u = h.uid(s);
          Chain c = h.to_chain(s);
          double se[6];
          h.walk(c, false, se);
          const int64_t t2 = h.id2(a, i);
          h.trk_azim[u] = a; h.trk_polar[u] = p; h.trk_xy[u] = i; h.trk_2d[u] = (int32_t)t2; h.trk_lz[u] = c.lz;
          h.trk_phi[u] = g.phi[a]; h.trk_theta[u] = h.theta[h.ap(a, p)];
          for (int k = 0; k < 3; k++) { h.trk_start[3 * u + k] = se[k]; h.trk_end[3 * u + k] = se[3 + k]; }
          h.trk_l0[u] = (se[0] - g.trk_start[2 * t2]) / cos(g.phi[a]);
          int64_t nf = -1, nb = -1; bool ff = true, bf = false; int bcf = 0, bcb = 0;
          const bool okf = h.link_of(s, c, true, &nf, &ff, &bcf);
          const bool okb = h.link_of(s, c, false, &nb, &bf, &bcb);
          if ((!okf && bcf != BC_VACUUM) || (!okb && bcb != BC_VACUUM)) {
#pragma omp atomic
            bad++;
          }
          h.trk_next_fwd[u] = nf; h.trk_next_bwd[u] = nb;
          h.trk_flags[u] = (ff ? 1 : 0) | (bf ? 2 : 0);
          h.trk_bc_fwd[u] = (uint8_t)bcf; h.trk_bc_bwd[u] = (uint8_t)bcb;
        }
      }
  }
  if (bad) { g.error = "3D track linking failed for " + std::to_string(bad) + " track ends"; return 1; }

  /* ---- axial mesh, FSR materials ---- */
  h.z_mesh.resize(h.n_axial + 1);
  for (int k = 0; k <= h.n_axial; k++) h.z_mesh[k] = h.zmin + width_z * k / h.n_axial;
  h.z_mesh[h.n_axial] = h.zmax;
  const int64_t nf3 = g.n_fsrs * h.n_axial;
  h.fsr_mat.resize(nf3);
  for (int64_t r = 0; r < g.n_fsrs; r++)
    for (int k = 0; k < h.n_axial; k++) h.fsr_mat[r * h.n_axial + k] = g.fsr_mat[r];
  h.fsr_volume.assign(nf3, 0.);
  if (!expand) return 0;

  /* ---- explicit 3D segments and volumes (SegmentationKernel / VolumeKernel, src/MOCKernel.cpp:80-162) ---- */
  h.trk_seg_offset.assign(n3 + 1, 0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t u = 0; u < n3; u++) {
    int64_t n = 0;
    trace_otf(g, h, h.trk_2d[u], h.trk_start[3 * u], h.trk_start[3 * u + 2], h.trk_theta[u],
              [&](double, int32_t, int32_t) { n++; });
    h.trk_seg_offset[u + 1] = n;
  }
  for (int64_t u = 0; u < n3; u++) h.trk_seg_offset[u + 1] += h.trk_seg_offset[u];
  const int64_t ns = h.trk_seg_offset[n3];
  h.seg_length.resize(ns); h.seg_fsr.resize(ns); h.seg_mat.resize(ns);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t u = 0; u < n3; u++) {
    int64_t o = h.trk_seg_offset[u];
    trace_otf(g, h, h.trk_2d[u], h.trk_start[3 * u], h.trk_start[3 * u + 2], h.trk_theta[u],
              [&](double len, int32_t fsr, int32_t mat) { h.seg_length[o] = len; h.seg_fsr[o] = fsr; h.seg_mat[o] = mat; o++; });
  }
  for (int64_t u = 0; u < n3; u++) {
    const int a = h.trk_azim[u], p = h.trk_polar[u];
    const double w = g.azim_weight[a] * g.azim_spacing[a] * polar_spacing[h.ap(a, p)] * polar_weight[h.ap(a, p)];
    for (int64_t s = h.trk_seg_offset[u]; s < h.trk_seg_offset[u + 1]; s++) h.fsr_volume[h.seg_fsr[s]] += w * h.seg_length[s];
  }
  return 0;
}

}  // namespace

/* ------------------------------- C ABI ------------------------------------ */
extern "C" {

struct b200_trackgen { Gen g; Gen3 h; bool is3d = false; };

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

/* 3D tracks (z-stacks, TrackGenerator3D) over the same 2D lattice extruded between z_min and z_max
 * in n_axial equal layers.  expand != 0 also produces the explicit 3D segments and FSR volumes on the
 * host (tests, small decks); otherwise the caller hands the 2D segments, the axial mesh and the
 * per-track start data to the device tracer (b200_upload_tracks_otf). */
b200_trackgen* b200_trackgen_create_3d(int nx, int ny, double pitch_x, double pitch_y, double xmin,
                                       double ymin, const int32_t* cell_type, const void* types, int n_types,
                                       int bc_xmin, int bc_xmax, int bc_ymin, int bc_ymax, int num_azim,
                                       double spacing, int num_polar, int polar_quad, int num_threads,
                                       double zmin, double zmax, int bc_zmin, int bc_zmax, int n_axial,
                                       double z_spacing, int expand, int* status) {
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
  h->is3d = true;
  h->h.zmin = zmin; h->h.zmax = zmax; h->h.bc_zmin = bc_zmin; h->h.bc_zmax = bc_zmax;
  h->h.n_axial = n_axial; h->h.z_spacing = z_spacing;
  /* the 2D stage only needs a valid polar set for its own (unused) 2D weights */
  const int saved_polar = g.num_polar;
  g.num_polar = 2;
  int rc = build(g, 0);
  g.num_polar = saved_polar;
  if (rc == 0) rc = build3d(g, h->h, polar_quad, expand);
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
  std::string n(name);
#define OUT(field)                                                             \
  if (n == #field) {                                                           \
    if (dst && !g.field.empty()) memcpy(dst, g.field.data(), g.field.size() * sizeof(g.field[0])); \
    return (int64_t)g.field.size();                                            \
  }
  if (h->is3d) {
    /* the 3D generator answers the plain names with the 3D data; the 2D stage is reachable as *2d* */
    {
      Gen3& g = h->h;
      OUT(seg_length) OUT(seg_fsr) OUT(seg_mat) OUT(trk_seg_offset) OUT(trk_next_fwd) OUT(trk_next_bwd)
      OUT(trk_azim) OUT(trk_polar) OUT(trk_xy) OUT(trk_2d) OUT(trk_lz) OUT(trk_flags) OUT(trk_bc_fwd) OUT(trk_bc_bwd)
      OUT(trk_phi) OUT(trk_theta) OUT(trk_start) OUT(trk_end) OUT(trk_l0) OUT(z_mesh) OUT(fsr_volume) OUT(fsr_mat)
      OUT(quad_weight) OUT(quad_sin_theta) OUT(quad_polar_spacing) OUT(quad_polar_weight)
      OUT(nl) OUT(nz) OUT(dl) OUT(dz) OUT(theta) OUT(per_stack) OUT(first_lz) OUT(cum)
    }
    Gen& g = h->g;
    OUT(quad_azim_spacing) OUT(quad_azim_weight)
    if (n.compare(0, 6, "seg2d_") == 0 || n.compare(0, 6, "trk2d_") == 0 || n.compare(0, 6, "fsr2d_") == 0)
      n = n.substr(0, 3) + n.substr(5);
    else return -1;
  }
  Gen& g = h->g;
  OUT(seg_length) OUT(seg_fsr) OUT(seg_mat) OUT(trk_seg_offset) OUT(trk_next_fwd) OUT(trk_next_bwd)
  OUT(trk_azim) OUT(trk_polar) OUT(trk_xy) OUT(trk_flags) OUT(trk_bc_fwd) OUT(trk_bc_bwd) OUT(trk_phi)
  OUT(trk_theta) OUT(trk_start) OUT(trk_end) OUT(quad_weight) OUT(quad_sin_theta) OUT(fsr_volume) OUT(fsr_mat)
  OUT(quad_azim_spacing) OUT(quad_azim_weight) OUT(quad_polar_spacing) OUT(quad_polar_weight)
  OUT(seg_surf_fwd) OUT(seg_surf_bwd) OUT(seg_cmfd_fwd) OUT(seg_cmfd_bwd) OUT(fsr_cell)
#undef OUT
  return -1;
}

}  // extern "C"
