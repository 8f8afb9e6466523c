/*
 * cmfd_impl.cuh - host side of the device CMFD (kernels: cmfd.cuh).  Included at the end of
 * b200moc.cu.  Holds the restatement of the reference's current-splitting rules
 * (Cmfd::getVertexSplitSurfaces / getEdgeSplitSurfaces, src/Cmfd.cpp:2348-2480), the tables
 * derived from them, the device state and the launch sequence of one solve.
 */
#pragma once

struct b200_cmfd {
  b200_cmfd_config cfg;
  int64_t n_cells = 0;
  bool configured = false, have_stencils = false, have_interp = false;
  bool in_loop = true;               /* the fused source iteration runs the CMFD solve (b200_cmfd_set_in_loop) */
  std::vector<double> h_wx, h_wy, h_wz;
  DevBuf<double> wx, wy, wz, azim_w, sin_theta, polar_w;
  DevBuf<int32_t> group_idx, cell_fsrs, fsr_cell, nbr, sv_src, se_src, st_cell, st_n;
  DevBuf<int64_t> cell_fsr_off, sv_off, se_off, st_off;
  DevBuf<double> st_w, st_own, ax_interp;
  DevBuf<double> rxn, volc, dift, xs_t, xs_nf, xs_chi, xs_s, old_flux, new_flux, dcoef, old_corr;
  DevBuf<double> diag, off, ain, mm, B, SO, SN, partials, cs, dq, qd;
  DevBuf<int32_t> slot_cell;
  int grid_threads = CMFD_GRID_THREADS;
  DevBuf<int> ci;
  int eigen_mode = 0, eigen_blocks = 1;     /* 0 one CTA, 1 cooperative grid, 2 one thread-block cluster */
  size_t eigen_smem = 0;
  int cluster_slots = 0, cluster_all_smem = 0, cluster_threads = CMFD_CLUSTER_THREADS;
  DevBuf<int32_t> nb_loc;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void release() {
    if (ev0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); ev0 = ev1 = nullptr; }
    wx.release(); wy.release(); wz.release(); azim_w.release(); sin_theta.release(); polar_w.release();
    group_idx.release(); cell_fsrs.release(); fsr_cell.release(); nbr.release(); sv_src.release(); se_src.release();
    st_cell.release(); st_n.release(); cell_fsr_off.release(); sv_off.release(); se_off.release(); st_off.release();
    st_w.release(); st_own.release(); ax_interp.release();
    rxn.release(); volc.release(); dift.release(); xs_t.release(); xs_nf.release(); xs_chi.release(); xs_s.release();
    old_flux.release(); new_flux.release(); dcoef.release(); old_corr.release();
    diag.release(); off.release(); ain.release(); mm.release(); B.release(); SO.release(); SN.release();
    partials.release(); cs.release(); ci.release(); nb_loc.release(); dq.release(); qd.release(); slot_cell.release();
  }
};

/* the eigenvalue kernel is compiled for the usual CMFD group counts (registers instead of loops), 0 = any */
typedef void (*cmfd_eigen_fn)(CmfdArgs, double);
template <int MODE>
static cmfd_eigen_fn cmfd_pick_eigen(int ncg) {
  switch (ncg) {
    case 1: return cmfd_eigen_kernel<MODE, 1>;
    case 2: return cmfd_eigen_kernel<MODE, 2>;
    case 4: return cmfd_eigen_kernel<MODE, 4>;
    case 7: return cmfd_eigen_kernel<MODE, 7>;
    default: return cmfd_eigen_kernel<MODE, 0>;
  }
}

typedef void (*cmfd_cluster_fn)(CmfdArgs, CmfdClusterArgs, double);
static cmfd_cluster_fn cmfd_pick_cluster(int ncg) {
  switch (ncg) {
    case 1: return cmfd_eigen_cluster_kernel<1>;
    case 2: return cmfd_eigen_cluster_kernel<2>;
    case 4: return cmfd_eigen_cluster_kernel<4>;
    case 7: return cmfd_eigen_cluster_kernel<7>;
    default: return nullptr;
  }
}

static void cmfd_destroy(b200_solver* s) {
  if (s->cmfd == nullptr) return;
  s->cmfd->release();
  delete s->cmfd;
  s->cmfd = nullptr;
}

/* ---- surfaces <-> directions (Cmfd::convertSurfaceToDirection / convertDirectionToSurface, Cmfd.cpp:5201-5290) ---- */
static void cmfd_surface_to_direction(int surface, int d[3]) {
  d[0] = d[1] = d[2] = 0;
  if (surface < CMFD_NF) {
    d[surface % 3] = 2 * (surface / 3) - 1;
  } else if (surface < CMFD_NFE) {
    surface -= CMFD_NF;
    const int skipped = 2 - surface / 4;
    surface %= 4;
    const int ind[2] = {surface % 2, surface / 2};
    int n = 0;
    for (int i = 0; i < 3; i++)
      if (i != skipped) d[i] = 2 * ind[n++] - 1;
  } else {
    surface -= CMFD_NFE;
    d[0] = 2 * (surface / 4) - 1;
    d[1] = 2 * ((surface / 2) % 2) - 1;
    d[2] = 2 * (surface % 2) - 1;
  }
}

static int cmfd_direction_to_surface(const int d[3]) {
  const int crossings = std::abs(d[0]) + std::abs(d[1]) + std::abs(d[2]);
  int surface = 0;
  if (crossings == 1) {
    for (int i = 0; i < 3; i++) surface += std::abs(d[i]) * (3 * ((d[i] + 1) / 2) + i);
  } else if (crossings == 2) {
    surface = CMFD_NF;
    int i1, i2;
    if (d[0] == 0) { i1 = d[1]; i2 = d[2]; surface += 8; }
    else if (d[1] == 0) { i1 = d[0]; i2 = d[2]; surface += 4; }
    else { i1 = d[0]; i2 = d[1]; }
    surface += 2 * ((i2 + 1) / 2) + (i1 + 1) / 2;
  } else {
    surface = CMFD_NFE + 4 * ((d[0] + 1) / 2) + 2 * ((d[1] + 1) / 2) + (d[2] + 1) / 2;
  }
  return surface;
}

struct CmfdMesh { int nx, ny, nz; int bc[6]; };

/* Cmfd::getCellNext(cell, face) on the whole mesh (Cmfd.cpp:2492-2580) */
static int cmfd_cell_next(const CmfdMesh& m, int cell, int face) {
  const int x = (cell % (m.nx * m.ny)) % m.nx, y = (cell % (m.nx * m.ny)) / m.nx, z = cell / (m.nx * m.ny);
  const bool per = m.bc[face] == B200_BC_PERIODIC;
  switch (face) {
    case 0: return x != 0 ? cell - 1 : (per ? cell + (m.nx - 1) : -1);
    case 1: return y != 0 ? cell - m.nx : (per ? cell + m.nx * (m.ny - 1) : -1);
    case 2: return z != 0 ? cell - m.nx * m.ny : (per ? cell + m.nx * m.ny * (m.nz - 1) : -1);
    case 3: return x != m.nx - 1 ? cell + 1 : (per ? cell - (m.nx - 1) : -1);
    case 4: return y != m.ny - 1 ? cell + m.nx : (per ? cell - m.nx * (m.ny - 1) : -1);
    default: return z != m.nz - 1 ? cell + m.nx * m.ny : (per ? cell - m.nx * m.ny * (m.nz - 1) : -1);
  }
}

/* where the current through an edge or a vertex of `cell` goes: half (a third) onto each face of the cell that
 * touches it and onto the matching surface of the cell behind that face - or back onto the cell itself at a
 * reflective boundary, nowhere at a vacuum boundary (Cmfd.cpp:2348-2480) */
static int cmfd_split_targets(const CmfdMesh& m, int cell, int surface, int out[6]) {
  const int idx[3] = {(cell % (m.nx * m.ny)) % m.nx, (cell % (m.nx * m.ny)) / m.nx, cell / (m.nx * m.ny)};
  const int lim[3] = {m.nx, m.ny, m.nz};
  int d[3];
  cmfd_surface_to_direction(surface, d);
  int n = 0;
  for (int i = 0; i < 3; i++) {
    if (d[i] == 0) continue;
    int partial_d[3] = {0, 0, 0}, rest_d[3] = {d[0], d[1], d[2]};
    partial_d[i] = d[i];
    rest_d[i] = 0;
    const int partial = cmfd_direction_to_surface(partial_d);
    const int rest = cmfd_direction_to_surface(rest_d);      /* an edge for a vertex, the other face for an edge */
    out[n++] = cell * CMFD_NS + partial;
    const int next = cmfd_cell_next(m, cell, partial);
    if ((idx[i] == 0 && d[i] == -1) || (idx[i] == lim[i] - 1 && d[i] == +1)) {
      if (m.bc[partial] == B200_BC_REFLECTIVE) out[n++] = cell * CMFD_NS + rest;
      else if (m.bc[partial] == B200_BC_PERIODIC) out[n++] = next * CMFD_NS + rest;
    } else {
      out[n++] = next * CMFD_NS + rest;
    }
  }
  return n;
}

extern "C" int b200_cmfd_split_targets(int32_t num_x, int32_t num_y, int32_t num_z, const int32_t* boundaries,
                                       int32_t cell, int32_t surface, int32_t* targets, int32_t* num_targets) {
  if (!boundaries || !targets || !num_targets) return fail("b200_cmfd_split_targets: null argument");
  if (num_x < 1 || num_y < 1 || num_z < 1 || cell < 0 || cell >= num_x * num_y * num_z)
    return fail("b200_cmfd_split_targets: cell %d outside the %d x %d x %d mesh", cell, num_x, num_y, num_z);
  if (surface < CMFD_NF || surface >= CMFD_NS)
    return fail("b200_cmfd_split_targets: surface %d is not an edge or a vertex", surface);
  CmfdMesh m{num_x, num_y, num_z, {0, 0, 0, 0, 0, 0}};
  for (int i = 0; i < 6; i++) m.bc[i] = boundaries[i];
  int out[6];
  *num_targets = cmfd_split_targets(m, cell, surface, out);
  for (int i = 0; i < *num_targets; i++) targets[i] = out[i];
  return 0;
}

/* invert "source splits onto targets" into CSR lists per destination (cell, surface < dest_per_cell) */
static void cmfd_build_split_table(const CmfdMesh& m, int first_src, int last_src, int dest_per_cell,
                                   std::vector<int64_t>& off, std::vector<int32_t>& src) {
  const int64_t n_cells = (int64_t)m.nx * m.ny * m.nz;
  off.assign(n_cells * dest_per_cell + 1, 0);
  int t[6];
  for (int pass = 0; pass < 2; pass++) {
    std::vector<int64_t> fill;
    if (pass == 1) {
      for (size_t i = 1; i < off.size(); i++) off[i] += off[i - 1];
      src.assign(off.back(), 0);
      fill.assign(off.begin(), off.end() - 1);
    }
    for (int64_t cell = 0; cell < n_cells; cell++)
      for (int sf = first_src; sf < last_src; sf++) {
        const int n = cmfd_split_targets(m, (int)cell, sf, t);
        for (int j = 0; j < n; j++) {
          const int64_t d = (int64_t)(t[j] / CMFD_NS) * dest_per_cell + t[j] % CMFD_NS;
          if (pass == 0) off[d + 1]++;
          else src[fill[d]++] = (int32_t)(cell * CMFD_NS + sf);
        }
      }
  }
}

extern "C" int b200_cmfd_configure(b200_solver* s, const b200_cmfd_config* cfg, const double* widths_x,
                                   const double* widths_y, const double* widths_z, const int32_t* group_indices,
                                   const int64_t* cell_fsr_offset, const int32_t* cell_fsrs, const double* azim_weight,
                                   const double* sin_theta, const double* polar_weight) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_cmfd_configure(c, cfg, widths_x, widths_y, widths_z, group_indices, cell_fsr_offset, cell_fsrs,
                                 azim_weight, sin_theta, polar_weight));
  if (!cfg || !widths_x || !widths_y || !widths_z || !group_indices || !cell_fsr_offset || !cell_fsrs)
    return fail("b200_cmfd_configure: null argument");
  if (!s->cmfd_on) return fail("b200_cmfd_configure: call b200_set_cmfd_groups first (the sweep tallies the currents)");
  if (cfg->num_x < 1 || cfg->num_y < 1 || cfg->num_z < 1) return fail("b200_cmfd_configure: empty mesh");
  const int64_t n_cells = (int64_t)cfg->num_x * cfg->num_y * cfg->num_z;
  if (n_cells * CMFD_NS != s->n_cmfd_slots)
    return fail("b200_cmfd_configure: %lld cells, but the current tally was sized for %lld", (long long)n_cells,
                (long long)(s->n_cmfd_slots / CMFD_NS));
  if (n_cells * CMFD_NS * (int64_t)cfg->num_cmfd_groups > INT32_MAX) return fail("b200_cmfd_configure: mesh too large for 32-bit current slots");
  if (cfg->num_cmfd_groups != s->ncg) return fail("b200_cmfd_configure: %d CMFD groups, b200_set_cmfd_groups said %d", cfg->num_cmfd_groups, s->ncg);
  if (s->cfg.precision != B200_PRECISION_DOUBLE && s->cfg.precision != B200_PRECISION_MIXED)
    return fail("b200_cmfd_configure: not available with this precision mode");
  const int ncg = s->ncg, G = s->G;
  if (group_indices[0] != 0 || group_indices[ncg] != G) return fail("b200_cmfd_configure: group_indices must run from 0 to %d", G);
  for (int e = 0; e < ncg; e++)
    if (group_indices[e + 1] <= group_indices[e]) return fail("b200_cmfd_configure: empty CMFD group %d", e);
  for (int i = 0; i < 6; i++)
    if (cfg->boundaries[i] < B200_BC_VACUUM || cfg->boundaries[i] > B200_BC_PERIODIC)
      return fail("b200_cmfd_configure: boundary %d of face %d is not VACUUM, REFLECTIVE or PERIODIC", cfg->boundaries[i], i);
  if (!cfg->linear_source && (!azim_weight || !sin_theta || !polar_weight || cfg->num_azim_2 < 1 || cfg->num_polar_2 < 1))
    return fail("b200_cmfd_configure: the quadrature of the Larsen factor is missing");
  if (cfg->linear_source && !s->linear) return fail("b200_cmfd_configure: linear_source on a flat-source solver");
  if (!(cfg->sor_factor > 0.)) return fail("b200_cmfd_configure: SOR factor %g", cfg->sor_factor);
  if (cell_fsr_offset[0] != 0) return fail("b200_cmfd_configure: cell_fsr_offset[0] != 0");

  if (s->cmfd == nullptr) s->cmfd = new b200_cmfd();
  b200_cmfd* c = s->cmfd;
  c->cfg = *cfg;
  c->n_cells = n_cells;
  c->have_stencils = c->have_interp = false;
  cudaStream_t st = s->stream;

  /* FSR -> cell (Cmfd::convertFSRIdToCmfdCell) */
  std::vector<int32_t> fsr_cell(s->n_fsr, -1);
  const int64_t n_listed = cell_fsr_offset[n_cells];
  for (int64_t i = 0; i < n_cells; i++) {
    if (cell_fsr_offset[i + 1] < cell_fsr_offset[i]) return fail("b200_cmfd_configure: cell_fsr_offset decreases at cell %lld", (long long)i);
    for (int64_t j = cell_fsr_offset[i]; j < cell_fsr_offset[i + 1]; j++) {
      const int32_t r = cell_fsrs[j];
      if (r < 0 || r >= s->n_fsr) return fail("b200_cmfd_configure: FSR %d of cell %lld outside [0,%lld)", r, (long long)i, (long long)s->n_fsr);
      if (fsr_cell[r] != -1) return fail("b200_cmfd_configure: FSR %d listed in cells %d and %lld", r, fsr_cell[r], (long long)i);
      fsr_cell[r] = (int32_t)i;
    }
  }
  CmfdMesh m{cfg->num_x, cfg->num_y, cfg->num_z, {0, 0, 0, 0, 0, 0}};
  for (int i = 0; i < 6; i++) m.bc[i] = cfg->boundaries[i];
  std::vector<int32_t> nbr(n_cells * CMFD_NF);
  for (int64_t i = 0; i < n_cells; i++)
    for (int f = 0; f < CMFD_NF; f++) nbr[i * CMFD_NF + f] = cmfd_cell_next(m, (int)i, f);
  std::vector<int64_t> sv_off, se_off;
  std::vector<int32_t> sv_src, se_src;
  cmfd_build_split_table(m, CMFD_NFE, CMFD_NS, CMFD_NFE, sv_off, sv_src);      /* vertices -> faces and edges */
  cmfd_build_split_table(m, CMFD_NF, CMFD_NFE, CMFD_NF, se_off, se_src);       /* edges -> faces */

  c->h_wx.assign(widths_x, widths_x + cfg->num_x);
  c->h_wy.assign(widths_y, widths_y + cfg->num_y);
  c->h_wz.assign(widths_z, widths_z + cfg->num_z);
  CU(c->wx.upload(widths_x, cfg->num_x, st));
  CU(c->wy.upload(widths_y, cfg->num_y, st));
  CU(c->wz.upload(widths_z, cfg->num_z, st));
  CU(c->group_idx.upload(group_indices, ncg + 1, st));
  CU(c->cell_fsr_off.upload(cell_fsr_offset, n_cells + 1, st));
  CU(c->cell_fsrs.upload(cell_fsrs, n_listed, st));
  CU(c->fsr_cell.upload(fsr_cell.data(), fsr_cell.size(), st));
  CU(c->nbr.upload(nbr.data(), nbr.size(), st));
  CU(c->sv_off.upload(sv_off.data(), sv_off.size(), st));
  CU(c->sv_src.upload(sv_src.data(), sv_src.size(), st));
  CU(c->se_off.upload(se_off.data(), se_off.size(), st));
  CU(c->se_src.upload(se_src.data(), se_src.size(), st));
  if (!cfg->linear_source) {
    CU(c->azim_w.upload(azim_weight, cfg->num_azim_2, st));
    CU(c->sin_theta.upload(sin_theta, (size_t)cfg->num_azim_2 * cfg->num_polar_2, st));
    CU(c->polar_w.upload(polar_weight, (size_t)cfg->num_azim_2 * cfg->num_polar_2, st));
  }
  const size_t nr = (size_t)n_cells * ncg;
  CU(c->rxn.alloc(nr)); CU(c->volc.alloc(n_cells)); CU(c->dift.alloc(nr)); CU(c->xs_t.alloc(nr)); CU(c->xs_nf.alloc(nr));
  CU(c->xs_chi.alloc(nr)); CU(c->xs_s.alloc(nr * ncg)); CU(c->old_flux.alloc(nr)); CU(c->new_flux.alloc(nr));
  CU(c->dcoef.alloc(nr * 3)); CU(c->old_corr.alloc(nr * CMFD_NF)); CU(c->diag.alloc(nr)); CU(c->off.alloc(nr * CMFD_NF));
  CU(c->ain.alloc(nr * ncg)); CU(c->mm.alloc(nr * ncg)); CU(c->B.alloc(nr)); CU(c->SO.alloc(nr)); CU(c->SN.alloc(nr));
  CU(c->cs.alloc(CS_COUNT)); CU(c->ci.alloc(CI_COUNT)); CU(c->dq.alloc(nr)); CU(c->qd.alloc(nr));
  CU(cudaMemsetAsync(c->old_corr.p, 0, nr * CMFD_NF * 8, st));
  CU(cudaMemsetAsync(c->ci.p, 0, CI_COUNT * sizeof(int), st));
  double cs0[CS_COUNT] = {0};
  cs0[CS_KEFF] = 1.0;                  /* Cmfd::Cmfd, Cmfd.cpp:69 */
  cs0[CS_THRESH] = 1e-5;               /* Cmfd.cpp:23 */
  CU(cudaMemcpyAsync(c->cs.p, cs0, sizeof cs0, cudaMemcpyHostToDevice, st));

  /* launch shape of the eigenvalue solve.  One thread-block cluster with the flux in distributed shared memory
   * while a thread has at most a few cells per colour (mode 2); a cooperative grid with the flux in L2 for larger
   * meshes or unusual group counts (mode 1); one CTA for tiny meshes (mode 0).  B200_CMFD_MODE overrides. */
  const int64_t n_slots = (int64_t)cfg->num_z * cfg->num_y * ((cfg->num_x + 1) / 2);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, s->cfg.device));
  {
    /* cell of every slot of the two colours (linalg.cpp:276-279: ix runs over (iy + iz + colour) % 2, +2, ...) */
    const int hx = (cfg->num_x + 1) / 2;
    std::vector<int32_t> slot_cell(2 * n_slots, -1);
    for (int colour = 0; colour < 2; colour++)
      for (int64_t idx = 0; idx < n_slots; idx++) {
        const int64_t rowi = idx / hx;
        const int k = (int)(idx - rowi * hx), iy = (int)(rowi % cfg->num_y), iz = (int)(rowi / cfg->num_y);
        const int ix = 2 * k + ((iy + iz + colour) & 1);
        if (ix < cfg->num_x) slot_cell[colour * n_slots + idx] = (int32_t)(rowi * cfg->num_x + ix);
      }
    CU(c->slot_cell.upload(slot_cell.data(), slot_cell.size(), st));
    CU(cudaStreamSynchronize(st));
  }
  /* measured on a B200 (tools/cmfd_tune.sh, profiles/r02_cmfd.md): a colour phase costs ~1.4 us + 0.01 us per cell
   * of the busiest SM on a cluster and ~3.9 us on a cooperative grid with one 128-thread CTA per SM */
  int mode = n_slots <= 64 ? 0 : (n_slots <= 4096 ? 2 : 1);
  if (const char* e = getenv("B200_CMFD_MODE")) mode = atoi(e);
  if (mode < 0 || mode > 2) return fail("B200_CMFD_MODE must be 0, 1 or 2");
  if (mode == 2 && cmfd_pick_cluster(ncg) == nullptr) mode = 1;
  c->eigen_smem = 0;
  c->eigen_blocks = 1;
  if (mode == 2) {
    /* the update of a cell is a chain of dependent FP64 operations: spread the cells over as many SMs as a
     * cluster spans rather than filling the threads of few CTAs */
    int C = (int)std::min<int64_t>(16, (n_slots + 31) / 32);
    if (const char* e = getenv("B200_CMFD_CLUSTER")) C = std::max(1, std::min(16, atoi(e)));
    const void* fn = (const void*)cmfd_pick_cluster(ncg);
    const size_t limit = (size_t)prop.sharedMemPerBlockOptin - 2048;
    bool placed = false;
    for (; C >= 1 && !placed; C = (C > 8 ? 8 : C / 2)) {
      const int S = (int)((n_slots + C - 1) / C);
      const size_t xb = (size_t)S * 2 * ncg * sizeof(double);
      if (xb > limit) { if (C == 1) break; continue; }
      const int all = 4 * xb <= limit;
      const size_t smem = all ? 4 * xb : xb;
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); if (C == 1) break; continue; }
      if (C > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(C); lc.blockDim = dim3(CMFD_CLUSTER_THREADS); lc.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      int n_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&n_clusters, fn, &lc) != cudaSuccess || n_clusters < 1) { cudaGetLastError(); if (C == 1) break; continue; }
      c->eigen_blocks = C; c->eigen_smem = smem; c->cluster_slots = S; c->cluster_all_smem = all;
      /* no more warps than cells of a colour: idle warps only lengthen the barriers */
      c->cluster_threads = std::max(64, std::min(CMFD_CLUSTER_THREADS, (S + 31) / 32 * 32));
      if (const char* e = getenv("B200_CMFD_CLUSTER_THREADS")) c->cluster_threads = std::max(32, std::min(CMFD_CLUSTER_THREADS, atoi(e) / 32 * 32));
      placed = true;
      break;
    }
    if (!placed) mode = 1;
  }
  if (mode == 2) {
    /* where the flux of every neighbour lives: owning CTA and offset in its shared memory */
    const int hx = (cfg->num_x + 1) / 2, S = c->cluster_slots;
    std::vector<int32_t> loc(n_cells * CMFD_NF, -1);
    for (int64_t i = 0; i < n_cells; i++)
      for (int f = 0; f < CMFD_NF; f++) {
        const int64_t nbc = nbr[i * CMFD_NF + f];
        if (nbc < 0) continue;
        const int ix = (int)(nbc % cfg->num_x), iy = (int)((nbc / cfg->num_x) % cfg->num_y), iz = (int)(nbc / ((int64_t)cfg->num_x * cfg->num_y));
        const int colour = (ix + iy + iz) & 1;
        const int64_t idx = ((int64_t)iz * cfg->num_y + iy) * hx + ix / 2;
        const int r = (int)(idx / S);
        loc[i * CMFD_NF + f] = (int32_t)((r << 26) | (int)((idx - (int64_t)r * S) * 2 + colour));
      }
    CU(c->nb_loc.upload(loc.data(), loc.size(), st));
    CU(cudaStreamSynchronize(st));
  }
  if (mode == 1 && !prop.cooperativeLaunch) mode = 0;
  c->eigen_mode = mode;
  if (mode == 0) {
    const size_t need = nr * sizeof(double);
    if (need <= (size_t)prop.sharedMemPerBlockOptin - 1024) {
      c->eigen_smem = need;
      CU(cudaFuncSetAttribute((const void*)cmfd_pick_eigen<0>(ncg), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    }
  } else if (mode == 1) {
    /* few warps per SM on as many SMs as the mesh can use (see above) */
    int threads = n_slots <= (int64_t)prop.multiProcessorCount * 128 ? 128 : CMFD_GRID_THREADS;
    if (const char* e = getenv("B200_CMFD_THREADS")) threads = std::max(32, std::min(CMFD_GRID_THREADS, atoi(e) / 32 * 32));
    c->grid_threads = threads;
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cmfd_pick_eigen<1>(ncg), threads, 0));
    if (per_sm < 1) return fail("b200_cmfd_configure: the cooperative CMFD kernel does not fit an SM");
    int64_t blocks = (n_slots + threads - 1) / threads;
    const int64_t cap = (int64_t)per_sm * prop.multiProcessorCount;
    if (blocks > cap) blocks = cap;
    if (const char* e = getenv("B200_CMFD_BLOCKS")) blocks = std::max<int64_t>(1, std::min<int64_t>(cap, atoll(e)));
    c->eigen_blocks = (int)blocks;
  }
  CU(c->partials.alloc((size_t)2 * std::max(1, c->eigen_blocks)));
  CU(cudaStreamSynchronize(st));       /* the host vectors above go out of scope */
  c->configured = true;
  return 0;
}

extern "C" int b200_cmfd_set_stencils(b200_solver* s, const int64_t* offset, const int32_t* cell, const double* weight,
                                      const double* own_weight, const int32_t* stencil_size) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_cmfd_set_stencils(c, offset, cell, weight, own_weight, stencil_size));
  if (s->cmfd == nullptr || !s->cmfd->configured) return fail("b200_cmfd_set_stencils: call b200_cmfd_configure first");
  if (!offset || !own_weight || !stencil_size) return fail("b200_cmfd_set_stencils: null argument");
  b200_cmfd* c = s->cmfd;
  const int64_t n = offset[s->n_fsr];
  if (offset[0] != 0 || n < 0 || (n > 0 && (!cell || !weight))) return fail("b200_cmfd_set_stencils: bad offsets");
  for (int64_t j = 0; j < n; j++)
    if (cell[j] < 0 || cell[j] >= c->n_cells) return fail("b200_cmfd_set_stencils: stencil cell %d outside the mesh", cell[j]);
  for (int64_t r = 0; r < s->n_fsr; r++)
    if (stencil_size[r] < 1) return fail("b200_cmfd_set_stencils: FSR %lld has an empty stencil", (long long)r);
  CU(c->st_off.upload(offset, s->n_fsr + 1, s->stream));
  CU(c->st_cell.upload(cell, n, s->stream));
  CU(c->st_w.upload(weight, n, s->stream));
  CU(c->st_own.upload(own_weight, s->n_fsr, s->stream));
  CU(c->st_n.upload(stencil_size, s->n_fsr, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  c->have_stencils = true;
  return 0;
}

extern "C" int b200_cmfd_set_axial_interpolants(b200_solver* s, const double* interpolants) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_cmfd_set_axial_interpolants(c, interpolants));
  if (s->cmfd == nullptr || !s->cmfd->configured) return fail("b200_cmfd_set_axial_interpolants: call b200_cmfd_configure first");
  if (!interpolants) return fail("b200_cmfd_set_axial_interpolants: null argument");
  CU(s->cmfd->ax_interp.upload(interpolants, (size_t)s->n_fsr * 3, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->cmfd->have_interp = true;
  return 0;
}

extern "C" int b200_cmfd_set_keff(b200_solver* s, double k_eff) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_cmfd_set_keff(c, k_eff));
  if (s->cmfd == nullptr || !s->cmfd->configured) return fail("b200_cmfd_set_keff: call b200_cmfd_configure first");
  CU(cudaMemcpyAsync(s->cmfd->cs.p + CS_KEFF, &k_eff, 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

static bool cmfd_in_loop(const b200_solver* s) {
  return s->cmfd != nullptr && s->cmfd->configured && s->cmfd->in_loop && s->cmfd_on;
}

extern "C" int b200_cmfd_set_in_loop(b200_solver* s, int32_t on) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_cmfd_set_in_loop(c, on));
  if (s->cmfd == nullptr || !s->cmfd->configured) return fail("b200_cmfd_set_in_loop: call b200_cmfd_configure first");
  s->cmfd->in_loop = on != 0;
  return 0;
}

/* start of a fused eigenvalue loop: Solver::initializeCmfd's threshold (Solver.cpp:1159) */
static int cmfd_loop_init(b200_solver* s, double tol) {
  if (!cmfd_in_loop(s)) return 0;
  const double t = tol * 1.e-1;
  CU(cudaMemcpyAsync(s->cmfd->cs.p + CS_THRESH, &t, 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

static CmfdArgs cmfd_args(b200_solver* s) {
  b200_cmfd* c = s->cmfd;
  CmfdArgs a;
  a.nx = c->cfg.num_x; a.ny = c->cfg.num_y; a.nz = c->cfg.num_z; a.ncg = s->ncg; a.G = s->G;
  a.n_cells = c->n_cells; a.n_fsr = s->n_fsr;
  for (int i = 0; i < 6; i++) a.bc[i] = c->cfg.boundaries[i];
  a.linear = c->cfg.linear_source; a.flux_limiting = c->cfg.flux_limiting; a.centroid = c->cfg.centroid_update;
  a.axial_interp = c->cfg.axial_interpolation; a.n_unbounded = c->cfg.num_unbounded_iterations;
  a.sor = c->cfg.sor_factor; a.relax = c->cfg.relaxation_factor; a.linalg_tol = c->cfg.linalg_tolerance;
  a.n_azim_2 = c->cfg.num_azim_2; a.n_polar_2 = c->cfg.num_polar_2;
  a.wx = c->wx.p; a.wy = c->wy.p; a.wz = c->wz.p;
  a.group_idx = c->group_idx.p; a.moc_to_cmfd = s->cmfd_group.p;
  a.cell_fsr_off = c->cell_fsr_off.p; a.cell_fsrs = c->cell_fsrs.p; a.fsr_cell = c->fsr_cell.p; a.nbr = c->nbr.p;
  a.azim_w = c->azim_w.p; a.sin_theta = c->sin_theta.p; a.polar_w = c->polar_w.p;
  a.fsr_mat = s->fsr_mat.p; a.vol = s->vol.p; a.sigma_t = s->sigma_t.p; a.sigma_s = s->sigma_s.p;
  a.nu_sigma_f = s->nu_sigma_f.p; a.chi = s->chi.p;
  a.phi = s->phi.p; a.phi_m = s->phi_m.p; a.cur = s->currents.p; a.scal = s->scal.p; a.iscal = s->iscal.p;
  a.sv_off = c->sv_off.p; a.sv_src = c->sv_src.p; a.se_off = c->se_off.p; a.se_src = c->se_src.p;
  a.rxn = c->rxn.p; a.volc = c->volc.p; a.dift = c->dift.p; a.xs_t = c->xs_t.p; a.xs_nf = c->xs_nf.p;
  a.xs_chi = c->xs_chi.p; a.xs_s = c->xs_s.p; a.old_flux = c->old_flux.p; a.new_flux = c->new_flux.p;
  a.dcoef = c->dcoef.p; a.old_corr = c->old_corr.p; a.diag = c->diag.p; a.off = c->off.p; a.ain = c->ain.p;
  a.mm = c->mm.p; a.B = c->B.p; a.SO = c->SO.p; a.SN = c->SN.p; a.partials = c->partials.p;
  a.cs = c->cs.p; a.ci = c->ci.p; a.x_in_smem = c->eigen_smem > 0;
  a.dq = c->dq.p; a.qd = c->qd.p; a.slot_cell = c->slot_cell.p;
  a.st_off = c->st_off.p; a.st_cell = c->st_cell.p; a.st_w = c->st_w.p; a.st_own = c->st_own.p; a.st_n = c->st_n.p;
  a.ax_interp = c->ax_interp.p;
  return a;
}

/* the seven launches of one Cmfd::computeKeff; moc_iteration < 0: the device-side iteration counter */
static int enqueue_cmfd(b200_solver* s, int moc_iteration, double source_threshold) {
  b200_cmfd* c = s->cmfd;
  if (c == nullptr || !c->configured) return fail("CMFD on the device: b200_cmfd_configure has not been called");
  if (!s->cmfd_on) return fail("CMFD on the device: the current tally is off (b200_set_cmfd_groups)");
  if (c->cfg.centroid_update && !c->have_stencils) return fail("CMFD on the device: centroid update without b200_cmfd_set_stencils");
  if (c->cfg.axial_interpolation && c->cfg.num_z >= 3 && !c->have_interp)
    return fail("CMFD on the device: axial interpolation without b200_cmfd_set_axial_interpolants");
  CmfdArgs a = cmfd_args(s);
  cudaStream_t st = s->stream;
  const int ncg = s->ncg;
  if (c->cfg.num_z > 1) {
    cmfd_split_kernel<<<grid_for(c->n_cells * CMFD_NFE * ncg, 256), 256, 0, st>>>(
        a.cur, a.sv_off, a.sv_src, c->n_cells * CMFD_NFE, CMFD_NFE, ncg, 3.0, s->iscal.p);
    CU(cudaGetLastError());
    s->n_launches++;
  }
  cmfd_split_kernel<<<grid_for(c->n_cells * CMFD_NF * ncg, 256), 256, 0, st>>>(
      a.cur, a.se_off, a.se_src, c->n_cells * CMFD_NF, CMFD_NF, ncg, 2.0, s->iscal.p);
  CU(cudaGetLastError());
  cmfd_collapse_kernel<<<grid_for(c->n_cells, 128), 128, 0, st>>>(a);
  CU(cudaGetLastError());
  cmfd_diffusion_kernel<<<grid_for(c->n_cells * ncg * 3, 256), 256, 0, st>>>(a);
  CU(cudaGetLastError());
  cmfd_matrix_kernel<<<grid_for(c->n_cells * ncg, 256), 256, 0, st>>>(a, moc_iteration);
  CU(cudaGetLastError());
  void* params[] = {(void*)&a, (void*)&source_threshold};
  if (c->eigen_mode == 2) {
    CmfdClusterArgs ca;
    ca.slots_per_cta = c->cluster_slots; ca.all_smem = c->cluster_all_smem; ca.nb_loc = c->nb_loc.p;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(c->eigen_blocks); lc.blockDim = dim3(c->cluster_threads); lc.dynamicSmemBytes = c->eigen_smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c->eigen_blocks; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    void* cparams[] = {(void*)&a, (void*)&ca, (void*)&source_threshold};
    CU(cudaLaunchKernelExC(&lc, (const void*)cmfd_pick_cluster(ncg), cparams));
  } else if (c->eigen_mode == 0)
    CU(cudaLaunchKernel((const void*)cmfd_pick_eigen<0>(ncg), dim3(1), dim3(CMFD_BLOCK_THREADS), params, c->eigen_smem, st));
  else
    CU(cudaLaunchCooperativeKernel((const void*)cmfd_pick_eigen<1>(ncg), dim3(c->eigen_blocks), dim3(c->grid_threads),
                                   params, 0, st));
  cmfd_update_kernel<<<grid_for(s->n_fsr, 256), 256, 0, st>>>(a, moc_iteration);
  CU(cudaGetLastError());
  s->n_launches += 6;
  return 0;
}

static int enqueue_cmfd_threshold(b200_solver* s) {
  if (s->cmfd == nullptr || !s->cmfd->configured) return 0;
  cmfd_threshold_kernel<<<1, 1, 0, s->stream>>>(s->cmfd->cs.p, s->scal.p, s->iscal.p);
  CU(cudaGetLastError());
  s->n_launches++;
  return 0;
}

extern "C" int b200_cmfd_solve(b200_solver* s, int32_t moc_iteration, double source_threshold, double* k_eff,
                               b200_cmfd_stats* stats) {
  NEED_FINAL(s);
  if (s->grp != nullptr) {            /* replicated on every shard: same bits everywhere */
    for (size_t i = 0; i < grp_shards(s).size(); i++) {
      b200_solver* c = grp_shards(s)[i];
      if (b200_cmfd_solve(c, moc_iteration, source_threshold, i == 0 ? k_eff : nullptr, i == 0 ? stats : nullptr)) return 1;
    }
    return 0;
  }
  if (clear_done(s)) return 1;
  const bool timed = stats != nullptr && !s->capturing;
  if (timed) {
    if (s->cmfd != nullptr && s->cmfd->ev0 == nullptr) { CU(cudaEventCreate(&s->cmfd->ev0)); CU(cudaEventCreate(&s->cmfd->ev1)); }
    if (s->cmfd != nullptr) CU(cudaEventRecord(s->cmfd->ev0, s->stream));
  }
  if (enqueue_cmfd(s, moc_iteration, source_threshold)) return 1;
  if (timed) CU(cudaEventRecord(s->cmfd->ev1, s->stream));
  if (k_eff == nullptr && stats == nullptr) return 0;
  double cs[CS_COUNT];
  int ci[CI_COUNT];
  CU(cudaMemcpyAsync(cs, s->cmfd->cs.p, sizeof cs, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(ci, s->cmfd->ci.p, sizeof ci, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (k_eff) *k_eff = cs[CS_KEFF];
  if (stats) {
    long long hi, lo;
    memcpy(&hi, &ci[CI_PF_MAX], 8);
    memcpy(&lo, &ci[CI_PF_MIN], 8);
    double lmax, lmin;
    memcpy(&lmax, &hi, 8);
    memcpy(&lmin, &lo, 8);
    stats->pf = lmax >= lmin ? exp(lmax) : exp(-lmin);
    stats->cmfd_res_1 = cs[CS_RES_1]; stats->cmfd_res_end = cs[CS_RES_END];
    stats->linear_res_1 = cs[CS_LIN_RES_1]; stats->linear_res_end = cs[CS_LIN_RES_END];
    stats->cmfd_iters = ci[CI_POWER_ITERS]; stats->linear_iters_1 = ci[CI_LIN_ITERS_1];
    stats->linear_iters_end = ci[CI_LIN_ITERS_END]; stats->linear_iters_total = ci[CI_LIN_TOTAL];
    stats->failed = ci[CI_FAIL]; stats->bad_tallies = ci[CI_BAD_TALLY];
    stats->device_ms = 0.;
    if (timed) { float ms = 0.f; CU(cudaEventElapsedTime(&ms, s->cmfd->ev0, s->cmfd->ev1)); stats->device_ms = ms; }
  }
  return 0;
}
