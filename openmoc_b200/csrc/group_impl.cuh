/*
 * group_impl.cuh - host side of the multi-device group (see group.cuh); included at the end of
 * b200moc.cu, where struct b200_solver and the launch helpers are complete.
 */
#pragma once

static std::vector<b200_solver*>& grp_shards(b200_solver* s) { return s->grp->shard; }

extern "C" int b200_set_devices(b200_solver* s, int32_t n_devices, const int32_t* devices) {
  NEED(s);
  if (n_devices < 1 || n_devices > GROUP_MAX || devices == nullptr)
    return fail("b200_set_devices: between 1 and %d devices", GROUP_MAX);
  if (s->have_tracks || s->have_fsrs || s->have_mats || s->grp != nullptr)
    return fail("b200_set_devices: call it right after b200_create, before any upload");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  for (int i = 0; i < n_devices; i++)
    if (devices[i] < 0 || devices[i] >= ndev) return fail("b200_set_devices: device %d out of range [0,%d)", devices[i], ndev);
  if (n_devices == 1 && devices[0] == s->cfg.device) return 0;          /* plain single-device solver */
  s->grp = new b200_group();
  s->grp->devices.assign(devices, devices + n_devices);
  return 0;
}

extern "C" int b200_get_num_devices(b200_solver* s, int32_t* n) {
  NEED(s);
  if (n) *n = s->grp ? (int32_t)s->grp->devices.size() : 1;
  return 0;
}

static void grp_destroy(b200_solver* s) {
  b200_group* g = s->grp;
  if (g == nullptr) return;
  for (size_t c = 0; c < g->shard.size(); c++) {
    if (g->shard[c] != nullptr) {
      cudaSetDevice(g->shard[c]->cfg.device);
      if (c < g->ev_a.size()) { cudaEventDestroy(g->ev_a[c]); cudaEventDestroy(g->ev_b[c]); }
      if (c < g->stage.size()) g->stage[c].release();
      b200_destroy(g->shard[c]);
    }
  }
  delete g;
  s->grp = nullptr;
}

/* ------------------------------------------------------------------------- */
/* the all-reduce of the shards' tallies over peer memory                      */
/* ------------------------------------------------------------------------- */
struct GrpTally { char kind; size_t n; };     /* kind: 'p' phi, 'x' fixed-point phi, 'm' moments, 'c' currents */

static void* grp_tally_ptr(b200_solver* c, char kind) {
  switch (kind) {
    case 'p': return c->phi.p;
    case 'x': return c->phi_fx.p;
    case 'm': return c->phi_m.p;
    case 'c': return c->currents.p;
  }
  return nullptr;
}

static int grp_reduce(b200_solver* s) {
  b200_group* g = s->grp;
  const int N = (int)g->shard.size();
  if (N < 2) return 0;
  b200_solver* c0 = g->shard[0];
  const size_t nphi = (size_t)c0->n_fsr * c0->G;
  std::vector<GrpTally> tallies;
  if (c0->cfg.deterministic) tallies.push_back({'x', (size_t)c0->n_fsr * c0->GP}); else tallies.push_back({'p', nphi});
  if (c0->linear) tallies.push_back({'m', 3 * nphi});
  if (c0->cmfd_on) tallies.push_back({'c', (size_t)c0->n_cmfd_slots * c0->ncg});
  size_t total = 0;
  for (const GrpTally& t : tallies) total += t.n;
  for (int c = 0; c < N; c++) {
    b200_solver* sc = g->shard[c];
    CU(cudaSetDevice(sc->cfg.device));
    if (g->stage[c].n < total) CU(g->stage[c].alloc(total));
    CU(cudaEventRecord(g->ev_a[c], sc->stream));
  }
  /* stage 1: every shard reduces its slice of every tally once all sweeps have finished */
  for (int c = 0; c < N; c++) {
    b200_solver* sc = g->shard[c];
    CU(cudaSetDevice(sc->cfg.device));
    for (int j = 0; j < N; j++)
      if (j != c) CU(cudaStreamWaitEvent(sc->stream, g->ev_a[j], 0));
    size_t off = 0;
    for (const GrpTally& t : tallies) {
      PeerPtrs src;
      for (int j = 0; j < N; j++) src.p[j] = grp_tally_ptr(g->shard[j], t.kind);
      const int64_t chunk = ((int64_t)t.n + N - 1) / N;
      const int64_t lo = std::min<int64_t>((int64_t)t.n, chunk * c), hi = std::min<int64_t>((int64_t)t.n, chunk * (c + 1));
      if (hi > lo) {
        const int nb = grid_for(hi - lo, 256, 148 * 8);
        if (t.kind == 'x')
          group_reduce_kernel<unsigned long long><<<nb, 256, 0, sc->stream>>>(src, N, (unsigned long long*)(g->stage[c].p + off), lo, hi, sc->iscal.p + SI_DONE);
        else
          group_reduce_kernel<double><<<nb, 256, 0, sc->stream>>>(src, N, g->stage[c].p + off, lo, hi, sc->iscal.p + SI_DONE);
        CU(cudaGetLastError());
        sc->n_launches++;
      }
      off += t.n;
    }
    CU(cudaEventRecord(g->ev_b[c], sc->stream));
  }
  /* stage 2: every shard gathers all slices */
  for (int c = 0; c < N; c++) {
    b200_solver* sc = g->shard[c];
    CU(cudaSetDevice(sc->cfg.device));
    for (int j = 0; j < N; j++)
      if (j != c) CU(cudaStreamWaitEvent(sc->stream, g->ev_b[j], 0));
    size_t off = 0;
    for (const GrpTally& t : tallies) {
      PeerPtrs stg;
      for (int j = 0; j < N; j++) stg.p[j] = g->stage[j].p + off;
      const int nb = grid_for((int64_t)t.n, 256, 148 * 8);
      if (t.kind == 'x')
        group_gather_kernel<unsigned long long><<<nb, 256, 0, sc->stream>>>(stg, N, (int64_t)t.n, (unsigned long long*)grp_tally_ptr(sc, t.kind), sc->iscal.p + SI_DONE);
      else
        group_gather_kernel<double><<<nb, 256, 0, sc->stream>>>(stg, N, (int64_t)t.n, (double*)grp_tally_ptr(sc, t.kind), sc->iscal.p + SI_DONE);
      CU(cudaGetLastError());
      sc->n_launches++;
      off += t.n;
    }
  }
  g->n_reduces++;
  CU(cudaSetDevice(s->cfg.device));
  return 0;
}

/* sweep on every shard, tallies summed: the group's b200_transport_sweep */
static int grp_sweep(b200_solver* s) {
  /* every shard writes the hand-offs that cross devices into the buffer the owner reads NEXT sweep: the
   * pointers are taken before any shard flips its double buffer */
  if (s->grp->cross_links) {
    std::vector<float*> next_in;
    for (b200_solver* c : grp_shards(s)) next_in.push_back(c->psi_other);
    for (b200_solver* c : grp_shards(s)) c->peer_out = next_in;
  }
  for (b200_solver* c : grp_shards(s)) {
    CU(cudaSetDevice(c->cfg.device));
    if (c->cfg.deterministic) c->defer_fx_convert = true;
    if (launch_sweep(c)) return 1;
  }
  if (grp_reduce(s)) return 1;
  for (b200_solver* c : grp_shards(s))
    if (c->cfg.deterministic) {
      CU(cudaSetDevice(c->cfg.device));
      fx_to_double_kernel<<<grid_for(c->n_fsr * c->G, 256), 256, 0, c->stream>>>(fsr_args(c), c->phi_fx.p, c->GP);
      CU(cudaGetLastError());
      c->n_launches++;
    }
  return 0;
}

static int grp_iteration(b200_solver* s, int i, int res_type, int loop_kind) {
  for (b200_solver* c : grp_shards(s)) {
    CU(cudaSetDevice(c->cfg.device));
    if (i != 0 && c->stabilize) { if (launch_stabilizing_flux(c)) return 1; }
    if (launch_sources(c, i, 0)) return 1;
  }
  if (grp_sweep(s)) return 1;
  for (b200_solver* c : grp_shards(s)) {
    CU(cudaSetDevice(c->cfg.device));
    if (enqueue_iteration_end(c, i, res_type, loop_kind)) return 1;
  }
  return 0;
}

static int grp_sync(b200_solver* s) {
  for (b200_solver* c : grp_shards(s)) {
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* finalize: shard the parked uploads and build one solver per shard           */
/* ------------------------------------------------------------------------- */
static int grp_finalize(b200_solver* s) {
  b200_group* g = s->grp;
  const int N = (int)g->devices.size();
  if (!g->have_explicit && !g->have_otf) return fail("b200_finalize: no tracks uploaded");
  if (g->weight.empty() || g->fsr_mat.empty() || g->sigma_t.empty())
    return fail("b200_finalize: tracks, quadrature, FSRs and materials must all be uploaded first");
  if (s->linear && !g->have_ls) return fail("b200_finalize: linear source requested but b200_upload_linear_source was not called");
  for (b200_solver* c : g->shard) if (c) { cudaSetDevice(c->cfg.device); b200_destroy(c); }
  g->shard.assign(N, nullptr);
  g->ids.assign(N, {});
  const int64_t nt = s->n_trk;

  /* peer access between every pair of distinct devices (an error here leaves the loads to fail loudly) */
  for (int a = 0; a < N; a++)
    for (int b = 0; b < N; b++)
      if (g->devices[a] != g->devices[b]) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, g->devices[a], g->devices[b]));
        if (!can) return fail("b200_finalize: device %d cannot access the memory of device %d (no P2P)", g->devices[a], g->devices[b]);
        CU(cudaSetDevice(g->devices[a]));
        cudaError_t e = cudaDeviceEnablePeerAccess(g->devices[b], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        (void)cudaGetLastError();
      }

  /* work per track: its segment count (on-the-fly tracks: counted on the first device) */
  std::vector<double> load(nt, 1.);
  b200_config cfg = s->cfg;
  if (g->have_explicit) {
    for (int64_t t = 0; t < nt; t++) load[t] = (double)(g->trk_off[t + 1] - g->trk_off[t]);
  } else {
    b200_solver* tmp = nullptr;
    b200_config c0 = cfg;
    c0.device = g->devices[0]; c0.n_tracks = 0; c0.n_segments = 0;
    if (b200_create(&c0, &tmp)) return 1;
    struct Guard { b200_solver* p; ~Guard() { if (p) b200_destroy(p); } } guard{tmp};
    if (b200_upload_otf_geometry(tmp, g->n_trk2d, g->n_seg2d, g->seg2d_len.data(), g->seg2d_ext.data(), g->trk2d_off.data(),
                                 g->n_ext, g->n_ext > 0 ? g->ext_off.data() : nullptr, g->ext_mesh.data(),
                                 g->n_ext > 0 ? g->ext_fsr.data() : nullptr, g->n_axial, g->theta.data())) return 1;
    std::vector<int32_t> cnt(nt);
    if (b200_otf_count_segments(tmp, nt, g->trk_2d.data(), g->trk_l0.data(), g->trk_z0.data(), g->trk_azim.data(),
                                g->trk_polar.data(), cnt.data())) return 1;
    for (int64_t t = 0; t < nt; t++) load[t] = (double)cnt[t];
    /* FSR volumes: traced once, the same bits go to every shard */
    if (g->have_voltrk && !g->have_volume) {
      if (b200_otf_compute_volumes(tmp, (int64_t)g->v_2d.size(), g->v_2d.data(), g->v_l0.data(), g->v_z0.data(), g->v_azim.data(),
                                   g->v_polar.data(), g->v_weight.data())) return 1;
      g->volume.resize(s->n_fsr);
      if (b200_get_volumes(tmp, g->volume.data(), s->n_fsr)) return 1;
      g->have_volume = true;
    }
  }
  if (!g->have_volume) return fail("b200_finalize: no FSR volumes (b200_upload_fsrs with volume = NULL needs b200_otf_compute_volumes)");
  /* Partition.  2D decks: whole chains when there are enough of them (no angular flux ever crosses a
   * shard); otherwise single tracks, dealt in a snake by decreasing load - any number of shards, also for
   * fully reflective decks whose tracks form a handful of cycles.  Whenever tracks are not sharded by
   * chain, the hand-offs that cross shards are stored by the sweep kernel straight into the owner's
   * start-flux buffer over peer memory (sweep.cuh: PeerOut).  B200_GROUP_PARTITION=chain|track|block overrides. */
  std::vector<int32_t> owner;
  const int64_t n_chains = partition_chains(nt, g->next_fwd.data(), g->next_bwd.data(), g->bc_fwd.data(), g->bc_bwd.data(),
                                            load.data(), N, owner);
  /* 3D decks: contiguous blocks of the Track uid order (azimuthal angle, 2D track, polar angle, z),
   * cut where the cumulative load crosses k/N: every shard then sweeps whole neighbouring z-stacks one
   * after the other like a single GPU does and its FSR rows stay in L2 (a chain deal spreads the millions
   * of short chains of a 3D deck with vacuum sides all over the core). */
  enum { CHAIN, TRACK, BLOCK } mode = s->cfg.solve_3d ? BLOCK : (n_chains < (int64_t)4 * N ? TRACK : CHAIN);
  if (const char* e = getenv("B200_GROUP_PARTITION")) {
    if (!strcmp(e, "track")) mode = TRACK;
    if (!strcmp(e, "chain")) mode = CHAIN;
    if (!strcmp(e, "block")) mode = BLOCK;
  }
  if (mode == CHAIN && n_chains < N) return fail("b200_finalize: %d shards but only %lld independent track chains", N, (long long)n_chains);
  if (mode == TRACK) {
    std::vector<int64_t> order(nt);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return load[a] > load[b]; });
    owner.assign(nt, 0);
    for (int64_t i = 0; i < nt; i++) {
      const int64_t pos = i % (2 * N);
      owner[order[i]] = (int32_t)(pos < N ? pos : 2 * N - 1 - pos);
    }
  } else if (mode == BLOCK) {
    double total = 0., cum = 0.;
    for (int64_t t = 0; t < nt; t++) total += load[t] + 1e-9;
    owner.assign(nt, 0);
    for (int64_t t = 0; t < nt; t++) {
      const double l = load[t] + 1e-9;
      cum += l;
      owner[t] = (int32_t)std::min<double>(N - 1, (cum - 0.5 * l) * N / total);
    }
  }
  g->cross_links = false;
  std::vector<int64_t> local(nt);
  for (int64_t t = 0; t < nt; t++) { local[t] = (int64_t)g->ids[owner[t]].size(); g->ids[owner[t]].push_back(t); }
  /* global one-to-one check of the hand-off table (b200_finalize of a shard only sees its own links) */
  {
    std::vector<uint8_t> fed(2 * nt, 0);
    for (int64_t t = 0; t < nt; t++)
      for (int d = 0; d < 2; d++) {
        const uint8_t bc = d == 0 ? g->bc_fwd[t] : g->bc_bwd[t];
        if (bc != B200_BC_REFLECTIVE && bc != B200_BC_PERIODIC) continue;
        const int64_t nx = d == 0 ? g->next_fwd[t] : g->next_bwd[t];
        if (nx < 0 || nx >= nt) return fail("b200_finalize: track %lld links to track %lld outside [0,%lld)", (long long)t, (long long)nx, (long long)nt);
        const int fwd = d == 0 ? (g->flags[t] & 1) : ((g->flags[t] >> 1) & 1);
        const int64_t slot = nx * 2 + (fwd ? 0 : 1);
        if (fed[slot]) return fail("b200_finalize: two track ends hand their flux to the same slot (track %lld)", (long long)nx);
        fed[slot] = 1;
      }
  }

  for (int c = 0; c < N; c++) {
    const std::vector<int64_t>& ids = g->ids[c];
    const int64_t n = (int64_t)ids.size();
    std::vector<int32_t> azim(n), polar(n);
    std::vector<int64_t> nf(n), nb(n);
    std::vector<uint8_t> fl(n), bf(n), bb(n);
    for (int64_t k = 0; k < n; k++) {
      const int64_t t = ids[k];
      azim[k] = g->trk_azim[t]; polar[k] = g->trk_polar[t]; fl[k] = g->flags[t]; bf[k] = g->bc_fwd[t]; bb[k] = g->bc_bwd[t];
      const bool lf = bf[k] == B200_BC_REFLECTIVE || bf[k] == B200_BC_PERIODIC, lb = bb[k] == B200_BC_REFLECTIVE || bb[k] == B200_BC_PERIODIC;
      nf[k] = lf ? local[g->next_fwd[t]] : -1;
      nb[k] = lb ? local[g->next_bwd[t]] : -1;
      /* a hand-off that leaves the shard: uploaded as an open end, patched into a peer store below */
      if (lf && owner[g->next_fwd[t]] != c) { nf[k] = -1; bf[k] = B200_BC_VACUUM; g->cross_links = true; }
      if (lb && owner[g->next_bwd[t]] != c) { nb[k] = -1; bb[k] = B200_BC_VACUUM; g->cross_links = true; }
    }
    b200_config cc = cfg;
    cc.device = g->devices[c];
    cc.n_tracks = n;
    cc.n_fsrs_global = s->n_fsr_global;
    std::vector<int64_t> off(n + 1, 0);
    std::vector<int64_t> seg_idx;
    if (g->have_explicit) {
      for (int64_t k = 0; k < n; k++) off[k + 1] = off[k] + (g->trk_off[ids[k] + 1] - g->trk_off[ids[k]]);
      cc.n_segments = off[n];
    } else {
      cc.n_segments = 0;
    }
    b200_solver* sc = nullptr;
    if (b200_create(&cc, &sc)) return 1;
    g->shard[c] = sc;
    sc->max_tau = s->max_tau;
    if (g->have_explicit) {
      std::vector<double> len(cc.n_segments);
      std::vector<int32_t> fsr(cc.n_segments);
      for (int64_t k = 0; k < n; k++) {
        const int64_t a = g->trk_off[ids[k]], m = off[k + 1] - off[k];
        std::copy(g->seg_length.begin() + a, g->seg_length.begin() + a + m, len.begin() + off[k]);
        std::copy(g->seg_fsr.begin() + a, g->seg_fsr.begin() + a + m, fsr.begin() + off[k]);
      }
      if (b200_upload_tracks(sc, len.data(), fsr.data(), off.data(), azim.data(), polar.data(), nf.data(), nb.data(), fl.data(),
                             bf.data(), bb.data())) return 1;
      if (s->linear) {
        std::vector<double> st((size_t)cc.n_segments * 3), dir((size_t)n * 3);
        for (int64_t k = 0; k < n; k++) {
          const int64_t a = g->trk_off[ids[k]], m = off[k + 1] - off[k];
          std::copy(g->seg_start.begin() + 3 * a, g->seg_start.begin() + 3 * (a + m), st.begin() + 3 * off[k]);
          for (int d = 0; d < 3; d++) dir[3 * k + d] = g->trk_dir[3 * ids[k] + d];
        }
        if (b200_upload_linear_source(sc, st.data(), dir.data(), g->lin_exp.data(), g->src_const.data())) return 1;
      }
      if (g->have_cmfd) {
        std::vector<int32_t> cf(cc.n_segments), cb(cc.n_segments);
        for (int64_t k = 0; k < n; k++) {
          const int64_t a = g->trk_off[ids[k]], m = off[k + 1] - off[k];
          std::copy(g->cmfd_fwd.begin() + a, g->cmfd_fwd.begin() + a + m, cf.begin() + off[k]);
          std::copy(g->cmfd_bwd.begin() + a, g->cmfd_bwd.begin() + a + m, cb.begin() + off[k]);
        }
        if (b200_upload_cmfd_surfaces(sc, cf.data(), cb.data())) return 1;
      }
    } else {
      if (b200_upload_otf_geometry(sc, g->n_trk2d, g->n_seg2d, g->seg2d_len.data(), g->seg2d_ext.data(), g->trk2d_off.data(),
                                   g->n_ext, g->n_ext > 0 ? g->ext_off.data() : nullptr, g->ext_mesh.data(),
                                   g->n_ext > 0 ? g->ext_fsr.data() : nullptr, g->n_axial, g->theta.data())) return 1;
      if (g->have_otf_cmfd &&
          b200_upload_otf_cmfd(sc, g->otf_surf_fwd.data(), g->otf_surf_bwd.data(), g->otf_fsr_cell.data(), g->otf_cmfd_nx,
                               g->otf_cmfd_ny, g->otf_cmfd_nz, g->otf_cmfd_z.data())) return 1;
      std::vector<int32_t> t2(n);
      std::vector<double> l0(n), z0(n);
      for (int64_t k = 0; k < n; k++) { t2[k] = g->trk_2d[ids[k]]; l0[k] = g->trk_l0[ids[k]]; z0[k] = g->trk_z0[ids[k]]; }
      if (b200_upload_tracks_otf(sc, t2.data(), l0.data(), z0.data(), azim.data(), polar.data(), nf.data(), nb.data(), fl.data(),
                                 bf.data(), bb.data(), nullptr)) return 1;
    }
    if (b200_upload_quadrature(sc, g->weight.data(), g->sin_theta.data())) return 1;
    if (b200_upload_fsrs(sc, g->volume.data(), g->fsr_mat.data())) return 1;
    if (b200_upload_materials(sc, g->sigma_t.data(), g->sigma_s.data(), g->fiss.data(), g->nu_sigma_f.data(),
                              g->have_sigma_f ? g->sigma_f.data() : nullptr, g->chi.data(), g->fissionable.data())) return 1;
    if (b200_finalize(sc)) return 1;
    sc->stabilize = s->stabilize; sc->stab_factor = s->stab_factor; sc->stab_type = s->stab_type;
    sc->neg_allowed = s->neg_allowed;
  }
  /* hand-offs that cross shards: destination = (owner + 1) << 48 | slot in the owner's buffer; the
   * destination slot is fed, so the owner must not copy its incoming flux through */
  if (g->cross_links) {
    std::vector<std::vector<int64_t>> out_slot(N);
    std::vector<std::vector<uint8_t>> carry(N);
    for (int c = 0; c < N; c++) { out_slot[c].assign(2 * g->ids[c].size(), -1); carry[c].assign(2 * g->ids[c].size(), 1); }
    for (int64_t t = 0; t < nt; t++)
      for (int d = 0; d < 2; d++) {
        const uint8_t bc = d == 0 ? g->bc_fwd[t] : g->bc_bwd[t];
        if (bc != B200_BC_REFLECTIVE && bc != B200_BC_PERIODIC) continue;
        const int64_t nx = d == 0 ? g->next_fwd[t] : g->next_bwd[t];
        const int fwd = d == 0 ? (g->flags[t] & 1) : ((g->flags[t] >> 1) & 1);
        const int co = owner[t], cd = owner[nx];
        const int64_t slot = local[nx] * 2 + (fwd ? 0 : 1);
        out_slot[co][local[t] * 2 + d] = cd == co ? slot : (((int64_t)(cd + 1) << PEER_SHIFT) | slot);
        carry[cd][slot] = 0;
      }
    for (int c = 0; c < N; c++) {
      b200_solver* sc = g->shard[c];
      CU(cudaSetDevice(sc->cfg.device));
      CU(sc->out_slot.upload(out_slot[c].data(), out_slot[c].size(), sc->stream));
      CU(sc->carry.upload(carry[c].data(), carry[c].size(), sc->stream));
      CU(cudaStreamSynchronize(sc->stream));
    }
  }
  /* totals and plumbing */
  s->n_seg = 0;
  for (b200_solver* c : g->shard) s->n_seg += c->n_seg;
  s->cfg.n_segments = s->n_seg;
  s->n_fissionable = g->shard[0]->n_fissionable;
  g->ev_a.resize(N); g->ev_b.resize(N); g->stage.resize(N);
  for (int c = 0; c < N; c++) {
    CU(cudaSetDevice(g->devices[c]));
    CU(cudaEventCreateWithFlags(&g->ev_a[c], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g->ev_b[c], cudaEventDisableTiming));
  }
  /* the big host copies are not needed again (a material refresh re-uploads its own tables) */
  std::vector<double>().swap(g->seg_length); std::vector<int32_t>().swap(g->seg_fsr);
  std::vector<double>().swap(g->seg_start); std::vector<int32_t>().swap(g->cmfd_fwd); std::vector<int32_t>().swap(g->cmfd_bwd);
  CU(cudaSetDevice(s->cfg.device));
  s->finalized = true;
  return 0;
}

/* start fluxes live with the tracks: gathered from / scattered to the shards by global track id */
static int grp_get_start_fluxes(b200_solver* s, float* out, int64_t n) {
  b200_group* g = s->grp;
  const int F = g->shard[0]->F;
  if (n != s->n_trk * 2 * (int64_t)F) return fail("b200_get_start_fluxes: size mismatch");
  for (size_t c = 0; c < g->shard.size(); c++) {
    const std::vector<int64_t>& ids = g->ids[c];
    std::vector<float> tmp(ids.size() * 2 * (size_t)F);
    if (b200_get_start_fluxes(g->shard[c], tmp.data(), (int64_t)tmp.size())) return 1;
    for (size_t k = 0; k < ids.size(); k++)
      std::copy(tmp.begin() + k * 2 * F, tmp.begin() + (k + 1) * 2 * F, out + ids[k] * 2 * F);
  }
  return 0;
}
static int grp_set_start_fluxes(b200_solver* s, const float* in, int64_t n) {
  b200_group* g = s->grp;
  const int F = g->shard[0]->F;
  if (n != s->n_trk * 2 * (int64_t)F) return fail("b200_set_start_fluxes: size mismatch");
  /* the bound on |psi| of the deterministic tally must be the same on every shard */
  double m = 0.;
  for (int64_t i = 0; i < n; i++) m = std::max(m, (double)std::fabs(in[i]));
  for (size_t c = 0; c < g->shard.size(); c++) {
    const std::vector<int64_t>& ids = g->ids[c];
    std::vector<float> tmp(ids.size() * 2 * (size_t)F);
    for (size_t k = 0; k < ids.size(); k++)
      std::copy(in + ids[k] * 2 * F, in + (ids[k] + 1) * 2 * F, tmp.begin() + k * 2 * F);
    b200_solver* sc = g->shard[c];
    if (b200_set_start_fluxes(sc, tmp.data(), (int64_t)tmp.size())) return 1;
    CU(cudaMemcpyAsync(sc->scal.p + SC_PSIMAX, &m, sizeof(double), cudaMemcpyHostToDevice, sc->stream));
    CU(cudaStreamSynchronize(sc->stream));
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* fused drivers                                                              */
/* ------------------------------------------------------------------------- */
static int grp_compute_eigenvalue(b200_solver* s, int max_iters, double tol, int res_type, int32_t* num_iterations) {
  if (res_type == B200_RES_FISSION_SOURCE && s->n_fissionable == 0)
    return fail("The Solver is unable to compute a FISSION_SOURCE residual without fissionable FSRs");
  for (b200_solver* c : grp_shards(s))
    if (b200_eigen_loop_init(c, max_iters, tol)) return 1;
  const int batch = 8;
  int done = 0, i = 0;
  b200_solver* c0 = grp_shards(s)[0];
  while (i < max_iters && !done) {
    const int end = std::min(max_iters, i + batch);
    for (; i < end; i++)
      if (grp_iteration(s, i, res_type, 1)) return 1;
    if (grp_sync(s)) return 1;
    if (fetch_scalars(c0)) return 1;
    done = c0->h_iscal[SI_DONE];
  }
  const int executed = c0->h_iscal[SI_EXEC];
  if (num_iterations) *num_iterations = c0->h_iscal[SI_ITERS];
  for (b200_solver* c : grp_shards(s)) {
    CU(cudaSetDevice(c->cfg.device));
    fix_psi_parity(c, i, executed);
    if (clear_done(c)) return 1;
    if (resolve_events(c)) return 1;
  }
  return 0;
}

static int grp_flux_source_loop(b200_solver* s, int max_iters, double tol, int res_type, bool sources_each_iter,
                                int32_t* num_iterations) {
  for (b200_solver* c : grp_shards(s)) {
    CU(cudaSetDevice(c->cfg.device));
    if (prepare_history(c, max_iters)) return 1;
    CU(cudaMemcpyAsync(c->scal.p + SC_TOL, &tol, 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (clear_done(c)) return 1;
  }
  b200_solver* c0 = grp_shards(s)[0];
  const int batch = 8;
  int done = 0, i = 0;
  while (i < max_iters && !done) {
    const int end = std::min(max_iters, i + batch);
    for (; i < end; i++) {
      if (sources_each_iter)
        for (b200_solver* c : grp_shards(s)) { CU(cudaSetDevice(c->cfg.device)); if (launch_sources(c, i, 0)) return 1; }
      if (grp_sweep(s)) return 1;
      for (b200_solver* c : grp_shards(s)) {
        CU(cudaSetDevice(c->cfg.device));
        if (launch_closure(c, 0, nullptr)) return 1;
        if (launch_residual(c, res_type, 0, 1, 2, i)) return 1;
      }
    }
    if (grp_sync(s)) return 1;
    if (fetch_scalars(c0)) return 1;
    done = c0->h_iscal[SI_DONE];
  }
  const int executed = c0->h_iscal[SI_EXEC], iters = c0->h_iscal[SI_ITERS];
  for (b200_solver* c : grp_shards(s)) {
    CU(cudaSetDevice(c->cfg.device));
    fix_psi_parity(c, i, executed);
    if (clear_done(c)) return 1;
    if (resolve_events(c)) return 1;
  }
  if (num_iterations) *num_iterations = done ? iters : max_iters;
  return 0;
}
