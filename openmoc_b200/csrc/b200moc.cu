/*
 * b200moc.cu - implementation of the C ABI declared in include/b200moc.h.
 *
 * Host side is deliberately thin: device buffers, launch geometry, the fused
 * source-iteration loop, and error translation.  All arithmetic on the hot path
 * lives in sweep.cuh / fsr_kernels.cuh.  There is NO CPU fallback: every entry
 * point fails with an error message if CUDA is unavailable.
 */
#include "../../include/b200moc.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "fsr_kernels.cuh"
#include "sweep.cuh"
#include "sweep_ls.cuh"
#include "otf.cuh"
#include "microbench.cuh"
#include "group.cuh"
#include "ls_prepass.cuh"
#include "cmfd.cuh"

using namespace b200;

/* ------------------------------------------------------------------------- */
/* errors                                                                     */
/* ------------------------------------------------------------------------- */
static thread_local std::string g_last_error;

static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return 1;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__,     \
                  __LINE__, #call);                                                    \
  } while (0)

#define NEED(s)                                                   \
  do {                                                            \
    if ((s) == nullptr) return fail("%s: null solver handle", __func__); \
    CU(cudaSetDevice((s)->cfg.device));                           \
  } while (0)

#define NEED_FINAL(s)                                                         \
  do {                                                                        \
    NEED(s);                                                                  \
    if (!(s)->finalized) return fail("%s: b200_finalize has not been called", __func__); \
  } while (0)

extern "C" const char* b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int b200_version(void) { return 100; }
extern "C" int b200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return -1;
  }
  return n;
}

/* ------------------------------------------------------------------------- */
/* solver object                                                              */
/* ------------------------------------------------------------------------- */
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    if (p != nullptr && n == count) return cudaSuccess;
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc((void**)&p, count * sizeof(T));
  }
  cudaError_t upload(const T* host, size_t count, cudaStream_t st) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess || count == 0) return e;
    return cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct b200_solver {
  b200_config cfg;
  int G = 0, NP = 0, F = 0, A2 = 0;
  int64_t n_trk = 0, n_seg = 0, n_fsr = 0, n_fsr_global = 0, n_fissionable = 0;
  int n_mat = 0;
  bool have_tracks = false, have_quad = false, have_fsrs = false, have_mats = false;
  bool finalized = false;
  cudaStream_t stream = nullptr;
  bool own_stream = false;

  /* host copies needed to derive tables */
  std::vector<int64_t> h_off, h_next_fwd, h_next_bwd;
  std::vector<int32_t> h_azim, h_polar, h_fsr_mat;
  std::vector<uint8_t> h_flags, h_bc_fwd, h_bc_bwd, h_fissionable;
  std::vector<double> h_weight, h_sin, h_sigma_t, h_sigma_s;

  /* device: tracks */
  DevBuf<double> seg_len;
  DevBuf<int32_t> seg_fsr;
  DevBuf<SegRec> seg_rec;
  bool seg_rec_ready = false;         /* the stream was written by the device tracer (no seg_len / seg_fsr copies) */
  bool volumes_from_tracer = false;
  /* axial on-the-fly tracing (otf.cuh) */
  bool otf = false;
  DevBuf<double> otf_seg2d_len, otf_mesh, otf_l0, otf_z0, otf_cos, otf_sin, otf_volw;
  DevBuf<int32_t> otf_seg2d_ext, otf_ext_fsr, otf_trk2d, otf_cls, otf_count;
  DevBuf<int64_t> otf_trk2d_off, otf_ext_off;
  int otf_n_axial = 0;
  double max_tau = 100.;              /* MAX_OPTICAL_LENGTH, src/constants.h:53 */
  DevBuf<double> otf_max_sigt;
  bool otf_split = false;             /* the stream was traced with the optical-length cuts */
  int64_t otf_n_trk2d = 0, otf_n_seg2d = 0, otf_n_ext = 0;
  /* CMFD surfaces from the device tracer (b200_upload_otf_cmfd) */
  bool otf_cmfd = false, otf_cmfd_filled = false;
  DevBuf<int8_t> otf_surf_fwd, otf_surf_bwd;
  DevBuf<int32_t> otf_fsr_cell;
  DevBuf<double> otf_cmfd_z;
  int otf_cmfd_nxy = 0;
  int n_rep = 1;                      /* tally replicas */
  bool capturing = false;             /* inside cudaStreamBeginCapture: no events, no host syncs */
  cudaGraphExec_t iter_graph = nullptr;   /* two fused source iterations (one per psi buffer parity) */
  int iter_graph_res = -1; const float* iter_graph_psi = nullptr; int64_t iter_graph_launches = 0;
  DevBuf<int64_t> trk_off, out_slot;
  DevBuf<int32_t> trk_class, order;
  DevBuf<uint8_t> carry;
  DevBuf<double> cls_w, cls_inv_sin;
  /* device: FSR + materials */
  DevBuf<int32_t> fsr_mat;
  DevBuf<double> vol, sigma_t, sigma_s, fiss, nu_sigma_f, sigma_f, sigma_a, chi, max_ratio, part3;
  DevBuf<float> leakage;
  DevBuf<uint8_t> fissionable;
  /* device: state */
  DevBuf<double> phi, phi_old, fixed, stab, scratch;
  DevBuf<unsigned long long> phi_fx, fx_bits;   /* deterministic mode */
  /* CMFD current tally */
  bool cmfd_on = false, have_cmfd_surf = false;
  int ncg = 0;
  int64_t n_cmfd_slots = 0;           /* n_cells * 26 */
  DevBuf<int32_t> cmfd_fwd, cmfd_bwd, cmfd_group;
  DevBuf<int2> seg_cmfd;
  DevBuf<double> currents;
  struct b200_cmfd* cmfd = nullptr;   /* CMFD solve on the device (cmfd_impl.cuh) */
  /* linear source */
  bool linear = false, have_ls = false;
  int nc = 3;
  DevBuf<double> ls_seg_start, ls_trk_dir, ls_lin_exp, ls_src_const, phi_m, mom_stage, fixed_m, stab_m;
  bool fixed_m_on = false;
  DevBuf<double4> seg_pos, qxyz;
  DevBuf<double2> qst;
  /* padded private copies of the sweep (sweep.cuh: pack_pad_kernel) */
  int GP = 0;                        /* row pitch in groups; == G: no padding */
  bool padded = false;
  DevBuf<double2> qst_pad;
  DevBuf<double4> qxyz_pad;
  DevBuf<double> tally_pad, tallym_pad;
  DevBuf<float> psi_a, psi_b;
  DevBuf<float4> f1tab;               /* B200_PRECISION_TABLE: {a, b, c, 0} per interval (sweep.cuh) */
  float* psi_start = nullptr;  /* what the next sweep reads (= reference _start_flux) */
  float* psi_other = nullptr;
  DevBuf<double> scal, partials, hist_k, hist_res;
  DevBuf<int> iscal;
  double* h_scal = nullptr;   /* pinned mirrors */
  int* h_iscal = nullptr;

  /* sweep launch geometry */
  int gpl = 1, lpi = 1, ipc = 32;   /* groups/thread, threads/item, items/CTA */
  bool smem_attr_set = false;
  bool defer_fx_convert = false;    /* multi-GPU deterministic mode: the host reduces the integers first */
  int64_t sweep_blocks = 0;

  /* options */
  bool fixed_on = false, stabilize = false, neg_allowed = false, balance = false;
  double stab_factor = 1.0;
  int stab_type = 0;

  /* stats */
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
  double sweep_ms = 0.;
  int64_t n_sweeps = 0, n_launches = 0;
  /* shard of a multi-device group: the other shards' start-flux buffers for hand-offs that cross devices */
  std::vector<float*> peer_out;       /* set by the group before every sweep (all shards flip their buffers together) */
  /* several devices behind this handle (group.cuh): the handle itself then owns no device state */
  struct b200_group* grp = nullptr;
};
typedef struct b200_group b200_group;

/* ------------------------------------------------------------------------- */
/* uploads are parked on the host until b200_finalize shards them             */
/* ------------------------------------------------------------------------- */
struct b200_group {
  std::vector<int> devices;
  std::vector<b200_solver*> shard;
  std::vector<std::vector<int64_t>> ids;          /* global track ids of every shard, ascending */
  /* explicit tracks */
  bool have_explicit = false, have_otf = false;
  std::vector<double> seg_length;
  std::vector<int32_t> seg_fsr;
  std::vector<int64_t> trk_off, next_fwd, next_bwd;
  std::vector<int32_t> trk_azim, trk_polar;
  std::vector<uint8_t> flags, bc_fwd, bc_bwd;
  /* axial on-the-fly tracks */
  bool have_geo = false;
  int64_t n_trk2d = 0, n_seg2d = 0, n_ext = 0;
  int32_t n_axial = 0;
  std::vector<double> seg2d_len, ext_mesh, theta, trk_l0, trk_z0;
  std::vector<int32_t> seg2d_ext, ext_fsr, trk_2d;
  std::vector<int64_t> trk2d_off, ext_off;
  bool have_otf_cmfd = false;
  std::vector<int8_t> otf_surf_fwd, otf_surf_bwd;
  std::vector<int32_t> otf_fsr_cell;
  std::vector<double> otf_cmfd_z;
  int otf_cmfd_nx = 0, otf_cmfd_ny = 0, otf_cmfd_nz = 0;
  /* volume tracks of b200_otf_compute_volumes */
  bool have_voltrk = false;
  std::vector<int32_t> v_2d, v_azim, v_polar;
  std::vector<double> v_l0, v_z0, v_weight;
  /* replicated tables */
  std::vector<double> weight, sin_theta, volume, sigma_t, sigma_s, fiss, nu_sigma_f, sigma_f, chi;
  bool have_volume = false, have_sigma_f = false;
  std::vector<int32_t> fsr_mat;
  std::vector<uint8_t> fissionable;
  /* linear source */
  bool have_ls = false;
  std::vector<double> seg_start, trk_dir, lin_exp, src_const;
  /* CMFD */
  bool have_cmfd = false;
  std::vector<int32_t> cmfd_fwd, cmfd_bwd;
  bool cross_links = false;        /* tracks sharded one by one: some hand-offs cross shards (peer stores) */
  /* all-reduce plumbing */
  std::vector<cudaEvent_t> ev_a, ev_b;
  std::vector<DevBuf<double>> stage;
  int64_t n_reduces = 0;
};

/* multi-device groups (group_impl.cuh, included at the end of this file) */
static std::vector<b200_solver*>& grp_shards(b200_solver* s);
static int grp_finalize(b200_solver* s);
static int grp_sweep(b200_solver* s);
static int grp_iteration(b200_solver* s, int i, int res_type, int loop_kind);
static int grp_sync(b200_solver* s);
static void grp_destroy(b200_solver* s);
static void cmfd_destroy(b200_solver* s);
static bool cmfd_in_loop(const b200_solver* s);
static int enqueue_cmfd(b200_solver* s, int moc_iteration, double source_threshold);
static int enqueue_cmfd_threshold(b200_solver* s);
static int cmfd_loop_init(b200_solver* s, double tol);
static int grp_get_start_fluxes(b200_solver* s, float* out, int64_t n);
static int grp_set_start_fluxes(b200_solver* s, const float* in, int64_t n);
static int grp_compute_eigenvalue(b200_solver* s, int max_iters, double tol, int res_type, int32_t* num_iterations);
static int grp_flux_source_loop(b200_solver* s, int max_iters, double tol, int res_type, bool sources_each_iter,
                                int32_t* num_iterations);
/* forward a call to every shard (replicated state) / to the first one (getters); `c` names the shard */
#define GRP_ALL(s, expr)                                              \
  do {                                                                \
    if ((s)->grp != nullptr) {                                        \
      for (b200_solver* c : grp_shards(s)) { if (expr) return 1; }    \
      return 0;                                                       \
    }                                                                 \
  } while (0)
#define GRP_FIRST(s, expr)                                            \
  do {                                                                \
    if ((s)->grp != nullptr) { b200_solver* c = grp_shards(s)[0]; return (expr); } \
  } while (0)

static FsrArgs fsr_args(b200_solver* s) {
  FsrArgs a;
  a.G = s->G;
  a.n_fsr = s->n_fsr;
  a.n_fsr_global = s->n_fsr_global;
  a.n_fissionable = s->n_fissionable;
  a.fsr_mat = s->fsr_mat.p;
  a.vol = s->vol.p;
  a.sigma_t = s->sigma_t.p;
  a.sigma_s = s->sigma_s.p;
  a.fiss = s->fiss.p;
  a.nu_sigma_f = s->nu_sigma_f.p;
  a.sigma_f = s->sigma_f.p;
  a.sigma_a = s->sigma_a.p;
  a.chi = s->chi.p;
  a.fissionable = s->fissionable.p;
  a.phi = s->phi.p;
  a.phi_old = s->phi_old.p;
  a.qst = s->qst.p;
  a.fixed = s->fixed_on ? s->fixed.p : nullptr;
  a.stab = s->stab.p;
  a.scal = s->scal.p;
  a.iscal = s->iscal.p;
  a.partials = s->partials.p;
  return a;
}

static LsArgs ls_args(b200_solver* s) {
  LsArgs l;
  l.nc = s->nc;
  l.solve_3d = s->cfg.solve_3d;
  l.lin_exp = s->ls_lin_exp.p;
  l.src_const = s->ls_src_const.p;
  l.phi_m = s->phi_m.p;
  l.qxyz = s->qxyz.p;
  l.fixed_m = s->fixed_m_on ? s->fixed_m.p : nullptr;
  return l;
}

static inline int grid_for(int64_t n, int threads, int cap = 148 * 8) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}


/* tables derived from the material data: sigma_t half of {q, sigma_t}, fissionable-FSR
 * count (Solver::countFissionableFSRs, src/Solver.cpp:882-892) and the YAMAMOTO
 * max |sigma_s(e,e)/sigma_t(e)| over the FSRs' materials (CPUSolver.cpp:2700-2716) */
static int refresh_material_tables(b200_solver* s) {
  s->n_fissionable = 0;
  for (int64_t r = 0; r < s->n_fsr; r++)
    if (s->h_fissionable[s->h_fsr_mat[r]]) s->n_fissionable++;
  std::vector<double> mr(s->G, 0.);
  std::vector<char> used(s->n_mat, 0);
  for (int64_t r = 0; r < s->n_fsr; r++) used[s->h_fsr_mat[r]] = 1;
  for (int m = 0; m < s->n_mat; m++) {
    if (!used[m]) continue;
    for (int e = 0; e < s->G; e++) {
      double ratio = std::fabs(s->h_sigma_s[((size_t)m * s->G + e) * s->G + e] / s->h_sigma_t[(size_t)m * s->G + e]);
      mr[e] = std::max(mr[e], ratio);
    }
  }
  CU(s->max_ratio.upload(mr.data(), s->G, s->stream));
  FsrArgs a = fsr_args(s);
  fill_sigma_t_kernel<<<grid_for((int64_t)s->n_fsr * s->G, 256, 1 << 30), 256, 0, s->stream>>>(a);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

/* ------------------------------------------------------------------------- */
/* create / destroy / uploads                                                 */
/* ------------------------------------------------------------------------- */
extern "C" int b200_create(const b200_config* cfg, b200_solver** out) {
  if (cfg == nullptr || out == nullptr) return fail("b200_create: null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("b200_create: no CUDA device available (%s); the B200 solver has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail("b200_create: device %d out of range [0,%d)", cfg->device, ndev);
  if (cfg->num_groups < 1) return fail("b200_create: num_groups=%d", cfg->num_groups);
  if (cfg->num_azim < 4 || cfg->num_azim % 4) return fail("b200_create: num_azim=%d must be a positive multiple of 4", cfg->num_azim);
  if (cfg->num_polar < 2 || cfg->num_polar % 2) return fail("b200_create: num_polar=%d must be even", cfg->num_polar);
  if (cfg->n_tracks < 0 || cfg->n_segments < 0 || cfg->n_fsrs < 1 || cfg->n_materials < 1)
    return fail("b200_create: negative or empty problem size");
  if (cfg->precision != B200_PRECISION_DOUBLE && cfg->precision != B200_PRECISION_MIXED && cfg->precision != B200_PRECISION_TABLE)
    return fail("b200_create: unknown precision %d", cfg->precision);
  if (cfg->deterministic && cfg->precision != B200_PRECISION_DOUBLE)
    return fail("b200_create: the deterministic tally needs B200_PRECISION_DOUBLE");
  if (cfg->linear_source && (cfg->deterministic || cfg->precision != B200_PRECISION_DOUBLE))
    return fail("b200_create: the linear-source solver supports B200_PRECISION_DOUBLE with the atomic tally only");
  if (cfg->linear_source && !cfg->solve_3d && cfg->num_polar / 2 > 3)
    return fail("b200_create: linear source supports at most 3 polar angles per 2D track in this build");
  if (cfg->linear_source && cfg->num_groups > 96)
    return fail("b200_create: linear source supports at most 96 energy groups in this build");
  if (!cfg->solve_3d && cfg->num_polar / 2 > 6)
    return fail("b200_create: %d polar angles per 2D track not supported (max 6)", cfg->num_polar / 2);
  if (cfg->num_groups > 256) return fail("b200_create: %d energy groups not supported (max 256)", cfg->num_groups);
  CU(cudaSetDevice(cfg->device));
  b200_solver* s = new b200_solver();
  s->cfg = *cfg;
  s->G = cfg->num_groups;
  s->NP = cfg->solve_3d ? 1 : cfg->num_polar / 2;
  s->F = s->G * s->NP;
  s->A2 = cfg->num_azim / 2;
  s->n_trk = cfg->n_tracks;
  s->n_seg = cfg->n_segments;
  s->n_fsr = cfg->n_fsrs;
  s->n_fsr_global = cfg->n_fsrs_global > 0 ? cfg->n_fsrs_global : cfg->n_fsrs;
  s->n_mat = cfg->n_materials;
  s->linear = cfg->linear_source != 0;
  s->nc = cfg->solve_3d ? 6 : 3;
  /* padded FSR rows (sweep.cuh: pack_pad_kernel), an experiment kept behind B200_PAD=1 */
  s->GP = s->G;
  {
    bool pad = false;     /* measured on the B200: 3D C5G7 48.7 -> 48.5 ms, 2D LS 0.655 -> 0.696 ms: off by default */
    if (const char* e = getenv("B200_PAD")) pad = atoi(e) != 0 && s->G % 8 != 0;
    if (pad) s->GP = (s->G + 7) & ~7;
    s->padded = s->GP != s->G;
  }
  e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete s; return fail("cudaStreamCreate: %s", cudaGetErrorString(e)); }
  s->own_stream = true;
  CU(s->scal.alloc(SC_COUNT_D));
  CU(s->iscal.alloc(SI_COUNT_I));
  CU(s->partials.alloc(MAX_PARTIALS));
  CU(cudaMemsetAsync(s->scal.p, 0, SC_COUNT_D * sizeof(double), s->stream));
  CU(cudaMemsetAsync(s->iscal.p, 0, SI_COUNT_I * sizeof(int), s->stream));
  CU(cudaMallocHost((void**)&s->h_scal, SC_COUNT_D * sizeof(double)));
  CU(cudaMallocHost((void**)&s->h_iscal, SI_COUNT_I * sizeof(int)));
  double one = 1.0;
  CU(cudaMemcpyAsync(s->scal.p + SC_KEFF, &one, sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  *out = s;
  return 0;
}

extern "C" int b200_destroy(b200_solver* s) {
  if (s == nullptr) return 0;
  if (s->grp != nullptr) grp_destroy(s);
  cudaSetDevice(s->cfg.device);
  cudaStreamSynchronize(s->stream);
  if (s->iter_graph != nullptr) cudaGraphExecDestroy(s->iter_graph);
  cmfd_destroy(s);
  for (auto& p : s->ev_pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  for (auto& p : s->ev_free) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  s->seg_len.release(); s->seg_fsr.release(); s->seg_rec.release(); s->trk_off.release(); s->out_slot.release();
  s->trk_class.release(); s->order.release(); s->carry.release(); s->cls_w.release();
  s->cls_inv_sin.release(); s->fsr_mat.release(); s->vol.release(); s->sigma_t.release();
  s->sigma_s.release(); s->fiss.release(); s->nu_sigma_f.release(); s->sigma_f.release();
  s->chi.release(); s->max_ratio.release(); s->sigma_a.release(); s->part3.release(); s->leakage.release(); s->fissionable.release(); s->phi.release();
  s->phi_fx.release(); s->fx_bits.release();
  s->ls_seg_start.release(); s->ls_trk_dir.release(); s->ls_lin_exp.release(); s->ls_src_const.release();
  s->phi_m.release(); s->mom_stage.release(); s->fixed_m.release(); s->stab_m.release(); s->seg_pos.release(); s->qxyz.release();
  s->cmfd_fwd.release(); s->cmfd_bwd.release(); s->cmfd_group.release(); s->seg_cmfd.release(); s->currents.release();
  s->otf_seg2d_len.release(); s->otf_mesh.release(); s->otf_l0.release(); s->otf_z0.release(); s->otf_cos.release();
  s->otf_sin.release(); s->otf_volw.release(); s->otf_seg2d_ext.release(); s->otf_ext_fsr.release(); s->otf_trk2d.release();
  s->otf_surf_fwd.release(); s->otf_surf_bwd.release(); s->otf_fsr_cell.release(); s->otf_cmfd_z.release();
  s->otf_max_sigt.release(); s->otf_cls.release(); s->otf_count.release(); s->otf_trk2d_off.release(); s->otf_ext_off.release();
  s->f1tab.release(); s->qst_pad.release(); s->qxyz_pad.release(); s->tally_pad.release(); s->tallym_pad.release();
  s->phi_old.release(); s->fixed.release(); s->stab.release(); s->scratch.release();
  s->qst.release(); s->psi_a.release(); s->psi_b.release(); s->scal.release();
  s->partials.release(); s->hist_k.release(); s->hist_res.release(); s->iscal.release();
  if (s->h_scal) cudaFreeHost(s->h_scal);
  if (s->h_iscal) cudaFreeHost(s->h_iscal);
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return 0;
}

extern "C" int b200_upload_tracks(b200_solver* s, const double* seg_length, const int32_t* seg_fsr,
                                  const int64_t* trk_seg_offset, const int32_t* trk_azim,
                                  const int32_t* trk_polar, const int64_t* trk_next_fwd,
                                  const int64_t* trk_next_bwd, const uint8_t* trk_flags,
                                  const uint8_t* trk_bc_fwd, const uint8_t* trk_bc_bwd) {
  NEED(s);
  if (!trk_seg_offset || !trk_azim || !trk_polar || !trk_next_fwd || !trk_next_bwd || !trk_flags ||
      !trk_bc_fwd || !trk_bc_bwd || (s->n_seg > 0 && (!seg_length || !seg_fsr)))
    return fail("b200_upload_tracks: null array");
  const int64_t nt = s->n_trk, ns = s->n_seg;
  if (trk_seg_offset[0] != 0 || trk_seg_offset[nt] != ns)
    return fail("b200_upload_tracks: trk_seg_offset must run from 0 to n_segments=%lld (got %lld..%lld)",
                (long long)ns, (long long)trk_seg_offset[0], (long long)trk_seg_offset[nt]);
  for (int64_t t = 0; t < nt; t++) {
    if (trk_seg_offset[t + 1] < trk_seg_offset[t])
      return fail("b200_upload_tracks: trk_seg_offset not monotone at track %lld", (long long)t);
    if (trk_azim[t] < 0 || trk_azim[t] >= s->A2)
      return fail("b200_upload_tracks: track %lld azim index %d outside [0,%d)", (long long)t, trk_azim[t], s->A2);
    if (s->cfg.solve_3d && (trk_polar[t] < 0 || trk_polar[t] >= s->cfg.num_polar))
      return fail("b200_upload_tracks: track %lld polar index %d outside [0,%d)", (long long)t, trk_polar[t], s->cfg.num_polar);
    const uint8_t bcs[2] = {trk_bc_fwd[t], trk_bc_bwd[t]};
    const int64_t nx[2] = {trk_next_fwd[t], trk_next_bwd[t]};
    for (int d = 0; d < 2; d++) {
      if (bcs[d] == B200_BC_INTERFACE)
        return fail("b200_upload_tracks: track %lld ends on a domain INTERFACE; spatial domain "
                    "decomposition is not supported by this build", (long long)t);
      if ((bcs[d] == B200_BC_REFLECTIVE || bcs[d] == B200_BC_PERIODIC) && (nx[d] < 0 || nx[d] >= nt))
        return fail("b200_upload_tracks: track %lld links to track %lld outside [0,%lld)",
                    (long long)t, (long long)nx[d], (long long)nt);
    }
  }
  for (int64_t i = 0; i < ns; i++)
    if (seg_fsr[i] < 0 || seg_fsr[i] >= s->n_fsr)
      return fail("b200_upload_tracks: segment %lld FSR id %d outside [0,%lld)", (long long)i, seg_fsr[i], (long long)s->n_fsr);
  if (s->grp != nullptr) {
    b200_group* g = s->grp;
    g->seg_length.assign(seg_length, seg_length + ns); g->seg_fsr.assign(seg_fsr, seg_fsr + ns);
    g->trk_off.assign(trk_seg_offset, trk_seg_offset + nt + 1);
    g->trk_azim.assign(trk_azim, trk_azim + nt); g->trk_polar.assign(trk_polar, trk_polar + nt);
    g->next_fwd.assign(trk_next_fwd, trk_next_fwd + nt); g->next_bwd.assign(trk_next_bwd, trk_next_bwd + nt);
    g->flags.assign(trk_flags, trk_flags + nt); g->bc_fwd.assign(trk_bc_fwd, trk_bc_fwd + nt); g->bc_bwd.assign(trk_bc_bwd, trk_bc_bwd + nt);
    g->have_explicit = true; g->have_otf = false;
    s->have_tracks = true; s->finalized = false;
    return 0;
  }
  s->seg_rec_ready = false;
  s->otf = false;
  CU(s->seg_len.upload(seg_length, ns, s->stream));
  CU(s->seg_fsr.upload(seg_fsr, ns, s->stream));
  CU(s->trk_off.upload(trk_seg_offset, nt + 1, s->stream));
  s->h_off.assign(trk_seg_offset, trk_seg_offset + nt + 1);
  s->h_azim.assign(trk_azim, trk_azim + nt);
  s->h_polar.assign(trk_polar, trk_polar + nt);
  s->h_next_fwd.assign(trk_next_fwd, trk_next_fwd + nt);
  s->h_next_bwd.assign(trk_next_bwd, trk_next_bwd + nt);
  s->h_flags.assign(trk_flags, trk_flags + nt);
  s->h_bc_fwd.assign(trk_bc_fwd, trk_bc_fwd + nt);
  s->h_bc_bwd.assign(trk_bc_bwd, trk_bc_bwd + nt);
  CU(cudaStreamSynchronize(s->stream));
  s->have_tracks = true;
  s->finalized = false;
  return 0;
}

extern "C" int b200_upload_quadrature(b200_solver* s, const double* weight, const double* sin_theta) {
  NEED(s);
  if (!weight || !sin_theta) return fail("b200_upload_quadrature: null array");
  const size_t n = (size_t)s->A2 * s->cfg.num_polar;
  for (size_t i = 0; i < n; i++)
    if (!(sin_theta[i] > 0.0) || !(sin_theta[i] <= 1.0 + 1e-12))
      return fail("b200_upload_quadrature: sin_theta[%zu]=%g outside (0,1]", i, sin_theta[i]);
  s->h_weight.assign(weight, weight + n);
  s->h_sin.assign(sin_theta, sin_theta + n);
  if (s->grp != nullptr) { s->grp->weight = s->h_weight; s->grp->sin_theta = s->h_sin; }
  s->have_quad = true;
  s->finalized = false;
  return 0;
}

extern "C" int b200_upload_fsrs(b200_solver* s, const double* volume, const int32_t* fsr_material) {
  NEED(s);
  if (!fsr_material) return fail("b200_upload_fsrs: null array");
  if (!volume && !s->volumes_from_tracer)
    return fail("b200_upload_fsrs: volume may only be NULL after b200_otf_compute_volumes");
  for (int64_t r = 0; r < s->n_fsr; r++)
    if (fsr_material[r] < 0 || fsr_material[r] >= s->n_mat)
      return fail("b200_upload_fsrs: FSR %lld material %d outside [0,%d)", (long long)r, fsr_material[r], s->n_mat);
  if (s->grp != nullptr) {
    b200_group* g = s->grp;
    if (volume) { g->volume.assign(volume, volume + s->n_fsr); g->have_volume = true; }
    g->fsr_mat.assign(fsr_material, fsr_material + s->n_fsr);
    s->h_fsr_mat = g->fsr_mat;
    s->have_fsrs = true; s->finalized = false;
    return 0;
  }
  if (volume) { CU(s->vol.upload(volume, s->n_fsr, s->stream)); s->volumes_from_tracer = false; }
  CU(s->fsr_mat.upload(fsr_material, s->n_fsr, s->stream));
  s->h_fsr_mat.assign(fsr_material, fsr_material + s->n_fsr);
  CU(cudaStreamSynchronize(s->stream));
  s->have_fsrs = true;
  s->finalized = false;
  return 0;
}

extern "C" int b200_upload_materials(b200_solver* s, const double* sigma_t, const double* sigma_s,
                                     const double* fiss_matrix, const double* nu_sigma_f,
                                     const double* sigma_f, const double* chi,
                                     const uint8_t* fissionable) {
  NEED(s);
  if (!sigma_t || !sigma_s || !fiss_matrix || !nu_sigma_f || !chi || !fissionable)
    return fail("b200_upload_materials: null array");
  const size_t nG = (size_t)s->n_mat * s->G, nGG = nG * s->G;
  for (size_t i = 0; i < nG; i++)
    if (!(sigma_t[i] > 0.0))
      return fail("b200_upload_materials: sigma_t[%zu]=%g must be positive", i, sigma_t[i]);
  if (s->grp != nullptr) {
    b200_group* g = s->grp;
    if (s->finalized) {          /* refresh (adjoint <-> forward): straight to the shards */
      for (b200_solver* c : grp_shards(s))
        if (b200_upload_materials(c, sigma_t, sigma_s, fiss_matrix, nu_sigma_f, sigma_f, chi, fissionable)) return 1;
      s->n_fissionable = grp_shards(s)[0]->n_fissionable;
      return 0;
    }
    g->sigma_t.assign(sigma_t, sigma_t + nG); g->sigma_s.assign(sigma_s, sigma_s + nGG); g->fiss.assign(fiss_matrix, fiss_matrix + nGG);
    g->nu_sigma_f.assign(nu_sigma_f, nu_sigma_f + nG); g->chi.assign(chi, chi + nG);
    g->fissionable.assign(fissionable, fissionable + s->n_mat);
    g->have_sigma_f = sigma_f != nullptr;
    if (sigma_f) g->sigma_f.assign(sigma_f, sigma_f + nG);
    s->have_mats = true;
    return 0;
  }
  /* sigma_f is optional (only computeFSRFissionRates(nu = false) reads it): NULL keeps a table
   * uploaded earlier - material refreshes (adjoint <-> forward) do not touch it - else zeros */
  std::vector<double> zeros;
  const bool keep_sigma_f = sigma_f == nullptr && s->sigma_f.n == nG;
  if (sigma_f == nullptr && !keep_sigma_f) { zeros.assign(nG, 0.); sigma_f = zeros.data(); }
  CU(s->sigma_t.upload(sigma_t, nG, s->stream));
  CU(s->sigma_s.upload(sigma_s, nGG, s->stream));
  CU(s->fiss.upload(fiss_matrix, nGG, s->stream));
  CU(s->nu_sigma_f.upload(nu_sigma_f, nG, s->stream));
  if (!keep_sigma_f) CU(s->sigma_f.upload(sigma_f, nG, s->stream));
  CU(s->chi.upload(chi, nG, s->stream));
  CU(s->fissionable.upload(fissionable, s->n_mat, s->stream));
  s->h_fissionable.assign(fissionable, fissionable + s->n_mat);
  s->h_sigma_t.assign(sigma_t, sigma_t + nG);
  s->h_sigma_s.assign(sigma_s, sigma_s + nGG);
  {
    /* absorption = total minus out-scatter, the default of Material::getSigmaA (src/Material.cpp:241-249) */
    std::vector<double> sa(nG);
    for (int m = 0; m < s->n_mat; m++)
      for (int g = 0; g < s->G; g++) {
        double v = sigma_t[(size_t)m * s->G + g];
        for (int gp = 0; gp < s->G; gp++) v -= sigma_s[((size_t)m * s->G + gp) * s->G + g];
        sa[(size_t)m * s->G + g] = v;
      }
    CU(s->sigma_a.upload(sa.data(), nG, s->stream));
  }
  CU(cudaStreamSynchronize(s->stream));
  s->have_mats = true;
  /* a finalized solver stays usable: only the material-derived tables are rebuilt
   * (adjoint <-> forward switches re-upload transposed matrices, Solver.cpp:806) */
  if (s->finalized) return refresh_material_tables(s);
  return 0;
}

extern "C" int b200_upload_linear_source(b200_solver* s, const double* seg_start, const double* trk_direction,
                                         const double* lin_exp_matrix, const double* source_constants) {
  NEED(s);
  if (!s->linear) return fail("b200_upload_linear_source: the solver was not created with linear_source = 1");
  if (!trk_direction || !lin_exp_matrix || !source_constants || (s->n_seg > 0 && !seg_start))
    return fail("b200_upload_linear_source: null array");
  if (s->grp != nullptr) {
    b200_group* g = s->grp;
    g->seg_start.assign(seg_start, seg_start + (size_t)s->n_seg * 3);
    g->trk_dir.assign(trk_direction, trk_direction + (size_t)s->n_trk * 3);
    g->lin_exp.assign(lin_exp_matrix, lin_exp_matrix + (size_t)s->n_fsr * s->nc);
    g->src_const.assign(source_constants, source_constants + (size_t)s->n_fsr * s->nc * s->G);
    g->have_ls = true; s->have_ls = true; s->finalized = false;
    return 0;
  }
  CU(s->ls_seg_start.upload(seg_start, (size_t)s->n_seg * 3, s->stream));
  CU(s->ls_trk_dir.upload(trk_direction, (size_t)s->n_trk * 3, s->stream));
  CU(s->ls_lin_exp.upload(lin_exp_matrix, (size_t)s->n_fsr * s->nc, s->stream));
  CU(s->ls_src_const.upload(source_constants, (size_t)s->n_fsr * s->nc * s->G, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->have_ls = true;
  s->finalized = false;
  return 0;
}

extern "C" int b200_upload_cmfd_surfaces(b200_solver* s, const int32_t* seg_cmfd_fwd, const int32_t* seg_cmfd_bwd) {
  NEED(s);
  if (s->n_seg > 0 && (!seg_cmfd_fwd || !seg_cmfd_bwd)) return fail("b200_upload_cmfd_surfaces: null array");
  if (s->grp != nullptr) {
    s->grp->cmfd_fwd.assign(seg_cmfd_fwd, seg_cmfd_fwd + s->n_seg);
    s->grp->cmfd_bwd.assign(seg_cmfd_bwd, seg_cmfd_bwd + s->n_seg);
    s->grp->have_cmfd = true; s->have_cmfd_surf = true;
    return 0;
  }
  CU(s->cmfd_fwd.upload(seg_cmfd_fwd, s->n_seg, s->stream));
  CU(s->cmfd_bwd.upload(seg_cmfd_bwd, s->n_seg, s->stream));
  CU(s->seg_cmfd.alloc((size_t)s->n_seg + 2 * SEG_PAD));
  build_segcmfd_kernel<<<grid_for(s->n_seg + 2 * SEG_PAD, 256), 256, 0, s->stream>>>(
      s->seg_cmfd.p, s->cmfd_fwd.p, s->cmfd_bwd.p, s->n_seg);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  s->cmfd_fwd.release(); s->cmfd_bwd.release();
  s->have_cmfd_surf = true;
  return 0;
}

extern "C" int b200_set_cmfd_groups(b200_solver* s, const int32_t* moc_to_cmfd_group, int32_t num_cmfd_groups,
                                    int64_t num_cmfd_cells) {
  NEED(s);
  if (s->grp != nullptr) {
    if (!s->finalized) return fail("b200_set_cmfd_groups: call b200_finalize first on a multi-device solver");
    for (b200_solver* c : grp_shards(s))
      if (b200_set_cmfd_groups(c, moc_to_cmfd_group, num_cmfd_groups, num_cmfd_cells)) return 1;
    s->cmfd_on = grp_shards(s)[0]->cmfd_on; s->ncg = grp_shards(s)[0]->ncg; s->n_cmfd_slots = grp_shards(s)[0]->n_cmfd_slots;
    return 0;
  }
  if (num_cmfd_groups <= 0) { s->cmfd_on = false; return 0; }      /* switches the tally off */
  if (!s->have_cmfd_surf) return fail("b200_set_cmfd_groups: b200_upload_cmfd_surfaces has not been called");
  if (!moc_to_cmfd_group || num_cmfd_cells <= 0) return fail("b200_set_cmfd_groups: bad argument");
  for (int e = 0; e < s->G; e++)
    if (moc_to_cmfd_group[e] < 0 || moc_to_cmfd_group[e] >= num_cmfd_groups)
      return fail("b200_set_cmfd_groups: MOC group %d maps to CMFD group %d outside [0,%d)", e, moc_to_cmfd_group[e], num_cmfd_groups);
  s->ncg = num_cmfd_groups;
  s->n_cmfd_slots = num_cmfd_cells * 26;     /* NUM_SURFACES, src/constants.h:119 */
  CU(s->cmfd_group.upload(moc_to_cmfd_group, s->G, s->stream));
  CU(s->currents.alloc((size_t)s->n_cmfd_slots * s->ncg));
  CU(cudaMemsetAsync(s->currents.p, 0, (size_t)s->n_cmfd_slots * s->ncg * 8, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->cmfd_on = true;
  return 0;
}

extern "C" int b200_get_cmfd_currents(b200_solver* s, double* out, int64_t n) {
  NEED(s);
  GRP_FIRST(s, b200_get_cmfd_currents(c, out, n));
  if (!s->cmfd_on) return fail("b200_get_cmfd_currents: CMFD tallies are off");
  if (n != s->n_cmfd_slots * s->ncg) return fail("b200_get_cmfd_currents: expected %lld values", (long long)(s->n_cmfd_slots * s->ncg));
  CU(cudaMemcpyAsync(out, s->currents.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

/* ------------------------------------------------------------------------- */
/* axial on-the-fly tracing (otf.cuh)                                          */
/* ------------------------------------------------------------------------- */
static OtfGeom otf_geom(b200_solver* s) {
  OtfGeom g;
  g.seg2d_len = s->otf_seg2d_len.p; g.seg2d_ext = s->otf_seg2d_ext.p; g.trk2d_off = s->otf_trk2d_off.p;
  g.ext_off = s->otf_n_ext > 0 ? s->otf_ext_off.p : nullptr;
  g.ext_mesh = s->otf_mesh.p; g.ext_fsr = s->otf_ext_fsr.p; g.n_axial = s->otf_n_axial;
  g.trk_2d = s->otf_trk2d.p; g.trk_l0 = s->otf_l0.p; g.trk_z0 = s->otf_z0.p; g.trk_class = s->otf_cls.p;
  g.cls_cos_theta = s->otf_cos.p; g.cls_sin_theta = s->otf_sin.p;
  g.seg2d_surf_fwd = s->otf_cmfd ? s->otf_surf_fwd.p : nullptr; g.seg2d_surf_bwd = s->otf_surf_bwd.p;
  g.fsr_cmfd_cell = s->otf_fsr_cell.p; g.cmfd_z = s->otf_cmfd_z.p; g.cmfd_nxy = s->otf_cmfd_nxy;
  g.n_trk = s->n_trk;
  g.fsr_max_sigma_t = nullptr; g.max_tau = s->max_tau;
  return g;
}

extern "C" int b200_upload_otf_geometry(b200_solver* s, int64_t n_tracks_2d, int64_t n_segments_2d,
                                        const double* seg2d_length, const int32_t* seg2d_extruded_fsr,
                                        const int64_t* trk2d_seg_offset, int64_t n_extruded_fsrs,
                                        const int64_t* ext_offset, const double* ext_mesh,
                                        const int32_t* ext_fsr_ids, int32_t n_axial_global,
                                        const double* theta /* [A/2][P] */) {
  NEED(s);
  if (!s->cfg.solve_3d) return fail("b200_upload_otf_geometry: axial on-the-fly tracing is for 3D solvers");
  if (n_tracks_2d < 1 || n_segments_2d < 0 || !seg2d_length || !seg2d_extruded_fsr || !trk2d_seg_offset || !ext_mesh || !theta)
    return fail("b200_upload_otf_geometry: null or empty argument");
  if (trk2d_seg_offset[0] != 0 || trk2d_seg_offset[n_tracks_2d] != n_segments_2d)
    return fail("b200_upload_otf_geometry: trk2d_seg_offset must run from 0 to n_segments_2d");
  if (s->grp != nullptr) {
    b200_group* g = s->grp;
    const bool glob = n_extruded_fsrs <= 0;
    if (glob && n_axial_global < 1) return fail("b200_upload_otf_geometry: a global axial mesh needs n_axial_global >= 1");
    if (!glob && (!ext_offset || !ext_fsr_ids)) return fail("b200_upload_otf_geometry: per-FSR axial meshes need ext_offset and ext_fsr_ids");
    g->n_trk2d = n_tracks_2d; g->n_seg2d = n_segments_2d; g->n_ext = glob ? 0 : n_extruded_fsrs; g->n_axial = glob ? n_axial_global : 0;
    g->seg2d_len.assign(seg2d_length, seg2d_length + n_segments_2d);
    g->seg2d_ext.assign(seg2d_extruded_fsr, seg2d_extruded_fsr + n_segments_2d);
    g->trk2d_off.assign(trk2d_seg_offset, trk2d_seg_offset + n_tracks_2d + 1);
    if (glob) {
      g->ext_mesh.assign(ext_mesh, ext_mesh + n_axial_global + 1);
    } else {
      const int64_t ntot = ext_offset[n_extruded_fsrs];
      g->ext_off.assign(ext_offset, ext_offset + n_extruded_fsrs + 1);
      g->ext_mesh.assign(ext_mesh, ext_mesh + ntot + n_extruded_fsrs);
      g->ext_fsr.assign(ext_fsr_ids, ext_fsr_ids + ntot);
    }
    g->theta.assign(theta, theta + (size_t)s->A2 * s->cfg.num_polar);
    g->have_geo = true;
    s->otf_n_trk2d = n_tracks_2d;
    return 0;            /* validated when the shards upload it */
  }
  const bool global = n_extruded_fsrs <= 0;
  int64_t n_ext_seen = 0;
  for (int64_t i = 0; i < n_segments_2d; i++) {
    if (seg2d_extruded_fsr[i] < 0) return fail("b200_upload_otf_geometry: negative extruded FSR id");
    n_ext_seen = std::max<int64_t>(n_ext_seen, seg2d_extruded_fsr[i] + 1);
    if (!(seg2d_length[i] >= 0.0)) return fail("b200_upload_otf_geometry: 2D segment %lld has length %g", (long long)i, seg2d_length[i]);
  }
  if (global) {
    if (n_axial_global < 1) return fail("b200_upload_otf_geometry: a global axial mesh needs n_axial_global >= 1");
    if (n_ext_seen * n_axial_global > s->n_fsr)
      return fail("b200_upload_otf_geometry: %lld extruded FSRs x %d layers exceed n_fsrs = %lld", (long long)n_ext_seen,
                  n_axial_global, (long long)s->n_fsr);
    for (int k = 0; k < n_axial_global; k++)
      if (!(ext_mesh[k + 1] > ext_mesh[k])) return fail("b200_upload_otf_geometry: the axial mesh must increase");
    CU(s->otf_mesh.upload(ext_mesh, n_axial_global + 1, s->stream));
    s->otf_n_ext = 0;
  } else {
    if (!ext_offset || !ext_fsr_ids) return fail("b200_upload_otf_geometry: per-FSR axial meshes need ext_offset and ext_fsr_ids");
    if (n_ext_seen > n_extruded_fsrs) return fail("b200_upload_otf_geometry: a 2D segment refers to extruded FSR %lld of %lld",
                                                  (long long)n_ext_seen - 1, (long long)n_extruded_fsrs);
    if (ext_offset[0] != 0) return fail("b200_upload_otf_geometry: ext_offset must start at 0");
    for (int64_t e = 0; e < n_extruded_fsrs; e++) {
      const int64_t n = ext_offset[e + 1] - ext_offset[e];
      if (n < 1) return fail("b200_upload_otf_geometry: extruded FSR %lld has no axial FSR", (long long)e);
      for (int64_t k = 0; k < n; k++) {
        const int32_t f = ext_fsr_ids[ext_offset[e] + k];
        if (f < 0 || f >= s->n_fsr) return fail("b200_upload_otf_geometry: FSR id %d outside [0,%lld)", f, (long long)s->n_fsr);
        if (!(ext_mesh[ext_offset[e] + e + k + 1] > ext_mesh[ext_offset[e] + e + k]))
          return fail("b200_upload_otf_geometry: the axial mesh of extruded FSR %lld must increase", (long long)e);
      }
    }
    const int64_t ntot = ext_offset[n_extruded_fsrs];
    CU(s->otf_ext_off.upload(ext_offset, n_extruded_fsrs + 1, s->stream));
    CU(s->otf_mesh.upload(ext_mesh, ntot + n_extruded_fsrs, s->stream));
    CU(s->otf_ext_fsr.upload(ext_fsr_ids, ntot, s->stream));
    s->otf_n_ext = n_extruded_fsrs;
  }
  s->otf_n_axial = global ? n_axial_global : 0;
  CU(s->otf_seg2d_len.upload(seg2d_length, n_segments_2d, s->stream));
  CU(s->otf_seg2d_ext.upload(seg2d_extruded_fsr, n_segments_2d, s->stream));
  CU(s->otf_trk2d_off.upload(trk2d_seg_offset, n_tracks_2d + 1, s->stream));
  const size_t ncls = (size_t)s->A2 * s->cfg.num_polar;
  std::vector<double> c(ncls), sn(ncls);
  for (size_t i = 0; i < ncls; i++) {
    if (!(theta[i] > 0.0 && theta[i] < M_PI) || theta[i] == M_PI_2)
      return fail("b200_upload_otf_geometry: polar angle %g of class %zu outside (0, pi) or horizontal", theta[i], i);
    c[i] = cos(theta[i]); sn[i] = sin(theta[i]);
  }
  CU(s->otf_cos.upload(c.data(), ncls, s->stream));
  CU(s->otf_sin.upload(sn.data(), ncls, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->otf_n_trk2d = n_tracks_2d; s->otf_n_seg2d = n_segments_2d;
  return 0;
}

struct OtfStarts { DevBuf<int32_t> trk2d, cls; DevBuf<double> l0, z0; };
static int otf_upload_track_starts(b200_solver* s, int64_t n, const int32_t* trk_2d, const double* trk_l0,
                                   const double* trk_z0, const int32_t* trk_azim, const int32_t* trk_polar,
                                   DevBuf<int32_t>& d_trk2d, DevBuf<double>& d_l0, DevBuf<double>& d_z0, DevBuf<int32_t>& d_cls) {
  std::vector<int32_t> cls(n);
  for (int64_t t = 0; t < n; t++) {
    if (trk_2d[t] < 0 || trk_2d[t] >= s->otf_n_trk2d)
      return fail("axial tracer: track %lld lies over 2D track %d outside [0,%lld)", (long long)t, trk_2d[t], (long long)s->otf_n_trk2d);
    if (trk_azim[t] < 0 || trk_azim[t] >= s->A2 || trk_polar[t] < 0 || trk_polar[t] >= s->cfg.num_polar)
      return fail("axial tracer: track %lld has angle indices (%d, %d) out of range", (long long)t, trk_azim[t], trk_polar[t]);
    cls[t] = trk_azim[t] * s->cfg.num_polar + trk_polar[t];
  }
  CU(d_trk2d.upload(trk_2d, n, s->stream));
  CU(d_l0.upload(trk_l0, n, s->stream));
  CU(d_z0.upload(trk_z0, n, s->stream));
  CU(d_cls.upload(cls.data(), n, s->stream));
  CU(cudaStreamSynchronize(s->stream));      /* cls is a local */
  return 0;
}

extern "C" int b200_otf_compute_volumes(b200_solver* s, int64_t n, const int32_t* trk_2d, const double* trk_l0,
                                        const double* trk_z0, const int32_t* trk_azim, const int32_t* trk_polar,
                                        const double* class_weight /* [A/2][P] */) {
  NEED(s);
  if (s->otf_n_trk2d == 0) return fail("b200_otf_compute_volumes: b200_upload_otf_geometry has not been called");
  if (n < 0 || (n > 0 && (!trk_2d || !trk_l0 || !trk_z0 || !trk_azim || !trk_polar)) || !class_weight)
    return fail("b200_otf_compute_volumes: null argument");
  if (s->grp != nullptr) {
    b200_group* g = s->grp;
    g->v_2d.assign(trk_2d, trk_2d + n); g->v_l0.assign(trk_l0, trk_l0 + n); g->v_z0.assign(trk_z0, trk_z0 + n);
    g->v_azim.assign(trk_azim, trk_azim + n); g->v_polar.assign(trk_polar, trk_polar + n);
    g->v_weight.assign(class_weight, class_weight + (size_t)s->A2 * s->cfg.num_polar);
    g->have_voltrk = true; g->have_volume = false;
    s->volumes_from_tracer = true;
    return 0;
  }
  /* the tracks given here (normally ALL tracks of the problem) are independent of the solver's own
   * (possibly sharded) track set: temporary buffers */
  OtfStarts tmp;
  struct Guard { OtfStarts& t; ~Guard() { t.trk2d.release(); t.cls.release(); t.l0.release(); t.z0.release(); } } guard{tmp};
  if (otf_upload_track_starts(s, n, trk_2d, trk_l0, trk_z0, trk_azim, trk_polar, tmp.trk2d, tmp.l0, tmp.z0, tmp.cls)) return 1;
  CU(s->otf_volw.upload(class_weight, (size_t)s->A2 * s->cfg.num_polar, s->stream));
  CU(s->vol.alloc(s->n_fsr));
  CU(cudaMemsetAsync(s->vol.p, 0, (size_t)s->n_fsr * 8, s->stream));
  OtfGeom g = otf_geom(s);
  g.trk_2d = tmp.trk2d.p; g.trk_l0 = tmp.l0.p; g.trk_z0 = tmp.z0.p; g.trk_class = tmp.cls.p;
  g.n_trk = n;
  if (n > 0) {
    otf_fill_kernel<<<grid_for(n, 128, 1 << 20), 128, 0, s->stream>>>(g, nullptr, nullptr, s->G, s->otf_volw.p, s->vol.p, nullptr);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(s->stream));
  s->volumes_from_tracer = true;
  return 0;
}

/* number of 3D segments of arbitrary tracks over the uploaded geometry (work estimate for a partition) */
extern "C" int b200_otf_count_segments(b200_solver* s, int64_t n, const int32_t* trk_2d, const double* trk_l0,
                                       const double* trk_z0, const int32_t* trk_azim, const int32_t* trk_polar,
                                       int32_t* counts) {
  NEED(s);
  if (s->grp != nullptr) return fail("b200_otf_count_segments: single-device solvers only");
  if (s->otf_n_trk2d == 0) return fail("b200_otf_count_segments: b200_upload_otf_geometry has not been called");
  if (n < 0 || (n > 0 && (!trk_2d || !trk_l0 || !trk_z0 || !trk_azim || !trk_polar || !counts)))
    return fail("b200_otf_count_segments: null argument");
  if (n == 0) return 0;
  OtfStarts tmp;
  struct Guard { OtfStarts& t; ~Guard() { t.trk2d.release(); t.cls.release(); t.l0.release(); t.z0.release(); } } guard{tmp};
  if (otf_upload_track_starts(s, n, trk_2d, trk_l0, trk_z0, trk_azim, trk_polar, tmp.trk2d, tmp.l0, tmp.z0, tmp.cls)) return 1;
  DevBuf<int32_t> cnt;
  CU(cnt.alloc(n));
  OtfGeom g = otf_geom(s);
  g.trk_2d = tmp.trk2d.p; g.trk_l0 = tmp.l0.p; g.trk_z0 = tmp.z0.p; g.trk_class = tmp.cls.p;
  g.n_trk = n;
  otf_count_kernel<<<grid_for(n, 128, 1 << 20), 128, 0, s->stream>>>(g, cnt.p);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(counts, cnt.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cnt.release();
  CU(e);
  return 0;
}

/* count + fill of the solver's own tracks.  with_cuts: cut pieces longer than the maximum optical length
 * (needs the FSR materials, so b200_finalize repeats the expansion when the cuts change the count) */
static int otf_expand(b200_solver* s, bool with_cuts) {
  const int64_t nt = s->n_trk;
  CU(s->otf_count.alloc(std::max<int64_t>(nt, 1)));
  OtfGeom g = otf_geom(s);
  if (with_cuts) g.fsr_max_sigma_t = s->otf_max_sigt.p;
  std::vector<int32_t> cnt(nt);
  if (nt > 0) {
    otf_count_kernel<<<grid_for(nt, 128, 1 << 20), 128, 0, s->stream>>>(g, s->otf_count.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(cnt.data(), s->otf_count.p, nt * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
  }
  CU(cudaStreamSynchronize(s->stream));
  int64_t ns = 0;
  for (int64_t t = 0; t < nt; t++) ns += cnt[t];
  if (with_cuts && ns == s->n_seg && s->seg_rec_ready && (!s->otf_cmfd || s->otf_cmfd_filled))
    return 0;                                                         /* nothing to cut: the stream stands */
  s->h_off.assign(nt + 1, 0);
  for (int64_t t = 0; t < nt; t++) s->h_off[t + 1] = s->h_off[t] + cnt[t];
  s->n_seg = ns;
  s->cfg.n_segments = ns;
  CU(s->trk_off.upload(s->h_off.data(), nt + 1, s->stream));
  /* pass 2: the device segment stream, written in place */
  s->seg_len.release(); s->seg_fsr.release();
  CU(s->seg_rec.alloc((size_t)ns + 2 * SEG_PAD));
  otf_pad_kernel<<<1, 2 * SEG_PAD, 0, s->stream>>>(s->seg_rec.p, ns);
  CU(cudaGetLastError());
  if (s->otf_cmfd) {
    /* the {forward, backward} CMFD surface stream of the sweep's current tally, padded like the records */
    CU(s->seg_cmfd.alloc((size_t)ns + 2 * SEG_PAD));
    CU(cudaMemsetAsync(s->seg_cmfd.p, 0xff, ((size_t)ns + 2 * SEG_PAD) * sizeof(int2), s->stream));
  }
  if (nt > 0) {
    otf_fill_kernel<<<grid_for(nt, 128, 1 << 20), 128, 0, s->stream>>>(g, s->trk_off.p, s->seg_rec.p + SEG_PAD, s->GP, nullptr, nullptr,
                                                                       s->otf_cmfd ? s->seg_cmfd.p + SEG_PAD : nullptr);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(s->stream));
  if (s->otf_cmfd) { s->otf_cmfd_filled = true; s->have_cmfd_surf = true; }
  s->otf_split = with_cuts;
  return 0;
}

/* CMFD with axially traced tracks: the tracer also produces segment::_cmfd_surface_fwd/_bwd of every 3D segment */
extern "C" int b200_upload_otf_cmfd(b200_solver* s, const int8_t* seg2d_surface_fwd, const int8_t* seg2d_surface_bwd,
                                    const int32_t* fsr_cmfd_cell, int32_t num_x, int32_t num_y, int32_t num_z,
                                    const double* z_planes) {
  NEED(s);
  if (s->otf_n_seg2d == 0 && (s->grp == nullptr || !s->grp->have_geo))
    return fail("b200_upload_otf_cmfd: b200_upload_otf_geometry has not been called");
  if (!seg2d_surface_fwd || !seg2d_surface_bwd || !fsr_cmfd_cell || !z_planes) return fail("b200_upload_otf_cmfd: null argument");
  if (num_x < 1 || num_y < 1 || num_z < 1) return fail("b200_upload_otf_cmfd: empty CMFD mesh");
  const int64_t ns2 = s->grp != nullptr ? s->grp->n_seg2d : s->otf_n_seg2d;
  const int64_t n_cells = (int64_t)num_x * num_y * num_z;
  for (int64_t i = 0; i < ns2; i++)
    if (seg2d_surface_fwd[i] < -1 || seg2d_surface_fwd[i] > 9 || seg2d_surface_bwd[i] < -1 || seg2d_surface_bwd[i] > 9)
      return fail("b200_upload_otf_cmfd: 2D segment %lld crosses surface %d / %d; a radial segment can only cross x / y faces and edges (0..9)",
                  (long long)i, seg2d_surface_fwd[i], seg2d_surface_bwd[i]);
  for (int64_t r = 0; r < s->n_fsr; r++)
    if (fsr_cmfd_cell[r] < 0 || fsr_cmfd_cell[r] >= n_cells)
      return fail("b200_upload_otf_cmfd: FSR %lld lies in CMFD cell %d outside [0,%lld)", (long long)r, fsr_cmfd_cell[r], (long long)n_cells);
  for (int k = 0; k < num_z; k++)
    if (!(z_planes[k + 1] > z_planes[k])) return fail("b200_upload_otf_cmfd: the z planes of the CMFD mesh must increase");
  if (s->grp != nullptr) {
    b200_group* g = s->grp;
    g->otf_surf_fwd.assign(seg2d_surface_fwd, seg2d_surface_fwd + ns2);
    g->otf_surf_bwd.assign(seg2d_surface_bwd, seg2d_surface_bwd + ns2);
    g->otf_fsr_cell.assign(fsr_cmfd_cell, fsr_cmfd_cell + s->n_fsr);
    g->otf_cmfd_z.assign(z_planes, z_planes + num_z + 1);
    g->otf_cmfd_nx = num_x; g->otf_cmfd_ny = num_y; g->otf_cmfd_nz = num_z;
    g->have_otf_cmfd = true;
    s->have_cmfd_surf = true;
    s->finalized = false;
    return 0;
  }
  CU(s->otf_surf_fwd.upload(seg2d_surface_fwd, ns2, s->stream));
  CU(s->otf_surf_bwd.upload(seg2d_surface_bwd, ns2, s->stream));
  CU(s->otf_fsr_cell.upload(fsr_cmfd_cell, s->n_fsr, s->stream));
  CU(s->otf_cmfd_z.upload(z_planes, num_z + 1, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->otf_cmfd_nxy = num_x * num_y;
  s->otf_cmfd = true;
  s->otf_cmfd_filled = false;       /* the next expansion (b200_upload_tracks_otf / b200_finalize) writes the surface stream */
  s->finalized = false;
  return 0;
}

/* Solver::setMaxOpticalLength / TrackGenerator::retrieveMaxOpticalLength: segments traced on the device
 * are cut at this optical length like the reference's on-the-fly kernels cut them */
extern "C" int b200_set_max_optical_length(b200_solver* s, double max_tau) {
  NEED(s);
  if (!(max_tau > 0.)) return fail("b200_set_max_optical_length: the maximum optical length must be positive");
  s->max_tau = max_tau;
  if (s->grp != nullptr && s->finalized)
    for (b200_solver* c : grp_shards(s)) c->max_tau = max_tau;
  s->finalized = false;
  return 0;
}

extern "C" int b200_upload_tracks_otf(b200_solver* s, const int32_t* trk_2d, const double* trk_l0, const double* trk_z0,
                                      const int32_t* trk_azim, const int32_t* trk_polar,
                                      const int64_t* trk_next_fwd, const int64_t* trk_next_bwd,
                                      const uint8_t* trk_flags, const uint8_t* trk_bc_fwd, const uint8_t* trk_bc_bwd,
                                      int64_t* n_segments_out) {
  NEED(s);
  if (s->otf_n_trk2d == 0) return fail("b200_upload_tracks_otf: b200_upload_otf_geometry has not been called");
  if (s->linear) return fail("b200_upload_tracks_otf: the linear source needs explicit segments (starting points) in this build");
  const int64_t nt = s->n_trk;
  if (nt > 0 && (!trk_2d || !trk_l0 || !trk_z0 || !trk_azim || !trk_polar || !trk_next_fwd || !trk_next_bwd ||
                 !trk_flags || !trk_bc_fwd || !trk_bc_bwd))
    return fail("b200_upload_tracks_otf: null array");
  if (s->grp != nullptr) {
    b200_group* g = s->grp;
    g->trk_2d.assign(trk_2d, trk_2d + nt); g->trk_l0.assign(trk_l0, trk_l0 + nt); g->trk_z0.assign(trk_z0, trk_z0 + nt);
    g->trk_azim.assign(trk_azim, trk_azim + nt); g->trk_polar.assign(trk_polar, trk_polar + nt);
    g->next_fwd.assign(trk_next_fwd, trk_next_fwd + nt); g->next_bwd.assign(trk_next_bwd, trk_next_bwd + nt);
    g->flags.assign(trk_flags, trk_flags + nt); g->bc_fwd.assign(trk_bc_fwd, trk_bc_fwd + nt); g->bc_bwd.assign(trk_bc_bwd, trk_bc_bwd + nt);
    g->have_otf = true; g->have_explicit = false;
    s->have_tracks = true; s->finalized = false;
    if (n_segments_out) *n_segments_out = 0;       /* known after b200_finalize: b200_get_num_segments */
    return 0;
  }
  for (int64_t t = 0; t < nt; t++) {
    const uint8_t bcs[2] = {trk_bc_fwd[t], trk_bc_bwd[t]};
    const int64_t nx[2] = {trk_next_fwd[t], trk_next_bwd[t]};
    for (int d = 0; d < 2; d++) {
      if (bcs[d] == B200_BC_INTERFACE)
        return fail("b200_upload_tracks_otf: track %lld ends on a domain INTERFACE; spatial domain "
                    "decomposition is not supported by this build", (long long)t);
      if ((bcs[d] == B200_BC_REFLECTIVE || bcs[d] == B200_BC_PERIODIC) && (nx[d] < 0 || nx[d] >= nt))
        return fail("b200_upload_tracks_otf: track %lld links to track %lld outside [0,%lld)",
                    (long long)t, (long long)nx[d], (long long)nt);
    }
  }
  if (otf_upload_track_starts(s, nt, trk_2d, trk_l0, trk_z0, trk_azim, trk_polar, s->otf_trk2d, s->otf_l0, s->otf_z0, s->otf_cls)) return 1;
  if (otf_expand(s, false)) return 1;
  const int64_t ns = s->n_seg;
  s->h_azim.assign(trk_azim, trk_azim + nt);
  s->h_polar.assign(trk_polar, trk_polar + nt);
  s->h_next_fwd.assign(trk_next_fwd, trk_next_fwd + nt);
  s->h_next_bwd.assign(trk_next_bwd, trk_next_bwd + nt);
  s->h_flags.assign(trk_flags, trk_flags + nt);
  s->h_bc_fwd.assign(trk_bc_fwd, trk_bc_fwd + nt);
  s->h_bc_bwd.assign(trk_bc_bwd, trk_bc_bwd + nt);
  CU(cudaStreamSynchronize(s->stream));
  s->seg_rec_ready = true;
  s->otf = true;
  s->have_tracks = true;
  s->finalized = false;
  if (n_segments_out) *n_segments_out = ns;
  return 0;
}

extern "C" int b200_get_num_segments(b200_solver* s, int64_t* n_segments) {
  NEED(s);
  if (n_segments) *n_segments = s->n_seg;
  return 0;
}

extern "C" int b200_get_segments(b200_solver* s, double* seg_length, int32_t* seg_fsr, int64_t n, int64_t* trk_seg_offset) {
  NEED(s);
  if (s->grp != nullptr) return fail("b200_get_segments: ask the shards of a multi-device solver");
  if (!s->have_tracks) return fail("b200_get_segments: no tracks uploaded");
  if (n != s->n_seg) return fail("b200_get_segments: %lld segments requested, the solver holds %lld", (long long)n, (long long)s->n_seg);
  if (trk_seg_offset) memcpy(trk_seg_offset, s->h_off.data(), (s->n_trk + 1) * sizeof(int64_t));
  if (n == 0) return 0;
  if (!s->seg_rec_ready) {
    if (seg_length) CU(cudaMemcpyAsync(seg_length, s->seg_len.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
    if (seg_fsr) CU(cudaMemcpyAsync(seg_fsr, s->seg_fsr.p, n * 4, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
  }
  struct Tmp { void* p = nullptr; ~Tmp() { if (p) cudaFree(p); } } dl, df;
  CU(cudaMalloc(&dl.p, n * 8));
  CU(cudaMalloc(&df.p, n * 4));
  otf_unpack_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(s->seg_rec.p + SEG_PAD, n, s->GP, (double*)dl.p, (int32_t*)df.p);
  CU(cudaGetLastError());
  if (seg_length) CU(cudaMemcpyAsync(seg_length, dl.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
  if (seg_fsr) CU(cudaMemcpyAsync(seg_fsr, df.p, n * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

extern "C" int b200_get_volumes(b200_solver* s, double* out, int64_t n) {
  NEED(s);
  GRP_FIRST(s, b200_get_volumes(c, out, n));
  if (n != s->n_fsr || s->vol.n != (size_t)s->n_fsr) return fail("b200_get_volumes: size mismatch or no volumes yet");
  CU(cudaMemcpyAsync(out, s->vol.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

/* Groups per thread (GPL) and threads per item (LPI = ceil(G/GPL) <= 32).  With the
 * flat thread->item mapping every LPI fills the warps, so prefer the smallest GPL
 * (fewest registers, most resident warps) whose slots are >= 95 % used. */
static void choose_lane_map(int G, int* gpl, int* lpi, int* ipc) {
  static const int cand[] = {1, 2, 3, 4, 7, 8};
  const char* force = getenv("B200_GPL");
  double best = -1.;
  *gpl = 0;
  for (int c : cand) {
    if (force != nullptr && atoi(force) != c) continue;
    int l = (G + c - 1) / c;
    if (l > 32 && force == nullptr) continue;   /* (a forced GPL may exceed one warp per item: flat mapping) */
    double util = (double)G / ((double)c * l);
    if (util >= 0.95) { *gpl = c; *lpi = l; break; }
    if (util > best) { best = util; *gpl = c; *lpi = l; }
  }
  if (*gpl == 0) { *gpl = 8; *lpi = (G + 7) / 8; }
  /* items per CTA: CTAs of at most 224 threads (the kernels' launch bound) */
  *ipc = 224 / *lpi;
  if (*ipc > 32) *ipc = 32;
  const char* fipc = getenv("B200_IPC");
  if (fipc != nullptr && atoi(fipc) > 0 && atoi(fipc) * *lpi <= 224) *ipc = atoi(fipc);
}

extern "C" int b200_finalize(b200_solver* s) {
  NEED(s);
  if (s->grp != nullptr) return grp_finalize(s);
  if (!s->have_tracks || !s->have_quad || !s->have_fsrs || !s->have_mats)
    return fail("b200_finalize: tracks, quadrature, FSRs and materials must all be uploaded first");
  const int64_t nt = s->n_trk;
  const int NP = s->NP, P = s->cfg.num_polar, A = s->cfg.num_azim;

  if (s->otf) {
    /* segments longer than the maximum optical length are cut (MOCKernel.cpp:216-268): now that the FSR
     * materials are known, recount with the cuts and retrace if they change anything */
    std::vector<double> mmax(s->n_mat, 0.), fmax(s->n_fsr);
    for (int m = 0; m < s->n_mat; m++)
      for (int e = 0; e < s->G; e++) mmax[m] = std::max(mmax[m], s->h_sigma_t[(size_t)m * s->G + e]);
    for (int64_t r = 0; r < s->n_fsr; r++) fmax[r] = mmax[s->h_fsr_mat[r]];
    CU(s->otf_max_sigt.upload(fmax.data(), s->n_fsr, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (otf_expand(s, true)) return 1;
  }

  /* angle classes: 2D -> azim; 3D -> (azim, polar).  2D inverse sines follow the
   * ExpEvaluator sharing rule (src/Solver.cpp:763-779): azim a >= A/4 uses A/2-1-a. */
  const int n_class = s->cfg.solve_3d ? s->A2 * P : s->A2;
  std::vector<double> cw((size_t)n_class * NP), cis((size_t)n_class * NP);
  for (int a = 0; a < s->A2; a++) {
    if (s->cfg.solve_3d) {
      for (int p = 0; p < P; p++) {
        cw[a * P + p] = s->h_weight[a * P + p];
        cis[a * P + p] = 1.0;  /* 3D segments carry the 3D length: F1(tau) directly (CPUSolver.cpp:2419-2433) */
      }
    } else {
      int ae = a;
      if (ae >= A / 4) ae = A / 2 - 1 - a;
      for (int p = 0; p < NP; p++) {
        cw[a * NP + p] = s->h_weight[a * P + p];
        cis[a * NP + p] = 1.0 / s->h_sin[ae * P + p];
      }
    }
  }
  std::vector<int32_t> cls(nt);
  for (int64_t t = 0; t < nt; t++)
    cls[t] = s->cfg.solve_3d ? s->h_azim[t] * P + s->h_polar[t] : s->h_azim[t];

  /* boundary hand-off table and copy-through flags */
  std::vector<int64_t> out_slot(2 * nt, -1);
  std::vector<uint8_t> carry(2 * nt, 1);
  for (int64_t t = 0; t < nt; t++) {
    for (int d = 0; d < 2; d++) {
      uint8_t bc = d == 0 ? s->h_bc_fwd[t] : s->h_bc_bwd[t];
      if (bc != B200_BC_REFLECTIVE && bc != B200_BC_PERIODIC) continue;
      int64_t nx = d == 0 ? s->h_next_fwd[t] : s->h_next_bwd[t];
      int next_is_fwd = d == 0 ? (s->h_flags[t] & 1) : ((s->h_flags[t] >> 1) & 1);
      int64_t slot = nx * 2 + (next_is_fwd ? 0 : 1);   /* _start_flux(next, !next_is_fwd) CPUSolver.cpp:2574-2588 */
      out_slot[t * 2 + d] = slot;
      /* two ends feeding one start-flux slot would race in the sweep (the reference's serial
       * memcpy lets the later track win); cyclic tracking never produces it */
      if (carry[slot] == 0)
        return fail("b200_finalize: two track ends hand their flux to the same slot (track %lld, %s start)",
                    (long long)nx, next_is_fwd ? "forward" : "backward");
      carry[slot] = 0;
    }
  }

  std::vector<int32_t> order(nt);
  std::iota(order.begin(), order.end(), 0);
  /* 2D decks: longest tracks first (stable sort): the last wave of CTAs is then made of short tracks
   * and the tail shrinks (+2 % on 82.8 M segments, more on the 10 M-segment shards of an 8-GPU
   * run); neighbours in the sorted order are still mostly neighbouring tracks, which cross the
   * same FSRs.  3D decks keep the Track uid order (azimuthal angle, 2D track, polar angle, position
   * in the z-stack): the tracks resident on the GPU at any time then belong to a few neighbouring
   * z-stacks, i.e. a thin slab of the core whose FSR rows stay in L2.  On the 3D C5G7 deck
   * (3.2 M FSRs, 540 MB of {q, sigma_t} and tally rows) the sorted order moves 292 GB through HBM
   * per sweep, the uid order 24.5 GB (ncu, profiles/r02_sweep3d.md): 80.7 -> 48.8 ms.
   * B200_ORDER=natural / sorted overrides. */
  {
    const char* o = getenv("B200_ORDER");
    bool sorted = !s->cfg.solve_3d;
    if (o != nullptr && strcmp(o, "natural") == 0) sorted = false;
    if (o != nullptr && strcmp(o, "sorted") == 0) sorted = true;
    if (sorted)
      std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return (s->h_off[x + 1] - s->h_off[x]) > (s->h_off[y + 1] - s->h_off[y]);
      });
  }

  CU(s->cls_w.upload(cw.data(), cw.size(), s->stream));
  CU(s->cls_inv_sin.upload(cis.data(), cis.size(), s->stream));
  CU(s->trk_class.upload(cls.data(), nt, s->stream));
  CU(s->out_slot.upload(out_slot.data(), 2 * nt, s->stream));
  CU(s->carry.upload(carry.data(), 2 * nt, s->stream));
  CU(s->order.upload(order.data(), nt, s->stream));

  /* state arrays (zero-initialised like CPUSolver::initializeFluxArrays, CPUSolver.cpp:281-370) */
  const size_t nphi = (size_t)s->n_fsr * s->G, npsi = (size_t)nt * 2 * s->F;
  /* tally replicas (sweep.cuh: SweepArgs::rep_mask): enough copies that FSRs x copies >= 16 Ki
   * rows, at most 64, at most 256 MB; B200_PHI_REPLICAS overrides (power of two) */
  {
    int R = 1;
    while ((int64_t)s->n_fsr * R < 16384 && R < 64) R *= 2;
    if (const char* e = getenv("B200_PHI_REPLICAS")) {
      const int v = atoi(e);
      if (v >= 1 && v <= 1024 && (v & (v - 1)) == 0) R = v;
    }
    const double nphi_pad_d = (double)s->n_fsr * s->GP;
    while (R > 1 && nphi_pad_d * 8.0 * (s->linear ? 4.0 : 1.0) * R > 256e6) R /= 2;
    while (R > 1 && nphi_pad_d * R * (s->linear ? 3.0 : 1.0) >= 4294967296.0) R /= 2;   /* 32-bit tally index incl. replica */
    s->n_rep = R;
  }
  CU(s->phi.alloc(nphi * s->n_rep)); CU(s->phi_old.alloc(nphi)); CU(s->fixed.alloc(nphi));
  CU(s->stab.alloc(nphi)); CU(s->qst.alloc(nphi)); CU(s->scratch.alloc(std::max(nphi, (size_t)s->n_fsr)));
  CU(s->psi_a.alloc(npsi)); CU(s->psi_b.alloc(npsi));
  CU(s->leakage.alloc(std::max<size_t>(nt, 1)));
  CU(cudaMemsetAsync(s->leakage.p, 0, std::max<size_t>(nt, 1) * 4, s->stream));
  CU(s->part3.alloc(3 * MAX_PARTIALS));
  const size_t nphi_pad = (size_t)s->n_fsr * s->GP;
  if (s->padded) {
    CU(s->qst_pad.alloc(nphi_pad));
    CU(cudaMemsetAsync(s->qst_pad.p, 0, nphi_pad * 16, s->stream));
    CU(s->tally_pad.alloc(nphi_pad * s->n_rep));
    CU(cudaMemsetAsync(s->tally_pad.p, 0, nphi_pad * s->n_rep * 8, s->stream));
  }
  if (s->cfg.deterministic) {
    CU(s->phi_fx.alloc(nphi_pad * s->n_rep)); CU(s->fx_bits.alloc(4));
    CU(cudaMemsetAsync(s->phi_fx.p, 0, nphi_pad * s->n_rep * 8, s->stream));
    CU(cudaMemsetAsync(s->fx_bits.p, 0, 4 * 8, s->stream));
  }
  CU(cudaMemsetAsync(s->phi.p, 0, nphi * s->n_rep * 8, s->stream));
  CU(cudaMemsetAsync(s->phi_old.p, 0, nphi * 8, s->stream));
  CU(cudaMemsetAsync(s->fixed.p, 0, nphi * 8, s->stream));
  CU(cudaMemsetAsync(s->stab.p, 0, nphi * 8, s->stream));
  CU(cudaMemsetAsync(s->qst.p, 0, nphi * 16, s->stream));
  if (npsi) {
    CU(cudaMemsetAsync(s->psi_a.p, 0, npsi * 4, s->stream));
    CU(cudaMemsetAsync(s->psi_b.p, 0, npsi * 4, s->stream));
  }
  s->psi_start = s->psi_a.p;
  s->psi_other = s->psi_b.p;

  /* device segment stream: padded 16-byte records with the FSR id premultiplied by G */
  if ((double)s->n_fsr * s->GP * (s->linear ? 3.0 : 1.0) >= 4294967296.0)
    return fail("b200_finalize: n_fsrs*G = %.3g exceeds the 32-bit tally index of this build", (double)s->n_fsr * s->GP);
  if (!s->seg_rec_ready) {
    if (s->seg_len.n != (size_t)s->n_seg) return fail("b200_finalize: tracks must be re-uploaded before finalize");
    CU(s->seg_rec.alloc((size_t)s->n_seg + 2 * SEG_PAD));
    build_segrec_kernel<<<grid_for(s->n_seg + 2 * SEG_PAD, 256), 256, 0, s->stream>>>(
        s->seg_rec.p, s->seg_len.p, s->seg_fsr.p, s->n_seg, s->GP);
    CU(cudaGetLastError());
  }

  if (s->linear) {
    if (!s->have_ls) return fail("b200_finalize: linear source requested but b200_upload_linear_source was not called");
    CU(s->seg_pos.alloc((size_t)s->n_seg + 2 * SEG_PAD));
    build_segpos_kernel<<<grid_for(s->n_seg + 2 * SEG_PAD, 256), 256, 0, s->stream>>>(
        s->seg_pos.p, s->ls_seg_start.p, s->n_seg);
    CU(cudaGetLastError());
    CU(s->phi_m.alloc(nphi * 3 * s->n_rep));
    CU(s->qxyz.alloc(nphi));
    CU(cudaMemsetAsync(s->phi_m.p, 0, nphi * 3 * s->n_rep * 8, s->stream));
    CU(cudaMemsetAsync(s->qxyz.p, 0, nphi * 32, s->stream));
    if (s->padded) {
      CU(s->qxyz_pad.alloc(nphi_pad));
      CU(cudaMemsetAsync(s->qxyz_pad.p, 0, nphi_pad * 32, s->stream));
      CU(s->tallym_pad.alloc(nphi_pad * 3 * s->n_rep));
      CU(cudaMemsetAsync(s->tallym_pad.p, 0, nphi_pad * 3 * s->n_rep * 8, s->stream));
    }
  }

  choose_lane_map(s->G, &s->gpl, &s->lpi, &s->ipc);
  /* 2D tracks: one group per thread whatever G is.  An item then spans several warps (and
   * CTAs) - the flat mapping does not care - and the kernel keeps its 72 registers / 4 CTAs per
   * SM: 2D C5G7 with 70 groups 4.5e11 -> 5.9e11 integrations/s.  3D tracks (one polar angle per
   * track) gain from sharing the per-segment work between 3 groups instead (3.5e11 vs 3.2e11). */
  if (!s->cfg.solve_3d && !s->linear && getenv("B200_GPL") == nullptr && s->G > 32) {
    s->gpl = 1; s->lpi = s->G; s->ipc = std::max(1, 224 / s->G);
  }
  if (s->linear) {          /* the LS kernel is instantiated for 1 or 3 groups per thread */
    s->gpl = s->G <= 32 ? 1 : 3;
    s->lpi = (s->G + s->gpl - 1) / s->gpl;
    s->ipc = std::min(32, 224 / s->lpi);
  }
  const int64_t n_items = 2 * nt;
  s->sweep_blocks = (n_items + s->ipc - 1) / s->ipc;
  if (s->cfg.precision == B200_PRECISION_TABLE) {
    if (s->gpl != 1 || s->NP > 3)
      return fail("B200_PRECISION_TABLE supports one group per thread and at most 3 polar angles per track (G = %d, NP = %d)", s->G, s->NP);
    /* quadratic Taylor expansion of F1(x) = (1 - exp(-x)) / x about the midpoint of every interval */
    std::vector<float4> tab(F1TAB_N);
    const double h = 1.0 / (double)F1TAB_INV_H;
    for (int i = 0; i < F1TAB_N; i++) {
      const double x = (i + 0.5) * h;
      double f0, f1, f2;          /* F1, F1', F1''/2 */
      if (x < 0.5) {
        f0 = f1 = f2 = 0.;
        double fact = 1.0;
        for (int k = 0; k < 30; k++) {
          fact *= (k + 1);
          const double ck = (k % 2 ? -1.0 : 1.0) / fact;        /* (-1)^k / (k+1)! */
          f0 += ck * std::pow(x, k);
          if (k >= 1) f1 += ck * k * std::pow(x, k - 1);
          if (k >= 2) f2 += ck * k * (k - 1) * std::pow(x, k - 2) * 0.5;
        }
      } else {
        const double ex = std::exp(-x);
        f0 = (1.0 - ex) / x;
        f1 = (ex * (x + 1.0) - 1.0) / (x * x);
        f2 = 0.5 * (2.0 - ex * (x * x + 2.0 * x + 2.0)) / (x * x * x);
      }
      tab[i] = make_float4((float)f0, (float)f1, (float)f2, 0.f);
    }
    CU(s->f1tab.upload(tab.data(), F1TAB_N, s->stream));
    CU(cudaStreamSynchronize(s->stream));
  }

  s->smem_attr_set = false;

  if (refresh_material_tables(s)) return 1;
  s->finalized = true;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* sweep dispatch                                                             */
/* ------------------------------------------------------------------------- */
typedef void (*sweep_fn)(const SweepArgs);

template <typename T, int NP, bool DET, bool CMFD>
static sweep_fn pick_gpl(int gpl) {
  switch (gpl) {
    case 1: return sweep_kernel<T, NP, 1, DET, CMFD>;
    case 2: return sweep_kernel<T, NP, 2, DET, CMFD>;
    case 3: return sweep_kernel<T, NP, 3, DET, CMFD>;
    case 4: return sweep_kernel<T, NP, 4, DET, CMFD>;
    case 7: return sweep_kernel<T, NP, 7, DET, CMFD>;
    case 8: return sweep_kernel<T, NP, 8, DET, CMFD>;
  }
  return nullptr;
}
template <typename T, bool DET, bool CMFD>
static sweep_fn pick_np(int np, int gpl) {
  switch (np) {
    case 1: return pick_gpl<T, 1, DET, CMFD>(gpl);
    case 2: return pick_gpl<T, 2, DET, CMFD>(gpl);
    case 3: return pick_gpl<T, 3, DET, CMFD>(gpl);
    case 4: return pick_gpl<T, 4, DET, CMFD>(gpl);
    case 5: return pick_gpl<T, 5, DET, CMFD>(gpl);
    case 6: return pick_gpl<T, 6, DET, CMFD>(gpl);
  }
  return nullptr;
}

static int take_events(b200_solver* s, cudaEvent_t* a, cudaEvent_t* b) {
  if (s->ev_free.empty()) {
    CU(cudaEventCreate(a));
    CU(cudaEventCreate(b));
  } else {
    *a = s->ev_free.back().first;
    *b = s->ev_free.back().second;
    s->ev_free.pop_back();
  }
  return 0;
}

static int resolve_events(b200_solver* s) {
  for (auto& p : s->ev_pending) {
    float ms = 0.f;
    CU(cudaEventSynchronize(p.second));
    CU(cudaEventElapsedTime(&ms, p.first, p.second));
    s->sweep_ms += ms;
    s->ev_free.push_back(p);
  }
  s->ev_pending.clear();
  return 0;
}

/* CPUSolver::transportSweep (src/CPUSolver.cpp:2338-2389): phi <- 0, start ->
 * boundary (here: read psi_start, write the other buffer), sweep. */
static int launch_sweep(b200_solver* s) {
  const size_t nphi = (size_t)s->n_fsr * s->G;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (!s->capturing) {
    if (take_events(s, &e0, &e1)) return 1;
    CU(cudaEventRecord(e0, s->stream));
  }
  const size_t nphi_pad = (size_t)s->n_fsr * s->GP;
  const bool pad_tally = s->padded && !s->cfg.deterministic;   /* the fixed-point tally is padded in place */
  if (s->padded) {
    /* sources into the padded private copy (its tallies were cleared by the previous unpack) */
    pack_pad_kernel<<<grid_for((int64_t)nphi_pad, 256), 256, 0, s->stream>>>(
        s->qst.p, s->qst_pad.p, s->linear ? s->qxyz.p : nullptr, s->linear ? s->qxyz_pad.p : nullptr, s->n_fsr, s->G, s->GP,
        s->iscal.p + SI_DONE);
    CU(cudaGetLastError());
    s->n_launches++;
  }
  if (!pad_tally || s->n_trk == 0) {
    /* flattenFSRFluxes(0) (CPUSolver.cpp:2347), skipped once the device-side loop has converged */
    zero_phi_kernel<<<grid_for(nphi, 256), 256, 0, s->stream>>>(s->phi.p, (int64_t)nphi, s->iscal.p);
    CU(cudaGetLastError());
    s->n_launches++;
    if (s->linear) {
      zero_phi_kernel<<<grid_for(nphi * 3, 256), 256, 0, s->stream>>>(s->phi_m.p, (int64_t)nphi * 3, s->iscal.p);
      CU(cudaGetLastError());
      s->n_launches++;
    }
  }
  if (s->cfg.deterministic) {
    FsrArgs fa = fsr_args(s);
    fx_bound_kernel<<<grid_for((int64_t)nphi, RED_THREADS), RED_THREADS, 0, s->stream>>>(fa, s->fx_bits.p);
    CU(cudaGetLastError());
    fx_scale_kernel<<<1, 1, 0, s->stream>>>(fa, s->fx_bits.p);
    CU(cudaGetLastError());
    s->n_launches += 2;
  }
  if (s->n_trk > 0) {
    SweepArgs a;
    a.seg = s->seg_rec.p + SEG_PAD; a.trk_off = s->trk_off.p;
    a.trk_class = s->trk_class.p; a.order = s->order.p; a.out_slot = s->out_slot.p;
    a.carry = s->carry.p; a.cls_w = s->cls_w.p; a.cls_inv_sin = s->cls_inv_sin.p;
    a.qst = s->padded ? s->qst_pad.p : s->qst.p; a.psi_in = s->psi_start; a.psi_out = s->psi_other;
    a.phi = pad_tally ? s->tally_pad.p : s->phi.p;
    for (int j = 0; j < 16; j++) a.peer_out.p[j] = j < (int)s->peer_out.size() ? s->peer_out[j] : nullptr;
    a.phi_fx = s->phi_fx.p; a.fx_scale = s->scal.p + SC_FXSCALE;
    a.rep_stride = (int64_t)nphi_pad; a.rep_mask = s->n_rep - 1;
    a.leakage = s->balance ? s->leakage.p : nullptr;
    a.seg_cmfd = nullptr; a.cmfd_group = nullptr; a.currents = nullptr; a.ncg = 0;
    if (s->cmfd_on) {
      a.seg_cmfd = s->seg_cmfd.p + SEG_PAD; a.cmfd_group = s->cmfd_group.p; a.currents = s->currents.p; a.ncg = s->ncg;
      zero_phi_kernel<<<grid_for(s->n_cmfd_slots * s->ncg, 256), 256, 0, s->stream>>>(
          s->currents.p, s->n_cmfd_slots * s->ncg, s->iscal.p);          /* Cmfd::zeroCurrents */
      CU(cudaGetLastError());
      s->n_launches++;
    }
    if (s->balance) {
      zero_float_kernel<<<grid_for(s->n_trk, 256), 256, 0, s->stream>>>(s->leakage.p, s->n_trk, s->iscal.p);
      CU(cudaGetLastError());
      s->n_launches++;
    }
    a.done = s->iscal.p + SI_DONE;
    a.n_items = 2 * s->n_trk; a.G = s->G; a.lpi = s->lpi; a.exact = (s->gpl * s->lpi == s->G) ? 1 : 0;
    {
      using C = F1Coef<double>;
      const double cf[11] = {C::d1, C::d2, C::d3, C::d4, C::d5, C::d6, C::p1, C::p2, C::p3, C::p4, C::p5};
      for (int k = 0; k < 11; k++) a.cf[k] = cf[k];
    }
    const bool mixed = s->cfg.precision == B200_PRECISION_MIXED;
    int nthr = s->lpi * s->ipc;
    int64_t nblk = s->sweep_blocks;
    if (!s->linear && (nthr < 192 || nthr > 224)) {     /* LPI does not pack 224 threads: let items straddle CTAs */
      nthr = 224;
      nblk = (2 * s->n_trk * (int64_t)s->lpi + nthr - 1) / nthr;
    }
    /* items may straddle CTAs as well as warps (no CTA-level cooperation), so any block size
     * works; B200_CTA overrides the default of LPI x IPC threads (tuning hook) */
    if (const char* e = getenv("B200_CTA")) {
      const int v = atoi(e);
      if (v >= 32 && v <= 1024 && !s->linear) {
        nthr = v;
        nblk = (2 * s->n_trk * (int64_t)s->lpi + nthr - 1) / nthr;
      }
    }
    if (s->linear) {
      if (s->balance) return fail("k_eff from the neutron balance is not available with the linear source in this build");
      SweepLSArgs la;
      la.f = a;
      la.seg_pos = s->seg_pos.p + SEG_PAD;
      la.trk_dir = s->ls_trk_dir.p;
      la.qxyz = s->padded ? s->qxyz_pad.p : s->qxyz.p;
      la.phi_m = s->padded ? s->tallym_pad.p : s->phi_m.p;
      /* expG_fractional coefficients p0..p5, d1..d6 (src/exponentials.h:113-127) */
      const double cg[12] = {0.5, 1.76558112351595e-1, 4.041584305811143e-2, 6.178333902037397e-3,
                             6.429894635552992e-4, 6.064409107557148e-5, 6.864462055546078e-1,
                             2.263358514260129e-1, 4.721469893686252e-2, 6.883236664917246e-3,
                             7.036272419147752e-4, 6.064409107557148e-5};
      for (int k = 0; k < 12; k++) la.cg[k] = cg[k];
      void (*lfn)(const SweepLSArgs) = nullptr;
      const int key = (s->cfg.solve_3d ? 100 : 0) + s->NP * 10 + s->gpl;
      switch (key) {
        case 111: lfn = sweep_ls_kernel<1, 1, true>; break;
        case 113: lfn = sweep_ls_kernel<1, 3, true>; break;
        case 11: lfn = sweep_ls_kernel<1, 1, false>; break;
        case 13: lfn = sweep_ls_kernel<1, 3, false>; break;
        case 21: lfn = sweep_ls_kernel<2, 1, false>; break;
        case 23: lfn = sweep_ls_kernel<2, 3, false>; break;
        case 31: lfn = sweep_ls_kernel<3, 1, false>; break;
        case 33: lfn = sweep_ls_kernel<3, 3, false>; break;
      }
      if (lfn == nullptr) return fail("no linear-source sweep kernel for NP=%d GPL=%d 3D=%d", s->NP, s->gpl, s->cfg.solve_3d);
      lfn<<<(unsigned)s->sweep_blocks, nthr, 0, s->stream>>>(la);
      CU(cudaGetLastError());
      s->n_launches++;
      if (s->padded) {
        unpack_pad_kernel<<<grid_for(nphi * 3, 256), 256, 0, s->stream>>>(s->tallym_pad.p, s->phi_m.p, s->n_fsr, s->G, s->GP, 3,
                                                                         s->n_rep, s->iscal.p + SI_DONE);
        CU(cudaGetLastError());
        s->n_launches++;
      } else if (s->n_rep > 1) {
        fold_replicas_kernel<double><<<grid_for(nphi * 3, 256), 256, 0, s->stream>>>(
            s->phi_m.p, (int64_t)nphi * 3, (int64_t)nphi * 3, s->n_rep);
        CU(cudaGetLastError());
        s->n_launches++;
      }
    } else {
    sweep_fn fn = nullptr;
    if (s->cfg.precision == B200_PRECISION_TABLE) {
      if (s->cmfd_on || s->cfg.deterministic)
        return fail("B200_PRECISION_TABLE: no CMFD current tally and no fixed-point tally in this build");
      switch (s->NP) {
        case 1: sweep_kernel_tab<1><<<(unsigned)nblk, nthr, 0, s->stream>>>(a, s->f1tab.p); break;
        case 2: sweep_kernel_tab<2><<<(unsigned)nblk, nthr, 0, s->stream>>>(a, s->f1tab.p); break;
        case 3: sweep_kernel_tab<3><<<(unsigned)nblk, nthr, 0, s->stream>>>(a, s->f1tab.p); break;
      }
      CU(cudaGetLastError());
      s->n_launches++;
    } else {
    if (s->cmfd_on) {
      if (mixed || s->cfg.deterministic)
        return fail("CMFD current tallies need B200_PRECISION_DOUBLE with the atomic tally in this build");
      fn = pick_np<double, false, true>(s->NP, s->gpl);
    } else if (s->cfg.deterministic) {
      fn = pick_np<double, true, false>(s->NP, s->gpl);
    } else {
      fn = mixed ? pick_np<float, false, false>(s->NP, s->gpl) : pick_np<double, false, false>(s->NP, s->gpl);
    }
    if (fn == nullptr) return fail("no sweep kernel for NP=%d GPL=%d", s->NP, s->gpl);
    fn<<<(unsigned)nblk, nthr, 0, s->stream>>>(a);
    CU(cudaGetLastError());
    s->n_launches++;
    }
    }
    if (pad_tally) {
      unpack_pad_kernel<<<grid_for(nphi, 256), 256, 0, s->stream>>>(s->tally_pad.p, s->phi.p, s->n_fsr, s->G, s->GP, 1, s->n_rep,
                                                                   s->iscal.p + SI_DONE);
      CU(cudaGetLastError());
      s->n_launches++;
    } else if (s->n_rep > 1) {
      if (s->cfg.deterministic)
        fold_replicas_kernel<unsigned long long><<<grid_for(nphi_pad, 256), 256, 0, s->stream>>>(
            s->phi_fx.p, (int64_t)nphi_pad, (int64_t)nphi_pad, s->n_rep);
      else
        fold_replicas_kernel<double><<<grid_for(nphi, 256), 256, 0, s->stream>>>(
            s->phi.p, (int64_t)nphi, (int64_t)nphi, s->n_rep);
      CU(cudaGetLastError());
      s->n_launches++;
    }
  }
  if (s->cfg.deterministic && !s->defer_fx_convert) {
    fx_to_double_kernel<<<grid_for(nphi, 256), 256, 0, s->stream>>>(fsr_args(s), s->phi_fx.p, s->GP);
    CU(cudaGetLastError());
    s->n_launches++;
  }
  if (!s->capturing) {
    CU(cudaEventRecord(e1, s->stream));
    s->ev_pending.push_back({e0, e1});
    if (s->ev_pending.size() >= 512) { if (resolve_events(s)) return 1; }
  }
  std::swap(s->psi_start, s->psi_other);
  s->n_sweeps++;
  return 0;
}

/* After a device-converged loop some enqueued iterations were no-ops on the GPU
 * but still flipped the host's psi double-buffer pointers: undo an odd surplus. */
static void fix_psi_parity(b200_solver* s, int enqueued, int executed) {
  if ((enqueued - executed) & 1) std::swap(s->psi_start, s->psi_other);
  s->n_sweeps -= (enqueued - executed);
}

/* ------------------------------------------------------------------------- */
/* step functions                                                             */
/* ------------------------------------------------------------------------- */
static int fetch_scalars(b200_solver* s) {
  CU(cudaMemcpyAsync(s->h_scal, s->scal.p, SC_COUNT_D * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(s->h_iscal, s->iscal.p, SI_COUNT_I * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

static int clear_done(b200_solver* s) {
  CU(cudaMemsetAsync(s->iscal.p, 0, SI_COUNT_I * sizeof(int), s->stream));
  return 0;
}

extern "C" int b200_zero_track_fluxes(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_zero_track_fluxes(c));
  const size_t npsi = (size_t)s->n_trk * 2 * s->F;
  if (npsi) {
    CU(cudaMemsetAsync(s->psi_a.p, 0, npsi * 4, s->stream));
    CU(cudaMemsetAsync(s->psi_b.p, 0, npsi * 4, s->stream));
  }
  CU(cudaMemsetAsync(s->scal.p + SC_PSIMAX, 0, sizeof(double), s->stream));    /* bound on |psi| (deterministic tally) */
  return 0;
}

extern "C" int b200_flatten_fsr_fluxes(b200_solver* s, double value) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_flatten_fsr_fluxes(c, value));
  const int64_t n = s->n_fsr * s->G;
  fill_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(s->phi.p, value, n);
  CU(cudaGetLastError());
  s->n_launches++;
  if (s->linear) CU(cudaMemsetAsync(s->phi_m.p, 0, (size_t)n * 3 * 8, s->stream));   /* CPULSSolver.cpp:342-354 */
  return 0;
}

extern "C" int b200_flatten_fsr_fluxes_chi_spectrum(b200_solver* s, int32_t material) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_flatten_fsr_fluxes_chi_spectrum(c, material));
  if (material < 0 || material >= s->n_mat)
    return fail("b200_flatten_fsr_fluxes_chi_spectrum: material %d outside [0,%d)", material, s->n_mat);
  fill_chi_kernel<<<grid_for(s->n_fsr * s->G, 256), 256, 0, s->stream>>>(fsr_args(s), material);
  CU(cudaGetLastError());
  s->n_launches++;
  return 0;
}

extern "C" int b200_store_fsr_fluxes(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_store_fsr_fluxes(c));
  CU(cudaMemcpyAsync(s->phi_old.p, s->phi.p, (size_t)s->n_fsr * s->G * 8, cudaMemcpyDeviceToDevice, s->stream));
  return 0;
}

static int launch_rate(b200_solver* s, int op) {
  FsrArgs a = fsr_args(s);
  const int nb = grid_for(s->n_fsr * s->G, RED_THREADS, MAX_PARTIALS);
  rate_partials_kernel<<<nb, RED_THREADS, 0, s->stream>>>(a);
  CU(cudaGetLastError());
  rate_finalize_kernel<<<1, RED_THREADS, 0, s->stream>>>(a, nb, op);
  CU(cudaGetLastError());
  s->n_launches += 2;
  return 0;
}

static int launch_scale_moments(b200_solver* s) {
  if (!s->linear) return 0;
  scale_moments_kernel<<<grid_for(s->n_fsr * s->G * 3, 256), 256, 0, s->stream>>>(fsr_args(s), s->phi_m.p);
  CU(cudaGetLastError());
  s->n_launches++;
  return 0;
}

static int launch_scale(b200_solver* s) {
  FsrArgs a = fsr_args(s);
  if (launch_scale_moments(s)) return 1;
  scale_phi_kernel<<<grid_for(s->n_fsr * s->G, 256), 256, 0, s->stream>>>(a);
  CU(cudaGetLastError());
  const int64_t npsi = s->n_trk * 2 * (int64_t)s->F;
  if (npsi) {
    /* the reference scales _start_flux and _boundary_flux (CPUSolver.cpp:1922-1928);
     * only the start buffer is ever read again */
    scale_psi_kernel<<<grid_for(npsi, 256), 256, 0, s->stream>>>(s->psi_start, npsi, s->scal.p, s->iscal.p);
    CU(cudaGetLastError());
  }
  s->n_launches += 2;
  return 0;
}

extern "C" int b200_normalize_fluxes(b200_solver* s, double* norm_factor) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_normalize_fluxes(c, norm_factor));
  if (clear_done(s)) return 1;
  if (launch_rate(s, 2)) return 1;
  if (launch_scale(s)) return 1;
  if (norm_factor != nullptr) {
    if (fetch_scalars(s)) return 1;
    *norm_factor = s->h_scal[SC_NORM];
  }
  return 0;
}

static int launch_sources(b200_solver* s, int iteration, int mode) {
  FsrArgs a = fsr_args(s);
  const int64_t n = s->n_fsr * s->G;
  sources_kernel<<<grid_for(n, 256, 1 << 30), 256, 0, s->stream>>>(a, iteration, mode, s->neg_allowed ? 1 : 0);
  CU(cudaGetLastError());
  s->n_launches++;
  if (s->linear && mode == 0) {
    sources_ls_kernel<<<grid_for(n, 256, 1 << 30), 256, 0, s->stream>>>(a, ls_args(s), iteration, s->neg_allowed ? 1 : 0);
    CU(cudaGetLastError());
    s->n_launches++;
  }
  return 0;
}

extern "C" int b200_compute_fsr_sources(b200_solver* s, int32_t iteration) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_compute_fsr_sources(c, iteration));
  if (clear_done(s)) return 1;
  return launch_sources(s, iteration, 0);
}
extern "C" int b200_compute_fsr_fission_sources(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_compute_fsr_fission_sources(c));
  if (clear_done(s)) return 1;
  return launch_sources(s, 0, 1);
}
extern "C" int b200_compute_fsr_scatter_sources(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_compute_fsr_scatter_sources(c));
  if (clear_done(s)) return 1;
  return launch_sources(s, 0, 2);
}

extern "C" int b200_transport_sweep(b200_solver* s) {
  NEED_FINAL(s);
  if (s->grp != nullptr) {
    for (b200_solver* c : grp_shards(s)) { CU(cudaSetDevice(c->cfg.device)); if (clear_done(c)) return 1; }
    return grp_sweep(s);
  }
  if (clear_done(s)) return 1;
  return launch_sweep(s);
}

static int launch_closure(b200_solver* s, int with_rate, int* n_partials) {
  FsrArgs a = fsr_args(s);
  const int nb = grid_for(s->n_fsr * s->G, RED_THREADS, MAX_PARTIALS);
  if (s->linear) closure_ls_kernel<<<nb, RED_THREADS, 0, s->stream>>>(a, ls_args(s), s->neg_allowed ? 1 : 0, with_rate);
  else closure_kernel<<<nb, RED_THREADS, 0, s->stream>>>(a, s->neg_allowed ? 1 : 0, with_rate);
  CU(cudaGetLastError());
  s->n_launches++;
  if (n_partials) *n_partials = nb;
  return 0;
}

extern "C" int b200_add_source_to_scalar_flux(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_add_source_to_scalar_flux(c));
  if (clear_done(s)) return 1;
  return launch_closure(s, 0, nullptr);
}

static int launch_balance_keff(b200_solver* s) {
  FsrArgs a = fsr_args(s);
  const int nb = grid_for(std::max<int64_t>(s->n_fsr * s->G, s->n_trk), RED_THREADS, MAX_PARTIALS);
  balance_partials_kernel<<<dim3(nb, 3), RED_THREADS, 0, s->stream>>>(a, s->leakage.p, s->n_trk, s->part3.p);
  CU(cudaGetLastError());
  balance_finalize_kernel<<<1, RED_THREADS, 0, s->stream>>>(a, s->part3.p, nb);
  CU(cudaGetLastError());
  s->n_launches += 2;
  return 0;
}

extern "C" int b200_set_keff_from_neutron_balance(b200_solver* s, int32_t on) {
  NEED(s);
  if (s->grp != nullptr && on) return fail("k_eff from the neutron balance is not available on a multi-device solver in this build");
  if (s->grp != nullptr) return 0;
  s->balance = on != 0;
  return 0;
}

extern "C" int b200_compute_keff(b200_solver* s, double* k_eff) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_compute_keff(c, k_eff));
  if (clear_done(s)) return 1;
  if (s->balance ? launch_balance_keff(s) : launch_rate(s, 1)) return 1;
  if (k_eff != nullptr) {
    if (fetch_scalars(s)) return 1;
    *k_eff = s->h_scal[SC_KEFF];
  }
  return 0;
}

static int launch_residual(b200_solver* s, int res_type, int scale_first, int store_after,
                           int loop_kind, int iteration) {
  FsrArgs a = fsr_args(s);
  const int nb = grid_for(s->n_fsr, RED_THREADS, MAX_PARTIALS);
  residual_kernel<<<nb, RED_THREADS, 0, s->stream>>>(a, res_type, scale_first, store_after);
  CU(cudaGetLastError());
  residual_finalize_kernel<<<1, RED_THREADS, 0, s->stream>>>(a, nb, res_type, loop_kind, iteration,
                                                            s->hist_k.p, s->hist_res.p);
  CU(cudaGetLastError());
  s->n_launches += 2;
  return 0;
}

extern "C" int b200_compute_residual(b200_solver* s, int32_t res_type, double* residual) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_compute_residual(c, res_type, residual));
  if (res_type < 0 || res_type > 2) return fail("b200_compute_residual: unknown residual type %d", res_type);
  if (res_type == B200_RES_FISSION_SOURCE && s->n_fissionable == 0)
    return fail("The Solver is unable to compute a FISSION_SOURCE residual without fissionable FSRs");
  if (clear_done(s)) return 1;
  if (launch_residual(s, res_type, 0, 0, 0, 0)) return 1;
  if (residual != nullptr) {      /* NULL: leave the value on the device, no host sync */
    if (fetch_scalars(s)) return 1;
    *residual = s->h_scal[SC_RESIDUAL];
  }
  return 0;
}

static int launch_stabilizing_flux(b200_solver* s) {
  stabilizing_flux_kernel<<<grid_for(s->n_fsr * s->G, 256), 256, 0, s->stream>>>(
      fsr_args(s), s->stab_type, s->stab_factor, s->max_ratio.p);
  CU(cudaGetLastError());
  s->n_launches++;
  if (s->linear) {            /* _stabilize_moments, the reference's default (CPULSSolver.cpp:19) */
    const size_t n3 = (size_t)s->n_fsr * s->G * 3;
    if (s->stab_m.n != n3) {
      CU(s->stab_m.alloc(n3));
      CU(cudaMemsetAsync(s->stab_m.p, 0, n3 * 8, s->stream));
    }
    stabilizing_moments_kernel<<<grid_for((int64_t)n3, 256), 256, 0, s->stream>>>(fsr_args(s), s->phi_m.p, s->stab_m.p,
                                                                                s->stab_type, s->stab_factor);
    CU(cudaGetLastError());
    s->n_launches++;
  }
  return 0;
}
static int launch_stabilize_flux(b200_solver* s) {
  stabilize_flux_kernel<<<grid_for(s->n_fsr * s->G, 256), 256, 0, s->stream>>>(
      fsr_args(s), s->stab_type, s->stab_factor, s->max_ratio.p);
  CU(cudaGetLastError());
  s->n_launches++;
  if (s->linear && s->stab_m.n == (size_t)s->n_fsr * s->G * 3) {
    stabilize_moments_kernel<<<grid_for((int64_t)s->stab_m.n, 256), 256, 0, s->stream>>>(fsr_args(s), s->phi_m.p, s->stab_m.p,
                                                                                        s->stab_type, s->stab_factor);
    CU(cudaGetLastError());
    s->n_launches++;
  }
  return 0;
}

extern "C" int b200_compute_stabilizing_flux(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_compute_stabilizing_flux(c));
  if (clear_done(s)) return 1;
  return launch_stabilizing_flux(s);
}
extern "C" int b200_stabilize_flux(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_stabilize_flux(c));
  if (clear_done(s)) return 1;
  return launch_stabilize_flux(s);
}

/* ------------------------------------------------------------------------- */
/* public Solver API                                                          */
/* ------------------------------------------------------------------------- */
extern "C" int b200_get_fluxes(b200_solver* s, double* out, int64_t n) {
  NEED_FINAL(s);
  GRP_FIRST(s, b200_get_fluxes(c, out, n));
  if (n != s->n_fsr * s->G)
    return fail("Unable to get FSR scalar fluxes since there are %d groups and %lld FSRs which does "
                "not match the requested %lld flux values", s->G, (long long)s->n_fsr, (long long)n);
  CU(cudaMemcpyAsync(out, s->phi.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
/* getFluxes + the current k_eff behind ONE host synchronisation (hosts that drive the solver step by
 * step, openmoc/krylov.py style, need both after every iteration) */
extern "C" int b200_get_fluxes_keff(b200_solver* s, double* out, int64_t n, double* k_eff) {
  NEED_FINAL(s);
  GRP_FIRST(s, b200_get_fluxes_keff(c, out, n, k_eff));
  if (n != s->n_fsr * s->G)
    return fail("Unable to get FSR scalar fluxes since there are %d groups and %lld FSRs which does "
                "not match the requested %lld flux values", s->G, (long long)s->n_fsr, (long long)n);
  CU(cudaMemcpyAsync(out, s->phi.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(s->h_scal, s->scal.p, SC_COUNT_D * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (k_eff) *k_eff = s->h_scal[SC_KEFF];
  return 0;
}
extern "C" int b200_set_fluxes(b200_solver* s, const double* in, int64_t n) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_set_fluxes(c, in, n));
  if (n != s->n_fsr * s->G)
    return fail("Unable to set an array with %lld flux values for %lld FSRs and %d groups",
                (long long)n, (long long)s->n_fsr, s->G);
  CU(cudaMemcpyAsync(s->phi.p, in, n * 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int b200_set_fixed_source_by_fsr(b200_solver* s, int64_t fsr_id, int32_t group, double source) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_set_fixed_source_by_fsr(c, fsr_id, group, source));
  if (group <= 0 || group > s->G)
    return fail("Unable to use fixed source for group %d in a %d energy group problem", group, s->G);
  if (fsr_id < 0 || fsr_id >= s->n_fsr)
    return fail("Unable to use fixed source for FSR %lld with only %lld FSRs in the geometry",
                (long long)fsr_id, (long long)s->n_fsr);
  CU(cudaMemcpyAsync(s->fixed.p + fsr_id * s->G + (group - 1), &source, 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->fixed_on = true;
  return 0;
}
/* CPULSSolver::setFixedSourceMomentByFSR (src/CPULSSolver.cpp:154-205): x, y, z moments of the fixed source */
extern "C" int b200_set_fixed_source_moments_by_fsr(b200_solver* s, int64_t fsr_id, int32_t group, double src_x,
                                                    double src_y, double src_z) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_set_fixed_source_moments_by_fsr(c, fsr_id, group, src_x, src_y, src_z));
  if (!s->linear) return fail("Fixed source moments need the linear-source solver");
  if (group <= 0 || group > s->G)
    return fail("Unable to use fixed source moments for group %d in a %d energy group problem", group, s->G);
  if (fsr_id < 0 || fsr_id >= s->n_fsr)
    return fail("Unable to use fixed source moments for FSR %lld with only %lld FSRs in the geometry",
                (long long)fsr_id, (long long)s->n_fsr);
  const size_t nphi = (size_t)s->n_fsr * s->G;
  if (s->fixed_m.n != 3 * nphi) {
    CU(s->fixed_m.alloc(3 * nphi));
    CU(cudaMemsetAsync(s->fixed_m.p, 0, 3 * nphi * 8, s->stream));
  }
  const double v[3] = {src_x, src_y, src_z};
  for (int cidx = 0; cidx < 3; cidx++)
    CU(cudaMemcpyAsync(s->fixed_m.p + cidx * nphi + fsr_id * s->G + (group - 1), &v[cidx], 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->fixed_m_on = true;
  return 0;
}
extern "C" int b200_reset_fixed_sources(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_reset_fixed_sources(c));
  CU(cudaMemsetAsync(s->fixed.p, 0, (size_t)s->n_fsr * s->G * 8, s->stream));
  if (s->fixed_m.n) CU(cudaMemsetAsync(s->fixed_m.p, 0, s->fixed_m.n * 8, s->stream));
  s->fixed_m_on = false;
  s->fixed_on = false;
  return 0;
}
extern "C" int b200_compute_fsr_fission_rates(b200_solver* s, double* out, int64_t n, int32_t nu) {
  NEED_FINAL(s);
  GRP_FIRST(s, b200_compute_fsr_fission_rates(c, out, n, nu));
  if (n != s->n_fsr) return fail("b200_compute_fsr_fission_rates: %lld values requested for %lld FSRs", (long long)n, (long long)s->n_fsr);
  fission_rates_kernel<<<grid_for(s->n_fsr, 256), 256, 0, s->stream>>>(fsr_args(s), s->scratch.p, nu);
  CU(cudaGetLastError());
  s->n_launches++;
  CU(cudaMemcpyAsync(out, s->scratch.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int b200_stabilize_transport(b200_solver* s, double factor, int32_t type) {
  NEED(s);
  if (type < 0 || type > 2) return fail("b200_stabilize_transport: unknown stabilization type %d", type);
  s->stabilize = true;
  s->stab_factor = factor;
  s->stab_type = type;
  if (s->grp != nullptr && s->finalized)
    for (b200_solver* c : grp_shards(s)) { c->stabilize = true; c->stab_factor = factor; c->stab_type = type; }
  return 0;
}
extern "C" int b200_allow_negative_fluxes(b200_solver* s, int32_t allowed) {
  NEED(s);
  s->neg_allowed = allowed != 0;
  if (s->grp != nullptr && s->finalized)
    for (b200_solver* c : grp_shards(s)) c->neg_allowed = s->neg_allowed;
  return 0;
}
extern "C" int b200_get_keff(b200_solver* s, double* k) {
  NEED(s);
  GRP_FIRST(s, b200_get_keff(c, k));
  if (fetch_scalars(s)) return 1;
  if (k) *k = s->h_scal[SC_KEFF];
  return 0;
}
extern "C" int b200_set_keff(b200_solver* s, double k) {
  NEED(s);
  GRP_ALL(s, b200_set_keff(c, k));
  CU(cudaMemcpyAsync(s->scal.p + SC_KEFF, &k, 8, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int b200_get_fsr_sources(b200_solver* s, double* out, int64_t n) {
  NEED_FINAL(s);
  GRP_FIRST(s, b200_get_fsr_sources(c, out, n));
  if (n != s->n_fsr * s->G) return fail("b200_get_fsr_sources: size mismatch");
  extract_q_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(s->qst.p, s->scratch.p, n);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, s->scratch.p, n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int b200_set_fsr_sources(b200_solver* s, const double* in, int64_t n) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_set_fsr_sources(c, in, n));
  if (n != s->n_fsr * s->G) return fail("b200_set_fsr_sources: size mismatch");
  CU(cudaMemcpyAsync(s->scratch.p, in, n * 8, cudaMemcpyHostToDevice, s->stream));
  insert_q_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(s->qst.p, s->scratch.p, n);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int b200_get_flux_moments(b200_solver* s, double* out, int64_t n) {
  NEED_FINAL(s);
  GRP_FIRST(s, b200_get_flux_moments(c, out, n));
  if (!s->linear) return fail("b200_get_flux_moments: not a linear-source solver");
  if (n != s->n_fsr * s->G * 3) return fail("b200_get_flux_moments: size mismatch");
  /* staging buffer in the reference's [r][c][e] order, kept for the next call (the CMFD path
   * moves the moments twice per iteration) */
  if (s->mom_stage.n < (size_t)n) CU(s->mom_stage.alloc(n));
  double* tmp = s->mom_stage.p;
  moments_to_ref_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(s->phi_m.p, tmp, s->n_fsr, s->G);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, tmp, n * 8, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int b200_set_flux_moments(b200_solver* s, const double* in, int64_t n) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_set_flux_moments(c, in, n));
  if (!s->linear) return fail("b200_set_flux_moments: not a linear-source solver");
  if (n != s->n_fsr * s->G * 3) return fail("b200_set_flux_moments: size mismatch");
  if (s->mom_stage.n < (size_t)n) CU(s->mom_stage.alloc(n));
  double* tmp = s->mom_stage.p;
  CU(cudaMemcpyAsync(tmp, in, n * 8, cudaMemcpyHostToDevice, s->stream));
  moments_from_ref_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(s->phi_m.p, tmp, s->n_fsr, s->G);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

extern "C" int b200_get_start_fluxes(b200_solver* s, float* out, int64_t n) {
  NEED_FINAL(s);
  if (s->grp != nullptr) return grp_get_start_fluxes(s, out, n);
  if (n != s->n_trk * 2 * (int64_t)s->F) return fail("b200_get_start_fluxes: size mismatch");
  if (n) CU(cudaMemcpyAsync(out, s->psi_start, n * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int b200_set_start_fluxes(b200_solver* s, const float* in, int64_t n) {
  NEED_FINAL(s);
  if (s->grp != nullptr) return grp_set_start_fluxes(s, in, n);
  if (n != s->n_trk * 2 * (int64_t)s->F) return fail("b200_set_start_fluxes: size mismatch");
  if (n) CU(cudaMemcpyAsync(s->psi_start, in, n * 4, cudaMemcpyHostToDevice, s->stream));
  double m = 0.;
  for (int64_t i = 0; i < n; i++) m = std::max(m, (double)std::fabs(in[i]));
  CU(cudaMemcpyAsync(s->scal.p + SC_PSIMAX, &m, sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

/* ------------------------------------------------------------------------- */
/* fused drivers                                                              */
/* ------------------------------------------------------------------------- */
/* one source iteration of Solver::computeEigenvalue (src/Solver.cpp:1614-1681) */
/* i < 0: captured into the CUDA graph (iterations >= 2, number read from the device counter) */
static int enqueue_iteration_begin(b200_solver* s, int i) {
  if (i != 0 && s->stabilize) { if (launch_stabilizing_flux(s)) return 1; }
  if (launch_sources(s, i, 0)) return 1;
  return launch_sweep(s);
}

static int enqueue_iteration_end(b200_solver* s, int i, int res_type, int loop_kind) {
  FsrArgs a = fsr_args(s);
  if (cmfd_in_loop(s)) {
    /* Solver.cpp:1624-1640 with CMFD: closure, Cmfd::computeKeff (k_eff and the prolongation), stabilisation,
     * normalisation from the updated flux */
    if (s->balance) return fail("k_eff from the neutron balance cannot be combined with CMFD (Solver.cpp:1627-1630)");
    if (launch_closure(s, 0, nullptr)) return 1;
    if (enqueue_cmfd(s, i, -1.0)) return 1;
    if (i != 0 && s->stabilize) { if (launch_stabilize_flux(s)) return 1; }
    if (launch_rate(s, 2)) return 1;
  } else if (s->balance) {
    if (launch_closure(s, 0, nullptr)) return 1;
    if (launch_balance_keff(s)) return 1;
    if (i != 0 && s->stabilize) { if (launch_stabilize_flux(s)) return 1; }
    if (launch_rate(s, 2)) return 1;
  } else if (!s->stabilize) {
    /* closure + the one nu-fission reduction that feeds both computeKeff and
     * normalizeFluxes (identical sums when no stabilisation sits in between) */
    int nb = 0;
    if (launch_closure(s, 1, &nb)) return 1;
    rate_finalize_kernel<<<1, RED_THREADS, 0, s->stream>>>(a, nb, 3);
    CU(cudaGetLastError());
    s->n_launches++;
  } else {
    if (launch_closure(s, 0, nullptr)) return 1;
    if (launch_rate(s, 1)) return 1;
    if (i != 0) { if (launch_stabilize_flux(s)) return 1; }
    if (launch_rate(s, 2)) return 1;
  }
  /* normalizeFluxes' scaling of phi fused into the residual pass, storeFSRFluxes after */
  if (launch_scale_moments(s)) return 1;
  const int64_t npsi = s->n_trk * 2 * (int64_t)s->F;
  if (npsi) {
    scale_psi_kernel<<<grid_for(npsi, 256), 256, 0, s->stream>>>(s->psi_start, npsi, s->scal.p, s->iscal.p);
    CU(cudaGetLastError());
    s->n_launches++;
  }
  if (launch_residual(s, res_type, 1, 1, loop_kind, i)) return 1;
  if (cmfd_in_loop(s)) return enqueue_cmfd_threshold(s);     /* Solver.cpp:1671-1675 */
  return 0;
}

static int enqueue_eigen_iteration(b200_solver* s, int i, int res_type, int loop_kind) {
  if (enqueue_iteration_begin(s, i)) return 1;
  return enqueue_iteration_end(s, i, res_type, loop_kind);
}

static int prepare_history(b200_solver* s, int max_iters) {
  CU(s->hist_k.alloc(std::max(max_iters, 1)));
  CU(s->hist_res.alloc(std::max(max_iters, 1)));
  return 0;
}

/* Launch-bound decks (a sweep of a few tens of microseconds followed by eight small FSR
 * kernels) replay the fused iteration as a CUDA graph: two iterations per graph, one for each
 * parity of the psi double buffer; the kernels read the iteration number from the device-side
 * counter (iteration argument -1).  B200_GRAPH=0/1 forces it off/on; by default decks with fewer
 * than 1e8 integrations per sweep use it. */
static bool want_graph(const b200_solver* s) {
  if (!s->own_stream) return false;      /* a borrowed stream may be the legacy default stream: no capture */
  if (cmfd_in_loop(s)) return false;     /* the CMFD solve may be a cooperative launch */
  if (const char* e = getenv("B200_GRAPH")) return atoi(e) != 0;
  return 2.0 * s->F * (double)s->n_seg < 1e8;
}
static void drop_iter_graph(b200_solver* s) {
  if (s->iter_graph != nullptr) cudaGraphExecDestroy(s->iter_graph);
  s->iter_graph = nullptr;
}
static int build_iter_graph(b200_solver* s, int res_type) {
  if (s->iter_graph != nullptr && s->iter_graph_res == res_type && s->iter_graph_psi == s->psi_start) return 0;
  drop_iter_graph(s);
  cudaGraph_t g = nullptr;
  const int64_t l0 = s->n_launches, w0 = s->n_sweeps;
  CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
  s->capturing = true;
  int rc = enqueue_eigen_iteration(s, -1, res_type, 1) || enqueue_eigen_iteration(s, -1, res_type, 1);
  s->capturing = false;
  cudaError_t e = cudaStreamEndCapture(s->stream, &g);
  s->iter_graph_launches = s->n_launches - l0;
  s->n_launches = l0; s->n_sweeps = w0;            /* nothing has run yet */
  if (rc) { if (g) cudaGraphDestroy(g); return 1; }
  CU(e);
  e = cudaGraphInstantiate(&s->iter_graph, g, 0);
  cudaGraphDestroy(g);
  CU(e);
  s->iter_graph_res = res_type;
  s->iter_graph_psi = s->psi_start;                /* the buffers are baked into the graph */
  return 0;
}

extern "C" int b200_compute_eigenvalue(b200_solver* s, int32_t max_iters, double tol, int32_t res_type,
                                       int32_t* num_iterations) {
  NEED_FINAL(s);
  if (res_type < 0 || res_type > 2) return fail("b200_compute_eigenvalue: unknown residual type %d", res_type);
  if (s->grp != nullptr) return grp_compute_eigenvalue(s, max_iters, tol, res_type, num_iterations);
  if (res_type == B200_RES_FISSION_SOURCE && s->n_fissionable == 0)
    return fail("The Solver is unable to compute a FISSION_SOURCE residual without fissionable FSRs");
  if (prepare_history(s, max_iters)) return 1;
  drop_iter_graph(s);     /* solver settings are baked into the captured launches: one graph per solve */
  /* _k_eff = 1, flux arrays zeroed, flat unit flux guess normalised and stored
   * (Solver.cpp:1566-1600, computeInitialFluxGuess :1710-1731) */
  double init[SC_COUNT_D] = {0};
  init[SC_KEFF] = 1.0; init[SC_KPREV] = 1.0; init[SC_TOL] = tol;
  CU(cudaMemcpyAsync(s->scal.p, init, sizeof init, cudaMemcpyHostToDevice, s->stream));
  if (cmfd_loop_init(s, tol)) return 1;
  if (clear_done(s)) return 1;
  if (b200_zero_track_fluxes(s)) return 1;
  CU(cudaMemsetAsync(s->phi_old.p, 0, (size_t)s->n_fsr * s->G * 8, s->stream));
  if (b200_flatten_fsr_fluxes(s, 1.0)) return 1;
  if (launch_rate(s, 2)) return 1;
  if (launch_scale(s)) return 1;
  if (b200_store_fsr_fluxes(s)) return 1;

  const int batch = 8;
  int done = 0, i = 0, flips = 0;      /* flips: host-side swaps of the psi buffers (plain launches) */
  const bool graph = want_graph(s) && max_iters > 4;
  while (i < max_iters && !done) {
    const int end = std::min(max_iters, i + batch);
    if (graph && i >= 2) {
      /* iterations 0 and 1 ran as plain launches (iteration 0 differs when stabilisation is on) */
      if (build_iter_graph(s, res_type)) return 1;
      cudaEvent_t e0, e1;
      if (take_events(s, &e0, &e1)) return 1;
      CU(cudaEventRecord(e0, s->stream));
      for (; i + 2 <= end; i += 2) {
        CU(cudaGraphLaunch(s->iter_graph, s->stream));
        s->n_launches += s->iter_graph_launches;
        s->n_sweeps += 2;
      }
      /* graph mode times whole iterations: the "sweep" split then includes the FSR kernels */
      CU(cudaEventRecord(e1, s->stream));
      s->ev_pending.push_back({e0, e1});
    }
    for (; i < end; i++, flips++)
      if (enqueue_eigen_iteration(s, i, res_type, 1)) return 1;
    if (fetch_scalars(s)) return 1;
    done = s->h_iscal[SI_DONE];
  }
  /* graph iterations come in pairs, so the parity of (enqueued - executed) is also the parity
   * of (host-side buffer swaps - executed) */
  (void)flips;
  fix_psi_parity(s, i, s->h_iscal[SI_EXEC]);
  if (num_iterations) *num_iterations = s->h_iscal[SI_ITERS];
  if (clear_done(s)) return 1;
  return resolve_events(s);
}

/* ---- the fused eigenvalue loop in pieces, for hosts that must reduce the FSR tally across
 *      GPUs in the middle of every iteration (multi-GPU angular decomposition) ---- */
extern "C" int b200_eigen_loop_init(b200_solver* s, int32_t max_iters, double tol) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_eigen_loop_init(c, max_iters, tol));
  if (prepare_history(s, max_iters)) return 1;
  double init[SC_COUNT_D] = {0};
  init[SC_KEFF] = 1.0; init[SC_KPREV] = 1.0; init[SC_TOL] = tol;
  CU(cudaMemcpyAsync(s->scal.p, init, sizeof init, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (cmfd_loop_init(s, tol)) return 1;
  if (clear_done(s)) return 1;
  if (b200_zero_track_fluxes(s)) return 1;
  CU(cudaMemsetAsync(s->phi_old.p, 0, (size_t)s->n_fsr * s->G * 8, s->stream));
  if (b200_flatten_fsr_fluxes(s, 1.0)) return 1;
  if (launch_rate(s, 2)) return 1;
  if (launch_scale(s)) return 1;
  return b200_store_fsr_fluxes(s);
}
/* Between begin and end the host all-reduces scalar_flux in place.  Once the loop has
 * converged the kernels of later iterations are no-ops, but the host's all-reduce still runs
 * and would sum the final flux over the ranks: begin saves it, end restores it. */
extern "C" int b200_iteration_begin(b200_solver* s, int32_t iteration) {
  NEED_FINAL(s);
  if (s->grp != nullptr) {
    for (b200_solver* c : grp_shards(s)) {
      CU(cudaSetDevice(c->cfg.device));
      if (iteration != 0 && c->stabilize) { if (launch_stabilizing_flux(c)) return 1; }
      if (launch_sources(c, iteration, 0)) return 1;
    }
    return grp_sweep(s);
  }
  if (enqueue_iteration_begin(s, iteration)) return 1;
  const int64_t n = s->n_fsr * s->G;
  copy_if_done_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(s->scratch.p, s->phi.p, n, s->iscal.p);
  CU(cudaGetLastError());
  return 0;
}
extern "C" int b200_iteration_end(b200_solver* s, int32_t iteration, int32_t res_type, int32_t check_convergence) {
  NEED_FINAL(s);
  if (res_type < 0 || res_type > 2) return fail("b200_iteration_end: unknown residual type %d", res_type);
  if (s->grp != nullptr) {
    for (b200_solver* c : grp_shards(s)) {
      CU(cudaSetDevice(c->cfg.device));
      if (enqueue_iteration_end(c, iteration, res_type, check_convergence ? 1 : 0)) return 1;
    }
    return 0;
  }
  if (check_convergence && iteration >= 0 && s->hist_k.n <= (size_t)iteration)
    return fail("b200_iteration_end: iteration %d beyond the max_iters given to b200_eigen_loop_init", iteration);
  const int64_t n = s->n_fsr * s->G;
  copy_if_done_kernel<<<grid_for(n, 256), 256, 0, s->stream>>>(s->phi.p, s->scratch.p, n, s->iscal.p);
  CU(cudaGetLastError());
  return enqueue_iteration_end(s, iteration, res_type, check_convergence ? 1 : 0);
}
extern "C" int b200_eigen_loop_status(b200_solver* s, int32_t enqueued, int32_t* done, int32_t* iterations,
                                      double* k_eff, double* residual) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_eigen_loop_status(c, enqueued, done, iterations, k_eff, residual));
  if (fetch_scalars(s)) return 1;
  if (done) *done = s->h_iscal[SI_DONE];
  if (iterations) *iterations = s->h_iscal[SI_ITERS];
  if (k_eff) *k_eff = s->h_scal[SC_KEFF];
  if (residual) *residual = s->h_scal[SC_RESIDUAL];
  if (s->h_iscal[SI_DONE]) {          /* loop over: undo the host-side psi flips of no-op iterations */
    fix_psi_parity(s, enqueued, s->h_iscal[SI_EXEC]);
    if (clear_done(s)) return 1;
  }
  return 0;
}

extern "C" int b200_iterate(b200_solver* s, int32_t n, int32_t res_type, double* k_eff, double* residual) {
  NEED_FINAL(s);
  if (res_type < 0 || res_type > 2) return fail("b200_iterate: unknown residual type %d", res_type);
  if (s->grp != nullptr) {
    for (b200_solver* c : grp_shards(s)) { CU(cudaSetDevice(c->cfg.device)); if (clear_done(c)) return 1; }
    for (int i = 0; i < n; i++)
      if (grp_iteration(s, 1000 + i, res_type, 0)) return 1;
    if (k_eff != nullptr || residual != nullptr) {
      if (grp_sync(s)) return 1;
      b200_solver* c0 = grp_shards(s)[0];
      if (fetch_scalars(c0)) return 1;
      if (k_eff) *k_eff = c0->h_scal[SC_KEFF];
      if (residual) *residual = c0->h_scal[SC_RESIDUAL];
    }
    return 0;
  }
  if (clear_done(s)) return 1;
  for (int i = 0; i < n; i++)
    if (enqueue_eigen_iteration(s, 1000 + i, res_type, 0)) return 1;
  if (k_eff != nullptr || residual != nullptr) {
    if (fetch_scalars(s)) return 1;
    if (k_eff) *k_eff = s->h_scal[SC_KEFF];
    if (residual) *residual = s->h_scal[SC_RESIDUAL];
  }
  return 0;
}

/* the common loop of computeFlux (Solver.cpp:1397-1413) and computeSource (:1492-1508) */
static int flux_source_loop(b200_solver* s, int max_iters, double tol, int res_type, bool sources_each_iter,
                            int32_t* num_iterations) {
  if (prepare_history(s, max_iters)) return 1;
  CU(cudaMemcpyAsync(s->scal.p + SC_TOL, &tol, 8, cudaMemcpyHostToDevice, s->stream));
  if (clear_done(s)) return 1;
  const int batch = 8;
  int done = 0, i = 0;
  while (i < max_iters && !done) {
    const int end = std::min(max_iters, i + batch);
    for (; i < end; i++) {
      if (sources_each_iter) { if (launch_sources(s, i, 0)) return 1; }
      if (launch_sweep(s)) return 1;
      if (launch_closure(s, 0, nullptr)) return 1;
      if (launch_residual(s, res_type, 0, 1, 2, i)) return 1;
    }
    if (fetch_scalars(s)) return 1;
    done = s->h_iscal[SI_DONE];
  }
  fix_psi_parity(s, i, s->h_iscal[SI_EXEC]);
  if (num_iterations) *num_iterations = done ? s->h_iscal[SI_ITERS] : max_iters;
  if (clear_done(s)) return 1;
  return resolve_events(s);
}

extern "C" int b200_compute_flux(b200_solver* s, int32_t max_iters, double tol, int32_t only_fixed_source,
                                 int32_t* num_iterations) {
  NEED_FINAL(s);
  if (s->grp != nullptr) {
    for (b200_solver* c : grp_shards(s)) {
      if (b200_set_keff(c, 1.0)) return 1;
      if (only_fixed_source) {
        if (b200_zero_track_fluxes(c)) return 1;
        if (b200_flatten_fsr_fluxes(c, 0.0)) return 1;
        if (b200_store_fsr_fluxes(c)) return 1;
      }
      if (clear_done(c)) return 1;
      if (launch_sources(c, 0, 0)) return 1;
    }
    return grp_flux_source_loop(s, max_iters, tol, B200_RES_SCALAR_FLUX, false, num_iterations);
  }
  if (b200_set_keff(s, 1.0)) return 1;
  if (only_fixed_source) {
    if (b200_zero_track_fluxes(s)) return 1;
    if (b200_flatten_fsr_fluxes(s, 0.0)) return 1;
    if (b200_store_fsr_fluxes(s)) return 1;
  }
  if (clear_done(s)) return 1;
  if (launch_sources(s, 0, 0)) return 1;
  return flux_source_loop(s, max_iters, tol, B200_RES_SCALAR_FLUX, false, num_iterations);
}

extern "C" int b200_compute_source(b200_solver* s, int32_t max_iters, double k_eff, double tol,
                                   int32_t res_type, int32_t* num_iterations) {
  NEED_FINAL(s);
  if (k_eff <= 0.)
    return fail("The Solver is unable to compute the source with keff = %f since it is not a positive value", k_eff);
  if (res_type < 0 || res_type > 2) return fail("b200_compute_source: unknown residual type %d", res_type);
  if (s->grp != nullptr) {
    for (b200_solver* c : grp_shards(s)) {
      if (b200_set_keff(c, k_eff)) return 1;
      if (b200_zero_track_fluxes(c)) return 1;
      if (b200_flatten_fsr_fluxes(c, 1.0)) return 1;
      if (b200_store_fsr_fluxes(c)) return 1;
    }
    return grp_flux_source_loop(s, max_iters, tol, res_type, true, num_iterations);
  }
  if (b200_set_keff(s, k_eff)) return 1;
  if (b200_zero_track_fluxes(s)) return 1;
  if (b200_flatten_fsr_fluxes(s, 1.0)) return 1;
  if (b200_store_fsr_fluxes(s)) return 1;
  return flux_source_loop(s, max_iters, tol, res_type, true, num_iterations);
}

/* ------------------------------------------------------------------------- */
/* exponential evaluator exposed for known-answer tests                        */
/* ------------------------------------------------------------------------- */
__global__ void eval_expf1_kernel(const double* __restrict__ x, double* __restrict__ out, int64_t n, int precision) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = precision == B200_PRECISION_MIXED ? (double)expF1((float)x[i]) : expF1(x[i]);
}

extern "C" int b200_eval_expF1(int32_t device, int32_t precision, const double* x, int64_t n, double* out) {
  if (!x || !out || n < 0) return fail("b200_eval_expF1: bad argument");
  CU(cudaSetDevice(device));
  struct Tmp { double* p = nullptr; ~Tmp() { if (p) cudaFree(p); } } dx, dout;   /* freed on every path */
  if (n == 0) return 0;
  CU(cudaMalloc((void**)&dx.p, n * 8));
  CU(cudaMalloc((void**)&dout.p, n * 8));
  CU(cudaMemcpy(dx.p, x, n * 8, cudaMemcpyHostToDevice));
  eval_expf1_kernel<<<grid_for(n, 256), 256>>>(dx.p, dout.p, n, precision);
  CU(cudaGetLastError());
  CU(cudaMemcpy(out, dout.p, n * 8, cudaMemcpyDeviceToHost));
  return 0;
}

/* ------------------------------------------------------------------------- */
/* linear-source pre-pass on the device (ls_prepass.cuh)                        */
/* ------------------------------------------------------------------------- */
extern "C" int b200_ls_prepass(int32_t device, int32_t num_groups, int32_t num_azim, int32_t num_polar, int32_t solve_3d,
                               int64_t n_tracks, int64_t n_segments, int64_t n_fsrs, int32_t n_materials,
                               const double* seg_length, const int32_t* seg_fsr, const double* seg_start,
                               const int64_t* trk_seg_offset, const int32_t* trk_azim, const int32_t* trk_polar,
                               const double* trk_phi, const double* trk_theta,
                               const double* azim_spacing, const double* azim_weight, const double* polar_spacing,
                               const double* polar_weight, const double* sin_theta,
                               const double* volume, const int32_t* fsr_material, const double* sigma_t,
                               double* lin_exp_matrix, double* source_constants, int32_t* n_flat_fsrs) {
  if (!seg_length || !seg_fsr || !seg_start || !trk_seg_offset || !trk_azim || !trk_polar || !trk_phi || !trk_theta ||
      !azim_spacing || !azim_weight || !polar_spacing || !polar_weight || !sin_theta || !volume || !fsr_material ||
      !sigma_t || !lin_exp_matrix || !source_constants)
    return fail("b200_ls_prepass: null argument");
  if (num_groups < 1 || n_tracks < 0 || n_segments < 0 || n_fsrs < 1 || n_materials < 1) return fail("b200_ls_prepass: bad size");
  CU(cudaSetDevice(device));
  const int nc = solve_3d ? 6 : 3, A2 = num_azim / 2;
  DevBuf<double> d_len, d_start, d_phi, d_theta, d_as, d_aw, d_ps, d_pw, d_st, d_vol, d_sig, d_lem, d_ilem, d_sc;
  DevBuf<int32_t> d_fsr, d_azim, d_polar, d_mat;
  DevBuf<int64_t> d_off;
  DevBuf<int> d_flat;
  struct Guard {
    std::vector<std::function<void()>> f;
    ~Guard() { for (auto& g : f) g(); }
  } guard;
  auto own = [&](auto& b) { guard.f.push_back([&b]() { b.release(); }); };
  own(d_len); own(d_start); own(d_phi); own(d_theta); own(d_as); own(d_aw); own(d_ps); own(d_pw); own(d_st); own(d_vol);
  own(d_sig); own(d_lem); own(d_ilem); own(d_sc); own(d_fsr); own(d_azim); own(d_polar); own(d_mat); own(d_off); own(d_flat);
  cudaStream_t st = nullptr;
  CU(d_len.upload(seg_length, n_segments, st)); CU(d_fsr.upload(seg_fsr, n_segments, st));
  CU(d_start.upload(seg_start, (size_t)n_segments * 3, st));
  CU(d_off.upload(trk_seg_offset, n_tracks + 1, st)); CU(d_azim.upload(trk_azim, n_tracks, st));
  CU(d_polar.upload(trk_polar, n_tracks, st)); CU(d_phi.upload(trk_phi, n_tracks, st)); CU(d_theta.upload(trk_theta, n_tracks, st));
  CU(d_as.upload(azim_spacing, A2, st)); CU(d_aw.upload(azim_weight, A2, st));
  CU(d_ps.upload(polar_spacing, (size_t)A2 * num_polar, st)); CU(d_pw.upload(polar_weight, (size_t)A2 * num_polar, st));
  CU(d_st.upload(sin_theta, (size_t)A2 * num_polar, st));
  CU(d_vol.upload(volume, n_fsrs, st)); CU(d_mat.upload(fsr_material, n_fsrs, st));
  CU(d_sig.upload(sigma_t, (size_t)n_materials * num_groups, st));
  CU(d_lem.alloc((size_t)n_fsrs * nc)); CU(d_ilem.alloc((size_t)n_fsrs * nc)); CU(d_sc.alloc((size_t)n_fsrs * nc * num_groups));
  CU(d_flat.alloc(1));
  CU(cudaMemset(d_lem.p, 0, (size_t)n_fsrs * nc * 8));
  CU(cudaMemset(d_sc.p, 0, (size_t)n_fsrs * nc * num_groups * 8));
  CU(cudaMemset(d_flat.p, 0, sizeof(int)));
  LsPrepassArgs a;
  a.G = num_groups; a.P = num_polar; a.solve_3d = solve_3d; a.nc = nc;
  a.n_trk = n_tracks; a.n_seg = n_segments; a.n_fsr = n_fsrs;
  a.seg_len = d_len.p; a.seg_fsr = d_fsr.p; a.seg_start = d_start.p; a.trk_off = d_off.p; a.trk_azim = d_azim.p;
  a.trk_polar = d_polar.p; a.trk_phi = d_phi.p; a.trk_theta = d_theta.p; a.azim_spacing = d_as.p; a.azim_weight = d_aw.p;
  a.polar_spacing = d_ps.p; a.polar_weight = d_pw.p; a.sin_theta = d_st.p; a.volume = d_vol.p; a.fsr_mat = d_mat.p;
  a.sigma_t = d_sig.p; a.lem = d_lem.p; a.src_const = d_sc.p;
  if (n_tracks > 0) {
    ls_prepass_kernel<<<grid_for(n_tracks, 128, 1 << 20), 128>>>(a);
    CU(cudaGetLastError());
  }
  ls_invert_kernel<<<grid_for(n_fsrs, 256), 256>>>(d_lem.p, d_vol.p, d_ilem.p, n_fsrs, solve_3d, d_flat.p);
  CU(cudaGetLastError());
  CU(cudaMemcpy(lin_exp_matrix, d_ilem.p, (size_t)n_fsrs * nc * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(source_constants, d_sc.p, (size_t)n_fsrs * nc * num_groups * 8, cudaMemcpyDeviceToHost));
  int nf = 0;
  CU(cudaMemcpy(&nf, d_flat.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (n_flat_fsrs) *n_flat_fsrs = nf;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* machine ceilings (microbench.cuh)                                           */
/* ------------------------------------------------------------------------- */
extern "C" int b200_measure_ceilings(int32_t device, int64_t table_rows, double* fp64_instr_per_s,
                                     double* red_f64_per_s) {
  if (table_rows < 1) return fail("b200_measure_ceilings: table_rows must be positive");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int sms = prop.multiProcessorCount;
  struct Tmp { void* p = nullptr; ~Tmp() { if (p) cudaFree(p); } } out, table;
  struct Ev { cudaEvent_t e = nullptr; ~Ev() { if (e) cudaEventDestroy(e); } } e0, e1;
  CU(cudaEventCreate(&e0.e)); CU(cudaEventCreate(&e1.e));
  float ms = 0.f;
  {
    const int threads = 224, blocks = sms * 4, iters = 2048;       /* 28 warps per SM like the sweep */
    CU(cudaMalloc(&out.p, (size_t)threads * blocks * 8));
    mb_fp64_kernel<<<blocks, threads>>>((double*)out.p, 16, 1.0000001);
    double best = 1e30;
    for (int rep = 0; rep < 3; rep++) {
      CU(cudaEventRecord(e0.e));
      mb_fp64_kernel<<<blocks, threads>>>((double*)out.p, iters, 1.0000001);
      CU(cudaEventRecord(e1.e));
      CU(cudaEventSynchronize(e1.e));
      CU(cudaEventElapsedTime(&ms, e0.e, e1.e));
      best = std::min(best, (double)ms);
    }
    if (fp64_instr_per_s) *fp64_instr_per_s = (double)threads * blocks * iters * 48.0 / (best * 1e-3);
  }
  {
    const int threads = 224, blocks = sms * 4, iters = 1000;
    CU(cudaMalloc(&table.p, (size_t)table_rows * 7 * 8));
    CU(cudaMemset(table.p, 0, (size_t)table_rows * 7 * 8));
    mb_red_kernel<<<blocks, threads>>>((double*)table.p, (int)table_rows, 8);
    double best = 1e30;
    for (int rep = 0; rep < 3; rep++) {
      CU(cudaEventRecord(e0.e));
      mb_red_kernel<<<blocks, threads>>>((double*)table.p, (int)table_rows, iters);
      CU(cudaEventRecord(e1.e));
      CU(cudaEventSynchronize(e1.e));
      CU(cudaEventElapsedTime(&ms, e0.e, e1.e));
      best = std::min(best, (double)ms);
    }
    if (red_f64_per_s) *red_f64_per_s = (double)threads * blocks * iters / (best * 1e-3);
  }
  CU(cudaGetLastError());
  return 0;
}

/* ------------------------------------------------------------------------- */
/* instrumentation / plumbing                                                 */
/* ------------------------------------------------------------------------- */
extern "C" int b200_get_sweep_stats(b200_solver* s, double* ms, int64_t* n_sweeps, int64_t* launches) {
  NEED(s);
  if (s->grp != nullptr) {
    double m = 0.; int64_t nl = 0, nsw = 0;
    for (b200_solver* c : grp_shards(s)) {
      double cm = 0.; int64_t cs = 0, cl = 0;
      if (b200_get_sweep_stats(c, &cm, &cs, &cl)) return 1;
      m = std::max(m, cm); nl += cl; nsw = cs;
    }
    if (ms) *ms = m;
    if (n_sweeps) *n_sweeps = nsw;
    if (launches) *launches = nl;
    return 0;
  }
  if (resolve_events(s)) return 1;
  if (ms) *ms = s->sweep_ms;
  if (n_sweeps) *n_sweeps = s->n_sweeps;
  if (launches) *launches = s->n_launches;
  return 0;
}
extern "C" int b200_reset_sweep_stats(b200_solver* s) {
  NEED(s);
  GRP_ALL(s, b200_reset_sweep_stats(c));
  if (resolve_events(s)) return 1;
  s->sweep_ms = 0.;
  s->n_sweeps = 0;
  s->n_launches = 0;
  return 0;
}
extern "C" int b200_synchronize(b200_solver* s) {
  NEED(s);
  GRP_ALL(s, b200_synchronize(c));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}
extern "C" int b200_device_pointer(b200_solver* s, const char* name, void** ptr, int64_t* n) {
  NEED_FINAL(s);
  GRP_FIRST(s, b200_device_pointer(c, name, ptr, n));
  if (!name || !ptr || !n) return fail("b200_device_pointer: null argument");
  const int64_t nphi = s->n_fsr * s->G;
  if (!strcmp(name, "scalar_flux")) { *ptr = s->phi.p; *n = nphi; }
  else if (!strcmp(name, "old_scalar_flux")) { *ptr = s->phi_old.p; *n = nphi; }
  else if (!strcmp(name, "reduced_sources")) { *ptr = s->qst.p; *n = 2 * nphi; }
  else if (!strcmp(name, "start_flux")) { *ptr = s->psi_start; *n = s->n_trk * 2 * (int64_t)s->F; }
  else if (!strcmp(name, "scalar_flux_moments")) {
    if (!s->linear) return fail("b200_device_pointer: 'scalar_flux_moments' exists for linear-source solvers only");
    *ptr = s->phi_m.p; *n = 3 * nphi;
  }
  else if (!strcmp(name, "scalar_flux_fixed")) {
    if (!s->cfg.deterministic) return fail("b200_device_pointer: 'scalar_flux_fixed' exists in deterministic mode only");
    *ptr = s->phi_fx.p; *n = s->n_fsr * s->GP;
  }
  else return fail("b200_device_pointer: unknown array '%s'", name);
  return 0;
}
/* deterministic multi-GPU: with defer=1 b200_transport_sweep leaves the tally in the int64
 * buffer ("scalar_flux_fixed") so that the host can sum it across ranks exactly;
 * b200_finish_fixed_tally then converts it into scalar_flux. */
extern "C" int b200_defer_fixed_tally(b200_solver* s, int32_t defer) {
  NEED(s);
  if (s->grp != nullptr) return fail("b200_defer_fixed_tally: a multi-device solver reduces its fixed-point tallies itself");
  if (!s->cfg.deterministic) return fail("b200_defer_fixed_tally: deterministic mode only");
  s->defer_fx_convert = defer != 0;
  return 0;
}
extern "C" int b200_finish_fixed_tally(b200_solver* s) {
  NEED_FINAL(s);
  GRP_ALL(s, b200_finish_fixed_tally(c));
  if (!s->cfg.deterministic) return fail("b200_finish_fixed_tally: deterministic mode only");
  fx_to_double_kernel<<<grid_for(s->n_fsr * s->G, 256), 256, 0, s->stream>>>(fsr_args(s), s->phi_fx.p, s->GP);
  CU(cudaGetLastError());
  s->n_launches++;
  return 0;
}

/* A host that captures its own CUDA graph around the split iteration (begin -> all-reduce -> end,
 * e.g. torch.cuda.graph with an NCCL all-reduce inside) brackets the capture with this: while set,
 * the engine records no timing events and issues no host synchronisation. */
extern "C" int b200_set_capturing(b200_solver* s, int32_t capturing) {
  NEED(s);
  GRP_ALL(s, b200_set_capturing(c, capturing));
  s->capturing = capturing != 0;
  return 0;
}

extern "C" int b200_set_stream(b200_solver* s, void* cuda_stream) {
  NEED(s);
  if (s->grp != nullptr) return fail("b200_set_stream: a multi-device solver owns one stream per device");
  CU(cudaStreamSynchronize(s->stream));
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  s->stream = (cudaStream_t)cuda_stream;
  s->own_stream = false;
  return 0;
}

#include "group_impl.cuh"
#include "cmfd_impl.cuh"
