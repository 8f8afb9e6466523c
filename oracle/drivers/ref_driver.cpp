/**
 * @file ref_driver.cpp
 * @brief Runs the UNMODIFIED reference CPUSolver / CPULSSolver on one of the
 *        restated input decks (models.cpp), prints its results in the format of
 *        the reference's regression harness (tests/testing_harness.py:158-207)
 *        and optionally dumps the flattened tracks as a B2TRK file.
 *        TEST INFRASTRUCTURE + CPU baseline; never part of the product path.
 *
 * Usage: ref_driver --model NAME [--dims 2|3] [--azim N] [--spacing S]
 *                   [--polar N] [--zspacing S] [--formation explicit|otf-tracks|otf-stacks]
 *                   [--quad ty|equal-angle|gl|equal-weight|leonard] [--groups70]
 *                   [--solver cpu|cpuls|b200|b200-fused|b200ls|both] [--ls (with both: linear source)] [--mode eigen|none] [--tol T]
 *        --solver both: CPUSolver and B200Solver in the same process on the same tracks,
 *        prints delta k_eff (pcm), max relative flux error and both sweep times.
 *                   [--max-iters N] [--threads N] [--res fission|flux|total]
 *                   [--axial N (c5g7-2d, dims 3: axial layers of the root lattice)]
 *                   [--devices 0,1,.. (--solver both: GPUs behind the one B200Solver; a device may repeat)]
 *                   [--cmfd-relax F (Cmfd::setCMFDRelaxationFactor)] [--cmfd-sor F] [--results-fsrs ("# FSRs:" line in --results)]
 *                   [--cmfd NXxNY[xNZ]] [--host-cmfd (B200 solvers: the reference's host Cmfd instead of the device CMFD)]
 *                   [--check-cmfd-split (compare the library's current-splitting tables with Cmfd's, no GPU needed)]
 *                   [--dump-tracks FILE] [--results FILE] [--json FILE] [--quiet] [--balance]
 *                   [--azim-sequence n1,n2,.. (tracks laid again with another azimuthal count before every solve)]
 *                   [--clone-materials (with --repeat: cells refilled with clones of their Materials before every solve)]
 *                   [--multisim-groups (hom-inf: 1-group, then 2-group data in the same Material, two solves)]
 *                   [--fission-rates (--json also holds computeFSRFissionRates with nu = false and nu = true)]
 *                   [--repeat N (the same eigenvalue solve N times on one solver; --results then holds one "Iters / keff" line per solve)]
 *                   [--restart (Solver::setRestartStatus(true) and a second computeEigenvalue)] [--otf-transport]
 *                   [--seg-zones z0,z1,.. (TrackGenerator3D::setSegmentationZones)]
 *                   [--cmfd-widths "x..;y..;z.." (Cmfd::setWidths; give --cmfd 1x1 as well)] [--cmfd-axial-interp 0|1|2]
 *                   [--stabilize F:T | --stabilize-sequence F:T,F:T,...] [--negative-water-scatter]
 *                   [--vacuum-mask M] [--periodic-mask M] (pwr-assembly: bit 0 xmin, 1 xmax, 2 ymin, 3 ymax)
 *                   [--cmfd-all-groups (no Cmfd::setGroupStructure)] [--symmetry (Geometry::useSymmetry(true, true, true))]
 *                   [--max-tau T (Solver::setMaxOpticalLength)] [--no-keff] [--results-tracks] [--results-segments]
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

#include "CPUSolver.h"
#include "CPULSSolver.h"
#include "Cmfd.h"
#include "TrackGenerator3D.h"
#include "log.h"

#include "models.h"
#include "../../openmoc_b200/cpp/b200_flatten.h"
#include "../../openmoc_b200/cpp/b200_cmfd_view.h"
#include "../../openmoc_b200/cpp/B200Solver.h"
#include "../../openmoc_b200/cpp/B200LSSolver.h"

/* the reference's private current-splitting rules (Cmfd::getVertexSplitSurfaces / getEdgeSplitSurfaces,
 * src/Cmfd.cpp:2348-2480), reached through an explicit instantiation: the checker of
 * b200_cmfd_split_targets */
namespace {
typedef void (Cmfd::*SplitFn)(int, int, std::vector<int>*);
template <typename Tag, typename Tag::type Member>
struct Rob { friend typename Tag::type get(Tag) { return Member; } };
struct TVertex { typedef SplitFn type; friend type get(TVertex); };
struct TEdge { typedef SplitFn type; friend type get(TEdge); };
template struct Rob<TVertex, &Cmfd::getVertexSplitSurfaces>;
template struct Rob<TEdge, &Cmfd::getEdgeSplitSurfaces>;

/* --check-cmfd-split NXxNYxNZ:b0b1b2b3b4b5 (boundaryType digit per face X_MIN..Z_MAX) */
int check_cmfd_split(const char* spec) {
  int nx = 1, ny = 1, nz = 1;
  char bcs[16] = "111111";
  sscanf(spec, "%dx%dx%d:%6s", &nx, &ny, &nz, bcs);
  Cmfd cmfd;
  cmfd.setNumX(nx); cmfd.setNumY(ny); cmfd.setNumZ(nz);
  int32_t bc[6];
  for (int s = 0; s < 6; s++) { bc[s] = bcs[s] - '0'; cmfd.setBoundary(s, (boundaryType)bc[s]); }
  long checked = 0, mismatches = 0;
  std::vector<int> ref;
  for (int cell = 0; cell < nx * ny * nz; cell++)
    for (int surf = NUM_FACES; surf < NUM_SURFACES; surf++) {
      if (surf < NUM_FACES + NUM_EDGES) (cmfd.*get(TEdge()))(cell, surf, &ref);
      else (cmfd.*get(TVertex()))(cell, surf, &ref);
      int32_t mine[6], n = 0;
      if (b200_cmfd_split_targets(nx, ny, nz, bc, cell, surf, mine, &n)) { fprintf(stderr, "%s\n", b200_last_error()); return 2; }
      bool same = (n == (int)ref.size());
      for (int j = 0; same && j < n; j++) same = (mine[j] == ref[j]);
      checked++;
      if (!same) mismatches++;
    }
  printf("{\"mesh\": \"%s\", \"checked\": %ld, \"mismatches\": %ld}\n", spec, checked, mismatches);
  return mismatches == 0 ? 0 : 1;
}
}  // namespace

static const char* arg(int argc, char** argv, const char* key, const char* dflt) {
  for (int i = 1; i < argc - 1; i++)
    if (!strcmp(argv[i], key)) return argv[i + 1];
  return dflt;
}
static bool flag(int argc, char** argv, const char* key) {
  for (int i = 1; i < argc; i++)
    if (!strcmp(argv[i], key)) return true;
  return false;
}

int main(int argc, char** argv) {
  if (strlen(arg(argc, argv, "--check-cmfd-split", "")) > 0) return check_cmfd_split(arg(argc, argv, "--check-cmfd-split", ""));
  std::string model_name = arg(argc, argv, "--model", "pin-cell");
  int dims = atoi(arg(argc, argv, "--dims", "2"));
  int num_azim = atoi(arg(argc, argv, "--azim", "4"));
  double spacing = atof(arg(argc, argv, "--spacing", "0.1"));
  int num_polar = atoi(arg(argc, argv, "--polar", dims == 3 ? "2" : "6"));
  double z_spacing = atof(arg(argc, argv, "--zspacing", "0.1"));
  std::string formation = arg(argc, argv, "--formation", "otf-tracks");
  std::string quad_name = arg(argc, argv, "--quad", "default");
  std::string solver_name = arg(argc, argv, "--solver", "cpu");
  std::string mode = arg(argc, argv, "--mode", "eigen");
  double tol = atof(arg(argc, argv, "--tol", "1e-5"));
  int max_iters = atoi(arg(argc, argv, "--max-iters", "500"));
  int threads = atoi(arg(argc, argv, "--threads", "1"));
  std::string res = arg(argc, argv, "--res", "fission");
  std::string dump_tracks = arg(argc, argv, "--dump-tracks", "");
  std::string results = arg(argc, argv, "--results", "");
  std::string json = arg(argc, argv, "--json", "");
  bool fluxes_in_results = !flag(argc, argv, "--no-fluxes");

  if (flag(argc, argv, "--quiet")) set_log_level("WARNING");
  else set_log_level("NORMAL");

  /* --stabilize FACTOR:TYPE (0 DIAGONAL, 1 YAMAMOTO, 2 GLOBAL): Solver::stabilizeTransport */
  double stab_factor = 0.; int stab_type = -1;
  {
    std::string st = arg(argc, argv, "--stabilize", "");
    if (!st.empty()) sscanf(st.c_str(), "%lf:%d", &stab_factor, &stab_type);
  }
  /* --stabilize-sequence F:T,F:T,...: one eigenvalue solve per entry on the same solver, results of the last one
   * (tests/test_transport_stabilization runs DIAGONAL, YAMAMOTO, GLOBAL in a row) */
  std::vector<std::pair<double, int> > stab_sequence;
  {
    std::string st = arg(argc, argv, "--stabilize-sequence", "");
    size_t pos = 0;
    while (pos < st.size()) {
      double f = 0.; int t = 0, used = 0;
      if (sscanf(st.c_str() + pos, "%lf:%d%n", &f, &t, &used) != 2) break;
      stab_sequence.push_back(std::make_pair(f, t));
      pos += used;
      if (pos < st.size() && st[pos] == ',') pos++;
    }
  }
  /* --max-tau X: Solver::setMaxOpticalLength (segments are cut at this optical length) */
  const double max_tau_arg = atof(arg(argc, argv, "--max-tau", "0"));
  set_axial_layers(atoi(arg(argc, argv, "--axial", "1")));
  set_boundary_masks(atoi(arg(argc, argv, "--vacuum-mask", "0")), atoi(arg(argc, argv, "--periodic-mask", "0")));
  Model md = build_model(model_name, dims);
  if (flag(argc, argv, "--groups70")) set_70_group_xs(md);
  /* tests/test_transport_stabilization: a large negative in-scatter in the moderator */
  if (flag(argc, argv, "--negative-water-scatter")) md.materials["Water"]->setSigmaSByGroup(-1., 4, 4);
  Geometry* geometry = md.geometry;
  /* tests/test_forward_3D_lattice_symmetry: Geometry::useSymmetry on the three axes */
  if (flag(argc, argv, "--symmetry")) geometry->useSymmetry(true, true, true);
  /* CMFD as in tests/test_cmfd_pwr_assembly / sample-input/benchmarks/c5g7/c5g7-2d.py:51-56 */
  std::string cmfd_arg = arg(argc, argv, "--cmfd", "");
  if (!cmfd_arg.empty()) {
    int nx = 1, ny = 1, nz = 1;
    sscanf(cmfd_arg.c_str(), "%dx%dx%d", &nx, &ny, &nz);
    Cmfd* cmfd = new Cmfd();
    cmfd->setSORRelaxationFactor(1.5);
    /* --cmfd-widths "x0,x1,..;y0,..;z0,.." instead of a uniform lattice (Cmfd::setWidths, tests/test_cmfd_axial_interpolation_*) */
    std::string cw = arg(argc, argv, "--cmfd-widths", "");
    if (!cw.empty()) {
      std::vector<std::vector<double> > widths(1);
      for (size_t pos = 0; pos < cw.size();) {
        if (cw[pos] == ';') { widths.push_back(std::vector<double>()); pos++; continue; }
        if (cw[pos] == ',') { pos++; continue; }
        char* end = NULL;
        widths.back().push_back(strtod(cw.c_str() + pos, &end));
        pos = end - cw.c_str();
      }
      cmfd->setWidths(widths);
    }
    else if (dims == 3) cmfd->setLatticeStructure(nx, ny, nz);
    else cmfd->setLatticeStructure(nx, ny);
    if (strlen(arg(argc, argv, "--cmfd-axial-interp", "")) > 0) cmfd->useAxialInterpolation(atoi(arg(argc, argv, "--cmfd-axial-interp", "0")));
    /* --cmfd-all-groups: no setGroupStructure, one CMFD group per MOC group (tests/test_split_segments_cmfd) */
    if (geometry->getNumEnergyGroups() == 7 && !flag(argc, argv, "--groups70") && !flag(argc, argv, "--cmfd-all-groups")) {
      std::vector<std::vector<int> > groups(2);
      for (int g = 1; g <= 3; g++) groups[0].push_back(g);
      for (int g = 4; g <= 7; g++) groups[1].push_back(g);
      cmfd->setGroupStructure(groups);
    }
    if (!flag(argc, argv, "--no-knearest")) cmfd->setKNearest(3);
    if (strlen(arg(argc, argv, "--cmfd-relax", "")) > 0) cmfd->setCMFDRelaxationFactor(atof(arg(argc, argv, "--cmfd-relax", "0.7")));
    if (strlen(arg(argc, argv, "--cmfd-sor", "")) > 0) cmfd->setSORRelaxationFactor(atof(arg(argc, argv, "--cmfd-sor", "1.5")));
    if (flag(argc, argv, "--no-flux-limiting")) cmfd->useFluxLimiting(false);   /* diagnostics */
    if (flag(argc, argv, "--rebalance")) cmfd->rebalanceSigmaT(true);          /* starting currents tallied every sweep */
    geometry->setCmfd(cmfd);
  }
  geometry->initializeFlatSourceRegions();

  Quadrature* quad = NULL;
  if (quad_name == "equal-angle") quad = new EqualAnglePolarQuad();
  else if (quad_name == "ty") quad = new TYPolarQuad();
  else if (quad_name == "gl") quad = new GLPolarQuad();
  else if (quad_name == "equal-weight") quad = new EqualWeightPolarQuad();
  else if (quad_name == "leonard") quad = new LeonardPolarQuad();
  if (quad != NULL) {
    /* without the azimuthal count generateTracks() silently discards a user
     * quadrature (src/TrackGenerator.cpp:802-806) */
    quad->setNumAzimAngles(num_azim);
    quad->setNumPolarAngles(num_polar);
  }

  TrackGenerator* tg;
  if (dims == 3) {
    TrackGenerator3D* tg3 = new TrackGenerator3D(geometry, num_azim, num_polar, spacing, z_spacing);
    if (formation == "explicit") tg3->setSegmentFormation(EXPLICIT_3D);
    else if (formation == "otf-stacks") tg3->setSegmentFormation(OTF_STACKS);
    else tg3->setSegmentFormation(OTF_TRACKS);
    /* --seg-zones z0,z1,...: TrackGenerator3D::setSegmentationZones (tests/test_axial_segmentation) */
    std::string zones = arg(argc, argv, "--seg-zones", "");
    if (!zones.empty()) {
      std::vector<double> z;
      for (size_t pos = 0; pos < zones.size();) {
        if (zones[pos] == ',') { pos++; continue; }
        char* end = NULL;
        z.push_back(strtod(zones.c_str() + pos, &end));
        pos = end - zones.c_str();
      }
      tg3->setSegmentationZones(z);
    }
    tg = tg3;
  } else {
    tg = new TrackGenerator(geometry, num_azim, spacing);
  }
  if (quad != NULL) tg->setQuadrature(quad);
  /* the harness ray-traces single-threaded in 2D "for FSR reproducibility"
   * (tests/testing_harness.py:83-85); OTF 3D needs tg threads == solver threads */
  tg->setNumThreads(dims == 3 ? threads : 1);
  tg->generateTracks();

  residualType rt = FISSION_SOURCE;
  if (res == "flux") rt = SCALAR_FLUX;
  else if (res == "total") rt = TOTAL_SOURCE;

  /* B200Solver in the reference's own process, next to CPUSolver on the same tracks */
  if (solver_name == "both") {
    long n_fsr = geometry->getNumFSRs();
    int G = geometry->getNumEnergyGroups();
    const bool ls = flag(argc, argv, "--ls");
    CPUSolver* cpu_p = ls ? new CPULSSolver(tg) : new CPUSolver(tg);
    CPUSolver& cpu = *cpu_p;
    cpu.setNumThreads(threads);
    cpu.setConvergenceThreshold(tol);
    if (flag(argc, argv, "--verbose")) cpu.setVerboseIterationReport();
    if (flag(argc, argv, "--balance")) cpu.setKeffFromNeutronBalance();
    if (stab_type >= 0) cpu.stabilizeTransport(stab_factor, (stabilizationType)stab_type);
    if (max_tau_arg > 0.) cpu.setMaxOpticalLength(max_tau_arg);
    cpu.computeEigenvalue(max_iters, rt);
    Timer timer;
    double cpu_sweep = timer.getSplit("Transport Sweep");
    double cpu_total = timer.getSplit("Total time");
    double cpu_cmfd = timer.getSplit("Total CMFD time");
    std::vector<FP_PRECISION> phi_cpu(n_fsr * G), phi_gpu(n_fsr * G);
    cpu.getFluxes(phi_cpu.data(), n_fsr * G);
    double k_cpu = cpu.getKeff();
    int it_cpu = cpu.getNumIterations();

    B200Solver* gpu_flat = ls ? NULL : new B200Solver(tg);
    B200LSSolver* gpu_ls = ls ? new B200LSSolver(tg) : NULL;
    Solver& gpu = ls ? *(Solver*)gpu_ls : *(Solver*)gpu_flat;
    if (ls) gpu_ls->setNumThreads(threads); else gpu_flat->setNumThreads(threads);
    if (flag(argc, argv, "--host-cmfd")) { if (ls) gpu_ls->setCmfdOnDevice(false); else gpu_flat->setCmfdOnDevice(false); }
    {
      /* --devices 0,1,...: several GPUs (or several shards on one GPU) behind the one B200Solver */
      std::string dl = arg(argc, argv, "--devices", "");
      std::vector<int> devs;
      for (size_t i = 0; i < dl.size();) {
        size_t j = dl.find(',', i);
        if (j == std::string::npos) j = dl.size();
        devs.push_back(atoi(dl.substr(i, j - i).c_str()));
        i = j + 1;
      }
      if (!devs.empty()) { if (ls) gpu_ls->setDevices(devs); else gpu_flat->setDevices(devs); }
    }
    gpu.setConvergenceThreshold(tol);
    if (flag(argc, argv, "--verbose")) gpu.setVerboseIterationReport();
    if (flag(argc, argv, "--balance")) gpu.setKeffFromNeutronBalance();
    if (stab_type >= 0) gpu.stabilizeTransport(stab_factor, (stabilizationType)stab_type);
    if (max_tau_arg > 0.) gpu.setMaxOpticalLength(max_tau_arg);
    gpu.computeEigenvalue(max_iters, rt);
    double gpu_sweep = timer.getSplit("Transport Sweep");
    double gpu_total = timer.getSplit("Total time");
    double gpu_cmfd = timer.getSplit("Total CMFD time");
    double cmfd_dev_ms = 0.; long cmfd_lin = 0;
    if (ls) gpu_ls->getCmfdStats(&cmfd_dev_ms, &cmfd_lin); else gpu_flat->getCmfdStats(&cmfd_dev_ms, &cmfd_lin);
    gpu.getFluxes(phi_gpu.data(), n_fsr * G);
    double err = 0.;
    for (long i = 0; i < n_fsr * G; i++) {
      double d = fabs(phi_gpu[i] - phi_cpu[i]) / fabs(phi_cpu[i]);
      if (d > err) err = d;
    }
    double dev_ms = 0.; long sweeps = 0;
    if (ls) gpu_ls->getSweepStats(&dev_ms, &sweeps); else gpu_flat->getSweepStats(&dev_ms, &sweeps);
    long n_seg = tg->getNumSegments();
    int F = (dims == 3) ? G : G * tg->getQuadrature()->getNumPolarAngles() / 2;
    printf("{\"model\": \"%s\", \"n_segments\": %ld, \"n_fsrs\": %ld, \"cpu_threads\": %d, "
           "\"cpu_keff\": %.12f, \"b200_keff\": %.12f, \"dk_pcm\": %.3e, \"max_rel_flux_err\": %.3e, "
           "\"cpu_iters\": %d, \"b200_iters\": %d, \"cpu_sweep_s\": %.6g, \"b200_sweep_s\": %.6g, "
           "\"b200_sweep_kernel_s\": %.6g, \"cpu_integrations_per_s\": %.4e, \"b200_integrations_per_s\": %.4e, "
           "\"cpu_total_s\": %.6g, \"b200_total_s\": %.6g, \"cmfd_on_device\": %s, \"cpu_cmfd_s\": %.6g, "
           "\"b200_cmfd_s\": %.6g, \"b200_cmfd_kernels_s\": %.6g, \"b200_cmfd_sor_iterations\": %ld}\n",
           model_name.c_str(), n_seg, n_fsr, threads, k_cpu, gpu.getKeff(), fabs(gpu.getKeff() - k_cpu) * 1e5, err,
           it_cpu, gpu.getNumIterations(), cpu_sweep, gpu_sweep, dev_ms * 1e-3,
           2.0 * F * n_seg * it_cpu / cpu_sweep, 2.0 * F * n_seg * gpu.getNumIterations() / gpu_sweep,
           cpu_total, gpu_total, (ls ? gpu_ls->isCmfdOnDevice() : gpu_flat->isCmfdOnDevice()) ? "true" : "false",
           cpu_cmfd, gpu_cmfd, cmfd_dev_ms * 1e-3, cmfd_lin);
    return 0;
  }

  Solver* solver;
  CPUSolver* cpu_solver = NULL;
  B200Solver* b200_solver = NULL;
  if (solver_name == "refgpu") {
    /* the reference's own GPUSolver, recompiled for sm_100a (oracle/Makefile), loaded late so
     * that ref_driver itself does not need a CUDA runtime */
    std::string self = argv[0];
    std::string dir = self.find('/') == std::string::npos ? "." : self.substr(0, self.rfind('/'));
    void* so = dlopen((dir + "/libopenmoc_refgpu.so").c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (so == NULL) { fprintf(stderr, "ref_driver: %s\n", dlerror()); return 2; }
    typedef Solver* (*factory_t)(TrackGenerator*, int, int);
    factory_t make = (factory_t)dlsym(so, "make_ref_gpu_solver");
    if (make == NULL) { fprintf(stderr, "ref_driver: %s\n", dlerror()); return 2; }
    solver = make(tg, atoi(arg(argc, argv, "--gpu-blocks", "0")), atoi(arg(argc, argv, "--gpu-threads", "0")));
  }
  else if (solver_name == "b200" || solver_name == "b200-fused") solver = b200_solver = new B200Solver(tg);
  else if (solver_name == "b200ls") {
    B200LSSolver* ls = new B200LSSolver(tg);
    ls->setNumThreads(threads);
    if (flag(argc, argv, "--host-cmfd")) ls->setCmfdOnDevice(false);
    solver = ls;
  }
  else if (solver_name == "cpuls") solver = cpu_solver = new CPULSSolver(tg);
  else solver = cpu_solver = new CPUSolver(tg);
  if (cpu_solver != NULL) cpu_solver->setNumThreads(threads);
  if (flag(argc, argv, "--otf-transport")) solver->setOTFTransport();     /* tests/test_OTF_transport (CPU solvers trace while they sweep) */
  if (b200_solver != NULL) b200_solver->setNumThreads(threads);   /* host side: flatten, Cmfd */
  if (b200_solver != NULL && flag(argc, argv, "--host-cmfd")) b200_solver->setCmfdOnDevice(false);
  solver->setConvergenceThreshold(tol);
  if (flag(argc, argv, "--balance")) solver->setKeffFromNeutronBalance();   /* Solver.cpp:2047 */

  /* --fixed-source g:value[,g:value...] on the model's source cell (Solver::setFixedSourceByCell) */
  std::string fixed = arg(argc, argv, "--fixed-source", "");
  if (!fixed.empty()) {
    if (md.source_cell == NULL) { fprintf(stderr, "ref_driver: model has no source cell\n"); return 2; }
    size_t pos = 0;
    while (pos < fixed.size()) {
      int g = 0; double v = 0.; int used = 0;
      if (sscanf(fixed.c_str() + pos, "%d:%lf%n", &g, &v, &used) != 2) break;
      solver->setFixedSourceByCell(md.source_cell, g, v);
      pos += used;
      if (pos < fixed.size() && fixed[pos] == ',') pos++;
    }
  }

  /* --fixed-moments g:x:y:z[,...] on the source cell (CPULSSolver::setFixedSourceMomentsByCell) */
  std::string fmom = arg(argc, argv, "--fixed-moments", "");
  if (!fmom.empty()) {
    CPULSSolver* ls_solver = dynamic_cast<CPULSSolver*>(solver);
    if (md.source_cell == NULL || ls_solver == NULL) { fprintf(stderr, "ref_driver: --fixed-moments needs a source cell and a linear-source solver\n"); return 2; }
    size_t pos = 0;
    while (pos < fmom.size()) {
      int g = 0; double x = 0., y = 0., z = 0.; int used = 0;
      if (sscanf(fmom.c_str() + pos, "%d:%lf:%lf:%lf%n", &g, &x, &y, &z, &used) != 4) break;
      ls_solver->setFixedSourceMomentsByCell(md.source_cell, g, x, y, z);
      pos += used;
      if (pos < fmom.size() && fmom[pos] == ',') pos++;
    }
  }
  if (flag(argc, argv, "--allow-negative")) solver->allowNegativeFluxes(true);
  if (stab_type >= 0) solver->stabilizeTransport(stab_factor, (stabilizationType)stab_type);
  if (max_tau_arg > 0.) solver->setMaxOpticalLength(max_tau_arg);     /* tests/test_split_segments */

  const int repeat = atoi(arg(argc, argv, "--repeat", "1"));
  std::vector<int> azim_sequence;
  {
    std::string st = arg(argc, argv, "--azim-sequence", "");
    for (size_t pos = 0; pos < st.size();) {
      if (st[pos] == ',') { pos++; continue; }
      char* end = NULL;
      azim_sequence.push_back((int)strtol(st.c_str() + pos, &end, 10));
      pos = end - st.c_str();
    }
  }
  std::string multisim;
  if (mode == "eigen" && !stab_sequence.empty()) {
    for (size_t i = 0; i < stab_sequence.size(); i++) {
      solver->stabilizeTransport(stab_sequence[i].first, (stabilizationType)stab_sequence[i].second);
      solver->computeEigenvalue(max_iters, rt);
    }
  } else if (mode == "eigen" && !azim_sequence.empty()) {
    /* tests/test_multisim_num_azim: the tracks are laid again with another number of azimuthal angles between
     * the solves, on the same TrackGenerator and the same solver */
    std::string counts;
    for (size_t i = 0; i < azim_sequence.size(); i++) {
      tg->setNumAzim(azim_sequence[i]);
      tg->generateTracks();
      solver->setTrackGenerator(tg);
      solver->computeEigenvalue(max_iters, rt);
      char line[96];
      snprintf(line, sizeof line, "# tracks: %ld\t# segments: %ld\n", (long)tg->getNumTracks(), (long)tg->getNumSegments());
      counts += line;
      snprintf(line, sizeof line, "Iters: %d\tkeff: %12.5E\n", solver->getNumIterations(), solver->getKeff());
      multisim += line;
    }
    multisim = counts + multisim;
  } else if (mode == "eigen" && flag(argc, argv, "--multisim-groups")) {
    /* tests/test_multisim_num_groups: the infinite medium solved with 1-group, then with 2-group data given to the
     * same Material object, on the same tracks and the same solver */
    Material* m = md.materials["infinite medium"];
    if (m == NULL) { fprintf(stderr, "ref_driver: --multisim-groups needs --model hom-inf\n"); return 2; }
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 0) {
        double nsf[1] = {0.0994076580}, ss[1] = {0.383259177}, chi[1] = {1.0}, st[1] = {0.452648699};
        m->setNumEnergyGroups(1);
        m->setNuSigmaF(nsf, 1); m->setSigmaS(ss, 1); m->setChi(chi, 1); m->setSigmaT(st, 1);
      } else {
        double nsf[2] = {0.0015, 0.325}, ss[4] = {0.1, 0.117, 0., 1.42}, chi[2] = {1.0, 0.0}, st[2] = {0.2208, 1.604};
        m->setNumEnergyGroups(2);
        m->setNuSigmaF(nsf, 2); m->setSigmaS(ss, 4); m->setChi(chi, 2); m->setSigmaT(st, 2);
      }
      solver->computeEigenvalue(max_iters, rt);
      char line[96];
      snprintf(line, sizeof line, "Iters: %d\tkeff: %12.5E\n", solver->getNumIterations(), solver->getKeff());
      multisim += line;
    }
  } else if (mode == "eigen" && repeat > 1) {
    /* tests/testing_harness.py:398-425 (MultiSimTestHarness): the same solve several times on one solver object */
    for (int i = 0; i < repeat; i++) {
      if (flag(argc, argv, "--clone-materials")) {
        /* tests/test_multisim_materials: every material cell is refilled with a clone of its Material before each solve */
        std::map<int, Cell*> cells = geometry->getAllMaterialCells();
        for (std::map<int, Cell*>::iterator it = cells.begin(); it != cells.end(); ++it)
          it->second->setFill(it->second->getFillMaterial()->clone());
      }
      solver->computeEigenvalue(max_iters, rt);
      char line[96];
      snprintf(line, sizeof line, "Iters: %d\tkeff: %12.5E\n", solver->getNumIterations(), solver->getKeff());
      multisim += line;
    }
  } else if (mode == "eigen") {
    if (solver_name == "b200-fused") b200_solver->computeEigenvalueFused(max_iters, rt);
    else solver->computeEigenvalue(max_iters, rt);
    if (flag(argc, argv, "--restart")) {          /* tests/test_cmfd_restart: a second solve that keeps the fluxes */
      solver->setRestartStatus(true);
      solver->computeEigenvalue(max_iters, rt);
    }
  } else if ((mode == "flux" || mode == "source") && repeat > 1) {
    /* tests/test_multisim_fixed_source: the fixed-source solve several times on one solver, iterations and fluxes of each */
    long nf = geometry->getNumFSRs() * geometry->getNumEnergyGroups();
    std::vector<FP_PRECISION> phi(nf);
    for (int i = 0; i < repeat; i++) {
      if (mode == "flux") solver->computeFlux(max_iters);
      else solver->computeSource(max_iters, 1.0, rt);
      char line[64];
      snprintf(line, sizeof line, "Iters: %d\nfluxes:\n", solver->getNumIterations());
      multisim += line;
      solver->getFluxes(phi.data(), nf);
      for (long j = 0; j < nf; j++) { snprintf(line, sizeof line, "%12.6E\n", phi[j]); multisim += line; }
    }
  } else if (mode == "flux") {
    solver->computeFlux(max_iters);                       /* tests/testing_harness.py:151 */
  } else if (mode == "source") {
    solver->computeSource(max_iters, 1.0, rt);            /* tests/testing_harness.py:153 */
  } else {
    solver->initializeSolver(FORWARD);
  }
  const bool solved = (mode == "eigen" || mode == "flux" || mode == "source");

  long n_fsr = geometry->getNumFSRs();
  int G = geometry->getNumEnergyGroups();
  long n_trk = tg->getNumTracks();
  long n_seg = tg->getNumSegments();
  int F = (dims == 3) ? G : G * tg->getQuadrature()->getNumPolarAngles() / 2;

  if (solved && !flag(argc, argv, "--quiet")) solver->printTimerReport();

  /* ---- results in the harness format ---- */
  if (solved && !results.empty() && !multisim.empty()) {
    FILE* f = fopen(results.c_str(), "w");
    fputs(multisim.c_str(), f);
    fclose(f);
  } else if (solved && !results.empty()) {
    FILE* f = fopen(results.c_str(), "w");
    fprintf(f, "# Iterations: %d\n", solver->getNumIterations());
    if (mode == "eigen" && !flag(argc, argv, "--no-keff")) fprintf(f, "keff: %12.5E\n", solver->getKeff());
    if (fluxes_in_results) {
      fprintf(f, "fluxes:\n");
      std::vector<FP_PRECISION> phi(n_fsr * G);
      solver->getFluxes(phi.data(), n_fsr * G);
      for (long i = 0; i < n_fsr * G; i++) fprintf(f, "%12.6E\n", phi[i]);
    }
    /* tests/testing_harness.py:190-203: num_fsrs / num_tracks / num_segments, in this order, after the fluxes */
    if (flag(argc, argv, "--results-fsrs")) fprintf(f, "# FSRs: %ld\n", n_fsr);
    if (flag(argc, argv, "--results-tracks")) fprintf(f, "# tracks: %ld\n", n_trk);
    if (flag(argc, argv, "--results-segments")) fprintf(f, "# segments: %ld\n", n_seg);
    fclose(f);
  }

  /* ---- machine-readable summary (full precision) ---- */
  if (!json.empty()) {
    FILE* f = fopen(json.c_str(), "w");
    double sweep_time = 0., total_time = 0.;
    if (solved) {
      Timer timer;  /* splits are static/shared in the reference's Timer */
      sweep_time = timer.getSplit("Transport Sweep");
      total_time = timer.getSplit("Total time");
    }
    fprintf(f, "{\"model\": \"%s\", \"dims\": %d, \"num_azim\": %d, \"spacing\": %.17g, "
               "\"num_polar\": %d, \"solver\": \"%s\", \"threads\": %d, \"tol\": %.17g,\n",
            model_name.c_str(), dims, num_azim, spacing,
            (int)tg->getQuadrature()->getNumPolarAngles(), solver_name.c_str(), threads, tol);
    fprintf(f, " \"n_tracks\": %ld, \"n_segments\": %ld, \"n_fsrs\": %ld, \"num_groups\": %d, "
               "\"fluxes_per_track\": %d,\n", n_trk, n_seg, n_fsr, G, F);
    if (solved) {
      int it = solver->getNumIterations();
      /* FSRs of the source cell, so that a flattened replay can place the same fixed source */
      if (md.source_cell != NULL) {
        fprintf(f, " \"source_fsrs\": [");
        bool first_fsr = true;
        for (long r = 0; r < n_fsr; r++)
          if (geometry->findCellContainingFSR(r) == md.source_cell) {
            fprintf(f, "%s%ld", first_fsr ? "" : ", ", r);
            first_fsr = false;
          }
        fprintf(f, "],\n");
      }
      {
        double cmfd_ms = 0.; long cmfd_lin = 0;
        B200LSSolver* lsp = dynamic_cast<B200LSSolver*>(solver);
        if (b200_solver != NULL) b200_solver->getCmfdStats(&cmfd_ms, &cmfd_lin);
        else if (lsp != NULL) lsp->getCmfdStats(&cmfd_ms, &cmfd_lin);
        Timer timer2;
        fprintf(f, " \"cmfd_time_s\": %.9g, \"cmfd_kernels_s\": %.9g, \"cmfd_sor_iterations\": %ld,\n",
                timer2.getSplit("Total CMFD time"), cmfd_ms * 1e-3, cmfd_lin);
      }
      fprintf(f, " \"iterations\": %d, \"keff\": %.17g, \"sweep_time_s\": %.9g, "
                 "\"total_time_s\": %.9g, \"integrations\": %.17g,\n",
              it, solver->getKeff(), sweep_time, total_time, 2.0 * F * (double)n_seg * it);
      std::vector<FP_PRECISION> phi(n_fsr * G);
      solver->getFluxes(phi.data(), n_fsr * G);
      fprintf(f, " \"fluxes\": [");
      for (long i = 0; i < n_fsr * G; i++) fprintf(f, "%s%.17g", i ? ", " : "", phi[i]);
      fprintf(f, "],\n");
      if (flag(argc, argv, "--fission-rates")) {
        /* Solver::computeFSRFissionRates with its default nu = false (sigma_f, not nu-sigma_f) and with nu = true */
        std::vector<double> rates(n_fsr);
        for (int nu = 0; nu < 2; nu++) {
          solver->computeFSRFissionRates(rates.data(), n_fsr, nu != 0);
          fprintf(f, " \"%s\": [", nu ? "nu_fission_rates" : "fission_rates");
          for (long i = 0; i < n_fsr; i++) fprintf(f, "%s%.17g", i ? ", " : "", rates[i]);
          fprintf(f, "],\n");
        }
      }
    }
    fprintf(f, " \"ok\": true}\n");
    fclose(f);
  }

  /* ---- flattened tracks (segments are final after the solve/initialize) ---- */
  if (!dump_tracks.empty()) {
    B200FlatTracks ft;
    /* --dump-device-otf: what the plug-in hands to the device tracer for an OTF deck (no 3D segment on the host) */
    b200_flatten(tg, &ft, true, flag(argc, argv, "--dump-device-otf"));
    B200CmfdView view;
    Cmfd* dumped_cmfd = solved ? geometry->getCmfd() : NULL;      /* the mesh is known once the solver initialised it */
    if (dumped_cmfd != NULL) b200_read_cmfd(dumped_cmfd, n_fsr, &view);
    b200_write_trackfile(ft, dump_tracks, dumped_cmfd != NULL ? &view : NULL);
    printf("[ref_driver] wrote %s: %ld tracks, %ld segments, %ld FSRs, G=%d F=%d\n",
           dump_tracks.c_str(), (long)ft.n_tracks, (long)ft.n_segments, (long)ft.n_fsrs,
           ft.num_groups, ft.fluxes_per_track);
  }
  return 0;
}
