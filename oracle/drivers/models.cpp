/**
 * @file models.cpp
 * @brief See models.h.  Our own C++ restatement of the reference's Python
 *        input decks (the Python module cannot be built here: no swig).
 */
#include "models.h"

#include <vector>

#include "Geometry.h"
#include "Universe.h"
#include "Cell.h"
#include "Surface.h"
#include "Material.h"
#include "c5g7_xs.h"

namespace {

std::map<std::string, Material*> make_c5g7_materials() {
  std::map<std::string, Material*> out;
  for (int i = 0; i < c5g7::num_materials; i++) {
    const c5g7::XS& x = c5g7::materials[i];
    Material* m = new Material(i + 1, x.name);
    m->setNumEnergyGroups(c5g7::G);
    m->setSigmaT(const_cast<double*>(x.sigma_t), c5g7::G);
    m->setSigmaS(const_cast<double*>(x.sigma_s), c5g7::G * c5g7::G);
    m->setSigmaF(const_cast<double*>(x.sigma_f), c5g7::G);
    m->setNuSigmaF(const_cast<double*>(x.nu_sigma_f), c5g7::G);
    m->setChi(const_cast<double*>(x.chi), c5g7::G);
    out[x.name] = m;
  }
  return out;
}

int g_axial_layers = 1;
int g_vacuum_mask = 0, g_periodic_mask = 0;

void fill_lattice(Lattice* lat, int ny, int nx, const std::vector<Universe*>& rows_top_down) {
  /* rows_top_down is row-major starting at the upper-left corner, exactly the
   * nested-list convention of Lattice::setUniverses (Universe.cpp:1604-1616) */
  std::vector<Universe*> tmp(rows_top_down);
  lat->setUniverses(1, ny, nx, tmp.data());
}

/* ---------------- tests/input_set.py:95-137 ---------------- */
Model pin_cell(int dims) {
  Model md;
  md.materials = make_c5g7_materials();
  ZCylinder* zcyl = new ZCylinder(0.0, 0.0, 1.0);
  XPlane* xmin = new XPlane(-2.0); XPlane* xmax = new XPlane(2.0);
  YPlane* ymin = new YPlane(-2.0); YPlane* ymax = new YPlane(2.0);
  /* --vacuum-mask 15: the deck of tests/test_krylov_forward (all four sides VACUUM) */
  xmin->setBoundaryType((g_vacuum_mask & 1) ? VACUUM : REFLECTIVE); xmax->setBoundaryType((g_vacuum_mask & 2) ? VACUUM : REFLECTIVE);
  ymin->setBoundaryType((g_vacuum_mask & 4) ? VACUUM : REFLECTIVE); ymax->setBoundaryType((g_vacuum_mask & 8) ? VACUUM : REFLECTIVE);

  Cell* fuel = new Cell();
  fuel->setFill(md.materials["UO2"]);
  fuel->addSurface(-1, zcyl);
  Cell* moderator = new Cell();
  moderator->setFill(md.materials["Water"]);
  moderator->addSurface(+1, zcyl);
  moderator->addSurface(+1, xmin); moderator->addSurface(-1, xmax);
  moderator->addSurface(+1, ymin); moderator->addSurface(-1, ymax);
  if (dims == 3) {
    ZPlane* zmin = new ZPlane(-2.0); ZPlane* zmax = new ZPlane(2.0);
    zmin->setBoundaryType(REFLECTIVE); zmax->setBoundaryType(REFLECTIVE);
    fuel->addSurface(+1, zmin); fuel->addSurface(-1, zmax);
    moderator->addSurface(+1, zmin); moderator->addSurface(-1, zmax);
  }
  Universe* root = new Universe();
  root->addCell(fuel);
  root->addCell(moderator);
  md.geometry = new Geometry();
  md.geometry->setRootUniverse(root);
  return md;
}

/* ---------------- tests/input_set.py:310-417 ---------------- */
Model simple_lattice(int dims) {
  Model md;
  md.materials = make_c5g7_materials();
  XPlane* xmin = new XPlane(-2.0); XPlane* xmax = new XPlane(2.0);
  YPlane* ymin = new YPlane(-2.0); YPlane* ymax = new YPlane(2.0);
  /* --vacuum-mask: bit 0 xmin, 1 xmax, 2 ymin, 3 ymax, 4 zmin (tests/test_OTF_transport switches three sides) */
  xmin->setBoundaryType((g_vacuum_mask & 1) ? VACUUM : REFLECTIVE); xmax->setBoundaryType((g_vacuum_mask & 2) ? VACUUM : REFLECTIVE);
  ymin->setBoundaryType((g_vacuum_mask & 4) ? VACUUM : REFLECTIVE); ymax->setBoundaryType((g_vacuum_mask & 8) ? VACUUM : REFLECTIVE);

  const double radii[3] = {0.4, 0.3, 0.2};
  Universe* pins[3];
  for (int i = 0; i < 3; i++) {
    ZCylinder* cyl = new ZCylinder(0.0, 0.0, radii[i]);
    Cell* fuel = new Cell();
    fuel->setNumRings(3);
    fuel->setNumSectors(8);
    fuel->setFill(md.materials["UO2"]);
    fuel->addSurface(-1, cyl);
    Cell* mod = new Cell();
    mod->setNumSectors(8);
    mod->setFill(md.materials["Water"]);
    mod->addSurface(+1, cyl);
    pins[i] = new Universe();
    pins[i]->addCell(fuel);
    pins[i]->addCell(mod);
  }

  Cell* lattice_cell = new Cell();
  Cell* root_cell = new Cell();
  root_cell->addSurface(+1, xmin); root_cell->addSurface(-1, xmax);
  root_cell->addSurface(+1, ymin); root_cell->addSurface(-1, ymax);
  if (dims == 3) {
    ZPlane* zmin = new ZPlane(-5.0); ZPlane* zmax = new ZPlane(5.0);
    zmin->setBoundaryType((g_vacuum_mask & 16) ? VACUUM : REFLECTIVE);
    zmax->setBoundaryType(VACUUM);
    root_cell->addSurface(+1, zmin); root_cell->addSurface(-1, zmax);
  }
  Universe* assembly = new Universe();
  Universe* root = new Universe();
  assembly->addCell(lattice_cell);
  root->addCell(root_cell);

  Lattice* lattice = new Lattice();
  lattice->setWidth(1.0, 1.0);
  fill_lattice(lattice, 2, 2, {pins[0], pins[1], pins[0], pins[2]});
  lattice_cell->setFill(lattice);

  Lattice* core = new Lattice();
  core->setWidth(2.0, 2.0);
  fill_lattice(core, 2, 2, {assembly, assembly, assembly, assembly});
  root_cell->setFill(core);

  md.geometry = new Geometry();
  md.geometry->setRootUniverse(root);
  return md;
}

/* ---------------- tests/input_set.py:420-559 (PwrAssemblyInput) ---------------- */
/* A 17 x 17 mixed-enrichment MOX assembly, every pin with 3 rings and 8 sectors in the fuel AND in the moderator (one
 * moderator Cell shared by the five pin universes, like the Python deck).  vacuum_mask / periodic_mask: bit 0 xmin,
 * 1 xmax, 2 ymin, 3 ymax (tests/test_cmfd_vacuum_boundary, tests/test_cmfd_periodic_boundaries). */
Model pwr_assembly(int, int vacuum_mask = 0, int periodic_mask = 0) {
  Model md;
  md.materials = make_c5g7_materials();
  {
    /* the Python decks read sample-input/c5g7-mgxs.h5, whose MOX-4.3% total cross section of group 7 is 0.682852;
     * the C++ decks under profile/models/c5g7 (what c5g7_xs.h restates) carry 0.68285 */
    Material* mox43 = md.materials["MOX-4.3%"];
    std::vector<double> st(mox43->getSigmaT(), mox43->getSigmaT() + c5g7::G);
    st[6] = 0.682852;
    mox43->setSigmaT(st.data(), c5g7::G);
  }
  ZCylinder* fuel_radius = new ZCylinder(0.0, 0.0, 0.54);
  XPlane* xmin = new XPlane(-10.71); XPlane* xmax = new XPlane(10.71);
  YPlane* ymin = new YPlane(-10.71); YPlane* ymax = new YPlane(10.71);
  Surface* sides[4] = {xmin, xmax, ymin, ymax};
  for (int i = 0; i < 4; i++)
    sides[i]->setBoundaryType((vacuum_mask >> i) & 1 ? VACUUM : (periodic_mask >> i) & 1 ? PERIODIC : REFLECTIVE);

  const char* fills[5] = {"MOX-4.3%", "MOX-7%", "MOX-8.7%", "Guide Tube", "Fission Chamber"};   /* template ids 1..5 */
  Cell* moderator = new Cell();
  moderator->setFill(md.materials["Water"]);
  moderator->addSurface(+1, fuel_radius);
  moderator->setNumRings(3);
  moderator->setNumSectors(8);
  Universe* pins[5];
  for (int i = 0; i < 5; i++) {
    Cell* fuel = new Cell();
    fuel->setFill(md.materials[fills[i]]);
    fuel->setNumRings(3);
    fuel->setNumSectors(8);
    fuel->addSurface(-1, fuel_radius);
    pins[i] = new Universe();
    pins[i]->addCell(fuel);
  }
  /* the Python deck creates the five fuel cells and universes first and adds the moderator afterwards */
  for (int i = 0; i < 5; i++) pins[i]->addCell(moderator);

  static const int tmpl[17][17] = {
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
    {1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1},
    {1, 2, 2, 2, 2, 4, 2, 2, 4, 2, 2, 4, 2, 2, 2, 2, 1},
    {1, 2, 2, 4, 2, 3, 3, 3, 3, 3, 3, 3, 2, 4, 2, 2, 1},
    {1, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 2, 1},
    {1, 2, 4, 3, 3, 4, 3, 3, 4, 3, 3, 4, 3, 3, 4, 2, 1},
    {1, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 1},
    {1, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 1},
    {1, 2, 4, 3, 3, 4, 3, 3, 5, 3, 3, 4, 3, 3, 4, 2, 1},
    {1, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 1},
    {1, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 1},
    {1, 2, 4, 3, 3, 4, 3, 3, 4, 3, 3, 4, 3, 3, 4, 2, 1},
    {1, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 2, 1},
    {1, 2, 2, 4, 2, 3, 3, 3, 3, 3, 3, 3, 2, 4, 2, 2, 1},
    {1, 2, 2, 2, 2, 4, 2, 2, 4, 2, 2, 4, 2, 2, 2, 2, 1},
    {1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1},
    {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}};
  std::vector<Universe*> rows;
  for (int i = 0; i < 17; i++)
    for (int j = 0; j < 17; j++) rows.push_back(pins[tmpl[i][j] - 1]);
  Lattice* assembly = new Lattice();
  assembly->setWidth(1.26, 1.26);
  fill_lattice(assembly, 17, 17, rows);

  Cell* root_cell = new Cell();
  root_cell->setFill(assembly);
  root_cell->addSurface(+1, xmin); root_cell->addSurface(-1, xmax);
  root_cell->addSurface(+1, ymin); root_cell->addSurface(-1, ymax);
  Universe* root = new Universe();
  root->addCell(root_cell);
  md.geometry = new Geometry();
  md.geometry->setRootUniverse(root);
  return md;
}

/* ---------------- tests/input_set.py:712-868 (AxialExtendedInput) ---------------- */
/* A 4 x 4 x 20 NON-UNIFORM lattice (gap, pin, pin, gap in x and y; twenty 1 cm layers), two pins replaced by the gap
 * material in one layer: extruded FSRs with different axial meshes.  The Clad material of sample-input/c5g7-mgxs.h5
 * (not among the seven of the C++ decks) is restated here from the file's datasets 'total' and 'scatter matrix'. */
Model axial_extended(int) {
  Model md;
  md.materials = make_c5g7_materials();
  {
    double total[7] = {0.13006, 0.30548, 0.32991, 0.2697, 0.27278, 0.27794, 0.29563};
    double scatter[49] = {0.097249, 0.032548, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.30398, 0.00077285, 0.0, 0.0, 0.0, 0.0,
                          0.0, 0.0, 0.32428, 0.00059405, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.2632, 0.0053135, 0.0, 0.0,
                          0.0, 0.0, 0.0, 0.0021268, 0.25395, 0.013908, 0.0, 0.0, 0.0, 0.0, 0.0, 0.01485, 0.24185,
                          0.016534, 0.0, 0.0, 0.0, 0.0, 0.0, 0.029899, 0.25716};
    double zeros[7] = {0., 0., 0., 0., 0., 0., 0.};
    Material* clad = new Material(100, "Clad");
    clad->setNumEnergyGroups(7);
    clad->setSigmaT(total, 7);
    clad->setSigmaS(scatter, 49);
    clad->setSigmaF(zeros, 7); clad->setNuSigmaF(zeros, 7); clad->setChi(zeros, 7);
    md.materials["Clad"] = clad;
  }
  const int num_sectors = 8;
  const double pin_pitch = 1.26, gap_size = 0.05;
  ZCylinder* c_fuel = new ZCylinder(0., 0., 0.54);
  ZCylinder* c_clad = new ZCylinder(0., 0., 0.57);
  ZCylinder* c_large = new ZCylinder(0., 0., 0.60);
  ZPlane* z_lo = new ZPlane(-1.0E10);
  ZPlane* z_hi = new ZPlane(1.0E10);

  Cell* fuel = new Cell(); Cell* clad = new Cell(); Cell* mod = new Cell(); Cell* gap = new Cell();
  Cell* large_fuel = new Cell(); Cell* large_mod = new Cell();
  fuel->addSurface(-1, c_fuel); fuel->addSurface(+1, z_lo); fuel->addSurface(-1, z_hi);
  fuel->setFill(md.materials["UO2"]); fuel->setNumSectors(num_sectors); fuel->setNumRings(1);
  clad->addSurface(+1, c_fuel); clad->addSurface(-1, c_clad); clad->addSurface(+1, z_lo); clad->addSurface(-1, z_hi);
  clad->setFill(md.materials["Clad"]); clad->setNumSectors(num_sectors);
  mod->addSurface(+1, c_clad); mod->addSurface(+1, z_lo); mod->addSurface(-1, z_hi);
  mod->setFill(md.materials["Water"]); mod->setNumSectors(num_sectors); mod->setNumRings(1);
  gap->addSurface(+1, z_lo); gap->addSurface(-1, z_hi);
  gap->setFill(md.materials["Clad"]);
  large_fuel->addSurface(-1, c_large); large_fuel->addSurface(+1, z_lo); large_fuel->addSurface(-1, z_hi);
  large_fuel->setFill(md.materials["UO2"]); large_fuel->setNumSectors(num_sectors);
  large_mod->addSurface(+1, c_large); large_mod->addSurface(+1, z_lo); large_mod->addSurface(-1, z_hi);
  large_mod->setFill(md.materials["Water"]); large_mod->setNumSectors(num_sectors);

  Universe* pin = new Universe(); pin->addCell(fuel); pin->addCell(clad); pin->addCell(mod);
  Universe* ugap = new Universe(); ugap->addCell(gap);
  Universe* large_pin = new Universe(); large_pin->addCell(large_fuel); large_pin->addCell(large_mod);   /* unused, as in the deck */

  const int n_z = 20;
  std::vector<double> wx = {gap_size, pin_pitch, pin_pitch, gap_size}, wy = wx, wz(n_z, 1.0);
  const double sx = 2 * gap_size + 2 * pin_pitch, sz = n_z * 1.0;
  Lattice* lattice = new Lattice();
  lattice->setWidths(wx, wy, wz);
  lattice->setOffset(sx / 2., sx / 2., sz / 2.);
  Universe* layer[16] = {ugap, ugap, ugap, ugap,  ugap, pin, pin, ugap,  ugap, pin, pin, ugap,  ugap, ugap, ugap, ugap};
  std::vector<Universe*> fill;
  for (int k = 0; k < n_z; k++) fill.insert(fill.end(), layer, layer + 16);
  fill[2 * 16 + 1 * 4 + 1] = ugap;                       /* fill_universes[2][1][1] = g: axially heterogeneous */
  lattice->setUniverses(n_z, 4, 4, fill.data());

  XPlane* xmin = new XPlane(0.); XPlane* xmax = new XPlane(sx);
  YPlane* ymin = new YPlane(0.); YPlane* ymax = new YPlane(sx);
  ZPlane* zmin = new ZPlane(0.); ZPlane* zmax = new ZPlane(sz);
  Surface* sides[6] = {xmin, xmax, ymin, ymax, zmin, zmax};
  for (int i = 0; i < 6; i++) sides[i]->setBoundaryType(REFLECTIVE);
  Cell* root_cell = new Cell();
  root_cell->setFill(lattice);
  root_cell->addSurface(+1, xmin); root_cell->addSurface(-1, xmax);
  root_cell->addSurface(+1, ymin); root_cell->addSurface(-1, ymax);
  root_cell->addSurface(+1, zmin); root_cell->addSurface(-1, zmax);
  Universe* root = new Universe();
  root->addCell(root_cell);
  md.geometry = new Geometry();
  md.geometry->setRootUniverse(root);
  return md;
}

/* ---------------- tests/input_set.py:32-92 ---------------- */
Model hom_inf(int, int vacuum_mask = 0) {   /* bit 0 xmin, 1 xmax, 2 ymin, 3 ymax set to VACUUM */
  Model md;
  double sigma_f[2] = {0.000625, 0.135416667};
  double nu_sigma_f[2] = {0.0015, 0.325};
  double sigma_s[4] = {0.1, 0.117, 0., 1.42};
  double chi[2] = {1.0, 0.0};
  double sigma_t[2] = {0.2208, 1.604};
  Material* m = new Material(1, "2-group infinite medium");
  m->setNumEnergyGroups(2);
  m->setSigmaF(sigma_f, 2); m->setNuSigmaF(nu_sigma_f, 2);
  m->setSigmaS(sigma_s, 4); m->setChi(chi, 2); m->setSigmaT(sigma_t, 2);
  md.materials["infinite medium"] = m;

  const double length = 2.5; const int n = 10;
  XPlane* xmin = new XPlane(-length / 2.); XPlane* xmax = new XPlane(length / 2.);
  YPlane* ymin = new YPlane(-length / 2.); YPlane* ymax = new YPlane(length / 2.);
  xmin->setBoundaryType((vacuum_mask & 1) ? VACUUM : REFLECTIVE);
  xmax->setBoundaryType((vacuum_mask & 2) ? VACUUM : REFLECTIVE);
  ymin->setBoundaryType((vacuum_mask & 4) ? VACUUM : REFLECTIVE);
  ymax->setBoundaryType((vacuum_mask & 8) ? VACUUM : REFLECTIVE);
  Cell* fill = new Cell();
  fill->setFill(m);
  Cell* root_cell = new Cell();
  root_cell->addSurface(+1, xmin); root_cell->addSurface(-1, xmax);
  root_cell->addSurface(+1, ymin); root_cell->addSurface(-1, ymax);
  Universe* fill_u = new Universe();
  fill_u->addCell(fill);
  Universe* root = new Universe();
  root->addCell(root_cell);
  Lattice* lat = new Lattice();
  lat->setWidth(length / n, length / n);
  fill_lattice(lat, n, n, std::vector<Universe*>(n * n, fill_u));
  root_cell->setFill(lat);
  md.geometry = new Geometry();
  md.geometry->setRootUniverse(root);
  return md;
}

/* ------- tests/test_compute_flux/test_compute_flux.py:18-79 (same deck in test_compute_source) ------- */
Model water_box(int) {
  Model md;
  md.materials = make_c5g7_materials();
  Material* water = md.materials["Water"];
  const double length = 2.5; const int n = 10;
  XPlane* xmin = new XPlane(-length / 2.); XPlane* xmax = new XPlane(length / 2.);
  YPlane* ymin = new YPlane(-length / 2.); YPlane* ymax = new YPlane(length / 2.);
  xmin->setBoundaryType(VACUUM); xmax->setBoundaryType(VACUUM);
  ymin->setBoundaryType(VACUUM); ymax->setBoundaryType(VACUUM);
  Cell* root_cell = new Cell();
  root_cell->addSurface(+1, xmin); root_cell->addSurface(-1, xmax);
  root_cell->addSurface(+1, ymin); root_cell->addSurface(-1, ymax);
  Cell* water_cell = new Cell();
  water_cell->setFill(water);
  Universe* water_u = new Universe();
  water_u->addCell(water_cell);
  Cell* source_cell = new Cell();
  source_cell->setFill(water);
  Universe* source_u = new Universe();
  source_u->addCell(source_cell);
  /* universes[0][int(lat_x)][int(lat_y)] with lat = (max - 0.5) / width = 3: fourth row from the
   * top, fourth column (the nested list is rows top-down) */
  std::vector<Universe*> rows(n * n, water_u);
  const double w = length / n;
  const int lat_x = (int)((length / 2. - 0.5) / w), lat_y = (int)((length / 2. - 0.5) / w);
  rows[lat_x * n + lat_y] = source_u;
  Lattice* lat = new Lattice();
  lat->setWidth(w, w);
  fill_lattice(lat, n, n, rows);
  root_cell->setFill(lat);
  Universe* root = new Universe();
  root->addCell(root_cell);
  md.geometry = new Geometry();
  md.geometry->setRootUniverse(root);
  md.source_cell = source_cell;
  return md;
}

/* ------- sample-input/benchmarks/c5g7/{surfaces,cells,universes,lattices,c5g7-2d}.py ------- */
Model c5g7_2d(int dims) {
  Model md;
  md.materials = make_c5g7_materials();
  std::map<std::string, Material*>& M = md.materials;

  XPlane* xmin = new XPlane(-32.13); XPlane* xmax = new XPlane(32.13);
  YPlane* ymin = new YPlane(-32.13); YPlane* ymax = new YPlane(32.13);
  xmin->setBoundaryType(REFLECTIVE); xmax->setBoundaryType(VACUUM);   /* surfaces.py:24-27 */
  ymin->setBoundaryType(VACUUM);     ymax->setBoundaryType(REFLECTIVE);
  ZCylinder* fuel_cyl = new ZCylinder(0.0, 0.0, 0.54);

  const int fuel_rings = 5, num_sectors = 4;   /* cells.py:6-8 */

  /* one shared moderator cell (8 sectors), cells.py:91,105 */
  Cell* moderator = new Cell();
  moderator->setFill(M["Water"]);
  moderator->setNumSectors(8);
  moderator->addSurface(+1, fuel_cyl);

  struct PinSpec { const char* mat; bool rings; };
  /* u, m, o, x: fuel pins without rings; g, f, p: 5 rings (cells.py:73-91) */
  const PinSpec specs[7] = {{"UO2", false}, {"MOX-4.3%", false}, {"MOX-7%", false},
                            {"MOX-8.7%", false}, {"Guide Tube", true},
                            {"Fission Chamber", true}, {"Water", true}};
  Universe* pin[7];
  for (int i = 0; i < 7; i++) {
    Cell* c = new Cell();
    c->setFill(M[specs[i].mat]);
    if (specs[i].rings) c->setNumRings(fuel_rings);
    c->setNumSectors(num_sectors);
    c->addSurface(-1, fuel_cyl);
    pin[i] = new Universe();
    pin[i]->addCell(c);
    pin[i]->addCell(moderator);
  }
  Universe *u = pin[0], *m = pin[1], *o = pin[2], *x = pin[3], *g = pin[4], *f = pin[5];

  /* plain reflector cell and its 3x3 refined mesh (lattices.py:52-58) */
  Cell* reflector = new Cell();
  reflector->setFill(M["Water"]);
  Universe* r = new Universe();
  r->addCell(reflector);
  Lattice* refined = new Lattice();
  refined->setWidth(1.26 / 3, 1.26 / 3, 100.);
  fill_lattice(refined, 3, 3, std::vector<Universe*>(9, r));
  Cell* refined_cell = new Cell();
  refined_cell->setFill(refined);
  Universe* a = new Universe();
  a->addCell(refined_cell);

  /* 17x17 assemblies (lattices.py:60-82 and 108-130) */
  Universe* uo2_t[17 * 17];
  Universe* mox_t[17 * 17];
  const char* guide_rows[17] = {
      ".................", ".................", ".....g..g..g.....", "...g.........g...",
      ".................", "..g..g..g..g..g..", ".................", ".................",
      "..g..g..f..g..g..", ".................", ".................", "..g..g..g..g..g..",
      ".................", "...g.........g...", ".....g..g..g.....", ".................",
      "................."};
  const char* mox_rows[17] = {
      "mmmmmmmmmmmmmmmmm", "mooooooooooooooom", "mooooooooooooooom", "mooooxxxxxxxoooom",
      "moooxxxxxxxxxooom", "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom",
      "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom", "mooxxxxxxxxxxxoom",
      "moooxxxxxxxxxooom", "mooooxxxxxxxoooom", "mooooooooooooooom", "mooooooooooooooom",
      "mmmmmmmmmmmmmmmmm"};
  for (int j = 0; j < 17; j++)
    for (int i = 0; i < 17; i++) {
      char gt = guide_rows[j][i];
      Universe* tube = (gt == 'g') ? g : (gt == 'f') ? f : NULL;
      uo2_t[j * 17 + i] = tube ? tube : u;
      char mc = mox_rows[j][i];
      mox_t[j * 17 + i] = tube ? tube : (mc == 'm') ? m : (mc == 'o') ? o : x;
    }

  auto make_assembly = [&](const std::vector<Universe*>& t) {
    Lattice* lat = new Lattice();
    lat->setWidth(1.26, 1.26, 100.);
    fill_lattice(lat, 17, 17, t);
    Cell* c = new Cell();
    c->setFill(lat);
    Universe* uni = new Universe();
    uni->addCell(c);
    return uni;
  };
  Universe* uu = make_assembly(std::vector<Universe*>(uo2_t, uo2_t + 289));
  Universe* mu = make_assembly(std::vector<Universe*>(mox_t, mox_t + 289));

  /* reflector assemblies (lattices.py:180-199) */
  std::vector<Universe*> right(289), bottom(289), corner(289);
  for (int j = 0; j < 17; j++)
    for (int i = 0; i < 17; i++) {
      right[j * 17 + i] = (i < 11) ? a : r;
      bottom[j * 17 + i] = (j < 11) ? a : r;
      corner[j * 17 + i] = (j < 11 && i < 11) ? a : r;
    }
  Universe* ri = make_assembly(right);
  Universe* rb = make_assembly(bottom);
  Universe* rc = make_assembly(corner);

  /* c5g7-2d.py:34-37 */
  Lattice* root_lat = new Lattice();
  if (dims == 3 && g_axial_layers > 1) {
    /* N identical axial layers: the 3 x 3 x N root lattice of profile/models/c5g7/c5g7-3d-cmfd.cpp */
    root_lat->setWidth(21.42, 21.42, 64.26 / g_axial_layers);
    std::vector<Universe*> all;
    for (int k = 0; k < g_axial_layers; k++)
      for (Universe* u9 : {uu, mu, ri, mu, uu, ri, rb, rb, rc}) all.push_back(u9);
    root_lat->setUniverses(g_axial_layers, 3, 3, all.data());
  } else {
    root_lat->setWidth(21.42, 21.42);
    fill_lattice(root_lat, 3, 3, {uu, mu, ri, mu, uu, ri, rb, rb, rc});
  }
  Cell* root_cell = new Cell();
  root_cell->addSurface(+1, xmin); root_cell->addSurface(-1, xmax);
  root_cell->addSurface(+1, ymin); root_cell->addSurface(-1, ymax);
  if (dims == 3) {
    /* axially uniform ("extruded") core between the planes the 3D decks use
     * (profile/models/c5g7/c5g7-3d-cmfd.cpp: -32.13 reflective, +32.13 vacuum) */
    ZPlane* zmin = new ZPlane(-32.13); ZPlane* zmax = new ZPlane(32.13);
    zmin->setBoundaryType(REFLECTIVE); zmax->setBoundaryType(VACUUM);
    root_cell->addSurface(+1, zmin); root_cell->addSurface(-1, zmax);
  }
  root_cell->setFill(root_lat);
  Universe* root = new Universe();
  root->addCell(root_cell);

  md.geometry = new Geometry();
  md.geometry->setRootUniverse(root);
  return md;
}

std::vector<double> linspace(double a, double b, int n) {
  std::vector<double> v(n);
  double step = (b - a) / (n - 1);
  for (int i = 0; i < n; i++) v[i] = a + i * step;   /* numpy: start + i*step */
  v[n - 1] = b;
  return v;
}

}  // namespace


void set_axial_layers(int n) { g_axial_layers = n < 1 ? 1 : n; }
void set_boundary_masks(int vacuum_mask, int periodic_mask) { g_vacuum_mask = vacuum_mask; g_periodic_mask = periodic_mask; }

Model build_model(const std::string& name, int dims) {
  if (name == "pin-cell") return pin_cell(dims);
  if (name == "simple-lattice") return simple_lattice(dims);
  if (name == "hom-inf") return hom_inf(dims);
  if (name == "gradient-1d") return hom_inf(dims, 1 | 2);     /* tests/test_1d_gradient: VACUUM in x */
  if (name == "gradient-2d") return hom_inf(dims, 1 | 8);     /* tests/test_2d_gradient: VACUUM on xmin, ymax */
  if (name == "water-box") return water_box(dims);
  if (name == "axial-extended") return axial_extended(dims);
  if (name == "pwr-assembly") return pwr_assembly(dims, g_vacuum_mask, g_periodic_mask);
  if (name == "c5g7-2d") return c5g7_2d(dims);
  log_printf(ERROR, "unknown model %s", name.c_str());
  return Model();
}


void set_70_group_xs(Model& model) {
  const int G = 70;
  std::vector<double> v;
  Material* uo2 = model.materials["UO2"];
  uo2->setNumEnergyGroups(G);
  v = linspace(0, 1, G); for (double& e : v) e *= 7; uo2->setNuSigmaF(v.data(), G);
  v = linspace(0, 1, G * G); for (double& e : v) e /= 1000; uo2->setSigmaS(v.data(), G * G);
  v.assign(G, 1 / 70.); uo2->setChi(v.data(), G);
  v = linspace(2, 3, G); uo2->setSigmaT(v.data(), G);

  Material* water = model.materials["Water"];
  water->setNumEnergyGroups(G);
  v.assign(G, 0.); water->setNuSigmaF(v.data(), G);
  v = linspace(1, 2, G * G); for (double& e : v) e /= 1000; water->setSigmaS(v.data(), G * G);
  v.assign(G, 0.); water->setChi(v.data(), G);
  v = linspace(3, 4, G); water->setSigmaT(v.data(), G);
}
