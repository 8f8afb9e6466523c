/**
 * @file models.h
 * @brief C++ restatements of the reference's test/sample input decks, built
 *        with the reference's own Geometry/Cell/Lattice classes.
 *        TEST INFRASTRUCTURE (oracle side) - never linked into the product.
 *
 *  pin-cell        tests/input_set.py:95-137   (PinCellInput)
 *  simple-lattice  tests/input_set.py:310-417  (SimpleLatticeInput, 2D or 3D)
 *  c5g7-2d         sample-input/benchmarks/c5g7/{surfaces,cells,universes,
 *                  lattices,c5g7-2d}.py
 *  hom-inf         tests/input_set.py:32-92    (HomInfMedInput)
 *  gradient-1d/2d  tests/test_1d_gradient, tests/test_2d_gradient: hom-inf with VACUUM sides
 *  water-box       tests/test_compute_flux/test_compute_flux.py:18-79 and test_compute_source:
 *                  HomInfMedInput's 10x10 lattice with VACUUM sides, filled with C5G7 water, one
 *                  lattice cell (the "source" cell) holding the fixed source
 */
#ifndef ORACLE_MODELS_H_
#define ORACLE_MODELS_H_

#include <map>
#include <string>

class Geometry;
class Material;

class Cell;
struct Model {
  Geometry* geometry;
  std::map<std::string, Material*> materials;
  Cell* source_cell = nullptr;      /* water-box: the cell Solver::setFixedSourceByCell is given */
};

/** dims = 2 or 3 (only simple-lattice and pin-cell honour 3). */
Model build_model(const std::string& name, int dims);
/** c5g7-2d with dims = 3: number of equal axial layers of the root lattice (3 x 3 x N, the structure of
 *  profile/models/c5g7/c5g7-3d-cmfd.cpp:520-535 with identical layers); default 1. */
void set_axial_layers(int n);
/** pwr-assembly: sides switched to VACUUM / PERIODIC (bit 0 xmin, 1 xmax, 2 ymin, 3 ymax), what
 *  tests/test_cmfd_vacuum_boundary and tests/test_cmfd_periodic_boundaries do to PwrAssemblyInput. */
void set_boundary_masks(int vacuum_mask, int periodic_mask);

/** Replace the UO2/Water data by the synthetic 70-group set of
 *  tests/test_forward_3D_lattice_70g/test_forward_3D_lattice_70g.py:43-61. */
void set_70_group_xs(Model& model);

#endif
