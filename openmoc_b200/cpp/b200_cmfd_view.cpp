#include "b200_cmfd_view.h"

#include <map>
#include <utility>

#include "Cmfd.h"

/* Access to private data members without touching the reference: an explicit instantiation may
 * name a private member ([temp.spec]/6), and the friend function it defines hands the pointer out. */
namespace {
template <typename Tag, typename Tag::type Member>
struct Rob {
  friend typename Tag::type get(Tag) { return Member; }
};
#define B200_MEMBER(Tag, Type, Name)                     \
  struct Tag { typedef Type Cmfd::*type; friend type get(Tag); }; \
  template struct Rob<Tag, &Cmfd::Name>;

typedef std::map<long, std::vector<std::pair<int, double> > > StencilMap;
B200_MEMBER(TSor, double, _SOR_factor)
B200_MEMBER(TRelax, double, _relaxation_factor)
B200_MEMBER(TKeff, double, _k_eff)
B200_MEMBER(TThresh, double, _source_convergence_threshold)
B200_MEMBER(TFluxLim, bool, _flux_limiting)
B200_MEMBER(TLinear, bool, _linear_source)
B200_MEMBER(TBalance, bool, _check_neutron_balance)
B200_MEMBER(TAxial, int, _use_axial_interpolation)
B200_MEMBER(TUnbounded, int, _num_unbounded_iterations)
B200_MEMBER(TKNearest, int, _k_nearest)
B200_MEMBER(TGroupIdx, int*, _group_indices)
B200_MEMBER(TWx, std::vector<double>, _cell_widths_x)
B200_MEMBER(TWy, std::vector<double>, _cell_widths_y)
B200_MEMBER(TWz, std::vector<double>, _cell_widths_z)
B200_MEMBER(TStencils, StencilMap, _k_nearest_stencils)
B200_MEMBER(TInterp, std::vector<double*>, _axial_interpolants)
B200_MEMBER(TConv, ConvergenceData*, _convergence_data)
#undef B200_MEMBER

/* Cmfd::getCellByStencil (src/Cmfd.cpp:3102-3160) on an undecomposed mesh */
int cell_by_stencil(const B200CmfdView& v, int cell, int stencil) {
  const int nx = v.num_x, ny = v.num_y;
  const int x = (cell % (nx * ny)) % nx, y = (cell % (nx * ny)) / nx;
  switch (stencil) {
    case 0: return (x != 0 && y != 0) ? cell - nx - 1 : -1;
    case 1: return y != 0 ? cell - nx : (v.boundaries[SURFACE_Y_MIN] == PERIODIC ? cell + nx * (ny - 1) : -1);
    case 2: return (x != nx - 1 && y != 0) ? cell - nx + 1 : -1;
    case 3: return x != 0 ? cell - 1 : (v.boundaries[SURFACE_X_MIN] == PERIODIC ? cell + (nx - 1) : -1);
    case 4: return cell;
    case 5: return x != nx - 1 ? cell + 1 : (v.boundaries[SURFACE_X_MAX] == PERIODIC ? cell - (nx - 1) : -1);
    case 6: return (x != 0 && y != ny - 1) ? cell + nx - 1 : -1;
    case 7: return y != ny - 1 ? cell + nx : (v.boundaries[SURFACE_Y_MAX] == PERIODIC ? cell - nx * (ny - 1) : -1);
    case 8: return (x != nx - 1 && y != ny - 1) ? cell + nx + 1 : -1;
  }
  return -1;
}
}  // namespace

double b200_cmfd_source_threshold(Cmfd* cmfd) { return cmfd->*get(TThresh()); }
double b200_cmfd_keff(Cmfd* cmfd) { return cmfd->*get(TKeff()); }
ConvergenceData* b200_cmfd_convergence_data(Cmfd* cmfd) { return cmfd->*get(TConv()); }

void b200_read_cmfd(Cmfd* cmfd, long num_fsrs, B200CmfdView* out) {
  B200CmfdView& v = *out;
  v.num_x = cmfd->getNumX();
  v.num_y = cmfd->getNumY();
  v.num_z = cmfd->getNumZ();
  v.num_cmfd_groups = cmfd->getNumCmfdGroups();
  for (int s = 0; s < NUM_FACES; s++) v.boundaries[s] = cmfd->getBoundary(s);
  v.linear_source = cmfd->*get(TLinear());
  v.flux_limiting = cmfd->*get(TFluxLim());
  v.centroid_update = cmfd->isCentroidUpdateOn();
  v.check_neutron_balance = cmfd->*get(TBalance());
  v.balance_sigma_t = cmfd->isSigmaTRebalanceOn();
  v.use_axial_interpolation = cmfd->*get(TAxial());
  v.num_unbounded_iterations = cmfd->*get(TUnbounded());
  v.k_nearest = cmfd->*get(TKNearest());
  v.sor_factor = cmfd->*get(TSor());
  v.relaxation_factor = cmfd->*get(TRelax());
  v.k_eff = cmfd->*get(TKeff());
  v.widths_x = cmfd->*get(TWx());
  v.widths_y = cmfd->*get(TWy());
  v.widths_z = cmfd->*get(TWz());
  const int* gi = cmfd->*get(TGroupIdx());
  v.group_indices.assign(gi, gi + v.num_cmfd_groups + 1);

  std::vector<std::vector<long> >* cells = cmfd->getCellFSRs();
  v.cell_fsr_offset.assign(cells->size() + 1, 0);
  v.cell_fsrs.clear();
  for (size_t i = 0; i < cells->size(); i++) {
    for (size_t j = 0; j < (*cells)[i].size(); j++) v.cell_fsrs.push_back((int32_t)(*cells)[i][j]);
    v.cell_fsr_offset[i + 1] = (int64_t)v.cell_fsrs.size();
  }

  /* stencils as Cmfd::getUpdateRatio walks them (src/Cmfd.cpp:3178-3210) */
  v.st_offset.clear(); v.st_cell.clear(); v.st_weight.clear(); v.st_own.clear(); v.st_size.clear();
  if (v.centroid_update) {
    StencilMap& st = cmfd->*get(TStencils());
    /* FSR -> cell from the lists above (Cmfd::convertFSRIdToCmfdCell searches them linearly) */
    std::vector<int32_t> fsr_cell(num_fsrs, -1);
    for (size_t i = 0; i + 1 < v.cell_fsr_offset.size(); i++)
      for (int64_t j = v.cell_fsr_offset[i]; j < v.cell_fsr_offset[i + 1]; j++)
        if (v.cell_fsrs[j] >= 0 && v.cell_fsrs[j] < num_fsrs) fsr_cell[v.cell_fsrs[j]] = (int32_t)i;
    v.st_offset.assign(num_fsrs + 1, 0);
    v.st_own.assign(num_fsrs, 1.0);
    v.st_size.assign(num_fsrs, 1);
    for (long r = 0; r < num_fsrs; r++) {
      StencilMap::iterator it = st.find(r);
      if (it != st.end() && !it->second.empty()) {
        const int cell = fsr_cell[r];
        const std::vector<std::pair<int, double> >& entries = it->second;
        v.st_size[r] = (int32_t)entries.size();
        v.st_own[r] = entries[0].second;
        for (size_t j = 0; j < entries.size(); j++) {
          if (entries[j].first == 4) continue;
          v.st_cell.push_back(cell_by_stencil(v, cell, entries[j].first));
          v.st_weight.push_back(entries[j].second);
        }
      }
      v.st_offset[r + 1] = (int64_t)v.st_cell.size();
    }
  }

  v.axial_interpolants.clear();
  if (v.use_axial_interpolation && v.num_z >= 3) {
    std::vector<double*>& ai = cmfd->*get(TInterp());
    v.axial_interpolants.assign((size_t)num_fsrs * 3, 0.);
    for (long r = 0; r < num_fsrs && r < (long)ai.size(); r++)
      for (int c = 0; c < 3; c++) v.axial_interpolants[r * 3 + c] = ai[r][c];
  }
}
