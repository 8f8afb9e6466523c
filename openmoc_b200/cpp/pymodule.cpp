/**
 * @file pymodule.cpp
 * @brief pybind11 module `_openmoc_b200`: B200Solver / B200LSSolver next to the reference classes an OpenMOC
 *        input script builds them from (Material, surfaces, Cell, Universe, Lattice, Geometry, Cmfd, the
 *        quadratures, TrackGenerator / TrackGenerator3D, CPUSolver / CPULSSolver), so that
 *        `B200Solver(track_generator)` is what it is in C++ - the role of the reference's SWIG modules
 *        (openmoc/swig/openmoc.i:167-169, openmoc/cuda/openmoc_cuda.i:52-56; `swig` is not in this image).
 *        Links the unmodified reference core (libopenmoc_ref.so, built from the reference sources by the same recipe
 *        that builds the checker: INTEGRATION.md) and libb200moc.so; it needs the reference headers.  Python-side conveniences (log, options,
 *        materials from the C5G7 cross-section file) live in openmoc_b200/openmoc.py.
 *
 * Ownership follows openmoc/swig/thisown.i: OpenMOC objects keep raw pointers to each other, so nothing created
 * from Python is ever deleted by Python (py::nodelete holders).
 */
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "CPULSSolver.h"
#include "CPUSolver.h"
#include "Cmfd.h"
#include "Geometry.h"
#include "Material.h"
#include "Quadrature.h"
#include "TrackGenerator3D.h"
#include "log.h"

#include "B200LSSolver.h"
#include "B200Solver.h"

namespace py = pybind11;
template <class T> using Keep = std::unique_ptr<T, py::nodelete>;
typedef py::array_t<double, py::array::c_style | py::array::forcecast> darray;

namespace {

/* the Solver interface every solver class exposes (src/Solver.h:433-585) */
template <class S, class PyClass>
void bind_solver(PyClass& c) {
  c.def("setConvergenceThreshold", [](S& s, double t) { s.setConvergenceThreshold(t); })
      .def("setNumThreads", [](S& s, int n) { s.setNumThreads(n); })
      .def("computeEigenvalue", [](S& s, int max_iters, residualType res) { s.computeEigenvalue(max_iters, res); },
           py::arg("max_iters") = 1000, py::arg("res_type") = FISSION_SOURCE, py::call_guard<py::gil_scoped_release>())
      .def("computeFlux", [](S& s, int max_iters, bool only_fixed) { s.computeFlux(max_iters, only_fixed); },
           py::arg("max_iters") = 1000, py::arg("only_fixed_source") = true)
      .def("computeSource", [](S& s, int max_iters, double k, residualType res) { s.computeSource(max_iters, k, res); },
           py::arg("max_iters") = 1000, py::arg("k_eff") = 1.0, py::arg("res_type") = TOTAL_SOURCE)
      .def("getKeff", [](S& s) { return s.getKeff(); })
      .def("getNumIterations", [](S& s) { return s.getNumIterations(); })
      .def("getTotalTime", [](S& s) { return s.getTotalTime(); })
      .def("getFlux", [](S& s, long fsr, int group) { return s.getFlux(fsr, group); })
      .def("getFSRSource", [](S& s, long fsr, int group) { return s.getFSRSource(fsr, group); })
      .def("getFluxes", [](S& s, long n) {
        py::array_t<double> out(n);
        s.getFluxes(out.mutable_data(), (int)n);
        return out;
      })
      .def("setFluxes", [](S& s, darray in) { s.setFluxes(in.mutable_data(), (int)in.size()); })
      .def("setFixedSourceByFSR", [](S& s, long fsr, int group, double q) { s.setFixedSourceByFSR(fsr, group, q); })
      .def("setFixedSourceByCell", [](S& s, Cell* cell, int group, double q) { s.setFixedSourceByCell(cell, group, q); })
      .def("setFixedSourceByMaterial", [](S& s, Material* m, int group, double q) { s.setFixedSourceByMaterial(m, group, q); })
      .def("resetFixedSources", [](S& s) { s.resetFixedSources(); })
      .def("computeFSRFissionRates", [](S& s, long n, bool nu) {
        py::array_t<double> out(n);
        s.computeFSRFissionRates(out.mutable_data(), n, nu);
        return out;
      }, py::arg("num_FSRs"), py::arg("nu") = false)
      .def("stabilizeTransport", [](S& s, double f, stabilizationType t) { s.stabilizeTransport(f, t); },
           py::arg("stabilization_factor"), py::arg("stabilization_type") = DIAGONAL)
      .def("allowNegativeFluxes", [](S& s, bool on) { s.allowNegativeFluxes(on); })
      .def("setKeffFromNeutronBalance", [](S& s) { s.setKeffFromNeutronBalance(); })
      .def("setVerboseIterationReport", [](S& s) { s.setVerboseIterationReport(); })
      .def("setMaxOpticalLength", [](S& s, double tau) { s.setMaxOpticalLength(tau); })
      .def("useExponentialIntrinsic", [](S& s) { s.useExponentialIntrinsic(); })
      .def("initializeSolver", [](S& s, solverMode m) { s.initializeSolver(m); })
      .def("fissionTransportSweep", [](S& s) { s.fissionTransportSweep(); })
      .def("scatterTransportSweep", [](S& s) { s.scatterTransportSweep(); })
      .def("printTimerReport", [](S& s) { s.printTimerReport(); })
      .def("getGeometry", [](S& s) { return s.getGeometry(); }, py::return_value_policy::reference)
      .def("getTrackGenerator", [](S& s) { return s.getTrackGenerator(); }, py::return_value_policy::reference);
}

template <class S, class PyClass>
void bind_b200(PyClass& c) {
  c.def("setNumDevices", [](S& s, int n) { s.setNumDevices(n); })
      .def("setDevices", [](S& s, std::vector<int> d) { s.setDevices(d); })
      .def("setCmfdOnDevice", [](S& s, bool on) { s.setCmfdOnDevice(on); })
      .def("isCmfdOnDevice", [](S& s) { return s.isCmfdOnDevice(); })
      .def("computeEigenvalueFused", [](S& s, int max_iters, residualType res) { s.computeEigenvalueFused(max_iters, res); },
           py::arg("max_iters") = 1000, py::arg("res_type") = FISSION_SOURCE)
      .def("getSweepStats", [](S& s) { double ms = 0.; long n = 0; s.getSweepStats(&ms, &n); return py::make_tuple(ms, n); })
      .def("getCmfdStats", [](S& s) { double ms = 0.; long n = 0; s.getCmfdStats(&ms, &n); return py::make_tuple(ms, n); });
}

void set_xs(Material& m, darray xs, void (Material::*setter)(double*, int)) { (m.*setter)(xs.mutable_data(), (int)xs.size()); }

}  // namespace

PYBIND11_MODULE(_openmoc_b200, m) {
  m.doc() = "B200Solver / B200LSSolver and the OpenMOC classes an input script needs (reference C++ core, unmodified)";

  /* std::logic_error from log_printf(ERROR) -> RuntimeError, like openmoc/swig/openmoc.i:106-112 */
  py::register_exception_translator([](std::exception_ptr p) {
    try { if (p) std::rethrow_exception(p); }
    catch (const std::logic_error& e) { PyErr_SetString(PyExc_RuntimeError, e.what()); }
  });

  py::enum_<boundaryType>(m, "boundaryType").value("VACUUM", VACUUM).value("REFLECTIVE", REFLECTIVE)
      .value("PERIODIC", PERIODIC).value("BOUNDARY_NONE", BOUNDARY_NONE).export_values();
  py::enum_<residualType>(m, "residualType").value("SCALAR_FLUX", SCALAR_FLUX).value("FISSION_SOURCE", FISSION_SOURCE)
      .value("TOTAL_SOURCE", TOTAL_SOURCE).export_values();
  py::enum_<stabilizationType>(m, "stabilizationType").value("DIAGONAL", DIAGONAL).value("YAMAMOTO", YAMAMOTO)
      .value("GLOBAL", GLOBAL).export_values();
  py::enum_<solverMode>(m, "solverMode").value("FORWARD", FORWARD).value("ADJOINT", ADJOINT).export_values();
  py::enum_<segmentationType>(m, "segmentationType").value("EXPLICIT_2D", EXPLICIT_2D).value("EXPLICIT_3D", EXPLICIT_3D)
      .value("OTF_TRACKS", OTF_TRACKS).value("OTF_STACKS", OTF_STACKS).export_values();

  m.def("set_log_level", [](const std::string& level) { set_log_level(level.c_str()); });
  m.def("log_printf", [](const std::string& level, const std::string& msg) {
    logLevel l = NORMAL;
    if (level == "DEBUG") l = DEBUG; else if (level == "INFO") l = INFO; else if (level == "TITLE") l = TITLE;
    else if (level == "HEADER") l = HEADER; else if (level == "SEPARATOR") l = SEPARATOR; else if (level == "RESULT") l = RESULT;
    else if (level == "WARNING") l = WARNING; else if (level == "ERROR") l = ERROR;
    log_printf(l, "%s", msg.c_str());
  });

  py::class_<Material, Keep<Material>>(m, "Material")
      .def(py::init<int, const char*>(), py::arg("id") = 0, py::arg("name") = "")
      .def("setNumEnergyGroups", &Material::setNumEnergyGroups)
      .def("getNumEnergyGroups", &Material::getNumEnergyGroups)
      .def("getName", &Material::getName)
      .def("getId", &Material::getId)
      .def("isFissionable", &Material::isFissionable)
      .def("setSigmaT", [](Material& s, darray xs) { set_xs(s, xs, &Material::setSigmaT); })
      .def("setSigmaS", [](Material& s, darray xs) { set_xs(s, xs, &Material::setSigmaS); })
      .def("setSigmaF", [](Material& s, darray xs) { set_xs(s, xs, &Material::setSigmaF); })
      .def("setNuSigmaF", [](Material& s, darray xs) { set_xs(s, xs, &Material::setNuSigmaF); })
      .def("setChi", [](Material& s, darray xs) { set_xs(s, xs, &Material::setChi); })
      .def("setSigmaTByGroup", &Material::setSigmaTByGroup)
      .def("setSigmaSByGroup", &Material::setSigmaSByGroup)
      .def("setNuSigmaFByGroup", &Material::setNuSigmaFByGroup)
      .def("setSigmaFByGroup", &Material::setSigmaFByGroup)
      .def("setChiByGroup", &Material::setChiByGroup);

  py::class_<Surface, Keep<Surface>>(m, "Surface")
      .def("setBoundaryType", &Surface::setBoundaryType)
      .def("getBoundaryType", &Surface::getBoundaryType)
      .def("getId", &Surface::getId)
      .def("getName", &Surface::getName);
  py::class_<XPlane, Surface, Keep<XPlane>>(m, "XPlane")
      .def(py::init<double, int, const char*>(), py::arg("x"), py::arg("id") = 0, py::arg("name") = "");
  py::class_<YPlane, Surface, Keep<YPlane>>(m, "YPlane")
      .def(py::init<double, int, const char*>(), py::arg("y"), py::arg("id") = 0, py::arg("name") = "");
  py::class_<ZPlane, Surface, Keep<ZPlane>>(m, "ZPlane")
      .def(py::init<double, int, const char*>(), py::arg("z"), py::arg("id") = 0, py::arg("name") = "");
  py::class_<ZCylinder, Surface, Keep<ZCylinder>>(m, "ZCylinder")
      .def(py::init<double, double, double, int, const char*>(), py::arg("x"), py::arg("y"), py::arg("radius"),
           py::arg("id") = 0, py::arg("name") = "");

  py::class_<Universe, Keep<Universe>>(m, "Universe")
      .def(py::init<int, const char*>(), py::arg("id") = -1, py::arg("name") = "")
      .def("addCell", &Universe::addCell)
      .def("getId", &Universe::getId)
      .def("getName", &Universe::getName);
  py::class_<Cell, Keep<Cell>>(m, "Cell")
      .def(py::init<int, const char*>(), py::arg("id") = 0, py::arg("name") = "")
      .def("setFill", [](Cell& c, Material* fill) { c.setFill(fill); })
      .def("setFill", [](Cell& c, Universe* fill) { c.setFill(fill); })
      .def("addSurface", &Cell::addSurface, py::arg("halfspace"), py::arg("surface"))
      .def("setNumRings", [](Cell& c, int n) { c.setNumRings(n); })
      .def("setNumSectors", &Cell::setNumSectors)
      .def("getId", &Cell::getId)
      .def("getName", &Cell::getName);
  py::class_<Lattice, Universe, Keep<Lattice>>(m, "Lattice")
      .def(py::init<int, const char*>(), py::arg("id") = -1, py::arg("name") = "")
      .def("setWidth", [](Lattice& l, double wx, double wy, double wz) { l.setWidth(wx, wy, wz); },
           py::arg("width_x"), py::arg("width_y"), py::arg("width_z") = std::numeric_limits<double>::infinity())
      .def("setOffset", [](Lattice& l, double x, double y, double z) { l.setOffset(x, y, z); },
           py::arg("x"), py::arg("y"), py::arg("z") = 0.0)
      /* nested lists [z][y][x] (or [y][x]) of universes, as openmoc/swig/typemaps.i accepts them */
      .def("setUniverses", [](Lattice& l, py::list rows) {
        std::vector<Universe*> flat;
        int nz = 1, ny = 0, nx = 0;
        py::list zs = rows;
        const bool three = py::isinstance<py::list>(rows[0]) && py::isinstance<py::list>(rows[0].cast<py::list>()[0]);
        if (!three) { zs = py::list(); zs.append(rows); }
        nz = (int)zs.size();
        for (auto zrow : zs) {
          py::list ys = zrow.cast<py::list>();
          ny = (int)ys.size();
          for (auto yrow : ys) {
            py::list xs = yrow.cast<py::list>();
            nx = (int)xs.size();
            for (auto u : xs) flat.push_back(u.cast<Universe*>());
          }
        }
        l.setUniverses(nz, ny, nx, flat.data());
      });

  py::class_<Cmfd, Keep<Cmfd>>(m, "Cmfd")
      .def(py::init<>())
      .def("setLatticeStructure", &Cmfd::setLatticeStructure, py::arg("num_x"), py::arg("num_y"), py::arg("num_z") = 1)
      .def("setGroupStructure", &Cmfd::setGroupStructure)
      .def("setSORRelaxationFactor", &Cmfd::setSORRelaxationFactor)
      .def("setCMFDRelaxationFactor", &Cmfd::setCMFDRelaxationFactor)
      .def("setKNearest", &Cmfd::setKNearest)
      .def("setCentroidUpdateOn", &Cmfd::setCentroidUpdateOn)
      .def("setFluxUpdateOn", &Cmfd::setFluxUpdateOn)
      .def("useFluxLimiting", &Cmfd::useFluxLimiting)
      .def("useAxialInterpolation", &Cmfd::useAxialInterpolation)
      .def("setNumUnboundedIterations", &Cmfd::setNumUnboundedIterations)
      .def("setWidths", &Cmfd::setWidths)
      .def("rebalanceSigmaT", &Cmfd::rebalanceSigmaT)
      .def("getNumCells", &Cmfd::getNumCells);

  py::class_<Geometry, Keep<Geometry>>(m, "Geometry")
      .def(py::init<>())
      .def("setRootUniverse", &Geometry::setRootUniverse)
      .def("setCmfd", &Geometry::setCmfd)
      .def("initializeFlatSourceRegions", &Geometry::initializeFlatSourceRegions)
      .def("getNumFSRs", &Geometry::getNumFSRs)
      .def("getNumEnergyGroups", &Geometry::getNumEnergyGroups)
      .def("getNumMaterials", &Geometry::getNumMaterials)
      .def("getNumCells", &Geometry::getNumCells)
      .def("getMinX", &Geometry::getMinX).def("getMaxX", &Geometry::getMaxX)
      .def("getMinY", &Geometry::getMinY).def("getMaxY", &Geometry::getMaxY);

  py::class_<Quadrature, Keep<Quadrature>>(m, "Quadrature")
      .def("setNumAzimAngles", &Quadrature::setNumAzimAngles)
      .def("setNumPolarAngles", &Quadrature::setNumPolarAngles)
      .def("getNumPolarAngles", &Quadrature::getNumPolarAngles)
      .def("getNumAzimAngles", &Quadrature::getNumAzimAngles);
  py::class_<TYPolarQuad, Quadrature, Keep<TYPolarQuad>>(m, "TYPolarQuad").def(py::init<>());
  py::class_<LeonardPolarQuad, Quadrature, Keep<LeonardPolarQuad>>(m, "LeonardPolarQuad").def(py::init<>());
  py::class_<GLPolarQuad, Quadrature, Keep<GLPolarQuad>>(m, "GLPolarQuad").def(py::init<>());
  py::class_<EqualWeightPolarQuad, Quadrature, Keep<EqualWeightPolarQuad>>(m, "EqualWeightPolarQuad").def(py::init<>());
  py::class_<EqualAnglePolarQuad, Quadrature, Keep<EqualAnglePolarQuad>>(m, "EqualAnglePolarQuad").def(py::init<>());

  py::class_<TrackGenerator, Keep<TrackGenerator>>(m, "TrackGenerator")
      .def(py::init<Geometry*, int, double>(), py::arg("geometry"), py::arg("num_azim"), py::arg("azim_spacing"))
      .def("setNumThreads", &TrackGenerator::setNumThreads)
      .def("setQuadrature", &TrackGenerator::setQuadrature)
      .def("setZCoord", &TrackGenerator::setZCoord)
      .def("generateTracks", &TrackGenerator::generateTracks, py::call_guard<py::gil_scoped_release>())
      .def("getNumTracks", &TrackGenerator::getNumTracks)
      .def("getNumSegments", &TrackGenerator::getNumSegments)
      .def("getGeometry", &TrackGenerator::getGeometry, py::return_value_policy::reference)
      .def("getQuadrature", &TrackGenerator::getQuadrature, py::return_value_policy::reference);
  py::class_<TrackGenerator3D, TrackGenerator, Keep<TrackGenerator3D>>(m, "TrackGenerator3D")
      .def(py::init<Geometry*, int, int, double, double>(), py::arg("geometry"), py::arg("num_azim"),
           py::arg("num_polar"), py::arg("azim_spacing"), py::arg("z_spacing"))
      .def("setSegmentFormation", &TrackGenerator3D::setSegmentFormation)
      .def("setSegmentationZones", &TrackGenerator3D::setSegmentationZones)
      .def("getNum3DTracks", &TrackGenerator3D::getNum3DTracks)
      .def("getNum3DSegments", &TrackGenerator3D::getNum3DSegments);

  py::class_<Solver, Keep<Solver>>(m, "Solver");
  {
    py::class_<CPUSolver, Solver, Keep<CPUSolver>> c(m, "CPUSolver");
    c.def(py::init<TrackGenerator*>(), py::arg("track_generator") = (TrackGenerator*)NULL);
    bind_solver<CPUSolver>(c);
  }
  {
    py::class_<CPULSSolver, CPUSolver, Keep<CPULSSolver>> c(m, "CPULSSolver");
    c.def(py::init<TrackGenerator*>(), py::arg("track_generator") = (TrackGenerator*)NULL);
  }
  {
    /* the drop-in: same constructor argument and methods as the reference's solvers (and its GPUSolver) */
    py::class_<B200Solver, Solver, Keep<B200Solver>> c(m, "B200Solver");
    c.def(py::init<TrackGenerator*, int, int>(), py::arg("track_generator") = (TrackGenerator*)NULL,
          py::arg("device") = 0, py::arg("precision") = 0);
    bind_solver<B200Solver>(c);
    bind_b200<B200Solver>(c);
  }
  {
    py::class_<B200LSSolver, Solver, Keep<B200LSSolver>> c(m, "B200LSSolver");
    c.def(py::init<TrackGenerator*, int>(), py::arg("track_generator") = (TrackGenerator*)NULL, py::arg("device") = 0);
    bind_solver<B200LSSolver>(c);
    bind_b200<B200LSSolver>(c);
  }
}
