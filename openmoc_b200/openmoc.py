"""`import openmoc_b200.openmoc as openmoc`: what an OpenMOC input script imports, with `B200Solver` /
`B200LSSolver` next to the reference classes.

The compiled half (`_openmoc_b200`, openmoc_b200/cpp/pymodule.cpp, pybind11) holds the UNMODIFIED reference C++
core's Material, surfaces, Cell, Universe, Lattice, Geometry, Cmfd, quadratures, TrackGenerator(3D), CPUSolver,
CPULSSolver and the B200 plug-in classes - the role of the reference's SWIG modules (openmoc/swig/openmoc.i,
openmoc/cuda/openmoc_cuda.i:52-56), which need `swig`.  This file adds the pure-Python helpers the sample inputs
use: `openmoc.log`, `openmoc.options.Options`, and `openmoc.materialize` reading the C5G7 cross sections from
openmoc_b200/data/c5g7_xs.json (the reference's loader, openmoc/materialize.py, needs h5py and the .h5 file).

A script of sample-input/ runs after two edits: the import line and the materials line.  See
examples/pin_cell_b200.py.
"""
from __future__ import annotations

import argparse
import json
import os
import types

try:
    from ._openmoc_b200 import *          # noqa: F401,F403  classes and enum values (REFLECTIVE, FISSION_SOURCE, ...)
    from . import _openmoc_b200 as _core
except ImportError as e:                   # pragma: no cover - build instructions instead of a bare ImportError
    raise ImportError(
        "openmoc_b200._openmoc_b200 is not built: it links the reference OpenMOC core, build it with "
        "`make -C oracle ref` where /root/reference is present (" + str(e) + ")") from e

_HERE = os.path.dirname(os.path.abspath(__file__))

# ---------------------------------------------------------------- openmoc.log (openmoc/log.py)
log = types.ModuleType("openmoc.log")
log.set_log_level = _core.set_log_level
log.py_printf = lambda level, msg, *args: _core.log_printf(level, msg % args if args else msg)

# ---------------------------------------------------------------- openmoc.options (openmoc/options.py:30-135)
options = types.ModuleType("openmoc.options")


class Options:
    """Command-line options of the sample inputs, same flags and defaults as openmoc/options.py."""

    def __init__(self, argv=None):
        p = argparse.ArgumentParser(add_help=True)
        p.add_argument("-a", "--num-azim", type=int, default=4)
        p.add_argument("-s", "--azim-spacing", type=float, default=0.1)
        p.add_argument("-p", "--num-polar", type=int, default=6)
        p.add_argument("-l", "--polar-spacing", "--z-spacing", dest="polar_spacing", type=float, default=1.5)
        p.add_argument("-i", "--max-iters", type=int, default=1000)
        p.add_argument("-c", "--tolerance", type=float, default=1e-5)
        p.add_argument("-t", "--num-omp-threads", type=int, default=os.cpu_count() or 1)
        ns, _ = p.parse_known_args(argv)
        self.num_azim, self.azim_spacing = ns.num_azim, ns.azim_spacing
        self.num_polar, self.polar_spacing = ns.num_polar, ns.polar_spacing
        self.max_iters, self.tolerance, self.num_omp_threads = ns.max_iters, ns.tolerance, ns.num_omp_threads


options.Options = Options

# ---------------------------------------------------------------- openmoc.materialize
materialize = types.ModuleType("openmoc.materialize")


def load_c5g7():
    """name -> Material with the C5G7 7-group cross sections (what load_from_hdf5('c5g7-mgxs.h5') returns)."""
    doc = json.load(open(os.path.join(_HERE, "data", "c5g7_xs.json")))
    G = doc["num_groups"]
    out = {}
    for i, (name, d) in enumerate(doc["materials"].items()):
        m = _core.Material(id=10000 + i, name=name)
        m.setNumEnergyGroups(G)
        m.setSigmaT(d["sigma_t"])
        m.setSigmaS(d["sigma_s"])
        m.setSigmaF(d["sigma_f"])
        m.setNuSigmaF(d["nu_sigma_f"])
        m.setChi(d["chi"])
        out[name] = m
    return out


materialize.load_c5g7 = load_c5g7
