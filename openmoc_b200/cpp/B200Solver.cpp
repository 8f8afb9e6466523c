/**
 * @file B200Solver.cpp
 * @brief Explicit instantiation of the flat-source plug-in (see B200SolverT.h).
 */
#include "B200Solver.h"

template class B200SolverT<Solver>;
