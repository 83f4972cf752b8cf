// fused solver, Lattice<3, 27>, double: one translation unit per instantiation (parallel build)
#include "solver_fused.cuh"

namespace lbm_impl {
SolverBase* make_fused_d3q27_f64() { return new Solver<lbm::Lattice<3, 27>, double>(); }
} // namespace lbm_impl
