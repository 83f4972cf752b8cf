// fused solver, Lattice<2, 9>, double: one translation unit per instantiation (parallel build)
#include "solver_fused.cuh"

namespace lbm_impl {
SolverBase* make_fused_d2q9_f64() { return new Solver<lbm::Lattice<2, 9>, double>(); }
} // namespace lbm_impl
