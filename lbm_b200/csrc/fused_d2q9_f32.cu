// fused solver, Lattice<2, 9>, float: one translation unit per instantiation (parallel build)
#include "solver_fused.cuh"

namespace lbm_impl {
SolverBase* make_fused_d2q9_f32() { return new Solver<lbm::Lattice<2, 9>, float>(); }
} // namespace lbm_impl
