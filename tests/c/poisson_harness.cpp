// poisson_harness.cpp -- TEST INFRASTRUCTURE: runs the kernel BODIES and the host set-up of lbm_b200/csrc/poisson.cuh on the CPU.
//
// The Poisson pipeline was written when no GPU time was left.  To check its logic anyway, this file compiles
// poisson.cuh with plain g++: __global__ / __device__ become nothing, blockIdx / threadIdx become variables that a loop walks over every
// thread of a launch, and the IEEE intrinsics become the plain operators (compiled with -ffp-contract=off, so a + b is __dadd_rn).
// What runs is the same source text the GPU executes -- k_init, k_cell, k_stream, k_neumann_value, k_ext_potential, k_dirichlet and
// poisson::prepare() -- in the launch order of PoissonSolver::step (solver.cu).  tests/test_poisson_harness.py compares the result
// with the reference's dumps bit for bit.  Never part of the product (only tests/ builds it).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define LBM_POISSON_HOST_HARNESS
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
struct SimDim { unsigned x = 0; };
static SimDim blockIdx, blockDim, threadIdx, gridDim;
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }

#include "../../lbm_b200/csrc/poisson.cuh"

namespace {
using namespace lbm;

template <class F>
void launch(int64_t n, F&& body) { // <<<blocks(n), 128>>>
  blockDim.x = 128;
  gridDim.x  = static_cast<unsigned>((n + 127) / 128);
  for(unsigned b = 0; b < gridDim.x; ++b)
    for(unsigned t = 0; t < 128; ++t) {
      blockIdx.x  = b;
      threadIdx.x = t;
      body();
    }
}

struct Harness {
  PlanInput          in;
  double             omega = 1;
  poisson::HostSetup setup;
  std::vector<double> f, fold, feq, vars, varsold;
  std::vector<poisson::Bc> bcs;
  std::vector<int64_t> nb64;
  std::string error;

  poisson::State state() {
    poisson::State s{};
    s.f = f.data(); s.fold = fold.data(); s.feq = feq.data(); s.vars = vars.data(); s.varsold = varsold.data();
    s.pull = setup.pull.data(); s.nghbr = nb64.data(); s.n = in.n;
    s.omega = setup.omega; s.om1 = setup.om1; s.dt_diff = setup.dt_diff; s.rate2 = setup.rate2;
    return s;
  }
};
} // namespace

extern "C" {

void* ph_create(int ndim, int ndist, int64_t n, const int64_t* nghbr, int stride, double omega, double dt, double rate) {
  auto* h = new Harness();
  if(!lattice_rt(ndim, ndist, &h->in.L)) { delete h; return nullptr; }
  h->in.n = n;
  h->omega = omega;
  h->in.poisson = true;
  h->in.poisson_dt = dt;
  h->in.poisson_rate = rate;
  const int QM = ndist - 1;
  // what lbm_b200_set_topology does (solver.cu): the first Q-1 columns, and all 8 for D2Q5
  h->in.nghbr.resize(static_cast<size_t>(n) * QM);
  for(int64_t c = 0; c < n; ++c)
    for(int j = 0; j < QM; ++j) h->in.nghbr[c * QM + j] = static_cast<int32_t>(nghbr[c * stride + j]);
  if(ndim == 2 && ndist == 5 && stride >= 8) {
    h->in.nghbr_wide.resize(static_cast<size_t>(n) * 8);
    for(int64_t c = 0; c < n; ++c)
      for(int j = 0; j < 8; ++j) h->in.nghbr_wide[c * 8 + j] = static_cast<int32_t>(nghbr[c * stride + j]);
  }
  return h;
}
void ph_destroy(void* p) { delete static_cast<Harness*>(p); }

void ph_add_bc(void* p, int neumann, const int64_t* cells, int64_t n, const double* values, double grad) {
  auto* h = static_cast<Harness*>(p);
  BcInput bc;
  bc.kind = neumann ? BC_POISSON_NEUMANN : BC_POISSON_DIRICHLET;
  bc.cells.assign(cells, cells + n);
  bc.values.assign(values, values + n);
  bc.grad = grad;
  h->in.bcs.push_back(bc);
}

// PoissonSolver::init: returns 0 or the error code of prepare()
int ph_init(void* p) {
  auto* h = static_cast<Harness*>(p);
  if(!poisson::prepare(h->in, h->omega, h->setup)) { h->error = h->setup.error; return h->setup.code; }
  const int64_t N = h->in.n;
  const int     Q = h->in.L.Q;
  h->f.assign(static_cast<size_t>(N) * Q, 0.0);
  h->fold = h->f;
  h->feq  = h->f;
  h->vars = h->setup.vars0;
  h->varsold.assign(static_cast<size_t>(N), 0.0);
  h->nb64.assign(h->in.nghbr.begin(), h->in.nghbr.end());
  h->bcs.clear();
  for(poisson::HostBc& hb : h->setup.bcs) {
    poisson::Bc b{};
    b.neumann = hb.neumann;
    b.n = static_cast<int64_t>(hb.cells.size());
    b.cells = hb.cells.data(); b.ext = hb.ext.data(); b.ext2 = hb.ext2.data(); b.values = hb.values.data();
    b.grad = hb.grad;
    h->bcs.push_back(b);
  }
  const poisson::State s = h->state();
  const poisson::Lat   L = h->setup.lat;
  launch(N, [&] { poisson::k_init(s, L); });
  return 0;
}

// PoissonSolver::step: the same launch sequence
void ph_step(void* p, int64_t nsteps) {
  auto* h = static_cast<Harness*>(p);
  const poisson::State s = h->state();
  const poisson::Lat   L = h->setup.lat;
  for(int64_t it = 0; it < nsteps; ++it) {
    std::memcpy(h->varsold.data(), h->vars.data(), h->vars.size() * sizeof(double)); // currToOldVars
    launch(s.n, [&] { poisson::k_cell(s, L); });
    launch(s.n, [&] { poisson::k_stream(s, L); });
    for(const poisson::Bc& b : h->bcs) {
      if(b.n == 0) continue;
      if(b.neumann) launch(b.n, [&] { poisson::k_neumann_value(s, L, b); });
      launch(b.n, [&] { poisson::k_ext_potential(s, L, b); });
      launch(b.n, [&] { poisson::k_dirichlet(s, L, b); });
    }
  }
}
void ph_potential(void* p, double* out) {
  auto* h = static_cast<Harness*>(p);
  const poisson::State s = h->state();
  const poisson::Lat   L = h->setup.lat;
  launch(s.n, [&] { poisson::k_potential(s, L, out); });
}
double* ph_array(void* p, int which) {
  auto* h = static_cast<Harness*>(p);
  switch(which) {
    case 0: return h->f.data();
    case 1: return h->fold.data();
    case 2: return h->vars.data();
    default: return h->varsold.data();
  }
}
const char* ph_error(void* p) { return static_cast<Harness*>(p)->error.c_str(); }

} // extern "C"
