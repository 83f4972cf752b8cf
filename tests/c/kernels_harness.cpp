// kernels_harness.cpp -- TEST INFRASTRUCTURE: the device code of lbm_b200/csrc/kernels.cuh executed on the CPU.
//
// g++ compiles the unmodified kernels.cuh against tests/c/fake_cuda/cuda_runtime.h.  A driver loop walks the thread indices of the
// kernels that need no barrier: k_gather_all (the gather of the fused kernel: link codes, chunk templates, wall descriptors,
// pressure-face chunks, ghost blocks), the per-cell update (gather_any + update_and_store: what a time step does to a cell), k_halo_pack /
// k_halo_unpack, k_velocity_pack and k_pressure_extrapolate (the velocity halo of the pressure boundary condition, which had no
// hardware run in round 1).  The device tables are built from lbm_b200_debug_plan exactly as Solver::init (solver.cu) builds them.
// tests/test_kernels_harness.py drives partitioned runs with it and compares with the single-domain oracle bit for bit.
#include <cstring>
#include <thread>
#include <vector>

#include "../../lbm_b200/csrc/kernels.cuh"
#include "../../include/lbm_b200.h"

namespace {
using namespace lbm;

template <class F>
void launch(int64_t n, int threads, F&& body) {
  blockDim.x = static_cast<unsigned>(threads);
  gridDim.x  = static_cast<unsigned>((n + threads - 1) / threads);
  for(unsigned b = 0; b < gridDim.x; ++b)
    for(unsigned t = 0; t < static_cast<unsigned>(threads); ++t) {
      blockIdx.x  = b;
      threadIdx.x = t;
      body();
    }
}

struct Ctx {
  int ndim = 0, ndist = 0;
  lbm_b200_plan_view v{};
  std::vector<AddEntryT<double>> wall, add;
  std::vector<CopySrcDev>        copy;
  std::vector<AbbDev<double>>    abb;
  std::vector<double>            uext[2], values, A, B, vars, vrecv, scratch;
  int    dyn = 0, first = 1;
  double omega = 1, omega_minus = 1, rates[27];
  int    coll = COLL_BGK;
  unsigned long long ticket = 0, ticket_next = 0;
  unsigned long long ticket2[2] = {0, 0}, ticket2_next[2] = {0, 0}; // the two launch classes of the overlapped path

  DevParams<double> params() {
    DevParams<double> p{};
    p.A = A.data();
    p.B = B.data();
    p.stride = v.npad;
    p.pr = PermRange{static_cast<int32_t>(v.perm_end), static_cast<int32_t>(v.gb_begin), static_cast<int32_t>(v.gb_end)};
    p.tmpl = v.tmpl;
    p.chunk_nb = v.chunk_nb;
    p.wall_desc = wall.data();
    p.chunk_abb_base = v.chunk_abb_base;
    p.chunk_abb = v.chunk_abb;
    p.n_fast_chunks = static_cast<int32_t>(v.n_fast_chunks);
    p.gen_begin = static_cast<int32_t>(v.gen_begin);
    p.n_gen = static_cast<int32_t>(v.n_gen);
    p.gen_stride = v.gen_stride;
    p.codes = v.codes;
    p.tabs.copytab = copy.data();
    p.tabs.addtab = add.data();
    p.tabs.abb = abb.data();
    p.tabs.uext = uext[dyn].data();
    p.tabs.values = values.data();
    p.tabs.stride = v.npad;
    p.tabs.pr = p.pr;
    p.omega = omega;
    p.om1 = 1 - omega;
    p.omega_minus = omega_minus;
    for(int i = 0; i < 27; ++i) p.rates[i] = rates[i];
    p.vars_out = vars.data();
    p.first = first;
    return p;
  }
};

template <class L>
void gather_all(Ctx& c, double* fold_out, double* mom_out) {
  const DevParams<double> p = c.params();
  const int32_t nc = static_cast<int32_t>(c.v.ghost_begin);
  launch(nc, 128, [&] { k_gather_all<L, double, true>(p, nc, fold_out, mom_out); });
}
// what a time step does to every owned cell (same device functions, without the persistent-CTA driver and its shared-memory stages)
template <class L, int COLL>
void update(Ctx& c) {
  const DevParams<double> p = c.params();
  for(int32_t cell = 0; cell < static_cast<int32_t>(c.v.ghost_begin); ++cell) {
    double fold[L::Q];
    gather_any<L, double, true>(p, p.A, cell, fold);
    update_and_store<L, double, true, COLL>(p, cell, fold);
  }
}
// The kernels of a time step THEMSELVES: k_step_generic (one thread per link-code cell) and k_step_fast, the persistent chunk CTAs
// (ticket counter, cp.async pipeline of whole chunks through shared memory, per-direction in-chunk layouts, in-place collision,
// 128-bit copy-out).  Every chunk CTA runs as kFastThreads OS threads with a real barrier behind __syncthreads(); cp.async is a memcpy
// (tests/c/fake_cuda); blocks run one after the other (legal for this kernel: a CTA that becomes resident late simply draws fewer
// tickets).  Launch geometry as Solver::one_step sets it, with a handful of persistent CTAs.
template <class L, int COLL>
void launch_step(const DevParams<double>& base, int64_t g0, int64_t ng, int64_t c0, int64_t ncnk, int persistent_ctas, unsigned long long* ticket,
                 unsigned long long* ticket_next) {
  DevParams<double> q = base;
  q.gen_off       = static_cast<int32_t>(g0);
  q.n_gen         = static_cast<int32_t>(ng);
  q.n_gen_blocks  = static_cast<int32_t>((ng + kThreads - 1) / kThreads);
  q.chunk_off     = static_cast<int32_t>(c0);
  q.n_fast_chunks = static_cast<int32_t>(ncnk);
  const int64_t ntiles = ncnk * FastCfg<L, double>::NSPLIT;
  q.n_fast_blocks = static_cast<int32_t>(ntiles < persistent_ctas ? ntiles : persistent_ctas);
  q.ticket        = ticket;
  q.ticket_base   = *ticket_next;
  *ticket_next += static_cast<unsigned long long>(ntiles) + static_cast<unsigned long long>(q.n_fast_blocks) * FastCfg<L, double>::PAST_END;
  launch(ng, kThreads, [&] { k_step_generic<L, double, true, COLL>(q); });
  blockDim.x = kFastThreads;
  gridDim.x  = static_cast<unsigned>(q.n_fast_blocks);
  pthread_barrier_init(&g_block_barrier, nullptr, kFastThreads);
  g_block_barrier_on = true;
  for(int b = 0; b < q.n_fast_blocks; ++b) {
    blockIdx.x = static_cast<unsigned>(b);
    std::vector<std::thread> th;
    for(int t = 0; t < kFastThreads; ++t)
      th.emplace_back([&q, t] {
        threadIdx.x = static_cast<unsigned>(t);
        k_step_fast<L, double, true, COLL>(q);
      });
    for(auto& x : th) x.join();
  }
  g_block_barrier_on = false;
  pthread_barrier_destroy(&g_block_barrier);
}
template <class L, int COLL>
void step_kernel(Ctx& c, int persistent_ctas) {
  launch_step<L, COLL>(c.params(), 0, c.v.n_gen, 0, c.v.n_fast_chunks, persistent_ctas, &c.ticket, &c.ticket_next);
}

// The overlapped path of Solver::one_step: two launches -- the outer cells (chunks / generic cells holding a population a peer needs,
// placed first in the device layout) with their own ticket counter, then the inner ones.
template <class L, int COLL>
void step_kernel_split(Ctx& c, int persistent_ctas) {
  const DevParams<double> base = c.params();
  launch_step<L, COLL>(base, 0, c.v.n_gen_outer, 0, c.v.n_fast_outer, persistent_ctas, &c.ticket2[1], &c.ticket2_next[1]);
  launch_step<L, COLL>(base, c.v.n_gen_outer, c.v.n_gen - c.v.n_gen_outer, c.v.n_fast_outer, c.v.n_fast_chunks - c.v.n_fast_outer, persistent_ctas,
                       &c.ticket2[0], &c.ticket2_next[0]);
}

template <class L>
void velocity_pack(Ctx& c, const int32_t* cells, int n, double* out) {
  const DevParams<double> p = c.params();
  launch(n, 128, [&] { k_velocity_pack<L, double, true>(p, cells, n, out); });
}
template <class L>
void pressure_extrapolate(Ctx& c) {
  const DevParams<double> p = c.params();
  const int n = static_cast<int>(c.abb.size());
  launch(n, 128, [&] { k_pressure_extrapolate<L, double, true>(p, n, c.uext[c.dyn ^ 1].data(), c.vrecv.data()); });
}

#define DISPATCH(c, call)                                                        \
  do {                                                                           \
    if((c)->ndim == 2) { using L = Lattice<2, 9>; call; }                        \
    else if((c)->ndist == 19) { using L = Lattice<3, 19>; call; }               \
    else { using L = Lattice<3, 27>; call; }                                     \
  } while(0)

} // namespace

extern "C" {

// solver = an inspection-only handle (device -1) whose set-up calls have been made; the harness keeps views into its plan
void* kh_create(lbm_b200_solver* solver, int ndim, int ndist, double omega) {
  auto* c = new Ctx();
  c->ndim = ndim;
  c->ndist = ndist;
  c->omega = omega;
  c->omega_minus = omega;
  for(double& r : c->rates) r = omega;
  if(lbm_b200_debug_plan(solver, &c->v) != 0) { delete c; return nullptr; }
  const lbm_b200_plan_view& v = c->v;
  const int QM = ndist - 1;
  // the conversions of Solver::init (solver.cu)
  for(int64_t k = 0; k < v.n_wall; ++k) {
    AddEntryT<double> e{};
    for(int d = 0; d < 3; ++d) e.v[d] = v.wall_desc[k * 4 + d];
    e.n = static_cast<int32_t>(v.wall_desc[k * 4 + 3]);
    c->wall.push_back(e);
  }
  if(c->wall.empty()) c->wall.resize(static_cast<size_t>(QM) * 27);
  for(int64_t k = 0; k < v.n_add; ++k) {
    AddEntryT<double> e{};
    for(int d = 0; d < 3; ++d) e.v[d] = v.addtab[k * 4 + d];
    e.n = static_cast<int32_t>(v.addtab[k * 4 + 3]);
    c->add.push_back(e);
  }
  for(int64_t k = 0; k < v.n_copy; ++k) c->copy.push_back({v.copytab[k * 2], v.copytab[k * 2 + 1]});
  for(int64_t k = 0; k < v.n_abb; ++k) c->abb.push_back({v.abb_cells[k * 3], v.abb_cells[k * 3 + 1], v.abb_cells[k * 3 + 2], v.abb_p[k]});
  for(int b = 0; b < 2; ++b) c->uext[b].assign(static_cast<size_t>(v.n_abb) * 3 + 3, 0.0);
  c->values.assign(v.values, v.values + v.n_values);
  c->A.assign(static_cast<size_t>(ndist) * v.npad, 0.0);
  c->B = c->A;
  c->vars.assign(static_cast<size_t>(ndim + 1) * v.npad, 0.0);
  c->scratch.assign(static_cast<size_t>(ndist) * v.npad, 0.0);
  return c;
}
void kh_destroy(void* p) { delete static_cast<Ctx*>(p); }
int64_t kh_npad(void* p) { return static_cast<Ctx*>(p)->v.npad; }
double* kh_A(void* p) { return static_cast<Ctx*>(p)->A.data(); }
double* kh_B(void* p) { return static_cast<Ctx*>(p)->B.data(); }
double* kh_vars(void* p) { return static_cast<Ctx*>(p)->vars.data(); }
double* kh_values(void* p) { return static_cast<Ctx*>(p)->values.data(); }
double* kh_uext(void* p, int next) { auto* c = static_cast<Ctx*>(p); return c->uext[next ? c->dyn ^ 1 : c->dyn].data(); }
void kh_set_first(void* p, int first) { static_cast<Ctx*>(p)->first = first; }
void kh_set_vrecv(void* p, const double* v, int64_t n) { static_cast<Ctx*>(p)->vrecv.assign(v, v + n); if(n == 0) static_cast<Ctx*>(p)->vrecv.assign(3, 0.0); }

// m_fold (SoA, device order) and its moments from the current A: k_gather_all
void kh_gather_all(void* p, double* fold_out, double* mom_out) { auto* c = static_cast<Ctx*>(p); DISPATCH(c, gather_all<L>(*c, fold_out, mom_out)); }
// one time step of the owned cells: A -> B (and vars), like the fused kernel
void kh_update(void* p) {
  auto* c = static_cast<Ctx*>(p);
  if(c->coll == COLL_TRT) DISPATCH(c, (update<L, COLL_TRT>(*c)));
  else if(c->coll == COLL_MRT) DISPATCH(c, (update<L, COLL_MRT>(*c)));
  else DISPATCH(c, (update<L, COLL_BGK>(*c)));
}
// outer launch, then inner launch (the overlapped path of a partitioned run)
void kh_step_kernel_split(void* p, int persistent_ctas) {
  auto* c = static_cast<Ctx*>(p);
  if(c->coll == COLL_TRT) DISPATCH(c, (step_kernel_split<L, COLL_TRT>(*c, persistent_ctas)));
  else if(c->coll == COLL_MRT) DISPATCH(c, (step_kernel_split<L, COLL_MRT>(*c, persistent_ctas)));
  else DISPATCH(c, (step_kernel_split<L, COLL_BGK>(*c, persistent_ctas)));
}
void kh_set_collision(void* p, int coll, double omega_minus, const double* rates) {
  auto* c = static_cast<Ctx*>(p);
  c->coll = coll;
  c->omega_minus = omega_minus;
  // MRT: the kernel takes the base rate in omega and (s_k - s0) / |row k|^2 per moment (Solver::params, solver_fused.cuh)
  for(int i = 0; i < 27; ++i) c->rates[i] = rates[i];
  if(coll == COLL_MRT) {
    const int d = c->ndim, q = c->ndist;
    const double s0 = mrt_base_rate(rates, q, d);
    c->omega = s0;
    for(int i = 0; i < 27; ++i) {
      double norm = 1;
      if(i < q) DISPATCH(c, norm = MrtBasis<L>::norm(i));
      c->rates[i] = (i > d && i < q) ? (rates[i] - s0) / norm : 0.0;
    }
  }
}
// the same step through the real kernel (generic blocks + persistent chunk CTAs)
void kh_step_kernel(void* p, int persistent_ctas) {
  auto* c = static_cast<Ctx*>(p);
  if(c->coll == COLL_TRT) DISPATCH(c, (step_kernel<L, COLL_TRT>(*c, persistent_ctas)));
  else if(c->coll == COLL_MRT) DISPATCH(c, (step_kernel<L, COLL_MRT>(*c, persistent_ctas)));
  else DISPATCH(c, (step_kernel<L, COLL_BGK>(*c, persistent_ctas)));
}
void kh_velocity_pack(void* p, const int32_t* cells, int n, double* out) { auto* c = static_cast<Ctx*>(p); DISPATCH(c, velocity_pack<L>(*c, cells, n, out)); }
void kh_pressure_extrapolate(void* p) { auto* c = static_cast<Ctx*>(p); DISPATCH(c, pressure_extrapolate<L>(*c)); }
void kh_halo_pack(void* p, const int64_t* index, int64_t n, double* out) {
  auto* c = static_cast<Ctx*>(p);
  launch(n, 256, [&] { k_halo_pack<double>(c->B.data(), index, n, out); });
}
void kh_halo_unpack(void* p, const int64_t* index, int64_t n, const double* in) {
  auto* c = static_cast<Ctx*>(p);
  launch(n, 256, [&] { k_halo_unpack<double>(c->B.data(), index, n, in); });
}
// Ar<double, true>::div_const against true division, for the three constant divisors of the equilibrium (cs^2, 2 cs^4, 2 cs^2): the number
// of operands whose results differ (must be 0: the scheme is correctly rounded; the sign of a zero quotient is the one exception)
int64_t kh_check_div_const(const double* a, int64_t n) {
  const double d[3] = {1.0 / 3.0, 2.0 * (1.0 / 3.0) * (1.0 / 3.0), 2.0 * (1.0 / 3.0)};
  int64_t bad = 0;
  for(int k = 0; k < 3; ++k) {
    const double y = 1.0 / d[k];
    for(int64_t i = 0; i < n; ++i) {
      const double q = Ar<double, true>::div_const(a[i], d[k], y), t = a[i] / d[k];
      bad += q != t ? 1 : 0; // numeric comparison: a = -0 gives +0 where IEEE gives -0; the quotient only ever enters a sum with 1
    }
  }
  return bad;
}
// end of a step: B becomes A, the dynamic buffers written for the next step become current (Solver::one_step)
void kh_swap(void* p, int dyn_written) {
  auto* c = static_cast<Ctx*>(p);
  c->A.swap(c->B);
  if(dyn_written) c->dyn ^= 1;
  c->first = 0;
}

} // extern "C"
