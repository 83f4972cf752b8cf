// solver.cu -- C ABI (include/lbm_b200.h) on top of the device plan (plan.hpp) and the kernels (kernels.cuh).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"
#include "kernels.cuh"
#include "plan.hpp"
#include "grid_box.hpp"
#include "nccl_dyn.hpp"
#include "sequential.cuh"
#include "poisson.cuh"
#include "partition.hpp"

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

#define NCCL_TRY(expr)                                                                                  \
  do {                                                                                                  \
    ncclResult_t r_ = (expr);                                                                           \
    if(r_ != ncclSuccess)                                                                               \
      return fail(LBM_B200_ECUDA, std::string(#expr) + ": " + lbm::nccl_api().GetErrorString(r_));      \
  } while(0)

#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if(e_ != cudaSuccess)                                                                               \
      return fail(LBM_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                  \
  } while(0)

template <class T>
struct DevBuf {
  T*     p = nullptr;
  size_t n = 0;
  ~DevBuf() { if(p) cudaFree(p); }
  cudaError_t alloc(size_t count) {
    if(p) cudaFree(p);
    p = nullptr;
    n = count;
    if(count == 0) return cudaSuccess;
    return cudaMalloc(&p, count * sizeof(T));
  }
  cudaError_t upload(const std::vector<T>& h) {
    cudaError_t e = alloc(h.size());
    if(e != cudaSuccess || h.empty()) return e;
    return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  }
  size_t bytes() const { return n * sizeof(T); }
};

struct SolverBase {
  virtual ~SolverBase() = default;
  virtual int init()                                                     = 0;
  virtual int step(int64_t n, float* ms_total, float* ms_main)           = 0;
  virtual int sync()                                                     = 0;
  virtual int residual(double* out, int32_t* diverged)                   = 0;
  virtual int get_populations(double* f, double* fold)                   = 0;
  virtual int set_populations(const double* f, const double* fold)       = 0;
  virtual int get_vars(double* vars, double* varsold)                    = 0;
  virtual int get_moments(double* m)                                     = 0;
  virtual void stats(lbm_b200_stats* st) const                           = 0;
  virtual int64_t owned() const                                          = 0;
  virtual int debug_plan(lbm_b200_plan_view* out) { (void)out; return fail(LBM_B200_EUNSUP, "no device plan for this solver kind"); }
  lbm_b200_config cfg{};
  lbm::PlanInput  in;
  cudaStream_t    stream = nullptr;
  ncclComm_t      comm   = nullptr;
  int             comm_rank = 0, comm_size = 1;
  bool            inited = false;
  int64_t         t      = 0;
};

template <class L, class Real>
struct Solver final : SolverBase {
  static constexpr int Q = L::Q, D = L::D, NVAR = L::D + 1;
  lbm::Plan plan;
  // device state
  DevBuf<Real>     f[2];      // populations, double buffered; f[cur] = m_f of the reference after t steps
  DevBuf<Real>     prev_fold; // only after set_populations: an explicit m_fold to start from
  DevBuf<Real>     vars[2];   // tracked m_vars / m_varsold
  DevBuf<Real>     scratch;   // [max(Q,NVAR)][npad] read-back staging
  DevBuf<uint16_t> d_tmpl;
  DevBuf<int32_t>  d_chunk_nb, d_codes, d_chunk_abb_base, d_chunk_abb;
  DevBuf<lbm::AddEntryT<Real>> d_wall;
  DevBuf<lbm::CopySrcDev>       d_copy;
  DevBuf<lbm::AddEntryT<Real>>  d_add;
  DevBuf<lbm::AbbDev<Real>>     d_abb;
  DevBuf<lbm::ForceDev<Real>>   d_force;
  DevBuf<lbm::PerPDev<Real>>    d_perp;
  DevBuf<lbm::VarFixDev<Real>>  d_varfix;
  DevBuf<Real>     d_uext[2], d_values[2];
  DevBuf<double>   d_partial;
  DevBuf<double>   stage;     // AoS staging for host transfers [n][Q]
  DevBuf<int32_t>  d_ref2dev;
  DevBuf<int64_t>  d_send_idx, d_recv_idx;
  DevBuf<unsigned long long> d_ticket; // [0]: inner / whole-domain launches, [1]: outer launches
  unsigned long long ticket_next[2] = {0, 0};
  DevBuf<Real>     d_sendbuf, d_recvbuf;
  DevBuf<int32_t>  d_vsend_cells;          // velocity halo of the pressure boundary condition
  DevBuf<Real>     d_vsendbuf, d_vrecvbuf; // 3 reals per item
  int64_t          halo_bytes = 0;
  int64_t          h2d_bytes = 0, d2h_bytes = 0;
  int cur = 0;       // f[cur] holds the current post-collision populations
  int dyn = 0;       // d_uext[dyn] / d_values[dyn] are the ones the next gather must use
  int vcur = 0;      // vars[vcur] = m_vars, vars[vcur^1] = m_varsold
  int64_t vars_step[2] = {-1, -1};
  bool    overlap_enabled = true, debug_identity = false;
  bool    first = true; // next step is step 0 of the reference loop (m_fold = initial condition)
  int     n_fast_blocks = 0, n_gen_blocks = 0, max_resident = 0;
  cudaStream_t comm_stream = nullptr;   // halo exchange runs here, overlapped with the update of the inner cells
  cudaEvent_t  ev_outer = nullptr, ev_halo = nullptr, ev_pack = nullptr;
  bool         halo_pending = false;
  int64_t launches = 0, launches_main = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm0 = nullptr, evm1 = nullptr;

  ~Solver() override {
    if(ev0) cudaEventDestroy(ev0);
    if(ev1) cudaEventDestroy(ev1);
    if(evm0) cudaEventDestroy(evm0);
    if(evm1) cudaEventDestroy(evm1);
    if(ev_outer) cudaEventDestroy(ev_outer);
    if(ev_halo) cudaEventDestroy(ev_halo);
    if(ev_pack) cudaEventDestroy(ev_pack);
    if(comm_stream) cudaStreamDestroy(comm_stream);
  }

  lbm::DevParams<Real> params(int src, int dst, Real* vars_out) const {
    lbm::DevParams<Real> p{};
    p.A = f[src].p;
    p.B = f[dst].p;
    p.stride = plan.npad;
    p.tmpl = d_tmpl.p;
    p.chunk_nb = d_chunk_nb.p;
    p.wall_desc = d_wall.p;
    p.chunk_abb_base = d_chunk_abb_base.p;
    p.chunk_abb = d_chunk_abb.p;
    p.n_fast_chunks = static_cast<int32_t>(plan.n_fast_chunks);
    p.chunk_off = 0;
    p.gen_off = 0;
    p.n_fast_blocks = n_fast_blocks;
    p.gen_begin = static_cast<int32_t>(plan.gen_begin);
    p.n_gen = static_cast<int32_t>(plan.n_gen);
    p.n_gen_blocks = n_gen_blocks;
    p.gen_stride = plan.gen_stride;
    p.codes = d_codes.p;
    p.tabs.copytab = d_copy.p;
    p.tabs.addtab = d_add.p;
    p.tabs.abb = d_abb.p;
    p.tabs.uext = d_uext[dyn].p;
    p.tabs.values = d_values[dyn].p;
    p.tabs.stride = plan.npad;
    p.omega = static_cast<Real>(cfg.omega);
    p.om1 = static_cast<Real>(1 - cfg.omega);
    p.omega_minus = static_cast<Real>(cfg.omega_minus);
    for(int i = 0; i < 27; ++i) p.rates[i] = static_cast<Real>(cfg.mrt_rates[i]);
    p.vars_out = vars_out;
    p.first = first ? 1 : 0;
    if(debug_identity) p.first = 1; // timing experiments only (LBM_B200_DEBUG_IDENTITY): every step reads its own cell, no gather
    return p;
  }

  template <bool STRICT, int COLL>
  static auto kernel_ptr() { return &lbm::k_step<L, Real, STRICT, COLL>; }

  using KernelFn = void (*)(const lbm::DevParams<Real>);
  KernelFn main_kernel() const {
    const bool strict = cfg.arithmetic == LBM_B200_STRICT;
    switch(cfg.collision) {
      case LBM_B200_TRT: return strict ? kernel_ptr<true, lbm::COLL_TRT>() : kernel_ptr<false, lbm::COLL_TRT>();
      case LBM_B200_MRT: return strict ? kernel_ptr<true, lbm::COLL_MRT>() : kernel_ptr<false, lbm::COLL_MRT>();
      default: return strict ? kernel_ptr<true, lbm::COLL_BGK>() : kernel_ptr<false, lbm::COLL_BGK>();
    }
  }

  int init() override {
    debug_identity = std::getenv("LBM_B200_DEBUG_IDENTITY") != nullptr;
    if(const char* e = std::getenv("LBM_B200_NO_OVERLAP")) overlap_enabled = e[0] == '0' || e[0] == 0;
    if(!lbm::build_plan(in, plan)) return fail(plan.error.find("order-dependent") != std::string::npos ? LBM_B200_EUNSUP : LBM_B200_EINVAL, plan.error);
    std::vector<int32_t>().swap(in.nghbr);
    CUDA_TRY(cudaSetDevice(cfg.device));
    const size_t npad = static_cast<size_t>(plan.npad);
    for(int b = 0; b < 2; ++b) {
      CUDA_TRY(f[b].alloc(npad * Q));
      CUDA_TRY(cudaMemset(f[b].p, 0, f[b].bytes()));
    }
    if(cfg.track_vars > 0) {
      for(int b = 0; b < 2; ++b) {
        CUDA_TRY(vars[b].alloc(npad * NVAR));
        CUDA_TRY(cudaMemset(vars[b].p, 0, vars[b].bytes()));
      }
    }
    CUDA_TRY(scratch.alloc(npad * (Q > NVAR ? Q : NVAR)));
    CUDA_TRY(cudaMemset(scratch.p, 0, scratch.bytes()));
    CUDA_TRY(d_tmpl.upload(plan.tmpl));
    CUDA_TRY(d_chunk_nb.upload(plan.chunk_nb));
    CUDA_TRY(d_chunk_abb_base.upload(plan.chunk_abb_base));
    CUDA_TRY(d_chunk_abb.upload(plan.chunk_abb));
    CUDA_TRY(d_codes.upload(plan.codes));
    CUDA_TRY(d_ref2dev.upload(plan.ref2dev));
    if(!in.peers.empty()) {
      if(comm == nullptr) return fail(LBM_B200_ESTATE, "halo lists set but lbm_b200_comm_init has not been called");
      CUDA_TRY(d_send_idx.upload(plan.send_index));
      CUDA_TRY(d_recv_idx.upload(plan.recv_index));
      CUDA_TRY(d_sendbuf.alloc(plan.send_index.size() + 1));
      CUDA_TRY(d_recvbuf.alloc(plan.recv_index.size() + 1));
      CUDA_TRY(d_vsend_cells.upload(plan.vsend_cells));
      CUDA_TRY(d_vsendbuf.alloc(plan.vsend_cells.size() * 3 + 3));
    }
    CUDA_TRY(d_vrecvbuf.alloc(static_cast<size_t>(plan.n_vrecv) * 3 + 3)); // never null: the pressure kernel takes the pointer
    CUDA_TRY(cudaMemset(d_vrecvbuf.p, 0, d_vrecvbuf.bytes()));
    {
      std::vector<lbm::CopySrcDev> h;
      for(auto& c : plan.copytab) h.push_back({c.cell, c.dir});
      CUDA_TRY(d_copy.upload(h));
    }
    {
      std::vector<lbm::AddEntryT<Real>> h;
      for(auto& a : plan.addtab) {
        lbm::AddEntryT<Real> e{};
        for(int d = 0; d < 3; ++d) e.v[d] = static_cast<Real>(a.v[d]);
        e.n = a.n;
        h.push_back(e);
      }
      CUDA_TRY(d_add.upload(h));
    }
    {
      std::vector<lbm::AddEntryT<Real>> h;
      for(auto& a : plan.wall_desc) {
        lbm::AddEntryT<Real> e{};
        for(int d = 0; d < 3; ++d) e.v[d] = static_cast<Real>(a.v[d]);
        e.n = a.n;
        h.push_back(e);
      }
      if(h.empty()) h.resize(static_cast<size_t>(Q - 1) * L::NSEL); // so that the pointer arithmetic in the kernel always has a valid base
      CUDA_TRY(d_wall.upload(h));
    }
    {
      std::vector<lbm::AbbDev<Real>> h;
      for(auto& a : plan.abb) h.push_back({a.cell, a.n1, a.n2, static_cast<Real>(a.p)});
      CUDA_TRY(d_abb.upload(h));
    }
    {
      std::vector<lbm::ForceDev<Real>> h;
      for(auto& a : plan.force) h.push_back({a.target, a.val, static_cast<Real>(a.p)});
      CUDA_TRY(d_force.upload(h));
    }
    {
      std::vector<lbm::PerPDev<Real>> h;
      for(auto& a : plan.perp) h.push_back({a.cell, a.vbase, static_cast<Real>(a.p)});
      CUDA_TRY(d_perp.upload(h));
    }
    {
      std::vector<lbm::VarFixDev<Real>> h;
      for(auto& a : plan.varfix) h.push_back({a.cell, a.var, a.abb, a.comp, static_cast<Real>(a.value)});
      CUDA_TRY(d_varfix.upload(h));
    }
    for(int b = 0; b < 2; ++b) {
      CUDA_TRY(d_uext[b].alloc(plan.abb.size() * 3 + 3));
      CUDA_TRY(cudaMemset(d_uext[b].p, 0, d_uext[b].bytes()));
    }
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(cudaEventCreate(&evm0));
    CUDA_TRY(cudaEventCreate(&evm1));

    // launch geometry: generic blocks first, then persistent fast blocks (a multiple of the SM count)
    n_gen_blocks = static_cast<int>((plan.n_gen + lbm::kThreads - 1) / lbm::kThreads);
    {
      cudaDeviceProp prop{};
      CUDA_TRY(cudaGetDeviceProperties(&prop, cfg.device));
      int per_sm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, main_kernel(), lbm::kThreads, 0));
      if(per_sm < 1) per_sm = 1;
      max_resident  = prop.multiProcessorCount * per_sm;
      int64_t want  = max_resident;
      if(want > plan.n_fast_chunks) want = plan.n_fast_chunks;
      n_fast_blocks = static_cast<int>(want);
    }
    if(!in.peers.empty()) {
      int prio_lo = 0, prio_hi = 0;
      CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      CUDA_TRY(cudaStreamCreateWithPriority(&comm_stream, cudaStreamNonBlocking, prio_hi));
      CUDA_TRY(cudaEventCreateWithFlags(&ev_outer, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&ev_halo, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&ev_pack, cudaEventDisableTiming));
    }

    // ---- initialCondition(): vars = 0, boundary presets, rho = 1, f = fold = feq   (solver.cpp:267-295)
    {
      std::vector<Real> v0(npad * NVAR, Real(0));
      for(size_t k = 0; k < plan.u0_cells.size(); ++k)
        for(int d = 0; d < D; ++d) v0[static_cast<size_t>(d) * npad + plan.u0_cells[k]] = static_cast<Real>(plan.u0_vals[k * D + d]);
      for(size_t c = 0; c < npad; ++c) v0[static_cast<size_t>(D) * npad + c] = plan.dev2ref[c] >= 0 ? Real(1) : Real(0);
      Real* d_v0 = cfg.track_vars > 0 ? vars[0].p : scratch.p;
      CUDA_TRY(cudaMemcpy(d_v0, v0.data(), v0.size() * sizeof(Real), cudaMemcpyHostToDevice));
      const int nb = static_cast<int>((npad + 255) / 256);
      if(cfg.arithmetic == LBM_B200_STRICT)
        lbm::k_init<L, Real, true><<<nb, 256, 0, stream>>>(f[0].p, d_v0, plan.npad, static_cast<int32_t>(npad));
      else
        lbm::k_init<L, Real, false><<<nb, 256, 0, stream>>>(f[0].p, d_v0, plan.npad, static_cast<int32_t>(npad));
      CUDA_TRY(cudaGetLastError());
      // padding cells must stay zero (rho preset 0 gives feq = 0)
      // slots nothing ever writes keep their initial m_fold value: fetch it once
      std::vector<double> values = plan.values;
      if(!plan.stale_ref.empty()) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        std::vector<Real> h(npad * Q);
        CUDA_TRY(cudaMemcpy(h.data(), f[0].p, h.size() * sizeof(Real), cudaMemcpyDeviceToHost));
        for(size_t k = 0; k < plan.stale_ref.size(); ++k) {
          const int64_t cell = plan.stale_ref[k] / Q;
          const int     dir  = static_cast<int>(plan.stale_ref[k] % Q);
          values[k + 1]      = static_cast<double>(h[static_cast<size_t>(dir) * npad + cell]);
        }
      }
      std::vector<Real> hv(values.size());
      for(size_t k = 0; k < values.size(); ++k) hv[k] = static_cast<Real>(values[k]);
      for(int b = 0; b < 2; ++b) CUDA_TRY(d_values[b].upload(hv));
    }
    CUDA_TRY(d_partial.alloc(static_cast<size_t>(NVAR) * 1024));
    CUDA_TRY(d_ticket.alloc(2));
    CUDA_TRY(cudaMemset(d_ticket.p, 0, d_ticket.bytes()));
    ticket_next[0] = ticket_next[1] = 0;
    CUDA_TRY(cudaStreamSynchronize(stream));
    cur = 0;
    dyn = 0;
    vcur = 0;
    vars_step[0] = 0;
    vars_step[1] = -1;
    first = true;
    t = 0;
    inited = true;
    return LBM_B200_OK;
  }

  bool want_vars(int64_t s) const {
    // kernel of reference step s produces m_vars as seen at loop index s+1
    if(cfg.track_vars <= 0) return false;
    if(cfg.track_vars == 1) return true;
    const int64_t k = cfg.track_vars;
    return (s + 1) % k == 0 || (s + 2) % k == 0;
  }

  // phase 1: forcing, periodic-with-pressure values; phase 2: pressure extrapolation (needs the velocity halo of THIS step when
  // a partition cut separates a pressure cell from its inward neighbours) and the m_vars fix-ups.  nd = the dynamic buffers
  // written for the next step; the caller flips `dyn` once both phases have run.
  template <bool STRICT>
  int aux_kernels(const lbm::DevParams<Real>& p, Real* vars_out, int nd, int phases, bool* dyn_written_out) {
    bool dyn_written = false;
    if((phases & 1) && d_force.n > 0) {
      const int n = static_cast<int>(d_force.n);
      lbm::k_forcing<L, Real, STRICT><<<(n + 127) / 128, 128, 0, stream>>>(p, d_force.p, n);
      ++launches;
    }
    if((phases & 1) && d_perp.n > 0) {
      const int n = static_cast<int>(d_perp.n);
      lbm::k_periodic_pressure<L, Real, STRICT><<<(n + 127) / 128, 128, 0, stream>>>(p, d_perp.p, n, d_values[nd].p);
      ++launches;
      dyn_written = true;
    }
    if((phases & 2) && d_abb.n > 0) {
      const int n = static_cast<int>(d_abb.n);
      lbm::k_pressure_extrapolate<L, Real, STRICT><<<(n + 127) / 128, 128, 0, stream>>>(p, n, d_uext[nd].p, d_vrecvbuf.p);
      ++launches;
      dyn_written = true;
    }
    if((phases & 2) && vars_out != nullptr && d_varfix.n > 0) {
      const int n = static_cast<int>(d_varfix.n);
      lbm::k_varfix<Real><<<(n + 127) / 128, 128, 0, stream>>>(d_varfix.p, n, d_uext[nd].p, vars_out, plan.npad);
      ++launches;
    }
    if(dyn_written) *dyn_written_out = true;
    CUDA_TRY(cudaGetLastError());
    return LBM_B200_OK;
  }

  // Outgoing populations of this step -> peers, theirs -> my ghost cells.  One pack kernel, one NCCL group of
  // send/recv pairs over NVLink, one unpack kernel, all on the solver's stream.
  // `vp` (only with a velocity halo): the parameters of this step, from which the sending side rebuilds the velocity of the
  // cells a peer's pressure boundary condition extrapolates from.
  int halo_exchange(Real* buf, cudaStream_t stream, cudaEvent_t after_pack = nullptr, const lbm::DevParams<Real>* vp = nullptr) {
    if(in.peers.empty()) return LBM_B200_OK;
    auto& nc = lbm::nccl_api();
    const int64_t ns = static_cast<int64_t>(plan.send_index.size()), nr = static_cast<int64_t>(plan.recv_index.size());
    if(ns > 0) {
      lbm::k_halo_pack<Real><<<static_cast<int>((ns + 255) / 256), 256, 0, stream>>>(buf, d_send_idx.p, ns, d_sendbuf.p);
      ++launches;
    }
    const bool with_velocity = vp != nullptr && has_velocity_halo();
    if(with_velocity && !plan.vsend_cells.empty()) {
      const int n = static_cast<int>(plan.vsend_cells.size());
      if(cfg.arithmetic == LBM_B200_STRICT) lbm::k_velocity_pack<L, Real, true><<<(n + 127) / 128, 128, 0, stream>>>(*vp, d_vsend_cells.p, n, d_vsendbuf.p);
      else lbm::k_velocity_pack<L, Real, false><<<(n + 127) / 128, 128, 0, stream>>>(*vp, d_vsend_cells.p, n, d_vsendbuf.p);
      ++launches;
    }
    if(after_pack != nullptr) CUDA_TRY(cudaEventRecord(after_pack, stream));
    const ncclDataType_t dt = sizeof(Real) == 8 ? ncclFloat64 : ncclFloat32;
    NCCL_TRY(nc.GroupStart());
    int64_t so = 0, ro = 0;
    for(size_t k = 0; k < in.peers.size(); ++k) {
      if(in.send_count[k] > 0) NCCL_TRY(nc.Send(d_sendbuf.p + so, static_cast<size_t>(in.send_count[k]), dt, in.peers[k], comm, stream));
      if(in.recv_count[k] > 0) NCCL_TRY(nc.Recv(d_recvbuf.p + ro, static_cast<size_t>(in.recv_count[k]), dt, in.peers[k], comm, stream));
      so += in.send_count[k];
      ro += in.recv_count[k];
    }
    if(with_velocity) { // second message per peer pair, matched in order inside the same group
      int64_t vso = 0, vro = 0;
      for(size_t k = 0; k < in.peers.size(); ++k) {
        const int64_t vs = in.vsend_count.empty() ? 0 : in.vsend_count[k], vr = in.vrecv_count.empty() ? 0 : in.vrecv_count[k];
        if(vs > 0) NCCL_TRY(nc.Send(d_vsendbuf.p + 3 * vso, static_cast<size_t>(3 * vs), dt, in.peers[k], comm, stream));
        if(vr > 0) NCCL_TRY(nc.Recv(d_vrecvbuf.p + 3 * vro, static_cast<size_t>(3 * vr), dt, in.peers[k], comm, stream));
        vso += vs;
        vro += vr;
        halo_bytes += 3 * (vs + vr) * static_cast<int64_t>(sizeof(Real));
      }
    }
    NCCL_TRY(nc.GroupEnd());
    if(nr > 0) {
      lbm::k_halo_unpack<Real><<<static_cast<int>((nr + 255) / 256), 256, 0, stream>>>(buf, d_recv_idx.p, nr, d_recvbuf.p);
      ++launches;
    }
    halo_bytes += (ns + nr) * static_cast<int64_t>(sizeof(Real));
    CUDA_TRY(cudaGetLastError());
    return LBM_B200_OK;
  }

  bool has_velocity_halo() const { return !in.vsend_cell.empty() || !in.vrecv_cell.empty(); }

  int one_step(bool time_main) {
    const int src = cur, dst = cur ^ 1;
    Real*     vout = nullptr;
    if(want_vars(t)) vout = vars[vcur ^ 1].p;
    lbm::DevParams<Real> p = params(src, dst, vout);
    if(prev_fold.p != nullptr) p.A = prev_fold.p; // explicit m_fold supplied by set_populations
    // a launch over generic cells [g0, g0+ng) and fast chunks [c0, c0+ncnk)
    auto launch = [&](int64_t g0, int64_t ng, int64_t c0, int64_t ncnk, int resident_cap, int cls) {
      lbm::DevParams<Real> q = p;
      q.gen_off       = static_cast<int32_t>(g0);
      q.n_gen         = static_cast<int32_t>(ng);
      q.n_gen_blocks  = static_cast<int>((ng + lbm::kThreads - 1) / lbm::kThreads);
      q.chunk_off     = static_cast<int32_t>(c0);
      q.n_fast_chunks = static_cast<int32_t>(ncnk);
      q.n_fast_blocks = static_cast<int32_t>(ncnk < resident_cap ? ncnk : resident_cap);
      q.ticket        = d_ticket.p + cls;
      q.ticket_base   = ticket_next[cls];
      ticket_next[cls] += static_cast<unsigned long long>(ncnk) + static_cast<unsigned long long>(q.n_fast_blocks);
      const int grid  = q.n_gen_blocks + q.n_fast_blocks;
      if(grid > 0) {
        main_kernel()<<<grid, lbm::kThreads, 0, stream>>>(q);
        ++launches;
        ++launches_main;
      }
    };
    const bool has_aux = d_force.n > 0 || d_perp.n > 0 || d_abb.n > 0 || (vout != nullptr && d_varfix.n > 0);
    // forcing and the periodic-with-pressure values write populations of buffer B that a peer may need: they must precede the pack.
    // The pressure extrapolation and the m_vars fix-ups only read buffer A and write uext / vars, so they do not stand in the way of
    // the overlap (unless the extrapolation itself waits for this step's velocity halo).
    const bool aux_before_exchange = d_force.n > 0 || d_perp.n > 0;
    const bool overlap = !in.peers.empty() && !aux_before_exchange && !has_velocity_halo() && overlap_enabled;
    if(halo_pending) { // ghosts of the buffer we are about to read were filled on the communication stream
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_halo, 0));
      halo_pending = false;
    }
    if(time_main) cudaEventRecord(evm0, stream);
    int rc = LBM_B200_OK;
    if(overlap) {
      // 1. outer cells (whatever a peer needs) with the whole GPU; 2. their populations are packed and travel (NCCL group on
      // the high-priority communication stream) while 3. the inner cells are updated.  The inner launch is released by the
      // same event that releases the NCCL kernel, so the (higher-priority, whole-SM-sized) NCCL CTAs are placed first and
      // the persistent inner CTAs fill the remaining SMs; with ticket scheduling late inner CTAs just take fewer chunks.
      launch(0, plan.n_gen_outer, 0, plan.n_fast_outer, max_resident, 1);
      CUDA_TRY(cudaEventRecord(ev_outer, stream));
      CUDA_TRY(cudaStreamWaitEvent(comm_stream, ev_outer, 0));
      rc = halo_exchange(f[dst].p, comm_stream, ev_pack);
      if(rc != LBM_B200_OK) return rc;
      CUDA_TRY(cudaEventRecord(ev_halo, comm_stream));
      halo_pending = true;
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_pack, 0));
      launch(plan.n_gen_outer, plan.n_gen - plan.n_gen_outer, plan.n_fast_outer, plan.n_fast_chunks - plan.n_fast_outer, max_resident, 0);
      if(time_main) cudaEventRecord(evm1, stream);
      CUDA_TRY(cudaGetLastError());
      if(has_aux) { // pressure boundary present: extrapolation + m_vars fix-ups behind the inner launch, the exchange still in flight
        const int nd = dyn ^ 1;
        bool      dyn_written = false;
        rc = cfg.arithmetic == LBM_B200_STRICT ? aux_kernels<true>(p, vout, nd, 2, &dyn_written) : aux_kernels<false>(p, vout, nd, 2, &dyn_written);
        if(rc != LBM_B200_OK) return rc;
        if(dyn_written) dyn = nd;
      }
    } else {
      launch(0, plan.n_gen, 0, plan.n_fast_chunks, max_resident, 0);
      if(time_main) cudaEventRecord(evm1, stream);
      CUDA_TRY(cudaGetLastError());
      const int  nd = dyn ^ 1;
      bool       dyn_written = false;
      const bool strict = cfg.arithmetic == LBM_B200_STRICT;
      auto aux = [&](int phases) { return strict ? aux_kernels<true>(p, vout, nd, phases, &dyn_written) : aux_kernels<false>(p, vout, nd, phases, &dyn_written); };
      if(has_velocity_halo()) {
        // the pressure extrapolation of this step reads velocities that arrive with this step's exchange
        rc = aux(1);
        if(rc != LBM_B200_OK) return rc;
        rc = halo_exchange(f[dst].p, stream, nullptr, &p);
        if(rc != LBM_B200_OK) return rc;
        rc = aux(2);
        if(rc != LBM_B200_OK) return rc;
      } else {
        rc = aux(3);
        if(rc != LBM_B200_OK) return rc;
        rc = halo_exchange(f[dst].p, stream);
        if(rc != LBM_B200_OK) return rc;
      }
      if(dyn_written) dyn = nd;
    }
    if(vout != nullptr) {
      vcur ^= 1;
      vars_step[vcur] = t + 1;
    }
    if(prev_fold.p != nullptr) {
      CUDA_TRY(cudaStreamSynchronize(stream));
      if(comm_stream) CUDA_TRY(cudaStreamSynchronize(comm_stream));
      prev_fold.alloc(0);
    }
    cur   = dst;
    first = false;
    ++t;
    return LBM_B200_OK;
  }

  int step(int64_t n, float* ms_total, float* ms_main) override {
    if(!inited) return fail(LBM_B200_ESTATE, "lbm_b200_step before lbm_b200_init");
    const bool timed = ms_total != nullptr;
    double     main_acc = 0;
    if(timed) CUDA_TRY(cudaEventRecord(ev0, stream));
    for(int64_t s = 0; s < n; ++s) {
      int rc = one_step(timed && ms_main != nullptr);
      if(rc != LBM_B200_OK) return rc;
      if(timed && ms_main != nullptr) {
        CUDA_TRY(cudaEventSynchronize(evm1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, evm0, evm1));
        main_acc += ms;
      }
    }
    if(halo_pending) { // the step is complete only when the ghosts have arrived
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_halo, 0));
      halo_pending = false;
    }
    if(timed) {
      CUDA_TRY(cudaEventRecord(ev1, stream));
      CUDA_TRY(cudaEventSynchronize(ev1));
      CUDA_TRY(cudaEventElapsedTime(ms_total, ev0, ev1));
      if(ms_main) *ms_main = static_cast<float>(main_acc);
    }
    return LBM_B200_OK;
  }

  int sync() override {
    if(comm_stream) CUDA_TRY(cudaStreamSynchronize(comm_stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LBM_B200_OK;
  }

  // SoA device array [width][npad] -> AoS host array [n][width] in reference cell order (and back).
  // The transposition runs on the device; the host side is one cudaMemcpy of the reference's own layout, so a
  // pinned caller buffer moves at PCIe speed.
  int ensure_stage() {
    if(stage.p == nullptr) CUDA_TRY(stage.alloc(static_cast<size_t>(plan.n) * Q));
    return LBM_B200_OK;
  }
  int download(const Real* dsrc, int width, double* out) {
    int rc = ensure_stage();
    if(rc) return rc;
    const int nb = static_cast<int>((plan.n + 255) / 256);
    lbm::k_pack_aos<Real><<<nb, 256, 0, stream>>>(dsrc, d_ref2dev.p, plan.n, width, stage.p, plan.npad);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, stage.p, sizeof(double) * static_cast<size_t>(plan.n) * width, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    d2h_bytes += static_cast<int64_t>(sizeof(double)) * plan.n * width;
    return LBM_B200_OK;
  }
  int upload_aos(const double* src, int width, Real* ddst) {
    int rc = ensure_stage();
    if(rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(stage.p, src, sizeof(double) * static_cast<size_t>(plan.n) * width, cudaMemcpyHostToDevice, stream));
    const int nb = static_cast<int>((plan.n + 255) / 256);
    lbm::k_unpack_aos<Real><<<nb, 256, 0, stream>>>(stage.p, d_ref2dev.p, plan.n, width, ddst, plan.npad);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    h2d_bytes += static_cast<int64_t>(sizeof(double)) * plan.n * width;
    return LBM_B200_OK;
  }

  int gather_all(Real* fold_out, Real* mom_out) {
    lbm::DevParams<Real> p = params(cur, cur ^ 1, nullptr);
    // owned cells only: ghost cells have no links of their own (their populations arrive by halo exchange)
    const int32_t nc = static_cast<int32_t>(plan.ghost_begin);
    const int nb = (nc + 127) / 128;
    if(prev_fold.p != nullptr) {
      p.A = prev_fold.p;
    }
    if(cfg.arithmetic == LBM_B200_STRICT) lbm::k_gather_all<L, Real, true><<<nb, 128, 0, stream>>>(p, nc, fold_out, mom_out);
    else lbm::k_gather_all<L, Real, false><<<nb, 128, 0, stream>>>(p, nc, fold_out, mom_out);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return LBM_B200_OK;
  }

  int get_populations(double* fo, double* foldo) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(fo != nullptr) {
      int rc = download(f[cur].p, Q, fo);
      if(rc) return rc;
    }
    if(foldo != nullptr) {
      if(prev_fold.p != nullptr) return download(prev_fold.p, Q, foldo);
      int rc = gather_all(scratch.p, nullptr);
      if(rc) return rc;
      return download(scratch.p, Q, foldo);
    }
    return LBM_B200_OK;
  }

  int set_populations(const double* fi, const double* foldi) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(foldi == nullptr) return fail(LBM_B200_EINVAL, "set_populations needs m_fold");
    CUDA_TRY(cudaStreamSynchronize(stream));
    int rc = LBM_B200_OK;
    if(fi != nullptr) { // m_f is not an input of the next step (the collision overwrites it, solver.cpp:603); kept for read-back only
      rc = upload_aos(fi, Q, f[cur].p);
      if(rc) return rc;
    }
    CUDA_TRY(prev_fold.alloc(static_cast<size_t>(plan.npad) * Q));
    CUDA_TRY(cudaMemset(prev_fold.p, 0, prev_fold.bytes()));
    rc = upload_aos(foldi, Q, prev_fold.p);
    if(rc) return rc;
    // slots nothing ever writes now keep the supplied m_fold value
    if(!plan.stale_ref.empty()) {
      std::vector<Real> hv(d_values[0].n, Real(0));
      CUDA_TRY(cudaMemcpy(hv.data(), d_values[dyn].p, hv.size() * sizeof(Real), cudaMemcpyDeviceToHost));
      for(size_t k = 0; k < plan.stale_ref.size(); ++k) {
        const int64_t ref = plan.dev2ref[plan.stale_ref[k] / Q];
        hv[k + 1]         = static_cast<Real>(foldi[ref * Q + plan.stale_ref[k] % Q]);
      }
      for(int b = 0; b < 2; ++b) CUDA_TRY(cudaMemcpy(d_values[b].p, hv.data(), hv.size() * sizeof(Real), cudaMemcpyHostToDevice));
    }
    first = true; // the next step consumes the supplied m_fold directly
    return LBM_B200_OK;
  }

  int get_vars(double* v, double* vo) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(cfg.track_vars <= 0) return fail(LBM_B200_ESTATE, "m_vars is not tracked (config.track_vars = 0)");
    if(vars_step[vcur] != t) return fail(LBM_B200_ESTATE, "m_vars of this step was not kept (track_vars interval)");
    if(v != nullptr) {
      int rc = download(vars[vcur].p, NVAR, v);
      if(rc) return rc;
    }
    if(vo != nullptr) {
      if(t == 0) {
        std::memset(vo, 0, sizeof(double) * static_cast<size_t>(plan.n) * NVAR); // solver.cpp:270
      } else {
        if(vars_step[vcur ^ 1] != t - 1) return fail(LBM_B200_ESTATE, "m_varsold of this step was not kept (track_vars interval)");
        return download(vars[vcur ^ 1].p, NVAR, vo);
      }
    }
    return LBM_B200_OK;
  }

  int get_moments(double* m) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    int rc = gather_all(nullptr, scratch.p);
    if(rc) return rc;
    return download(scratch.p, NVAR, m);
  }

  int residual(double* out, int32_t* diverged) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(cfg.track_vars <= 0) return fail(LBM_B200_ESTATE, "residual needs config.track_vars");
    if(vars_step[vcur] != t || (t > 0 && vars_step[vcur ^ 1] != t - 1))
      return fail(LBM_B200_ESTATE, "m_vars / m_varsold of this step were not kept (track_vars interval)");
    const int nb = 592; // 4 x 148 SMs
    lbm::k_residual<Real><<<nb, 256, 0, stream>>>(vars[vcur].p, vars[vcur ^ 1].p, plan.npad, plan.ghost_begin, NVAR, d_partial.p);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    std::vector<double> h(static_cast<size_t>(NVAR) * nb);
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaMemcpy(h.data(), d_partial.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    int bad = 0;
    for(int v = 0; v < NVAR; ++v) {
      double s = 0;
      for(int b = 0; b < nb; ++b) s += h[static_cast<size_t>(v) * nb + b];
      out[v] = s;
      if(std::isnan(s) || std::isinf(s)) bad = 1;
    }
    if(comm != nullptr && comm_size > 1) {
      // partitioned run: the residual of the whole domain = sum over the ranks' owned cells, one
      // ncclAllReduce of NVAR + 1 doubles (the last one carries the NaN/Inf flag).  Collective: every rank calls residual().
      double h2[NVAR + 1];
      for(int v = 0; v < NVAR; ++v) h2[v] = (std::isnan(out[v]) || std::isinf(out[v])) ? 0.0 : out[v];
      h2[NVAR] = bad;
      CUDA_TRY(cudaMemcpyAsync(d_partial.p, h2, sizeof(h2), cudaMemcpyHostToDevice, stream));
      NCCL_TRY(lbm::nccl_api().AllReduce(d_partial.p, d_partial.p, NVAR + 1, ncclFloat64, ncclSum, comm, stream));
      CUDA_TRY(cudaMemcpyAsync(h2, d_partial.p, sizeof(h2), cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      bad = h2[NVAR] > 0.0 ? 1 : 0;
      for(int v = 0; v < NVAR; ++v) out[v] = bad ? std::numeric_limits<double>::quiet_NaN() : h2[v];
    }
    if(diverged) *diverged = bad;
    return LBM_B200_OK;
  }

  int64_t owned() const override { return plan.n_owned; }

  // host-side layout planning only (tests inspect it on machines without a GPU); the views stay valid until destroy
  std::vector<double> dbg_add, dbg_wall, dbg_abb;
  std::vector<int32_t> dbg_copy, dbg_abb_cells;
  int debug_plan(lbm_b200_plan_view* v) override {
    if(!inited) {
      if(in.nghbr.empty()) return fail(LBM_B200_ESTATE, "lbm_b200_set_topology has not been called");
      lbm::PlanInput copy = in; // keep the inputs: a later lbm_b200_init must still see them
      if(!lbm::build_plan(copy, plan)) return fail(LBM_B200_EINVAL, plan.error);
    }
    std::memset(v, 0, sizeof(*v));
    v->n = plan.n; v->n_owned = plan.n_owned; v->npad = plan.npad; v->chunk = plan.CH; v->nsel = plan.L.NSEL;
    v->n_fast_chunks = plan.n_fast_chunks; v->n_fast_outer = plan.n_fast_outer; v->gen_begin = plan.gen_begin; v->n_gen = plan.n_gen;
    v->n_gen_outer = plan.n_gen_outer; v->gen_stride = plan.gen_stride; v->ghost_begin = plan.ghost_begin;
    v->n_ghost_blocks = plan.n_ghost_blocks; v->n_values_static = plan.n_values_static;
    v->ref2dev = plan.ref2dev.data(); v->tmpl = plan.tmpl.data(); v->chunk_nb = plan.chunk_nb.data(); v->codes = plan.codes.data();
    dbg_copy.clear();
    for(auto& c : plan.copytab) { dbg_copy.push_back(c.cell); dbg_copy.push_back(c.dir); }
    v->copytab = dbg_copy.data(); v->n_copy = static_cast<int64_t>(plan.copytab.size());
    dbg_add.clear();
    for(auto& a : plan.addtab) { dbg_add.push_back(a.v[0]); dbg_add.push_back(a.v[1]); dbg_add.push_back(a.v[2]); dbg_add.push_back(a.n); }
    v->addtab = dbg_add.data(); v->n_add = static_cast<int64_t>(plan.addtab.size());
    dbg_wall.clear();
    for(auto& a : plan.wall_desc) { dbg_wall.push_back(a.v[0]); dbg_wall.push_back(a.v[1]); dbg_wall.push_back(a.v[2]); dbg_wall.push_back(a.n); }
    v->wall_desc = dbg_wall.data(); v->n_wall = static_cast<int64_t>(plan.wall_desc.size());
    dbg_abb.clear(); dbg_abb_cells.clear();
    for(auto& a : plan.abb) { dbg_abb.push_back(a.p); dbg_abb_cells.push_back(a.cell); dbg_abb_cells.push_back(a.n1); dbg_abb_cells.push_back(a.n2); }
    v->abb_p = dbg_abb.data(); v->abb_cells = dbg_abb_cells.data(); v->n_abb = static_cast<int64_t>(plan.abb.size());
    v->values = plan.values.data(); v->n_values = static_cast<int64_t>(plan.values.size());
    v->stale_ref = plan.stale_ref.data(); v->n_stale = static_cast<int64_t>(plan.stale_ref.size());
    v->send_index = plan.send_index.data(); v->n_send = static_cast<int64_t>(plan.send_index.size());
    v->recv_index = plan.recv_index.data(); v->n_recv = static_cast<int64_t>(plan.recv_index.size());
    v->chunk_abb_base = plan.chunk_abb_base.data(); v->chunk_abb = plan.chunk_abb.data();
    v->n_chunk_abb_rows = static_cast<int64_t>(plan.chunk_abb.size() / (plan.CH > 0 ? plan.CH : 1));
    v->vsend_cells = plan.vsend_cells.data(); v->n_vsend = static_cast<int64_t>(plan.vsend_cells.size()); v->n_vrecv = plan.n_vrecv;
    return LBM_B200_OK;
  }

  void stats(lbm_b200_stats* st) const override {
    std::memset(st, 0, sizeof(*st));
    st->ncells        = plan.n_owned;
    st->cells_fast    = plan.n_fast_chunks * plan.CH;
    st->cells_generic = plan.n_owned - st->cells_fast;
    st->cells_ghost   = plan.n_ghost;
    st->halo_bytes    = halo_bytes;
    st->chunk_cells   = plan.CH;
    st->slots_bc      = plan.slots_bc;
    st->slots_stale   = plan.slots_stale;
    st->device_bytes  = static_cast<int64_t>(f[0].bytes() + f[1].bytes() + vars[0].bytes() + vars[1].bytes() + scratch.bytes() + d_codes.bytes()
                                             + d_chunk_nb.bytes() + d_tmpl.bytes());
    st->launches      = launches;
    st->launches_main = launches_main;
    st->bytes_per_cell_alg = 2.0 * Q * sizeof(Real);
    st->h2d_bytes = h2d_bytes;
    st->d2h_bytes = d2h_bytes;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// Reference-order pipeline (sequential.cuh) behind the same SolverBase interface: used when a configuration contains a
// wet-node wall, whose result depends on the order in which boundary conditions touch a cell.
template <class L>
struct SequentialSolver final : SolverBase {
  static constexpr int Q = L::Q, D = L::D, NV = L::D + 1, QM = L::Q - 1;
  DevBuf<double>  d_f, d_fold, d_feq, d_vars, d_varsold, d_scratch, d_partial;
  DevBuf<int32_t> d_pull;
  DevBuf<int64_t> d_nghbr;
  DevBuf<lbm::ForceEntry> d_force;
  std::vector<lbm::seq::BcDev> bcs;
  // owners of the per-BC device arrays
  std::vector<std::unique_ptr<DevBuf<int64_t>>> keep_i64;
  std::vector<std::unique_ptr<DevBuf<double>>>  keep_f64;
  std::vector<std::unique_ptr<DevBuf<int>>>     keep_i32;
  int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;

  template <class T, class Keep>
  const T* up(Keep& keep, const std::vector<T>& h, cudaError_t* err) {
    keep.emplace_back(new DevBuf<T>());
    std::vector<T> tmp = h;
    if(tmp.empty()) tmp.resize(1);
    cudaError_t e = keep.back()->upload(tmp);
    if(e != cudaSuccess) *err = e;
    return keep.back()->p;
  }

  lbm::seq::State state() const {
    lbm::seq::State s{};
    s.f = d_f.p; s.fold = d_fold.p; s.feq = d_feq.p; s.vars = d_vars.p; s.varsold = d_varsold.p;
    s.pull = d_pull.p; s.nghbr = d_nghbr.p; s.stride = QM; s.n = in.n;
    s.omega = cfg.omega; s.om1 = 1 - cfg.omega; s.omega_minus = cfg.omega_minus;
    for(int i = 0; i < 27; ++i) s.rates[i] = cfg.mrt_rates[i];
    return s;
  }
  static int blocks(int64_t n) { return static_cast<int>((n + 127) / 128); }

  int init() override {
    const lbm::LatticeRT& LR = in.L;
    const int64_t N = in.n;
    if(cfg.precision != LBM_B200_FP64) return fail(LBM_B200_EUNSUP, "wet-node wall boundary conditions run in fp64 only");
    if(!in.peers.empty() || in.n_ghost > 0) return fail(LBM_B200_EUNSUP, "wet-node wall boundary conditions are not partitioned yet");
    if(in.nghbr.empty()) return fail(LBM_B200_ESTATE, "no topology set");
    CUDA_TRY(cudaSetDevice(cfg.device));
    auto NB = [&](int64_t c, int j) -> int64_t { return in.nghbr[static_cast<size_t>(c) * QM + j]; };
    // inverse of the push table (the highest source wins, like the reference's serial loop)
    std::vector<int32_t> pull(static_cast<size_t>(N) * QM, -1);
    for(int64_t c = 0; c < N; ++c)
      for(int j = 0; j < QM; ++j) {
        const int64_t t = NB(c, j);
        if(t >= 0) pull[static_cast<size_t>(t) * QM + j] = static_cast<int32_t>(c);
      }
    std::vector<int64_t> nb64(in.nghbr.begin(), in.nghbr.end());
    CUDA_TRY(d_pull.upload(pull));
    CUDA_TRY(d_nghbr.upload(nb64));
    const size_t nq = static_cast<size_t>(N) * Q, nv = static_cast<size_t>(N) * NV;
    CUDA_TRY(d_f.alloc(nq)); CUDA_TRY(d_fold.alloc(nq)); CUDA_TRY(d_feq.alloc(nq));
    CUDA_TRY(d_vars.alloc(nv)); CUDA_TRY(d_varsold.alloc(nv)); CUDA_TRY(d_scratch.alloc(nq));
    CUDA_TRY(d_partial.alloc(static_cast<size_t>(NV) * 64));
    CUDA_TRY(cudaMemset(d_varsold.p, 0, d_varsold.bytes()));

    std::vector<char>   periodic(static_cast<size_t>(N), 0); // CellProperties::periodic, set in boundary-condition order
    std::vector<double> vars0(nv, 0.0);
    cudaError_t cerr = cudaSuccess;
    for(const lbm::BcInput& bc : in.bcs) {
      lbm::seq::BcDev b{};
      const int64_t n = static_cast<int64_t>(bc.cells.size());
      b.kind = bc.kind;
      b.n = n;
      b.cells = up<int64_t>(keep_i64, bc.cells, &cerr);
      b.normals = up<double>(keep_f64, bc.normals, &cerr);
      for(int d = 0; d < 3; ++d) b.value[d] = bc.value[d];
      b.has_pressure = !std::isnan(bc.pressure);
      b.pressure = b.has_pressure ? bc.pressure : 0.0;
      b.has_velocity = bc.has_velocity ? 1 : 0;
      std::vector<int64_t> first(static_cast<size_t>(n), 1);
      for(int64_t k = 0; k < n; ++k)
        for(int64_t j = 0; j < k; ++j)
          if(bc.cells[j] == bc.cells[k]) { first[k] = 0; break; }
      b.first_of_cell = up<int64_t>(keep_i64, first, &cerr);
      if(bc.kind == lbm::BC_WALL_BB_TANGENTIAL) {
        if(D != 2) return fail(LBM_B200_EINVAL, "tangential wall velocity is implemented for 2D only (reference: bnd_wall.h:52-54)");
        std::vector<double> wv(static_cast<size_t>(n) * Q, 0.0);
        for(int64_t k = 0; k < n; ++k) {
          const double* nrm = &bc.normals[k * D];
          for(int id = 0; id < Q; ++id) {
            if(!lbm::in_direction(LR, nrm, id)) continue;
            const int    inside = LR.opp[id];
            const double tdot = nrm[1] * LR.c[inside][0] + nrm[0] * LR.c[inside][1];
            const double ndot = nrm[0] * LR.c[inside][0] + nrm[1] * LR.c[inside][1];
            const double nn   = std::sqrt(double(LR.c[inside][0] * LR.c[inside][0] + LR.c[inside][1] * LR.c[inside][1]));
            const bool parallel = std::abs(std::acos(ndot / nn) - 3.14159265358979323846) < 10 * lbm::kEps;
            if(!parallel) wv[k * Q + inside] = bc.tangential * tdot;
          }
        }
        b.wallval = up<double>(keep_f64, wv, &cerr);
      } else if(bc.kind == lbm::BC_DIRICHLET_BB) {
        for(int64_t k = 0; k < n; ++k)
          for(int d = 0; d < D; ++d) vars0[bc.cells[k] * NV + d] = bc.value[d]; // initCnd, bnd_dirichlet.h:44-50
      } else if(bc.kind == lbm::BC_PRESSURE) {
        std::vector<int64_t> n1(static_cast<size_t>(n)), n2(static_cast<size_t>(n));
        std::vector<char> seen(static_cast<size_t>(N), 0);
        for(int64_t k = 0; k < n; ++k) {
          const double* nrm = &bc.normals[k * D];
          int ins = -1;
          for(int d = 0; d < D && ins < 0; ++d) {
            if(nrm[d] < 0) ins = 2 * d + 1;
            else if(nrm[d] > 0) ins = 2 * d;
          }
          if(ins < 0) return fail(LBM_B200_EINVAL, "pressure boundary: zero normal");
          n1[k] = NB(bc.cells[k], ins);
          n2[k] = n1[k] >= 0 ? NB(n1[k], ins) : -1;
          if(n1[k] < 0 || n2[k] < 0) return fail(LBM_B200_EINVAL, "pressure boundary: cell without two inward neighbours");
          if(seen[n1[k]] || seen[n2[k]])
            return fail(LBM_B200_EUNSUP, "pressure boundary: inward neighbour is an earlier entry of the same surface (order-dependent in the reference)");
          seen[bc.cells[k]] = 1;
        }
        b.n1 = up<int64_t>(keep_i64, n1, &cerr);
        b.n2 = up<int64_t>(keep_i64, n2, &cerr);
      } else if(bc.kind == lbm::BC_PERIODIC) {
        if(in.center.empty()) return fail(LBM_B200_EINVAL, "periodic boundary condition needs set_geometry");
        std::vector<int64_t> link(static_cast<size_t>(n) * Q, -1);
        std::vector<int>     ldist(static_cast<size_t>(n) * Q, 0), nset(static_cast<size_t>(n), 0);
        for(int64_t k = 0; k < n; ++k) {
          int     setd[27], ns = 0;
          int64_t links[27];
          std::string err;
          if(!lbm::periodic_links(in, bc, k, setd, links, &ns, &err)) return fail(LBM_B200_EINVAL, err);
          nset[k] = ns;
          for(int id = 0; id < ns; ++id) { link[k * Q + id] = links[id]; ldist[k * Q + id] = setd[id]; }
        }
        b.link = up<int64_t>(keep_i64, link, &cerr);
        b.linkdist = up<int>(keep_i32, ldist, &cerr);
        b.nset = up<int>(keep_i32, nset, &cerr);
        for(int64_t c : bc.cells) periodic[c] = 1;     // surfA.setProperty(periodic), bnd.h:219
        for(int64_t c : bc.connected) periodic[c] = 1; // surfB
      } else if(bc.kind >= lbm::BC_WALL_EQ) {
        if(bc.kind == lbm::BC_WALL_NEBB && !(D == 2 && Q == 9)) return fail(LBM_B200_EINVAL, "Not implemented for this distribution!"); // bnd_wall.h:325-328
        // LBMBnd_wallWetnode, bnd_wetnode.h:24-63
        std::vector<int>    lim_n(static_cast<size_t>(n), 0), lim_dist(static_cast<size_t>(n) * Q, 0);
        std::vector<double> lim_const(static_cast<size_t>(n) * Q, 0.0);
        std::vector<int64_t> c2b(static_cast<size_t>(n)), ext(static_cast<size_t>(n), -1);
        for(int64_t k = 0; k < n; ++k) {
          const int64_t c   = bc.cells[k];
          const double* nrm = &bc.normals[k * D];
          double*       cst = &lim_const[k * Q];
          int           m   = 0;
          for(int dir = 0; dir < QM; ++dir) {
            const int  op  = LR.opp[dir];
            const bool per = periodic[c] != 0;
            double dot = 0;
            for(int d = 0; d < D; ++d) dot += nrm[d] * LR.c[dir][d];
            const bool has_opp = per || NB(c, op) != -1;
            if(dot >= lbm::kEps && has_opp) {
              lim_dist[k * Q + m++] = dir;
              cst[dir] = 2;
            } else if(std::abs(dot) <= lbm::kEps && has_opp) {
              lim_dist[k * Q + m++] = dir;
              cst[dir] = (NB(c, dir) == -1 && !per) ? 2 : 1;
            }
          }
          int sumC = 0;
          for(int i = 0; i < Q; ++i) sumC = static_cast<int>(sumC + cst[i]);
          if(sumC != Q - 1) m = 0;
          else { lim_dist[k * Q + m++] = Q - 1; cst[Q - 1] = 1; }
          lim_n[k] = m;
          int64_t idx = k;
          for(int64_t j = n - 1; j > k; --j)
            if(bc.cells[j] == c) { idx = j; break; }
          c2b[k] = idx;
          if(bc.kind == lbm::BC_WALL_NEEM) {
            int ex = -1;
            for(int d = 0; d < D && ex < 0; ++d) {
              if(nrm[d] < 0) ex = 2 * d + 1;
              else if(nrm[d] > 0) ex = 2 * d;
            }
            ext[k] = ex < 0 ? -1 : NB(c, ex);
            if(ext[k] < 0) return fail(LBM_B200_EINVAL, "No valid extrapolation cellId"); // bnd_wall.h:243-246
          }
        }
        if(bc.kind == lbm::BC_WALL_NEEM) {
          std::vector<char> mine(static_cast<size_t>(N), 0);
          for(int64_t c : bc.cells) mine[c] = 1;
          for(int64_t e : ext)
            if(mine[e]) return fail(LBM_B200_EUNSUP, "NEEM wall: extrapolation cell lies on the same surface (order-dependent in the reference)");
        }
        b.lim_n = up<int>(keep_i32, lim_n, &cerr);
        b.lim_dist = up<int>(keep_i32, lim_dist, &cerr);
        b.lim_const = up<double>(keep_f64, lim_const, &cerr);
        b.cell2bnd = up<int64_t>(keep_i64, c2b, &cerr);
        b.ext = up<int64_t>(keep_i64, ext, &cerr);
      }
      bcs.push_back(b);
    }
    if(cerr != cudaSuccess) return fail(LBM_B200_ECUDA, cudaGetErrorString(cerr));
    // forcing pairs, solver.cpp:651-693
    if(in.forcing) {
      if(in.center.empty()) return fail(LBM_B200_EINVAL, "forcing needs set_geometry");
      std::vector<lbm::ForceEntry> fe;
      for(int64_t a : in.inlet) {
        const int64_t val = NB(a, 1);
        if(val < 0) return fail(LBM_B200_EINVAL, "forcing: inlet cell without +x neighbour");
        for(int64_t o : in.outlet)
          if(std::abs(in.center[val * D + 1] - in.center[o * D + 1]) < lbm::kEps) fe.push_back({static_cast<int32_t>(o), static_cast<int32_t>(val), 1.0});
      }
      for(int64_t o : in.outlet) {
        const int64_t val = NB(o, 0);
        if(val < 0) return fail(LBM_B200_EINVAL, "forcing: outlet cell without -x neighbour");
        for(int64_t a : in.inlet)
          if(std::abs(in.center[a * D + 1] - in.center[val * D + 1]) < lbm::kEps) fe.push_back({static_cast<int32_t>(a), static_cast<int32_t>(val), 1.0 + in.gradient});
      }
      std::vector<char> target(static_cast<size_t>(N), 0);
      for(auto& e : fe) {
        if(target[e.target]) return fail(LBM_B200_EUNSUP, "forcing: a cell is forced twice (order-dependent in the reference)");
        target[e.target] = 1;
      }
      for(auto& e : fe)
        if(target[e.val]) return fail(LBM_B200_EUNSUP, "forcing: value cell is itself a forced cell (order-dependent in the reference)");
      CUDA_TRY(d_force.upload(fe));
    }
    CUDA_TRY(cudaMemcpy(d_vars.p, vars0.data(), nv * sizeof(double), cudaMemcpyHostToDevice));
    lbm::seq::k_init<L><<<blocks(N), 128, 0, stream>>>(state());
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<int32_t>().swap(in.nghbr);
    t = 0;
    inited = true;
    return LBM_B200_OK;
  }

  template <int COLL>
  void launch_cell(const lbm::seq::State& s) { lbm::seq::k_cell<L, COLL><<<blocks(s.n), 128, 0, stream>>>(s); }

  int step(int64_t n, float* ms_total, float* ms_main) override {
    if(!inited) return fail(LBM_B200_ESTATE, "lbm_b200_step before lbm_b200_init");
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if(ms_total != nullptr) {
      CUDA_TRY(cudaEventCreate(&e0));
      CUDA_TRY(cudaEventCreate(&e1));
      CUDA_TRY(cudaEventRecord(e0, stream));
    }
    const lbm::seq::State s = state();
    for(int64_t it = 0; it < n; ++it) {
      CUDA_TRY(cudaMemcpyAsync(d_varsold.p, d_vars.p, d_vars.bytes(), cudaMemcpyDeviceToDevice, stream)); // currToOldVars
      if(cfg.collision == LBM_B200_TRT) launch_cell<lbm::COLL_TRT>(s);
      else if(cfg.collision == LBM_B200_MRT) launch_cell<lbm::COLL_MRT>(s);
      else launch_cell<lbm::COLL_BGK>(s);
      ++launches;
      if(d_force.n > 0) {
        lbm::seq::k_forcing<L><<<blocks(static_cast<int64_t>(d_force.n)), 128, 0, stream>>>(s, d_force.p, static_cast<int>(d_force.n));
        ++launches;
      }
      for(const auto& b : bcs)
        if((b.kind == lbm::BC_PRESSURE || b.kind == lbm::BC_PERIODIC) && b.n > 0) {
          lbm::seq::k_pre_apply<L><<<blocks(b.n), 128, 0, stream>>>(s, b);
          ++launches;
        }
      lbm::seq::k_stream<L><<<blocks(s.n), 128, 0, stream>>>(s);
      ++launches;
      for(const auto& b : bcs) {
        if(b.n == 0 || b.kind == lbm::BC_PERIODIC) continue;
        int nphase = 1, first_phase = 0;
        if(b.kind == lbm::BC_PRESSURE) first_phase = 1, nphase = 1;
        if(b.kind == lbm::BC_WALL_NEEM) nphase = 4;
        if(b.kind == lbm::BC_WALL_NEBB) nphase = b.has_velocity ? 3 : 4;
        for(int ph = first_phase; ph < first_phase + nphase; ++ph) {
          if(b.kind == lbm::BC_WALL_NEBB) {
            if constexpr(D == 2 && Q == 9) lbm::seq::k_nebb<<<blocks(b.n), 128, 0, stream>>>(s, b, ph);
          } else {
            lbm::seq::k_apply<L><<<blocks(b.n), 128, 0, stream>>>(s, b, ph);
          }
          ++launches;
        }
      }
      ++t;
    }
    CUDA_TRY(cudaGetLastError());
    if(ms_total != nullptr) {
      CUDA_TRY(cudaEventRecord(e1, stream));
      CUDA_TRY(cudaEventSynchronize(e1));
      CUDA_TRY(cudaEventElapsedTime(ms_total, e0, e1));
      if(ms_main) *ms_main = *ms_total;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    return LBM_B200_OK;
  }

  int sync() override {
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LBM_B200_OK;
  }
  int d2h(const double* src, double* dst, size_t count) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    d2h_bytes += static_cast<int64_t>(count * sizeof(double));
    return LBM_B200_OK;
  }
  int get_populations(double* fo, double* foldo) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    const size_t nq = static_cast<size_t>(in.n) * Q;
    if(fo != nullptr) { int rc = d2h(d_f.p, fo, nq); if(rc) return rc; }
    if(foldo != nullptr) { int rc = d2h(d_fold.p, foldo, nq); if(rc) return rc; }
    return LBM_B200_OK;
  }
  int set_populations(const double* fi, const double* foldi) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(foldi == nullptr) return fail(LBM_B200_EINVAL, "set_populations needs m_fold");
    const size_t nq = static_cast<size_t>(in.n) * Q;
    if(fi != nullptr) CUDA_TRY(cudaMemcpyAsync(d_f.p, fi, nq * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_fold.p, foldi, nq * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    h2d_bytes += static_cast<int64_t>((fi != nullptr ? 2 : 1) * nq * sizeof(double));
    return LBM_B200_OK;
  }
  int get_vars(double* v, double* vo) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    const size_t nv = static_cast<size_t>(in.n) * NV;
    if(v != nullptr) { int rc = d2h(d_vars.p, v, nv); if(rc) return rc; }
    if(vo != nullptr) { int rc = d2h(d_varsold.p, vo, nv); if(rc) return rc; }
    return LBM_B200_OK;
  }
  int get_moments(double* m) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    lbm::seq::k_moments<L><<<blocks(in.n), 128, 0, stream>>>(state(), d_scratch.p);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return d2h(d_scratch.p, m, static_cast<size_t>(in.n) * NV);
  }
  int residual(double* out, int32_t* diverged) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    const int nb = 64;
    lbm::seq::k_residual_aos<<<nb, 256, 0, stream>>>(d_vars.p, d_varsold.p, in.n, NV, d_partial.p);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    std::vector<double> h(static_cast<size_t>(NV) * nb);
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaMemcpy(h.data(), d_partial.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    int bad = 0;
    for(int v = 0; v < NV; ++v) {
      double sum = 0;
      for(int b = 0; b < nb; ++b) sum += h[static_cast<size_t>(v) * nb + b];
      out[v] = sum;
      if(std::isnan(sum) || std::isinf(sum)) bad = 1;
    }
    if(diverged) *diverged = bad;
    return LBM_B200_OK;
  }
  int64_t owned() const override { return in.n; }
  void stats(lbm_b200_stats* st) const override {
    std::memset(st, 0, sizeof(*st));
    st->ncells = in.n;
    st->cells_generic = in.n; // none on the fused chunk path
    st->device_bytes = static_cast<int64_t>(d_f.bytes() * 4 + d_vars.bytes() * 2);
    st->launches = launches;
    st->launches_main = 0;
    st->bytes_per_cell_alg = 2.0 * Q * sizeof(double);
    st->h2d_bytes = h2d_bytes;
    st->d2h_bytes = d2h_bytes;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// Poisson equation types (poisson.cuh) behind the same SolverBase interface: one variable per cell, the reference's passes in
// the reference's order, Dirichlet / Neumann NEEM conditions as phase kernels in LBMBndManager order.
struct PoissonSolver final : SolverBase {
  lbm::poisson::Lat lat{};
  DevBuf<double>  d_f, d_fold, d_feq, d_vars, d_varsold, d_scratch, d_partial;
  DevBuf<int32_t> d_pull;
  DevBuf<int64_t> d_nghbr;
  std::vector<lbm::poisson::Bc> bcs;
  std::vector<std::unique_ptr<DevBuf<int64_t>>> keep_i64;
  std::vector<std::unique_ptr<DevBuf<double>>>  keep_f64;
  int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;

  lbm::poisson::HostSetup setup;

  lbm::poisson::State state() const {
    lbm::poisson::State s{};
    s.f = d_f.p; s.fold = d_fold.p; s.feq = d_feq.p; s.vars = d_vars.p; s.varsold = d_varsold.p;
    s.pull = d_pull.p; s.nghbr = d_nghbr.p; s.n = in.n;
    s.omega = setup.omega;
    s.om1   = setup.om1;
    s.dt_diff = setup.dt_diff;
    s.rate2   = setup.rate2;
    return s;
  }
  static int blocks(int64_t n) { return static_cast<int>((n + 127) / 128); }

  template <class T, class Keep>
  T* up(Keep& keep, const std::vector<T>& h, cudaError_t* err) {
    keep.emplace_back(new DevBuf<T>());
    std::vector<T> tmp = h;
    if(tmp.empty()) tmp.resize(1);
    cudaError_t e = keep.back()->upload(tmp);
    if(e != cudaSuccess) *err = e;
    return keep.back()->p;
  }

  int init() override {
    const int     Q = in.L.Q;
    const int64_t N = in.n;
    if(cfg.precision != LBM_B200_FP64) return fail(LBM_B200_EUNSUP, "the Poisson equation types run in fp64 only");
    if(cfg.collision != LBM_B200_BGK) return fail(LBM_B200_EINVAL, "Invalid equation configuration!");
    if(!in.peers.empty() || in.n_ghost > 0) return fail(LBM_B200_EUNSUP, "the Poisson equation types are not partitioned");
    if(in.forcing) return fail(LBM_B200_EINVAL, "forcing is a Navier-Stokes feature");
    if(in.nghbr.empty()) return fail(LBM_B200_ESTATE, "no topology set");
    // everything derived from the caller's tables (lattice constants, inverse push table, extrapolation cells, order-hazard checks,
    // initial potential): host code shared with the CPU harness of the tests
    if(!lbm::poisson::prepare(in, cfg.omega, setup)) return fail(setup.code, setup.error);
    lat = setup.lat;
    CUDA_TRY(cudaSetDevice(cfg.device));
    std::vector<int64_t> nb64(in.nghbr.begin(), in.nghbr.end());
    CUDA_TRY(d_pull.upload(setup.pull));
    CUDA_TRY(d_nghbr.upload(nb64));
    const size_t nq = static_cast<size_t>(N) * Q;
    CUDA_TRY(d_f.alloc(nq)); CUDA_TRY(d_fold.alloc(nq)); CUDA_TRY(d_feq.alloc(nq));
    CUDA_TRY(d_vars.alloc(N)); CUDA_TRY(d_varsold.alloc(N)); CUDA_TRY(d_scratch.alloc(N));
    CUDA_TRY(d_partial.alloc(64));
    CUDA_TRY(cudaMemset(d_varsold.p, 0, d_varsold.bytes()));
    cudaError_t cerr = cudaSuccess;
    for(const lbm::poisson::HostBc& hb : setup.bcs) {
      lbm::poisson::Bc b{};
      b.neumann = hb.neumann;
      b.n = static_cast<int64_t>(hb.cells.size());
      b.cells = up<int64_t>(keep_i64, hb.cells, &cerr);
      b.ext = up<int64_t>(keep_i64, hb.ext, &cerr);
      b.ext2 = up<int64_t>(keep_i64, hb.ext2, &cerr);
      b.values = up<double>(keep_f64, hb.values, &cerr);
      b.grad = hb.grad;
      bcs.push_back(b);
    }
    if(cerr != cudaSuccess) return fail(LBM_B200_ECUDA, cudaGetErrorString(cerr));
    CUDA_TRY(cudaMemcpy(d_vars.p, setup.vars0.data(), setup.vars0.size() * sizeof(double), cudaMemcpyHostToDevice));
    lbm::poisson::k_init<<<blocks(N), 128, 0, stream>>>(state(), lat);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<int32_t>().swap(in.nghbr);
    std::vector<int32_t>().swap(setup.pull);
    t = 0;
    inited = true;
    return LBM_B200_OK;
  }

  int step(int64_t n, float* ms_total, float* ms_main) override {
    if(!inited) return fail(LBM_B200_ESTATE, "lbm_b200_step before lbm_b200_init");
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if(ms_total != nullptr) {
      CUDA_TRY(cudaEventCreate(&e0));
      CUDA_TRY(cudaEventCreate(&e1));
      CUDA_TRY(cudaEventRecord(e0, stream));
    }
    const lbm::poisson::State s = state();
    for(int64_t it = 0; it < n; ++it) {
      CUDA_TRY(cudaMemcpyAsync(d_varsold.p, d_vars.p, d_vars.bytes(), cudaMemcpyDeviceToDevice, stream)); // currToOldVars
      lbm::poisson::k_cell<<<blocks(s.n), 128, 0, stream>>>(s, lat);
      lbm::poisson::k_stream<<<blocks(s.n), 128, 0, stream>>>(s, lat);
      launches += 2;
      for(const auto& b : bcs) {
        if(b.n == 0) continue;
        if(b.neumann) {
          lbm::poisson::k_neumann_value<<<blocks(b.n), 128, 0, stream>>>(s, lat, b);
          ++launches;
        }
        lbm::poisson::k_ext_potential<<<blocks(b.n), 128, 0, stream>>>(s, lat, b);
        lbm::poisson::k_dirichlet<<<blocks(b.n), 128, 0, stream>>>(s, lat, b);
        launches += 2;
      }
      ++t;
    }
    CUDA_TRY(cudaGetLastError());
    if(ms_total != nullptr) {
      CUDA_TRY(cudaEventRecord(e1, stream));
      CUDA_TRY(cudaEventSynchronize(e1));
      CUDA_TRY(cudaEventElapsedTime(ms_total, e0, e1));
      if(ms_main) *ms_main = *ms_total;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    return LBM_B200_OK;
  }

  int sync() override {
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LBM_B200_OK;
  }
  int d2h(const double* src, double* dst, size_t count) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    d2h_bytes += static_cast<int64_t>(count * sizeof(double));
    return LBM_B200_OK;
  }
  int get_populations(double* fo, double* foldo) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    const size_t nq = static_cast<size_t>(in.n) * in.L.Q;
    if(fo != nullptr) { int rc = d2h(d_f.p, fo, nq); if(rc) return rc; }
    if(foldo != nullptr) { int rc = d2h(d_fold.p, foldo, nq); if(rc) return rc; }
    return LBM_B200_OK;
  }
  int set_populations(const double* fi, const double* foldi) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(foldi == nullptr) return fail(LBM_B200_EINVAL, "set_populations needs m_fold");
    const size_t nq = static_cast<size_t>(in.n) * in.L.Q;
    if(fi != nullptr) CUDA_TRY(cudaMemcpyAsync(d_f.p, fi, nq * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_fold.p, foldi, nq * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    h2d_bytes += static_cast<int64_t>((fi != nullptr ? 2 : 1) * nq * sizeof(double));
    return LBM_B200_OK;
  }
  int get_vars(double* v, double* vo) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(v != nullptr) { int rc = d2h(d_vars.p, v, static_cast<size_t>(in.n)); if(rc) return rc; }
    if(vo != nullptr) { int rc = d2h(d_varsold.p, vo, static_cast<size_t>(in.n)); if(rc) return rc; }
    return LBM_B200_OK;
  }
  int get_moments(double* m) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    lbm::poisson::k_potential<<<blocks(in.n), 128, 0, stream>>>(state(), lat, d_scratch.p);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return d2h(d_scratch.p, m, static_cast<size_t>(in.n));
  }
  int residual(double* out, int32_t* diverged) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    const int nb = 64;
    lbm::poisson::k_residual<<<nb, 256, 0, stream>>>(d_vars.p, d_varsold.p, in.n, d_partial.p);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    double h[64];
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaMemcpy(h, d_partial.p, sizeof(h), cudaMemcpyDeviceToHost));
    double sum = 0;
    for(int b = 0; b < nb; ++b) sum += h[b];
    out[0] = sum;
    if(diverged) *diverged = (std::isnan(sum) || std::isinf(sum)) ? 1 : 0;
    return LBM_B200_OK;
  }
  int64_t owned() const override { return in.n; }
  void stats(lbm_b200_stats* st) const override {
    std::memset(st, 0, sizeof(*st));
    st->ncells = in.n;
    st->cells_generic = in.n;
    st->device_bytes = static_cast<int64_t>(d_f.bytes() * 3 + d_vars.bytes() * 3);
    st->launches = launches;
    st->bytes_per_cell_alg = 2.0 * in.L.Q * sizeof(double);
    st->h2d_bytes = h2d_bytes;
    st->d2h_bytes = d2h_bytes;
  }
};

SolverBase* make_sequential(const lbm_b200_config& c) {
  if(c.ndim == 2 && c.ndist == 9) return new SequentialSolver<lbm::Lattice<2, 9>>();
  if(c.ndim == 3 && c.ndist == 19) return new SequentialSolver<lbm::Lattice<3, 19>>();
  if(c.ndim == 3 && c.ndist == 27) return new SequentialSolver<lbm::Lattice<3, 27>>();
  return nullptr;
}

SolverBase* make_solver(const lbm_b200_config& c) {
  const bool dbl = c.precision == LBM_B200_FP64;
  if((c.ndim == 1 && c.ndist == 3) || (c.ndim == 2 && c.ndist == 5)) return new PoissonSolver(); // Poisson-only lattices
#ifdef LBM_EXPERIMENT_D3Q19_F64
  // tuning builds: one instantiation only, to keep compile times short
  if(c.ndim == 3 && c.ndist == 19 && dbl) return new Solver<lbm::Lattice<3, 19>, double>();
  return nullptr;
#endif
#ifdef LBM_EXPERIMENT_D3Q27_F64
  if(c.ndim == 3 && c.ndist == 27 && dbl) return new Solver<lbm::Lattice<3, 27>, double>();
  return nullptr;
#endif
  if(c.ndim == 2 && c.ndist == 9) return dbl ? static_cast<SolverBase*>(new Solver<lbm::Lattice<2, 9>, double>()) : new Solver<lbm::Lattice<2, 9>, float>();
  if(c.ndim == 3 && c.ndist == 19) return dbl ? static_cast<SolverBase*>(new Solver<lbm::Lattice<3, 19>, double>()) : new Solver<lbm::Lattice<3, 19>, float>();
  if(c.ndim == 3 && c.ndist == 27) return dbl ? static_cast<SolverBase*>(new Solver<lbm::Lattice<3, 27>, double>()) : new Solver<lbm::Lattice<3, 27>, float>();
  return nullptr;
}

} // namespace

struct lbm_b200_solver {
  std::unique_ptr<SolverBase> impl;
};
struct lbm_b200_partition {
  lbm::Partition p;
};

extern "C" {

void lbm_b200_default_config(lbm_b200_config* cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = LBM_B200_ABI_VERSION;
  cfg->ndim        = 2;
  cfg->ndist       = 9;
  cfg->precision   = LBM_B200_FP64;
  cfg->collision   = LBM_B200_BGK;
  cfg->arithmetic  = LBM_B200_STRICT;
  cfg->device      = 0;
  cfg->track_vars  = 1;
  cfg->omega       = 1.0;
  cfg->omega_minus = 1.0;
  for(double& r : cfg->mrt_rates) r = 1.0;
}

int lbm_b200_create(const lbm_b200_config* cfg, int64_t ncells, lbm_b200_solver** out) {
  if(cfg == nullptr || out == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  *out = nullptr;
  if(cfg->abi_version != LBM_B200_ABI_VERSION) return fail(LBM_B200_EINVAL, "ABI version mismatch");
  if(ncells <= 0) return fail(LBM_B200_EINVAL, "ncells must be positive");
  if(cfg->precision != LBM_B200_FP64 && cfg->precision != LBM_B200_FP32) return fail(LBM_B200_EINVAL, "unknown precision");
  if(cfg->collision < LBM_B200_BGK || cfg->collision > LBM_B200_MRT) return fail(LBM_B200_EINVAL, "Invalid equation configuration!"); // constants.h:75
  if(cfg->arithmetic != LBM_B200_STRICT && cfg->arithmetic != LBM_B200_FAST) return fail(LBM_B200_EINVAL, "unknown arithmetic policy");
  if(!(cfg->omega > 0.0) || !(cfg->omega < 2.0)) return fail(LBM_B200_EINVAL, "omega must be in (0,2)");
  SolverBase* s = make_solver(*cfg);
  if(s == nullptr) return fail(LBM_B200_EINVAL, "Unsupported model"); // solverExe.h:90
  s->cfg = *cfg;
  if(!lbm::lattice_rt(cfg->ndim, cfg->ndist, &s->in.L)) {
    delete s;
    return fail(LBM_B200_EINVAL, "Unsupported model");
  }
  s->in.n = ncells;
  // no silent CPU path: a missing device is an error right here.  device == -1 creates an INSPECTION-ONLY handle: it accepts
  // the set-up calls and lbm_b200_debug_plan (host-side layout planning, no arithmetic), and refuses init / step / read-back.
  int ndev = 0;
  cudaError_t e = cfg->device == -1 ? cudaSuccess : cudaGetDeviceCount(&ndev);
  if(cfg->device < -1 || (cfg->device >= 0 && (e != cudaSuccess || ndev <= cfg->device))) {
    delete s;
    return fail(LBM_B200_ECUDA, std::string("no CUDA device ") + std::to_string(cfg->device) + ": " + cudaGetErrorString(e));
  }
  *out = new lbm_b200_solver{std::unique_ptr<SolverBase>(s)};
  return LBM_B200_OK;
}

void lbm_b200_destroy(lbm_b200_solver* s) {
  if(s != nullptr && s->impl && s->impl->comm != nullptr) lbm::nccl_api().CommDestroy(s->impl->comm);
  delete s;
}

#define CHECK_HANDLE(s)                                               \
  if((s) == nullptr || !(s)->impl) return fail(LBM_B200_EINVAL, "null solver handle")
#define CHECK_NOT_INITED(s) \
  if((s)->impl->inited) return fail(LBM_B200_ESTATE, "topology and boundary conditions are frozen after lbm_b200_init")

int lbm_b200_set_topology(lbm_b200_solver* s, const int64_t* nghbr, int32_t stride) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(nghbr == nullptr || stride < s->impl->in.L.Q - 1) return fail(LBM_B200_EINVAL, "bad neighbour table");
  auto&         in = s->impl->in;
  const int     QM = in.L.Q - 1;
  const int64_t N  = in.n;
  in.nghbr.resize(static_cast<size_t>(N) * QM);
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for(int64_t c = 0; c < N; ++c)
    for(int j = 0; j < QM; ++j) {
      const int64_t t = nghbr[c * stride + j];
      if(t < -1 || t >= N) bad |= 1;
      in.nghbr[static_cast<size_t>(c) * QM + j] = static_cast<int32_t>(t);
    }
  if(bad) {
    std::vector<int32_t>().swap(in.nghbr);
    return fail(LBM_B200_EINVAL, "neighbour id out of range");
  }
  in.nghbr_wide.clear();
  if(in.L.D == 2 && in.L.Q == 5 && stride >= 8) {
    in.nghbr_wide.resize(static_cast<size_t>(N) * 8);
    for(int64_t c = 0; c < N; ++c)
      for(int j = 0; j < 8; ++j) {
        const int64_t t = nghbr[c * stride + j];
        if(t < -1 || t >= N) return fail(LBM_B200_EINVAL, "neighbour id out of range");
        in.nghbr_wide[static_cast<size_t>(c) * 8 + j] = static_cast<int32_t>(t);
      }
  }
  return LBM_B200_OK;
}

int lbm_b200_set_geometry(lbm_b200_solver* s, const double* center, const double* bbmin, const double* bbmax, double cell_length) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(center == nullptr || bbmin == nullptr || bbmax == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  auto& in = s->impl->in;
  in.center.assign(center, center + static_cast<size_t>(in.n) * in.L.D);
  for(int d = 0; d < in.L.D; ++d) {
    in.bbmin[d] = bbmin[d];
    in.bbmax[d] = bbmax[d];
  }
  in.cell_length = cell_length;
  return LBM_B200_OK;
}

static int add_bc(lbm_b200_solver* s, lbm::BcInput& bc, const int64_t* cells, const double* normals, int64_t n) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(n < 0 || (n > 0 && (cells == nullptr || normals == nullptr))) return fail(LBM_B200_EINVAL, "bad cell list");
  auto& in = s->impl->in;
  for(int64_t k = 0; k < n; ++k)
    if(cells[k] < 0 || cells[k] >= in.n) return fail(LBM_B200_EINVAL, "boundary cell id out of range");
  bc.cells.assign(cells, cells + n);
  bc.normals.assign(normals, normals + n * in.L.D);
  in.bcs.push_back(std::move(bc));
  return LBM_B200_OK;
}

int lbm_b200_add_wall_bb(lbm_b200_solver* s, const int64_t* cells, const double* normals, int64_t n, double tangential) {
  lbm::BcInput bc;
  // bnd.h:232-240: |tangentialVelocity| > eps selects the moving-wall instantiation
  bc.kind       = std::abs(tangential) > lbm::kEps ? lbm::BC_WALL_BB_TANGENTIAL : lbm::BC_WALL_BB;
  bc.tangential = tangential;
  return add_bc(s, bc, cells, normals, n);
}

int lbm_b200_add_dirichlet_bb(lbm_b200_solver* s, const int64_t* cells, const double* normals, int64_t n, const double* value) {
  CHECK_HANDLE(s);
  if(value == nullptr) return fail(LBM_B200_EINVAL, "null value");
  lbm::BcInput bc;
  bc.kind = lbm::BC_DIRICHLET_BB;
  for(int d = 0; d < s->impl->in.L.D; ++d) bc.value[d] = value[d];
  return add_bc(s, bc, cells, normals, n);
}

int lbm_b200_add_wall_wetnode(lbm_b200_solver* s, int32_t model, const int64_t* cells, const double* normals, int64_t n,
                              int32_t has_velocity, const double* velocity) {
  CHECK_HANDLE(s);
  if(model < LBM_B200_WALL_EQUILIBRIUM || model > LBM_B200_WALL_NEBB) return fail(LBM_B200_EINVAL, "Invalid wall boundary model");
  if(has_velocity && velocity == nullptr) return fail(LBM_B200_EINVAL, "null velocity");
  lbm::BcInput bc;
  bc.kind = model == LBM_B200_WALL_EQUILIBRIUM ? lbm::BC_WALL_EQ : (model == LBM_B200_WALL_NEEM ? lbm::BC_WALL_NEEM : lbm::BC_WALL_NEBB);
  bc.has_velocity = has_velocity != 0;
  for(int d = 0; d < s->impl->in.L.D; ++d) bc.value[d] = has_velocity ? velocity[d] : 0.0;
  return add_bc(s, bc, cells, normals, n);
}

int lbm_b200_set_poisson(lbm_b200_solver* s, double dt, double rate) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(!(dt > 0.0) || std::isnan(rate)) return fail(LBM_B200_EINVAL, "bad Poisson parameters");
  auto& in = s->impl->in;
  if(!((in.L.D == 1 && in.L.Q == 3) || (in.L.D == 2 && (in.L.Q == 5 || in.L.Q == 9)))) return fail(LBM_B200_EINVAL, "Unsupported model");
  in.poisson      = true;
  in.poisson_dt   = dt;
  in.poisson_rate = rate;
  return LBM_B200_OK;
}

int lbm_b200_add_poisson_neem(lbm_b200_solver* s, int32_t neumann, const int64_t* cells, const double* normals, int64_t n,
                              const double* values, double gradient) {
  CHECK_HANDLE(s);
  if(n > 0 && values == nullptr) return fail(LBM_B200_EINVAL, "null values");
  lbm::BcInput bc;
  bc.kind = neumann ? lbm::BC_POISSON_NEUMANN : lbm::BC_POISSON_DIRICHLET;
  bc.grad = gradient;
  if(n > 0) bc.values.assign(values, values + n);
  return add_bc(s, bc, cells, normals, n);
}

int lbm_b200_add_pressure(lbm_b200_solver* s, const int64_t* cells, const double* normals, int64_t n, double pressure) {
  lbm::BcInput bc;
  bc.kind     = lbm::BC_PRESSURE;
  bc.pressure = pressure;
  return add_bc(s, bc, cells, normals, n);
}

int lbm_b200_add_periodic(lbm_b200_solver* s, const int64_t* cells, const double* normals, int64_t n, const int64_t* connected,
                          int64_t nconnected, double pressure) {
  CHECK_HANDLE(s);
  if(nconnected <= 0 || connected == nullptr) return fail(LBM_B200_EINVAL, "Invalid connected surface"); // bnd_periodic.h:180
  for(int64_t k = 0; k < nconnected; ++k)
    if(connected[k] < 0 || connected[k] >= s->impl->in.n) return fail(LBM_B200_EINVAL, "connected cell id out of range");
  lbm::BcInput bc;
  bc.kind     = lbm::BC_PERIODIC;
  bc.pressure = pressure;
  bc.connected.assign(connected, connected + nconnected);
  return add_bc(s, bc, cells, normals, n);
}

int lbm_b200_set_forcing(lbm_b200_solver* s, const int64_t* inlet, int64_t ninlet, const int64_t* outlet, int64_t noutlet, double gradient) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(ninlet < 0 || noutlet < 0 || (ninlet > 0 && inlet == nullptr) || (noutlet > 0 && outlet == nullptr)) return fail(LBM_B200_EINVAL, "bad forcing lists");
  auto& in = s->impl->in;
  for(int64_t k = 0; k < ninlet; ++k)
    if(inlet[k] < 0 || inlet[k] >= in.n) return fail(LBM_B200_EINVAL, "inlet cell id out of range");
  for(int64_t k = 0; k < noutlet; ++k)
    if(outlet[k] < 0 || outlet[k] >= in.n) return fail(LBM_B200_EINVAL, "outlet cell id out of range");
  in.forcing = true;
  in.inlet.assign(inlet, inlet + ninlet);
  in.outlet.assign(outlet, outlet + noutlet);
  in.gradient = gradient;
  return LBM_B200_OK;
}

int lbm_b200_set_ghosts(lbm_b200_solver* s, int64_t nghost) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(nghost < 0 || nghost >= s->impl->in.n) return fail(LBM_B200_EINVAL, "bad ghost count");
  s->impl->in.n_ghost = nghost;
  return LBM_B200_OK;
}

int lbm_b200_set_halo(lbm_b200_solver* s, int32_t npeers, const int32_t* peers, const int64_t* send_count, const int64_t* send_cell,
                      const int32_t* send_dir, const int64_t* recv_count, const int64_t* recv_cell, const int32_t* recv_dir) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(npeers < 0 || (npeers > 0 && (!peers || !send_count || !recv_count))) return fail(LBM_B200_EINVAL, "bad halo description");
  auto& in = s->impl->in;
  in.peers.assign(peers, peers + npeers);
  in.send_count.assign(send_count, send_count + npeers);
  in.recv_count.assign(recv_count, recv_count + npeers);
  int64_t ns = 0, nr = 0;
  for(int k = 0; k < npeers; ++k) {
    if(send_count[k] < 0 || recv_count[k] < 0) return fail(LBM_B200_EINVAL, "negative halo count");
    ns += send_count[k];
    nr += recv_count[k];
  }
  if((ns > 0 && (!send_cell || !send_dir)) || (nr > 0 && (!recv_cell || !recv_dir))) return fail(LBM_B200_EINVAL, "null halo list");
  in.send_cell.assign(send_cell, send_cell + ns);
  in.send_dir.assign(send_dir, send_dir + ns);
  in.recv_cell.assign(recv_cell, recv_cell + nr);
  in.recv_dir.assign(recv_dir, recv_dir + nr);
  return LBM_B200_OK;
}

int lbm_b200_set_vars_halo(lbm_b200_solver* s, const int64_t* send_count, const int64_t* send_cell, const int64_t* recv_count,
                           const int64_t* recv_cell) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  auto& in = s->impl->in;
  const size_t np = in.peers.size();
  if(np == 0) return fail(LBM_B200_ESTATE, "lbm_b200_set_vars_halo: call lbm_b200_set_halo first (the peer list is shared)");
  if(send_count == nullptr || recv_count == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  int64_t ns = 0, nr = 0;
  for(size_t k = 0; k < np; ++k) {
    if(send_count[k] < 0 || recv_count[k] < 0) return fail(LBM_B200_EINVAL, "negative velocity halo count");
    ns += send_count[k];
    nr += recv_count[k];
  }
  if((ns > 0 && send_cell == nullptr) || (nr > 0 && recv_cell == nullptr)) return fail(LBM_B200_EINVAL, "null velocity halo list");
  in.vsend_count.assign(send_count, send_count + np);
  in.vrecv_count.assign(recv_count, recv_count + np);
  in.vsend_cell.assign(send_cell, send_cell + ns);
  in.vrecv_cell.assign(recv_cell, recv_cell + nr);
  return LBM_B200_OK;
}

int lbm_b200_comm_unique_id(char* out128) {
  if(out128 == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  std::string err;
  if(!lbm::nccl_api().load(&err)) return fail(LBM_B200_ECUDA, err);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_TRY(lbm::nccl_api().GetUniqueId(&id));
  std::memcpy(out128, &id, 128);
  return LBM_B200_OK;
}

int lbm_b200_comm_init(lbm_b200_solver* s, const char* id128, int32_t rank, int32_t nranks) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(id128 == nullptr || nranks < 1 || rank < 0 || rank >= nranks) return fail(LBM_B200_EINVAL, "bad communicator description");
  std::string err;
  if(!lbm::nccl_api().load(&err)) return fail(LBM_B200_ECUDA, err);
  CUDA_TRY(cudaSetDevice(s->impl->cfg.device));
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  NCCL_TRY(lbm::nccl_api().CommInitRank(&s->impl->comm, nranks, id, rank));
  s->impl->comm_rank = rank;
  s->impl->comm_size = nranks;
  return LBM_B200_OK;
}

int lbm_b200_box_rows(int32_t ndim, const int64_t* shape, const int32_t* periodic, const int64_t* cells, int64_t ncells, int64_t* nghbr,
                      int32_t stride, double* center) {
  if(shape == nullptr || periodic == nullptr || nghbr == nullptr || (ncells > 0 && cells == nullptr)) return fail(LBM_B200_EINVAL, "null argument");
  std::string err;
  if(!lbm::box_rows(ndim, shape, periodic, cells, ncells, nghbr, stride, center, &err)) return fail(LBM_B200_EINVAL, err);
  return LBM_B200_OK;
}

int lbm_b200_partition_create(int64_t ncells_global, int32_t ndim, int32_t ndist, int32_t stride, int32_t rank, int32_t world,
                              lbm_b200_rows_fn rows_fn, void* user, int32_t npressure, const int64_t* const* pressure_cells,
                              const double* const* pressure_normals, const int64_t* pressure_count, lbm_b200_partition** out) {
  if(out == nullptr || rows_fn == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  *out = nullptr;
  if(npressure < 0 || (npressure > 0 && (!pressure_cells || !pressure_normals || !pressure_count))) return fail(LBM_B200_EINVAL, "bad pressure surface lists");
  if(ndim < 1 || ndim > 3) return fail(LBM_B200_EINVAL, "bad dimension");
  std::vector<lbm::PressureSurface> ps;
  for(int k = 0; k < npressure; ++k) {
    if(pressure_count[k] < 0 || (pressure_count[k] > 0 && (!pressure_cells[k] || !pressure_normals[k]))) return fail(LBM_B200_EINVAL, "bad pressure surface lists");
    ps.push_back({pressure_cells[k], pressure_normals[k], pressure_count[k]});
  }
  auto* part = new lbm_b200_partition();
  if(!lbm::build_partition(ncells_global, ndist, stride, ndim, rank, world, rows_fn, user, ps, part->p)) {
    const std::string msg = part->p.error;
    delete part;
    return fail(LBM_B200_EINVAL, msg);
  }
  *out = part;
  return LBM_B200_OK;
}

int lbm_b200_partition_get(const lbm_b200_partition* p, lbm_b200_partition_view* v) {
  if(p == nullptr || v == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  const lbm::Partition& P = p->p;
  std::memset(v, 0, sizeof(*v));
  v->lo = P.lo; v->hi = P.hi; v->n_owned = P.n_owned(); v->n_ghost = P.n_ghost();
  v->ghosts = P.ghosts.data(); v->nghbr = P.nghbr.data(); v->stride = P.stride; v->npeers = static_cast<int32_t>(P.peers.size());
  v->peers = P.peers.data();
  v->send_count = P.send_count.data(); v->recv_count = P.recv_count.data(); v->send_cell = P.send_cell.data(); v->recv_cell = P.recv_cell.data();
  v->send_dir = P.send_dir.data(); v->recv_dir = P.recv_dir.data();
  v->vsend_count = P.vsend_count.data(); v->vrecv_count = P.vrecv_count.data(); v->vsend_cell = P.vsend_cell.data(); v->vrecv_cell = P.vrecv_cell.data();
  return LBM_B200_OK;
}

int64_t lbm_b200_partition_restrict(const lbm_b200_partition* p, const int64_t* cells, int64_t n, int64_t* local_cells, int64_t* index) {
  if(p == nullptr || n < 0 || (n > 0 && cells == nullptr)) return -1;
  int64_t m = 0;
  for(int64_t k = 0; k < n; ++k) {
    if(cells[k] < p->p.lo || cells[k] >= p->p.hi) continue;
    if(local_cells != nullptr) local_cells[m] = cells[k] - p->p.lo;
    if(index != nullptr) index[m] = k;
    ++m;
  }
  return m;
}

int lbm_b200_partition_apply(const lbm_b200_partition* p, lbm_b200_solver* s) {
  if(p == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  const lbm::Partition& P = p->p;
  int rc = lbm_b200_set_ghosts(s, P.n_ghost());
  if(rc != LBM_B200_OK) return rc;
  rc = lbm_b200_set_halo(s, static_cast<int32_t>(P.peers.size()), P.peers.data(), P.send_count.data(), P.send_cell.data(), P.send_dir.data(),
                         P.recv_count.data(), P.recv_cell.data(), P.recv_dir.data());
  if(rc != LBM_B200_OK) return rc;
  if(!P.vsend_cell.empty() || !P.vrecv_cell.empty())
    rc = lbm_b200_set_vars_halo(s, P.vsend_count.data(), P.vsend_cell.data(), P.vrecv_count.data(), P.vrecv_cell.data());
  return rc;
}

void lbm_b200_partition_destroy(lbm_b200_partition* p) { delete p; }

int lbm_b200_set_stream(lbm_b200_solver* s, void* cuda_stream) {
  CHECK_HANDLE(s);
  s->impl->stream = static_cast<cudaStream_t>(cuda_stream);
  return LBM_B200_OK;
}

int lbm_b200_init(lbm_b200_solver* s) {
  CHECK_HANDLE(s);
  CHECK_NOT_INITED(s);
  if(s->impl->in.nghbr.empty()) return fail(LBM_B200_ESTATE, "lbm_b200_set_topology has not been called");
  if(s->impl->cfg.device < 0) return fail(LBM_B200_ECUDA, "inspection-only handle (device -1): there is no CPU compute path");
  bool wet = false, poisson_bc = false;
  for(const lbm::BcInput& bc : s->impl->in.bcs) {
    if(bc.kind >= lbm::BC_POISSON_DIRICHLET) poisson_bc = true;
    else wet = wet || bc.kind >= lbm::BC_WALL_EQ;
  }
  const bool poisson_lattice = s->impl->in.L.Q == 3 || s->impl->in.L.Q == 5;
  if(s->impl->in.poisson || poisson_lattice || poisson_bc) {
    if(!s->impl->in.poisson) return fail(LBM_B200_ESTATE, "D1Q3 / D2Q5 and the NEEM Dirichlet / Neumann conditions belong to the Poisson equation: call lbm_b200_set_poisson");
    if(wet) return fail(LBM_B200_EINVAL, "wall boundary conditions do not exist for the Poisson equation types");
    if(dynamic_cast<PoissonSolver*>(s->impl.get()) == nullptr) {
      SolverBase* q = new PoissonSolver();
      q->cfg    = s->impl->cfg;
      q->in     = std::move(s->impl->in);
      q->stream = s->impl->stream;
      s->impl.reset(q);
    }
    return s->impl->init();
  }
  if(wet) {
    // order-dependent boundary conditions: hand the same inputs to the reference-order pipeline (sequential.cuh)
    SolverBase* q = make_sequential(s->impl->cfg);
    if(q == nullptr) return fail(LBM_B200_EINVAL, "Unsupported model");
    q->cfg    = s->impl->cfg;
    q->in     = std::move(s->impl->in);
    q->stream = s->impl->stream;
    s->impl.reset(q);
  }
  return s->impl->init();
}

int lbm_b200_step(lbm_b200_solver* s, int64_t nsteps) {
  CHECK_HANDLE(s);
  if(nsteps < 0) return fail(LBM_B200_EINVAL, "negative step count");
  return s->impl->step(nsteps, nullptr, nullptr);
}

int lbm_b200_step_timed(lbm_b200_solver* s, int64_t nsteps, float* ms_total, float* ms_main) {
  CHECK_HANDLE(s);
  if(nsteps < 0 || ms_total == nullptr) return fail(LBM_B200_EINVAL, "bad argument");
  return s->impl->step(nsteps, ms_total, ms_main);
}

int lbm_b200_synchronize(lbm_b200_solver* s) {
  CHECK_HANDLE(s);
  return s->impl->sync();
}

int lbm_b200_residual(lbm_b200_solver* s, double* out, int32_t* diverged) {
  CHECK_HANDLE(s);
  if(out == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  return s->impl->residual(out, diverged);
}

int lbm_b200_get_populations(lbm_b200_solver* s, double* f, double* fold) {
  CHECK_HANDLE(s);
  return s->impl->get_populations(f, fold);
}

int lbm_b200_set_populations(lbm_b200_solver* s, const double* f, const double* fold) {
  CHECK_HANDLE(s);
  return s->impl->set_populations(f, fold);
}

int lbm_b200_get_vars(lbm_b200_solver* s, double* vars, double* varsold) {
  CHECK_HANDLE(s);
  return s->impl->get_vars(vars, varsold);
}

int lbm_b200_get_moments(lbm_b200_solver* s, double* moments) {
  CHECK_HANDLE(s);
  if(moments == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  return s->impl->get_moments(moments);
}

int64_t lbm_b200_steps_done(const lbm_b200_solver* s) { return (s && s->impl) ? s->impl->t : -1; }

int lbm_b200_get_stats(const lbm_b200_solver* s, lbm_b200_stats* out) {
  CHECK_HANDLE(s);
  if(out == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  if(!s->impl->inited) return fail(LBM_B200_ESTATE, "not initialised");
  s->impl->stats(out);
  return LBM_B200_OK;
}

int64_t lbm_b200_sfc_index(int32_t ndim, const double* x, int32_t level) {
  if(x == nullptr || ndim < 1 || ndim > 4 || level < 0 || ndim * level > 62) return -1;
  return lbm::sfc_index_unit(ndim, x, level);
}

int64_t lbm_b200_box_ncells(int32_t ndim, const int64_t* shape) {
  if(shape == nullptr || ndim < 2 || ndim > 3) return -1;
  int64_t n = 1;
  for(int d = 0; d < ndim; ++d) {
    if(shape[d] <= 0) return -1;
    n *= shape[d];
  }
  return n;
}

int lbm_b200_box_topology(int32_t ndim, const int64_t* shape, const int32_t* periodic, int64_t* nghbr, int32_t stride, double* center,
                          int64_t* coords) {
  if(shape == nullptr || periodic == nullptr || nghbr == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  std::string err;
  if(!lbm::box_topology(ndim, shape, periodic, nghbr, stride, center, coords, &err)) return fail(LBM_B200_EINVAL, err);
  return LBM_B200_OK;
}

int lbm_b200_debug_plan(lbm_b200_solver* s, lbm_b200_plan_view* out) {
  CHECK_HANDLE(s);
  if(out == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  if(s->impl->in.poisson || s->impl->in.L.Q < 9) return fail(LBM_B200_EUNSUP, "the Poisson equation types have no fused device plan");
  for(const lbm::BcInput& bc : s->impl->in.bcs)
    if(bc.kind >= lbm::BC_WALL_EQ) return fail(LBM_B200_EUNSUP, "configurations with wet-node walls have no fused device plan");
  return s->impl->debug_plan(out);
}

const char* lbm_b200_last_error(void) { return g_error.c_str(); }
int         lbm_b200_abi_version(void) { return LBM_B200_ABI_VERSION; }

} // extern "C"
