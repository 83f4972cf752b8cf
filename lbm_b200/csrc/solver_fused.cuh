// solver_fused.cuh -- the fused solver (device plan + k_step_generic / k_step_fast + auxiliary kernels + halo exchange) behind the
// SolverBase interface; instantiated by fused_*.cu.
#pragma once
#include "solver_common.cuh"
#include "output.cuh"

namespace lbm_impl {

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
  DevBuf<int32_t>  d_ref2dev, d_dev2ref;
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
  bool    overlap_enabled = true, debug_identity = false, extrap_side_enabled = false;
  bool    first = true; // next step is step 0 of the reference loop (m_fold = initial condition)
  int     n_fast_blocks = 0, n_gen_blocks = 0, max_resident = 0;
  cudaStream_t comm_stream = nullptr;   // halo exchange runs here, overlapped with the update of the inner cells
  cudaStream_t gen_stream = nullptr;    // link-code cells are updated here, next to the persistent chunk CTAs
  cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t  ev_outer = nullptr, ev_halo = nullptr, ev_pack = nullptr;
  bool         halo_pending = false;
  int64_t launches = 0, launches_main = 0;
  // ---- peer-to-peer halo over NVLink (CUDA IPC): every rank owns a mailbox = flag words + two receive buffers (exchange parity);
  // peers write into it with device-to-device copies on the copy engines and then raise their flag
  static constexpr int kMaxP2PPeers = 40;
  struct P2PBlob {
    cudaIpcMemHandle_t mem;
    int32_t rank, npeers, real_bytes, pad;
    int64_t slot_elems;
    int32_t peers[kMaxP2PPeers];
    int64_t recv_off[kMaxP2PPeers], recv_cnt[kMaxP2PPeers];
  };
  static_assert(sizeof(P2PBlob) <= LBM_B200_P2P_BLOB, "blob size");
  static constexpr size_t kFlagBytes = 512;
  DevBuf<unsigned char> mailbox;
  DevBuf<unsigned long long> d_outer_done;
  DevBuf<unsigned long long*> d_flag_ptrs;
  DevBuf<int> d_p2p_err;
  std::vector<void*>  p2p_peer_base;   // opened mailboxes, one per peer in list order
  std::vector<size_t> p2p_peer_off;    // byte offset of my segment inside the peer's receive buffer
  std::vector<size_t> p2p_peer_slot;   // bytes of one receive buffer of the peer
  int64_t p2p_slot_elems = 0;
  unsigned long long p2p_seq = 0, outer_total = 0;
  bool p2p_on = false;
  cudaEvent_t ev_step = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm0 = nullptr, evm1 = nullptr; // evm0 / evm1 point into ev_ring while a timed run records
  std::vector<cudaEvent_t> ev_ring;

  ~Solver() override {
    if(ev0) cudaEventDestroy(ev0);
    if(ev1) cudaEventDestroy(ev1);
    for(cudaEvent_t e : ev_ring) cudaEventDestroy(e);
    if(ev_outer) cudaEventDestroy(ev_outer);
    if(ev_halo) cudaEventDestroy(ev_halo);
    if(ev_pack) cudaEventDestroy(ev_pack);
    if(ev_step) cudaEventDestroy(ev_step);
    for(void* b : p2p_peer_base) if(b) cudaIpcCloseMemHandle(b);
    if(comm_stream) cudaStreamDestroy(comm_stream);
    if(ev_fork) cudaEventDestroy(ev_fork);
    if(ev_join) cudaEventDestroy(ev_join);
    if(gen_stream) cudaStreamDestroy(gen_stream);
  }

  lbm::DevParams<Real> params(int src, int dst, Real* vars_out) const {
    lbm::DevParams<Real> p{};
    p.A = f[src].p;
    p.B = f[dst].p;
    p.stride = plan.npad;
    p.pr = plan.perm_range();
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
    p.tabs.pr = plan.perm_range();
    p.omega = static_cast<Real>(cfg.omega);
    p.om1 = static_cast<Real>(1 - cfg.omega);
    p.omega_minus = static_cast<Real>(cfg.omega_minus);
    // MRT: base rate s0 in p.omega, (s_k - s0) / |row k|^2 per moment (kernels.cuh: Phys::mrt_row; lattice.h: mrt_base_rate)
    if(cfg.collision == LBM_B200_MRT) {
      const double s0 = lbm::mrt_base_rate(cfg.mrt_rates, Q, D);
      p.omega = static_cast<Real>(s0);
      for(int i = 0; i < 27; ++i) p.rates[i] = (i > D && i < Q) ? static_cast<Real>((cfg.mrt_rates[i] - s0) / lbm::MrtBasis<L>::norm(i)) : Real(0);
    }
    p.vars_out = vars_out;
    p.first = first ? 1 : 0;
    if(debug_identity) p.first = 1; // timing experiments only (LBM_B200_DEBUG_IDENTITY): every step reads its own cell, no gather
    return p;
  }

  using KernelFn = void (*)(const lbm::DevParams<Real>);
  // the two kernels of a time step: [0] persistent chunk CTAs (k_step_fast), [1] link-code cells (k_step_generic)
  template <bool STRICT, int COLL>
  static KernelFn kernel_ptr(int which) {
    return which == 0 ? static_cast<KernelFn>(&lbm::k_step_fast<L, Real, STRICT, COLL>) : static_cast<KernelFn>(&lbm::k_step_generic<L, Real, STRICT, COLL>);
  }
  KernelFn main_kernel(int which) const {
    const bool strict = cfg.arithmetic == LBM_B200_STRICT;
    switch(cfg.collision) {
      case LBM_B200_TRT: return strict ? kernel_ptr<true, lbm::COLL_TRT>(which) : kernel_ptr<false, lbm::COLL_TRT>(which);
      case LBM_B200_MRT: return strict ? kernel_ptr<true, lbm::COLL_MRT>(which) : kernel_ptr<false, lbm::COLL_MRT>(which);
      default: return strict ? kernel_ptr<true, lbm::COLL_BGK>(which) : kernel_ptr<false, lbm::COLL_BGK>(which);
    }
  }
  static constexpr int kFastSmem = lbm::FastCfg<L, Real>::SMEM_BYTES;

  int init() override {
    debug_identity = std::getenv("LBM_B200_DEBUG_IDENTITY") != nullptr;
    extrap_side_enabled = std::getenv("LBM_B200_EXTRAP_SIDE") != nullptr;
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
    CUDA_TRY(d_dev2ref.upload(plan.dev2ref));
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
    CUDA_TRY(cudaStreamCreateWithFlags(&gen_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));

    // launch geometry: generic blocks first, then persistent fast blocks (a multiple of the SM count)
    n_gen_blocks = static_cast<int>((plan.n_gen + lbm::kThreads - 1) / lbm::kThreads);
    {
      cudaDeviceProp prop{};
      CUDA_TRY(cudaGetDeviceProperties(&prop, cfg.device));
      // the chunk kernel stages whole chunks in shared memory: opt in to the large carve-out
      CUDA_TRY(cudaFuncSetAttribute(reinterpret_cast<const void*>(main_kernel(0)), cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmem));
      int per_sm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, main_kernel(0), lbm::kFastThreads, kFastSmem));
      if(per_sm < 1) return fail(LBM_B200_ECUDA, "the chunk kernel does not fit on this device (shared memory)");
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
      CUDA_TRY(cudaEventCreateWithFlags(&ev_step, cudaEventDisableTiming));
      // mailbox of the peer-to-peer halo (used once lbm_b200_p2p_import has mapped the peers)
      int64_t nrecv = 0;
      for(int64_t v : in.recv_count) nrecv += v;
      p2p_slot_elems = (nrecv + 63) / 64 * 64;
      CUDA_TRY(mailbox.alloc(kFlagBytes + 2 * static_cast<size_t>(p2p_slot_elems) * sizeof(Real)));
      CUDA_TRY(cudaMemset(mailbox.p, 0, mailbox.bytes()));
      CUDA_TRY(d_outer_done.alloc(1));
      CUDA_TRY(cudaMemset(d_outer_done.p, 0, sizeof(unsigned long long)));
      CUDA_TRY(d_p2p_err.alloc(1));
      CUDA_TRY(cudaMemset(d_p2p_err.p, 0, sizeof(int)));
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
        lbm::k_init<L, Real, true><<<nb, 256, 0, stream>>>(f[0].p, d_v0, plan.npad, static_cast<int32_t>(npad), plan.perm_range());
      else
        lbm::k_init<L, Real, false><<<nb, 256, 0, stream>>>(f[0].p, d_v0, plan.npad, static_cast<int32_t>(npad), plan.perm_range());
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
          values[k + 1]      = static_cast<double>(h[static_cast<size_t>(plan.pop_index(dir, cell))]);
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

  // phases (bit mask) -- 1: forcing, periodic-with-pressure values; 2: pressure extrapolation; 4: m_vars fix-ups.  Old wording:
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
    if((phases & 4) && vars_out != nullptr && d_varfix.n > 0) {
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

  int p2p_export(void* out) override {
    if(!inited) return fail(LBM_B200_ESTATE, "lbm_b200_p2p_export: call after lbm_b200_init");
    if(in.peers.empty()) return fail(LBM_B200_ESTATE, "lbm_b200_p2p_export: this rank has no halo lists");
    if(in.peers.size() > static_cast<size_t>(kMaxP2PPeers)) return fail(LBM_B200_EUNSUP, "peer-to-peer halo: too many neighbour ranks");
    if(has_velocity_halo()) return fail(LBM_B200_EUNSUP, "peer-to-peer halo: the velocity halo of a pressure boundary across a cut travels over NCCL");
    P2PBlob b{};
    CUDA_TRY(cudaIpcGetMemHandle(&b.mem, mailbox.p));
    b.rank = comm_rank;
    b.npeers = static_cast<int32_t>(in.peers.size());
    b.real_bytes = static_cast<int32_t>(sizeof(Real));
    b.slot_elems = p2p_slot_elems;
    int64_t off = 0;
    for(size_t k = 0; k < in.peers.size(); ++k) {
      b.peers[k] = in.peers[k];
      b.recv_off[k] = off;
      b.recv_cnt[k] = in.recv_count[k];
      off += in.recv_count[k];
    }
    std::memset(out, 0, LBM_B200_P2P_BLOB);
    std::memcpy(out, &b, sizeof(b));
    return LBM_B200_OK;
  }

  int p2p_import(int32_t nranks, const void* blobs) override {
    if(!inited) return fail(LBM_B200_ESTATE, "lbm_b200_p2p_import: call after lbm_b200_init");
    if(p2p_on) return fail(LBM_B200_ESTATE, "lbm_b200_p2p_import: already imported");
    if(has_velocity_halo()) return fail(LBM_B200_EUNSUP, "peer-to-peer halo: the velocity halo of a pressure boundary across a cut travels over NCCL");
    const unsigned char* base = static_cast<const unsigned char*>(blobs);
    std::vector<unsigned long long*> flag_ptrs;
    for(size_t k = 0; k < in.peers.size(); ++k) {
      const int r = in.peers[k];
      if(r < 0 || r >= nranks) return fail(LBM_B200_EINVAL, "peer-to-peer halo: peer rank out of range");
      P2PBlob b;
      std::memcpy(&b, base + static_cast<size_t>(r) * LBM_B200_P2P_BLOB, sizeof(b));
      if(b.rank != r || b.real_bytes != static_cast<int32_t>(sizeof(Real))) return fail(LBM_B200_EINVAL, "peer-to-peer halo: blob does not belong to that rank / precision");
      int j = -1;
      for(int q = 0; q < b.npeers; ++q) if(b.peers[q] == comm_rank) j = q;
      if(j < 0 || b.recv_cnt[j] != in.send_count[k]) return fail(LBM_B200_EINVAL, "peer-to-peer halo: the peer's receive list does not match this rank's send list");
      void* pb = nullptr;
      CUDA_TRY(cudaIpcOpenMemHandle(&pb, b.mem, cudaIpcMemLazyEnablePeerAccess));
      p2p_peer_base.push_back(pb);
      p2p_peer_off.push_back(static_cast<size_t>(b.recv_off[j]) * sizeof(Real));
      p2p_peer_slot.push_back(static_cast<size_t>(b.slot_elems) * sizeof(Real));
      flag_ptrs.push_back(reinterpret_cast<unsigned long long*>(pb) + j);
    }
    CUDA_TRY(d_flag_ptrs.upload(flag_ptrs));
    p2p_on = true;
    return LBM_B200_OK;
  }

  // Outgoing populations of this step -> the peers' mailboxes, theirs -> my ghost cells, without a single SM-sized kernel: one small
  // pack kernel, device-to-device copies over NVLink on the copy engines, a one-warp kernel that raises my flag in every peer's
  // mailbox, a one-warp kernel that waits for the peers' flags, one small unpack kernel.  All of them fit beside the persistent chunk
  // CTAs, so the exchange really runs while the inner tiles are updated.
  int p2p_exchange(Real* buf, cudaStream_t st) {
    const int64_t ns = static_cast<int64_t>(plan.send_index.size()), nr = static_cast<int64_t>(plan.recv_index.size());
    const unsigned long long seq = ++p2p_seq;
    const size_t parity = static_cast<size_t>(seq & 1);
    if(ns > 0) {
      lbm::k_halo_pack<Real><<<static_cast<int>((ns + 127) / 128), 128, 0, st>>>(buf, d_send_idx.p, ns, d_sendbuf.p);
      ++launches;
    }
    int64_t so = 0;
    for(size_t k = 0; k < in.peers.size(); ++k) {
      if(in.send_count[k] > 0) {
        unsigned char* dstp = static_cast<unsigned char*>(p2p_peer_base[k]) + kFlagBytes + parity * p2p_peer_slot[k] + p2p_peer_off[k];
        CUDA_TRY(cudaMemcpyAsync(dstp, d_sendbuf.p + so, static_cast<size_t>(in.send_count[k]) * sizeof(Real), cudaMemcpyDeviceToDevice, st));
      }
      so += in.send_count[k];
    }
    const int np = static_cast<int>(in.peers.size());
    lbm::k_p2p_signal<<<1, 64, 0, st>>>(d_flag_ptrs.p, np, seq);
    lbm::k_p2p_wait<<<1, 64, 0, st>>>(reinterpret_cast<const unsigned long long*>(mailbox.p), np, seq, d_p2p_err.p);
    launches += 2;
    if(nr > 0) {
      const Real* rb = reinterpret_cast<const Real*>(mailbox.p + kFlagBytes) + parity * static_cast<size_t>(p2p_slot_elems);
      lbm::k_halo_unpack<Real><<<static_cast<int>((nr + 127) / 128), 128, 0, st>>>(buf, d_recv_idx.p, nr, rb);
      ++launches;
    }
    halo_bytes += (ns + nr) * static_cast<int64_t>(sizeof(Real));
    CUDA_TRY(cudaGetLastError());
    return LBM_B200_OK;
  }

  bool has_velocity_halo() const { return !in.vsend_cell.empty() || !in.vrecv_cell.empty(); }

  // a launch over generic cells [g0, g0+ng) and fast chunks [c0, c0+ncnk): the link-code cells on the side stream (they are
  // latency bound and few), the persistent chunk CTAs on the solver's stream; both read buffer A and write disjoint cells of B
  // extrap_nd >= 0: also run the pressure extrapolation of this step (it only reads buffer A and writes d_uext[extrap_nd]) on the side stream
  int launch_main(const lbm::DevParams<Real>& p, int64_t g0, int64_t ng, int64_t c0, int64_t ncnk, int resident_cap, int cls, int extrap_nd = -1,
                  int64_t outer_chunks = 0) {
    lbm::DevParams<Real> q = p;
    q.gen_off       = static_cast<int32_t>(g0);
    q.n_gen         = static_cast<int32_t>(ng);
    q.n_gen_blocks  = static_cast<int>((ng + lbm::kThreads - 1) / lbm::kThreads);
    q.chunk_off     = static_cast<int32_t>(c0);
    q.n_fast_chunks = static_cast<int32_t>(ncnk);
    const int64_t ntiles = ncnk * lbm::FastCfg<L, Real>::NSPLIT; // the persistent CTAs draw tiles (whole chunks or their halves)
    q.n_fast_blocks = static_cast<int32_t>(ntiles < resident_cap ? ntiles : resident_cap);
    q.outer_done    = outer_chunks > 0 ? d_outer_done.p : nullptr;
    q.n_outer_tiles = static_cast<int32_t>(outer_chunks * lbm::FastCfg<L, Real>::NSPLIT);
    q.ticket        = d_ticket.p + cls;
    q.ticket_base   = ticket_next[cls];
    ticket_next[cls] += static_cast<unsigned long long>(ntiles)
                        + static_cast<unsigned long long>(q.n_fast_blocks) * lbm::FastCfg<L, Real>::PAST_END;
    const bool extrap = extrap_nd >= 0 && d_abb.n > 0;
    const bool side = q.n_fast_blocks > 0 && (q.n_gen_blocks > 0 || extrap);
    if(side) {
      CUDA_TRY(cudaEventRecord(ev_fork, stream));
      CUDA_TRY(cudaStreamWaitEvent(gen_stream, ev_fork, 0));
    }
    if(q.n_gen_blocks > 0) {
      main_kernel(1)<<<q.n_gen_blocks, lbm::kThreads, 0, side ? gen_stream : stream>>>(q);
      ++launches;
      ++launches_main;
    }
    if(extrap) {
      const int n = static_cast<int>(d_abb.n);
      cudaStream_t st = side ? gen_stream : stream;
      if(cfg.arithmetic == LBM_B200_STRICT) lbm::k_pressure_extrapolate<L, Real, true><<<(n + 127) / 128, 128, 0, st>>>(p, n, d_uext[extrap_nd].p, d_vrecvbuf.p);
      else lbm::k_pressure_extrapolate<L, Real, false><<<(n + 127) / 128, 128, 0, st>>>(p, n, d_uext[extrap_nd].p, d_vrecvbuf.p);
      ++launches;
    }
    if(q.n_fast_blocks > 0) {
      main_kernel(0)<<<q.n_fast_blocks, lbm::kFastThreads, kFastSmem, stream>>>(q);
      ++launches;
      ++launches_main;
    }
    if(side) {
      CUDA_TRY(cudaEventRecord(ev_join, gen_stream));
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_join, 0));
    }
    return LBM_B200_OK;
  }

  int one_step(bool time_main) {
    const int src = cur, dst = cur ^ 1;
    Real*     vout = nullptr;
    if(want_vars(t)) vout = vars[vcur ^ 1].p;
    lbm::DevParams<Real> p = params(src, dst, vout);
    if(prev_fold.p != nullptr) p.A = prev_fold.p; // explicit m_fold supplied by set_populations
    auto launch = [&](int64_t g0, int64_t ng, int64_t c0, int64_t ncnk, int resident_cap, int cls, int extrap_nd = -1) -> int {
      return launch_main(p, g0, ng, c0, ncnk, resident_cap, cls, extrap_nd);
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
    if(overlap && p2p_on) {
      // Peer-to-peer halo: ONE launch over all tiles, the outer ones first in ticket order.  The communication stream waits (a one-warp
      // kernel) until the device-side counter says they are written, then moves them; the persistent CTAs just carry on with the inner
      // tiles.  No split launch, no SM-sized communication kernel.
      CUDA_TRY(cudaEventRecord(ev_step, stream));
      CUDA_TRY(cudaStreamWaitEvent(comm_stream, ev_step, 0));
      const bool gen_side = plan.n_gen > 0 && plan.n_fast_chunks > 0;
      rc = launch_main(p, 0, plan.n_gen, 0, plan.n_fast_chunks, max_resident, 0, -1, plan.n_fast_outer);
      if(rc != LBM_B200_OK) return rc;
      if(time_main) cudaEventRecord(evm1, stream);
      if(gen_side) CUDA_TRY(cudaStreamWaitEvent(comm_stream, ev_join, 0)); // the link-code cells (all of them, they are few) are written
      else if(plan.n_fast_chunks == 0) { CUDA_TRY(cudaEventRecord(ev_outer, stream)); CUDA_TRY(cudaStreamWaitEvent(comm_stream, ev_outer, 0)); }
      if(plan.n_fast_outer > 0) {
        outer_total += static_cast<unsigned long long>(plan.n_fast_outer) * lbm::FastCfg<L, Real>::NSPLIT;
        lbm::k_wait_counter<<<1, 32, 0, comm_stream>>>(d_outer_done.p, outer_total, d_p2p_err.p);
        ++launches;
      }
      rc = p2p_exchange(f[dst].p, comm_stream);
      if(rc != LBM_B200_OK) return rc;
      CUDA_TRY(cudaEventRecord(ev_halo, comm_stream));
      halo_pending = true;
      if(has_aux) { // pressure boundary present: extrapolation + m_vars fix-ups behind the launch, the exchange in flight
        const int nd = dyn ^ 1;
        bool      dyn_written = false;
        rc = cfg.arithmetic == LBM_B200_STRICT ? aux_kernels<true>(p, vout, nd, 2 | 4, &dyn_written) : aux_kernels<false>(p, vout, nd, 2 | 4, &dyn_written);
        if(rc != LBM_B200_OK) return rc;
        if(dyn_written) dyn = nd;
      }
    } else if(overlap) {
      // 1. outer cells (whatever a peer needs) with the whole GPU; 2. their populations are packed and travel (NCCL group on
      // the high-priority communication stream) while 3. the inner cells are updated.  The inner launch is released by the
      // same event that releases the NCCL kernel, so the (higher-priority, whole-SM-sized) NCCL CTAs are placed first and
      // the persistent inner CTAs fill the remaining SMs; with ticket scheduling late inner CTAs just take fewer chunks.
      rc = launch(0, plan.n_gen_outer, 0, plan.n_fast_outer, max_resident, 1);
      if(rc != LBM_B200_OK) return rc;
      CUDA_TRY(cudaEventRecord(ev_outer, stream));
      CUDA_TRY(cudaStreamWaitEvent(comm_stream, ev_outer, 0));
      rc = halo_exchange(f[dst].p, comm_stream, ev_pack);
      if(rc != LBM_B200_OK) return rc;
      CUDA_TRY(cudaEventRecord(ev_halo, comm_stream));
      halo_pending = true;
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_pack, 0));
      // pressure boundary present: the extrapolation runs beside the inner launch (side stream), the m_vars fix-ups behind it
      const int nd = dyn ^ 1;
      rc = launch(plan.n_gen_outer, plan.n_gen - plan.n_gen_outer, plan.n_fast_outer, plan.n_fast_chunks - plan.n_fast_outer, max_resident, 0,
                  extrap_side_enabled ? nd : -1);
      if(rc != LBM_B200_OK) return rc;
      if(time_main) cudaEventRecord(evm1, stream);
      CUDA_TRY(cudaGetLastError());
      if(has_aux) {
        bool      dyn_written = extrap_side_enabled && d_abb.n > 0;
        const int phases = extrap_side_enabled ? 4 : (2 | 4);
        rc = cfg.arithmetic == LBM_B200_STRICT ? aux_kernels<true>(p, vout, nd, phases, &dyn_written) : aux_kernels<false>(p, vout, nd, phases, &dyn_written);
        if(rc != LBM_B200_OK) return rc;
        if(dyn_written) dyn = nd;
      }
    } else {
      const int  nd = dyn ^ 1;
      // The pressure extrapolation depends on buffer A only and could run beside the main kernels (launch_main: extrap_nd); measured on
      // B200 (3D step / sphere workloads, 256^3) that is no faster than running it behind them -- it rebuilds m_fold of two neighbours
      // per entry and competes with the bandwidth-bound chunk kernel -- so it stays behind (LBM_B200_EXTRAP_SIDE=1 switches it on).
      const bool extrap_side = extrap_side_enabled && !has_velocity_halo() && d_abb.n > 0;
      rc = launch(0, plan.n_gen, 0, plan.n_fast_chunks, max_resident, 0, extrap_side ? nd : -1);
      if(rc != LBM_B200_OK) return rc;
      if(time_main) cudaEventRecord(evm1, stream);
      CUDA_TRY(cudaGetLastError());
      bool       dyn_written = extrap_side;
      const bool strict = cfg.arithmetic == LBM_B200_STRICT;
      auto aux = [&](int phases) { return strict ? aux_kernels<true>(p, vout, nd, phases, &dyn_written) : aux_kernels<false>(p, vout, nd, phases, &dyn_written); };
      if(has_velocity_halo()) {
        // the pressure extrapolation of this step reads velocities that arrive with this step's exchange
        rc = aux(1);
        if(rc != LBM_B200_OK) return rc;
        rc = halo_exchange(f[dst].p, stream, nullptr, &p);
        if(rc != LBM_B200_OK) return rc;
        rc = aux(2 | 4);
        if(rc != LBM_B200_OK) return rc;
      } else {
        rc = aux(extrap_side ? (1 | 4) : (1 | 2 | 4));
        if(rc != LBM_B200_OK) return rc;
        rc = (p2p_on && !in.peers.empty()) ? p2p_exchange(f[dst].p, stream) : halo_exchange(f[dst].p, stream);
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
    // per-launch timing: one event pair per step, recorded on the stream and read only after the loop -- the host never waits inside
    // the timed region, so launches run ahead of the device exactly as in an untimed run
    const bool per_launch = timed && ms_main != nullptr;
    if(per_launch)
      while(static_cast<int64_t>(ev_ring.size()) < 2 * n) {
        cudaEvent_t e = nullptr;
        CUDA_TRY(cudaEventCreate(&e));
        ev_ring.push_back(e);
      }
    for(int64_t s = 0; s < n; ++s) {
      if(per_launch) { evm0 = ev_ring[2 * s]; evm1 = ev_ring[2 * s + 1]; }
      int rc = one_step(per_launch);
      if(rc != LBM_B200_OK) return rc;
    }
    if(halo_pending) { // the step is complete only when the ghosts have arrived
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_halo, 0));
      halo_pending = false;
    }
    if(timed) {
      CUDA_TRY(cudaEventRecord(ev1, stream));
      CUDA_TRY(cudaEventSynchronize(ev1));
      CUDA_TRY(cudaEventElapsedTime(ms_total, ev0, ev1));
      if(per_launch) {
        for(int64_t s = 0; s < n; ++s) {
          float ms = 0;
          CUDA_TRY(cudaEventElapsedTime(&ms, ev_ring[2 * s], ev_ring[2 * s + 1]));
          main_acc += ms;
        }
        *ms_main = static_cast<float>(main_acc);
      }
    }
    return LBM_B200_OK;
  }

  int sync() override {
    if(comm_stream) CUDA_TRY(cudaStreamSynchronize(comm_stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if(p2p_on) {
      int e = 0;
      CUDA_TRY(cudaMemcpy(&e, d_p2p_err.p, sizeof(int), cudaMemcpyDeviceToHost));
      if(e != 0) return fail(LBM_B200_ECUDA, e == 1 ? "peer-to-peer halo: the outer tiles of a step never completed" : "peer-to-peer halo: a peer never signalled its populations");
    }
    return LBM_B200_OK;
  }

  // SoA device array [width][npad] -> AoS host array [n][width] in reference cell order (and back).
  // The transposition runs on the device; the host side is one cudaMemcpy of the reference's own layout, so a
  // pinned caller buffer moves at PCIe speed.
  int ensure_stage() {
    if(stage.p == nullptr) CUDA_TRY(stage.alloc(static_cast<size_t>(plan.n) * Q));
    return LBM_B200_OK;
  }
  // pop: dsrc / ddst is a population array (per-direction in-chunk layouts), otherwise a plain per-cell SoA array
  int download(const Real* dsrc, int width, double* out, bool pop) {
    int rc = ensure_stage();
    if(rc) return rc;
    const int nb = static_cast<int>((plan.n + 255) / 256);
    if(pop) lbm::k_pack_aos<L, Real, true><<<nb, 256, 0, stream>>>(dsrc, d_ref2dev.p, plan.n, width, stage.p, plan.npad, plan.perm_range());
    else lbm::k_pack_aos<L, Real, false><<<nb, 256, 0, stream>>>(dsrc, d_ref2dev.p, plan.n, width, stage.p, plan.npad, plan.perm_range());
    ++launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, stage.p, sizeof(double) * static_cast<size_t>(plan.n) * width, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    d2h_bytes += static_cast<int64_t>(sizeof(double)) * plan.n * width;
    return LBM_B200_OK;
  }
  int upload_aos(const double* src, int width, Real* ddst, bool pop) {
    int rc = ensure_stage();
    if(rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(stage.p, src, sizeof(double) * static_cast<size_t>(plan.n) * width, cudaMemcpyHostToDevice, stream));
    const int nb = static_cast<int>((plan.n + 255) / 256);
    if(pop) {
      // population arrays: one thread per destination position, so that every direction's array is written with full coalescing
      const int nbp = static_cast<int>((plan.npad + 255) / 256);
      lbm::k_unpack_aos_pop<L, Real><<<nbp, 256, 0, stream>>>(stage.p, d_dev2ref.p, plan.npad, ddst, plan.npad, plan.perm_range());
    } else {
      lbm::k_unpack_aos<L, Real, false><<<nb, 256, 0, stream>>>(stage.p, d_ref2dev.p, plan.n, width, ddst, plan.npad, plan.perm_range());
    }
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
      int rc = download(f[cur].p, Q, fo, true);
      if(rc) return rc;
    }
    if(foldo != nullptr) {
      if(prev_fold.p != nullptr) return download(prev_fold.p, Q, foldo, true);
      int rc = gather_all(scratch.p, nullptr);
      if(rc) return rc;
      return download(scratch.p, Q, foldo, false);
    }
    return LBM_B200_OK;
  }

  int set_populations(const double* fi, const double* foldi) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(foldi == nullptr) return fail(LBM_B200_EINVAL, "set_populations needs m_fold");
    CUDA_TRY(cudaStreamSynchronize(stream));
    if(comm_stream) CUDA_TRY(cudaStreamSynchronize(comm_stream));
    halo_pending = false;
    int rc = LBM_B200_OK;
    if(fi != nullptr) {
      // m_f is not an input of the next step (the collision overwrites it, solver.cpp:603); kept for read-back only, with the
      // supplied m_fold in a buffer of its own until the next step has consumed it
      rc = upload_aos(fi, Q, f[cur].p, true);
      if(rc) return rc;
      CUDA_TRY(prev_fold.alloc(static_cast<size_t>(plan.npad) * Q));
      CUDA_TRY(cudaMemset(prev_fold.p, 0, prev_fold.bytes()));
      rc = upload_aos(foldi, Q, prev_fold.p, true);
    } else {
      // no m_f given: the supplied m_fold takes its place in the current buffer (no allocation, one pass); until the next step
      // lbm_b200_get_populations reports it as m_f as well
      prev_fold.alloc(0);
      rc = upload_aos(foldi, Q, f[cur].p, true);
    }
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

  // Device side of LBMSolver::output (output.cuh): moments pass, cell filter, 15-decimal rounding and base64 on the device; the host
  // receives the text of the NVAR <DataArray> payloads.  keep: one byte per OWNED reference cell (nullptr: all).
  DevBuf<int32_t> d_out_sel;
  DevBuf<double>  d_out_col;
  DevBuf<char>    d_out_text;
  DevBuf<int>     d_out_slow;
  std::vector<uint8_t> out_keep_cached;
  int64_t              out_n_cached = -1;
  int encode_output(const uint8_t* keep, char* text, int64_t capacity, int64_t* offsets) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    const int64_t no = plan.n_owned;
    // selection list (device cells of the kept reference cells), rebuilt only when the filter changes
    const bool same = out_n_cached >= 0 && (keep == nullptr ? out_keep_cached.empty() : (out_keep_cached.size() == static_cast<size_t>(no)
                                                                                         && std::memcmp(out_keep_cached.data(), keep, static_cast<size_t>(no)) == 0));
    if(!same) {
      std::vector<int32_t> sel;
      sel.reserve(static_cast<size_t>(no));
      for(int64_t c = 0; c < no; ++c)
        if(keep == nullptr || keep[c]) sel.push_back(plan.ref2dev[c]);
      out_n_cached = -1; // nothing is cached until the buffers below exist
      if(sel.empty()) return fail(LBM_B200_EINVAL, "ERROR: Invalid call to encodeLE() with length = 0"); // base64.h:219-223
      if(keep == nullptr) out_keep_cached.clear(); else out_keep_cached.assign(keep, keep + no);
      CUDA_TRY(d_out_sel.upload(sel));
      CUDA_TRY(d_out_col.alloc(sel.size()));
      CUDA_TRY(d_out_text.alloc(static_cast<size_t>(NVAR) * lbm::out::base64_chars(static_cast<int64_t>(sel.size()))));
      CUDA_TRY(d_out_slow.alloc(1));
      out_n_cached = static_cast<int64_t>(sel.size());
    }
    const int64_t n = out_n_cached, chars = lbm::out::base64_chars(n);
    if(capacity < chars * NVAR) return fail(LBM_B200_EINVAL, "lbm_b200_encode_output: text buffer too small");
    // the moments output() recomputes (solver.cpp:336): one pass of the step kernels without population stores
    if(halo_pending) {
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_halo, 0));
      halo_pending = false;
    }
    lbm::DevParams<Real> p = params(cur, cur ^ 1, scratch.p);
    if(prev_fold.p != nullptr) p.A = prev_fold.p;
    p.B = nullptr;
    int rc = launch_main(p, 0, plan.n_gen, 0, plan.n_fast_chunks, max_resident, 0);
    if(rc) return rc;
    CUDA_TRY(cudaMemsetAsync(d_out_slow.p, 0, sizeof(int), stream));
    const int nbk = static_cast<int>((n + 255) / 256);
    const int64_t ngroups = chars / 4;
    for(int v = 0; v < NVAR; ++v) {
      lbm::out::k_output_column<Real><<<nbk, 256, 0, stream>>>(scratch.p, plan.npad, v, d_out_sel.p, n, d_out_col.p, d_out_slow.p);
      lbm::out::k_base64_field<<<static_cast<int>((ngroups + 255) / 256), 256, 0, stream>>>(d_out_col.p, n, static_cast<unsigned long long>(n) * 8ull,
                                                                                           d_out_text.p + static_cast<size_t>(v) * chars, ngroups);
      launches += 2;
      offsets[v] = static_cast<int64_t>(v) * chars;
    }
    offsets[NVAR] = static_cast<int64_t>(NVAR) * chars;
    CUDA_TRY(cudaGetLastError());
    int slow = 0;
    CUDA_TRY(cudaMemcpyAsync(text, d_out_text.p, static_cast<size_t>(NVAR) * chars, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(&slow, d_out_slow.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    d2h_bytes += static_cast<int64_t>(NVAR) * chars;
    if(slow) return fail(LBM_B200_EUNSUP, "a value is outside the range of the device's exact 15-decimal rounding (not finite or |x| >= 9.007): use the host writer");
    return LBM_B200_OK;
  }

  int get_vars(double* v, double* vo) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(cfg.track_vars <= 0) return fail(LBM_B200_ESTATE, "m_vars is not tracked (config.track_vars = 0)");
    if(vars_step[vcur] != t) return fail(LBM_B200_ESTATE, "m_vars of this step was not kept (track_vars interval)");
    if(v != nullptr) {
      int rc = download(vars[vcur].p, NVAR, v, false);
      if(rc) return rc;
    }
    if(vo != nullptr) {
      if(t == 0) {
        std::memset(vo, 0, sizeof(double) * static_cast<size_t>(plan.n) * NVAR); // solver.cpp:270
      } else {
        if(vars_step[vcur ^ 1] != t - 1) return fail(LBM_B200_ESTATE, "m_varsold of this step was not kept (track_vars interval)");
        return download(vars[vcur ^ 1].p, NVAR, vo, false);
      }
    }
    return LBM_B200_OK;
  }

  // The moments output() wants (solver.cpp:336: updateMacroscopicValues of the current m_fold) are what the next time step's kernels
  // compute on their way: run them once with the population stores switched off (B = nullptr) and m_vars going to the scratch buffer.
  int get_moments(double* m) override {
    if(!inited) return fail(LBM_B200_ESTATE, "not initialised");
    if(halo_pending) {
      CUDA_TRY(cudaStreamWaitEvent(stream, ev_halo, 0));
      halo_pending = false;
    }
    lbm::DevParams<Real> p = params(cur, cur ^ 1, scratch.p);
    if(prev_fold.p != nullptr) p.A = prev_fold.p;
    p.B = nullptr;
    int rc = launch_main(p, 0, plan.n_gen, 0, plan.n_fast_chunks, max_resident, 0);
    if(rc) return rc;
    CUDA_TRY(cudaGetLastError());
    return download(scratch.p, NVAR, m, false);
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
  std::vector<int32_t> dbg_copy, dbg_abb_cells, dbg_layout;
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
    v->perm_end = plan.perm_end; v->gb_begin = plan.gb_begin; v->gb_end = plan.gb_end;
    dbg_layout.assign(plan.L.lay, plan.L.lay + plan.L.Q);
    v->layout = dbg_layout.data();
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
} // namespace lbm_impl
