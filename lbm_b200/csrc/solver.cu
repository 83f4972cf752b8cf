// solver.cu -- C ABI (include/lbm_b200.h) on top of the device plan (plan.hpp) and the kernels (kernels.cuh).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lbm_b200.h"
#include "solver_common.cuh"
#include "grid_box.hpp"
#include "nccl_dyn.hpp"
#include "sequential.cuh"
#include "poisson.cuh"
#include "partition.hpp"

namespace lbm_impl {
thread_local std::string g_error;
}
using namespace lbm_impl;

namespace {

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
    if(cfg.collision == LBM_B200_MRT) { // base rate + per-moment differences, as Solver::params (solver_fused.cuh)
      const double s0 = lbm::mrt_base_rate(cfg.mrt_rates, Q, D);
      s.omega = s0;
      for(int i = 0; i < 27; ++i) s.rates[i] = (i > D && i < Q) ? (cfg.mrt_rates[i] - s0) / lbm::MrtBasis<L>::norm(i) : 0.0;
    }
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
        // The reference applies the entries one after the other, so when several entries write one slot (in 3D many entries resolve to
        // the same first connected cell, bnd_periodic.h:75-91) the LAST one wins.  The kernel runs one thread per entry: only the
        // winning writer of every slot may remain, decided here once.
        {
          const bool with_p = !std::isnan(bc.pressure);
          std::unordered_map<int64_t, int64_t> last; // slot (or linked cell) -> packed (entry, index in entry)
          for(int64_t k = 0; k < n; ++k) {
            if(with_p) last[link[k * Q]] = k;        // bnd_periodic.h:101-108: all Q populations of the first linked cell
            else for(int id = 0; id < nset[k]; ++id) last[link[k * Q + id] * Q + ldist[k * Q + id]] = k * Q + id;
          }
          for(int64_t k = 0; k < n; ++k) {
            if(with_p) { if(last[link[k * Q]] != k) nset[k] = -1; }
            else for(int id = 0; id < nset[k]; ++id)
              if(last[link[k * Q + id] * Q + ldist[k * Q + id]] != k * Q + id) ldist[k * Q + id] = -1;
          }
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
  if(c.ndim == 2 && c.ndist == 9) return dbl ? make_fused_d2q9_f64() : make_fused_d2q9_f32();
  if(c.ndim == 3 && c.ndist == 19) return dbl ? make_fused_d3q19_f64() : make_fused_d3q19_f32();
  if(c.ndim == 3 && c.ndist == 27) return dbl ? make_fused_d3q27_f64() : make_fused_d3q27_f32();
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

int lbm_b200_mrt_moments(int32_t ndim, int32_t ndist, int32_t* kinds) {
  if(kinds == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  if(ndim == 2 && ndist == 9) { for(int k = 0; k < 9; ++k) kinds[k] = lbm::MrtBasis<lbm::Lattice<2, 9>>::kind(k); return LBM_B200_OK; }
  if(ndim == 3 && ndist == 19) { for(int k = 0; k < 19; ++k) kinds[k] = lbm::MrtBasis<lbm::Lattice<3, 19>>::kind(k); return LBM_B200_OK; }
  if(ndim == 3 && ndist == 27) { for(int k = 0; k < 27; ++k) kinds[k] = lbm::MrtBasis<lbm::Lattice<3, 27>>::kind(k); return LBM_B200_OK; }
  return fail(LBM_B200_EINVAL, "MRT moment bases exist for D2Q9, D3Q19 and D3Q27");
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

int lbm_b200_p2p_export(lbm_b200_solver* s, void* blob) {
  CHECK_HANDLE(s);
  if(blob == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  return s->impl->p2p_export(blob);
}

int lbm_b200_p2p_import(lbm_b200_solver* s, int32_t nranks, const void* blobs) {
  CHECK_HANDLE(s);
  if(blobs == nullptr || nranks < 1) return fail(LBM_B200_EINVAL, "bad argument");
  return s->impl->p2p_import(nranks, blobs);
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
    // everything that can be refused is refused BEFORE the handle changes hands: a failed init leaves the set-up calls intact
    if(s->impl->cfg.precision != LBM_B200_FP64) return fail(LBM_B200_EUNSUP, "the Poisson equation types run in fp64 only");
    if(s->impl->in.n_ghost > 0 || !s->impl->in.peers.empty() || s->impl->comm != nullptr)
      return fail(LBM_B200_EUNSUP, "the Poisson equation types are not partitioned");
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
    if(s->impl->cfg.precision != LBM_B200_FP64) return fail(LBM_B200_EUNSUP, "wet-node walls (equilibrium / NEEM / NEBB) run in fp64 only");
    if(s->impl->in.n_ghost > 0 || !s->impl->in.peers.empty() || s->impl->comm != nullptr)
      return fail(LBM_B200_EUNSUP, "wet-node wall boundary conditions are not partitioned yet");
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

int64_t lbm_b200_output_chars(int64_t n_values) { return n_values > 0 ? (8 + 8 * n_values + 2) / 3 * 4 : 0; }

int lbm_b200_encode_output(lbm_b200_solver* s, const uint8_t* keep, char* text, int64_t capacity, int64_t* offsets) {
  CHECK_HANDLE(s);
  if(text == nullptr || offsets == nullptr) return fail(LBM_B200_EINVAL, "null argument");
  return s->impl->encode_output(keep, text, capacity, offsets);
}

int lbm_b200_host_alloc(void** ptr, int64_t bytes) {
  if(ptr == nullptr || bytes <= 0) return fail(LBM_B200_EINVAL, "lbm_b200_host_alloc: null pointer or non-positive size");
  *ptr = nullptr;
  const cudaError_t e = cudaMallocHost(ptr, static_cast<size_t>(bytes));
  if(e != cudaSuccess) {
    *ptr = nullptr;
    return fail(LBM_B200_ECUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
  }
  return LBM_B200_OK;
}

int lbm_b200_host_free(void* ptr) {
  if(ptr == nullptr) return LBM_B200_OK;
  const cudaError_t e = cudaFreeHost(ptr);
  return e == cudaSuccess ? LBM_B200_OK : fail(LBM_B200_ECUDA, std::string("cudaFreeHost: ") + cudaGetErrorString(e));
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
