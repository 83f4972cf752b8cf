// lbm_solver.hpp -- host-side mirror of the reference's solver interface, driving the CUDA core through the C ABI.
//
// Same names, argument meaning and error behaviour as the reference:
//   Runnable        /root/reference/src/interface/solver_interface.h:6-24  (init / initBenchmark / run / grid / transferGrid)
//   GridGenerator   /root/reference/src/gridgenerator/gridGenerator.cpp:72-229 (configuration keys :124-143,240-281)
//   LBMSolver       /root/reference/src/lbm/solver.cpp: loadConfiguration :71-163, run :176-214, writeInfo :217-230,
//                   convergenceCondition :233-263, output :323-384, compareToAnalyticalResult :388-482
//   boundary set-up /root/reference/src/lbm/bnd/bnd.h:71-142 (lexicographic order, empty surfaces skipped, dummies)
// The time step itself, the boundary kernels, forcing and the residual run on the GPU (include/lbm_b200.h); there is no CPU
// path here.  TERMM(code, msg) of the reference (src/common/term.h:37) becomes a TermError that main() turns into the same
// stderr text and exit status.
#pragma once
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../../include/lbm_b200.h"
#include "grid.hpp"
#include "json.hpp"
#include "run_log.hpp"
#include "vtk_writer.hpp"
#include "expr.hpp"
#include "uniform_grid.hpp"
#include <thread>
#include <unistd.h>

namespace lbmhost {

struct TermError : std::runtime_error {
  int code;
  TermError(int c, const std::string& where, const std::string& msg) : std::runtime_error("Error in " + where + ": " + msg), code(c) {}
};
#define LBMHOST_STR2(x) #x
#define LBMHOST_STR(x) LBMHOST_STR2(x)
#define TERMM(code, msg) throw ::lbmhost::TermError((code), std::string(__FILE__ ":" LBMHOST_STR(__LINE__)), (msg))

// gcem::sqrt as the reference uses it in compareToAnalyticalResult (src/lbm/solver.cpp:445): a Newton-Raphson iteration that
// treats |x| < epsilon as zero (vendored gcem, external/gcem_incl/sqrt.hpp:36-73).  Restated because it decides pass/fail:
// for the Couette cases the squared error sum is ~1e-24, the reference's L2 error is therefore exactly 0 and the cases'
// errorL2 thresholds (e.g. 1e-13) rely on that.
inline double gcem_sqrt(double x) {
  const double eps = std::numeric_limits<double>::epsilon();
  if(std::isnan(x) || x < 0) return std::numeric_limits<double>::quiet_NaN();
  if(std::isinf(x)) return x;
  if(eps > std::abs(x)) return 0.0;
  if(eps > std::abs(1.0 - x)) return x;
  double m = 1.0;
  while(x > 4.0) { x /= 4.0; m *= 2.0; }
  double xn = x / 2.0;
  for(int count = 0;; ++count) {
    if(std::abs(xn - x / xn) / (1.0 + xn) < eps || count >= 100) break;
    xn = 0.5 * (xn + x / xn);
  }
  return m * xn;
}

// One process per GPU.  The reference starts its ranks with mpirun and reads MPI_Comm_rank / MPI_Comm_size (src/main.cpp:283-300); this
// host has no MPI and takes rank / world size / local rank from the launcher's environment instead: LBM_B200_RANK / LBM_B200_WORLD /
// LBM_B200_LOCAL_RANK, else torchrun's RANK / WORLD_SIZE / LOCAL_RANK, else Open MPI's or PMI's variables (so `mpirun -np N lbm` works
// when an MPI launcher is around).  The 128-byte NCCL id travels through a file (LBM_B200_ID_FILE, default /tmp/lbm_b200_id_<port or ppid>).
struct RankInfo {
  int         rank = 0, world = 1, local = 0;
  std::string id_file;
  static int env_int(const char* name, int fallback) {
    const char* v = std::getenv(name);
    return v != nullptr && *v != 0 ? std::atoi(v) : fallback;
  }
  // Only the `lbm` executable reads the launcher's environment (main.cpp switches this on): code that embeds the host library -- the
  // Python tests and bench.py run under torchrun, whose WORLD_SIZE must not turn a grid build into a partitioned run -- stays single-rank
  // unless it sets the rank explicitly.
  static bool& use_environment() {
    static bool on = false;
    return on;
  }
  static RankInfo from_env() {
    RankInfo r;
    if(!use_environment()) return r;
    const char* sets[4][3] = {{"LBM_B200_RANK", "LBM_B200_WORLD", "LBM_B200_LOCAL_RANK"}, {"RANK", "WORLD_SIZE", "LOCAL_RANK"},
                              {"OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_RANK"}, {"PMI_RANK", "PMI_SIZE", "MPI_LOCALRANKID"}};
    for(auto& s : sets) {
      if(std::getenv(s[1]) == nullptr) continue;
      r.world = env_int(s[1], 1);
      r.rank  = env_int(s[0], 0);
      r.local = env_int(s[2], r.rank);
      break;
    }
    if(r.world < 1 || r.rank < 0 || r.rank >= r.world) throw std::runtime_error("Invalid rank / world size in the environment");
    const char* f = std::getenv("LBM_B200_ID_FILE");
    if(f != nullptr) r.id_file = f;
    else {
      const char* port = std::getenv("MASTER_PORT");
      r.id_file = std::string("/tmp/lbm_b200_id_") + (port != nullptr ? std::string(port) : std::to_string(static_cast<long>(getppid())));
      // a per-launch nonce when the launcher provides one: two runs that reuse a port never read each other's file
      for(const char* name : {"LBM_B200_RUN_ID", "TORCHELASTIC_RUN_ID", "SLURM_JOB_ID"}) {
        const char* v = std::getenv(name);
        if(v != nullptr && v[0] != 0) {
          std::string clean;
          for(const char* c = v; *c; ++c) clean += (std::isalnum(static_cast<unsigned char>(*c)) ? *c : '_');
          r.id_file += "_" + clean.substr(0, 64);
          break;
        }
      }
    }
    return r;
  }
};

class GridInterface {
 public:
  virtual ~GridInterface()       = default;
  virtual int     dim() const     = 0;
  virtual int64_t noCells() const = 0;
  virtual int     maxLvl() const  = 0;
};

class Runnable {
 public:
  virtual ~Runnable() = default;
  virtual void                 init(int argc, char** argv, std::string config_file) = 0;
  virtual void                 initBenchmark(int argc, char** argv)                 = 0;
  virtual int64_t              run()                                                = 0;
  virtual const GridInterface& grid() const                                         = 0;
  virtual void                 transferGrid(const GridInterface& grid)              = 0;
};

class GeneratedGrid final : public GridInterface {
 public:
  GridGen gen;
  int     dim() const override { return gen.ndim; }
  int64_t noCells() const override { return static_cast<int64_t>(gen.cells.size()); }
  int     maxLvl() const override { return gen.level; }
};

class GridGenerator final : public Runnable {
 public:
  void init(int argc, char** argv, std::string config_file) override {
    RunRecord& rec = RunRecord::get();
    rec.timers.start(rec.gridTotal);
    rec.timers.start(rec.gridInit);
    const RankInfo rank = RankInfo::from_env();
    if(rec.enabled) rec.grid_log.open("gridgen_log", argc, argv, rank.rank, rank.world); // gridGenerator.cpp:22-29
    m_config = Json::parse_file(config_file);
    std::cout << "Grid generator started ||>" << std::endl;
    rec.grid_log("Grid generator started ||>");
    rec.grid_log("Loading configuration file [" + config_file + "]");
    rec.timers.stop(rec.gridInit);
  }
  void initBenchmark(int /*argc*/, char** /*argv*/) override {
    // gridGenerator.cpp:43-54: 3D, uniform level 5, default cube geometry
    m_config = Json::parse(R"({"dim":3,"partitionLevel":5,"uniformLevel":5,"maxNoCells":100000,
      "geometry":{"cube":{"type":"box","A":[0.0,0.0,0.0],"B":[1.0,1.0,1.0]}}})");
  }
  int64_t run() override {
    RunRecord& rec = RunRecord::get();
    rec.timers.start(rec.gridCreate);
    m_grid.gen.configure(m_config);
    rec.grid_log("Generating a grid[" + std::to_string(m_grid.gen.ndim) + "D]");
    if(RankInfo::from_env().world > 1) {
      // partitioned run: no rank builds the whole tree; the solver asks the on-demand provider (uniform_grid.hpp) for its own rows
      std::cout << "    * partitioned run: grid rows are generated per rank" << std::endl;
      rec.grid_log("    * partitioned run: grid rows are generated per rank");
      finish(rec);
      return 0;
    }
    m_grid.gen.generate();
    std::cout << "    * grid has " << m_grid.noCells() << " cells" << std::endl;
    rec.grid_log("      * grid has " + std::to_string(m_grid.noCells()) + " cells");
    const long long maxc = m_config.opt_int("maxNoCells", -1);
    if(maxc >= 0 && m_grid.noCells() > maxc) TERMM(-1, "Out of memory!"); // cartesiangrid_generation.h:336
    finish(rec);
    return 0;
  }
  const GridInterface& grid() const override { return m_grid; }
  void transferGrid(const GridInterface&) override { TERMM(-1, "Not implemented!"); }
  const Json& config() const { return m_config; }

 private:
  // gridGenerator.cpp:88-100: the generator's log ends with the timer table as it stands when the generator is done
  static void finish(RunRecord& rec) {
    rec.timers.stop(rec.gridCreate);
    rec.timers.stop(rec.gridTotal);
    rec.grid_log("Grid generator finished <||");
    rec.timers.display(rec.grid_log);
    rec.grid_log.close();
  }
  Json          m_config;
  GeneratedGrid m_grid;
};

class LBMGrid final : public GridInterface {
 public:
  SolverGrid g;
  int     dim() const override { return g.ndim; }
  int64_t noCells() const override { return g.n; }
  int     maxLvl() const override { return g.max_level; }
};

class LBMSolver final : public Runnable {
 public:
  ~LBMSolver() override {
    lbm_b200_host_free(m_encBuf);
    if(m_gpu != nullptr) lbm_b200_destroy(m_gpu);
    if(m_part != nullptr) lbm_b200_partition_destroy(m_part);
  }

  void init(int argc, char** argv, std::string config_file) override {
    RunRecord& rec = RunRecord::get();
    rec.createLbmTimers();
    rec.timers.start(rec.lbmTotal);
    rec.timers.start(rec.lbmInit);
    m_configFile = config_file;
    {
      const RankInfo rank = RankInfo::from_env();
      if(rec.enabled) rec.lbm_log.open("lbm_log", argc, argv, rank.rank, rank.world); // solver.cpp:13-20
    }
    const Json all = Json::parse_file(config_file);
    if(!all.has("solver")) TERMM(-1, "The required configuration value is missing: solver");
    m_cfg = all.at("solver");
    // solverExe.h:19-93
    m_model = m_cfg.opt_str("model", "D2Q9");
    m_equation = m_cfg.opt_str("equation", "navierstokes");
    if(m_equation != "navierstokes" && m_equation != "poisson") {
      if(m_equation == "navierstokespoisson") TERMM(-1, "Unsupported equation type"); // like the reference: solverExe.h:43,53,69 (every Navier_Stokes_Poisson case is commented out)
      TERMM(-1, "Invalid equation configuration!"); // constants.h:36-47
    }
    const bool poisson = m_equation == "poisson";
    if(m_model == "D2Q9") { m_ndim = 2; m_ndist = 9; }
    else if(m_model == "D3Q19") { m_ndim = 3; m_ndist = 19; }
    else if(m_model == "D3Q27") { m_ndim = 3; m_ndist = 27; }
    else if(m_model == "D1Q3") { m_ndim = 1; m_ndist = 3; }
    else if(m_model == "D2Q5") { m_ndim = 2; m_ndist = 5; }
    else TERMM(-1, "Invalid model configuration!");
    // solverExe.h:29-90: D1Q3 / D2Q5 exist for the Poisson equation only; the 3D models are this host's extension (Navier-Stokes only)
    if((m_ndist == 3 || m_ndist == 5) != poisson && !(poisson && m_ndist == 9)) TERMM(-1, "Unsupported model");
    m_rank = RankInfo::from_env();
    m_startTime = ::time(nullptr);
    if(m_rank.world > 1 && m_rank.rank == 0) std::remove(m_rank.id_file.c_str()); // nothing stale may be picked up by the other ranks
    std::cout << m_ndim << "D LBM Solver started ||>" << std::endl;
    rec.lbm_log(std::to_string(m_ndim) + "D LBM Solver started ||>"); // solver.cpp:29-30
    rec.lbm_log("Loading configuration file [" + config_file + "]");
  }

  void initBenchmark(int /*argc*/, char** /*argv*/) override {
    // the reference: TERMM(-1, "Not implemented!") (solver.cpp:39-46). Here: the synthetic cube of SURVEY.md section 8d S3.
    m_benchmark = true;
    m_ndim = 3; m_ndist = 19; m_model = "D3Q19";
  }

  void transferGrid(const GridInterface& grid) override {
    say("Transferring " + std::to_string(m_ndim) + "D Grid to LBM solver");
    if(grid.dim() != m_ndim) TERMM(-1, "Invalid configuration the grid dimensionality is not matching!");
    const auto* gen = dynamic_cast<const GeneratedGrid*>(&grid);
    if(gen == nullptr) TERMM(-1, "transferGrid expects the generator's grid");
    try {
      if(m_rank.world > 1) {
        m_ugrid.configure(Json::parse_file(m_configFile));
        m_grid.g.ndim = m_ugrid.ndim;
        m_grid.g.max_level = m_ugrid.level;
        m_grid.g.n = 0;
      } else {
        m_grid.g.load(gen->gen, m_cfg);
      }
    } catch(const std::runtime_error& e) {
      TERMM(-1, e.what());
    }
  }

  // multi-GPU set-up of one rank up to (not including) the NCCL bootstrap; with device == -1 an inspection-only handle that the tests
  // compare with the Python path's plan (tests/test_host_partitioned.py).  The caller owns the returned handle.
  lbm_b200_solver* buildPartitioned(int device) {
    loadConfiguration();
    setupGpuPartitioned(device);
    lbm_b200_solver* h = m_gpu;
    m_gpu = nullptr;
    return h;
  }
  void setRank(int rank, int world) { m_rank.rank = rank; m_rank.world = world; m_rank.local = rank; }

  const GridInterface& grid() const override { return m_grid; }
  const SolverGrid& solverGrid() const { return m_grid.g; }
  int noVars() const { return m_equation == "poisson" ? 1 : m_ndim + 1; }

  // LBMBndManager::addPeriodicBndry marks the cells of both surfaces of a periodic boundary CONDITION as periodic
  // (src/lbm/bnd/bnd.h:217-221, CellProperties::periodic = bit 0); dummies (generateBndry:false) are left alone
  void markBoundaryProperties() {
    SolverGrid& g = m_grid.g;
    for(const auto& gk : m_cfg.at("boundary").obj) {
      const size_t nkeys = gk.second.size();
      for(const auto& sk : gk.second.obj) {
        const Json& bc = sk.second;
        if(bc.opt_str("type", "") != "periodic" || !bc.opt_bool("generateBndry", true)) continue;
        const Surface* a = g.find(nkeys > 1 ? gk.first + "_" + sk.first : gk.first);
        const Surface* b = g.find(bc.at("connection").as_string());
        if(a == nullptr || a->cells.empty() || b == nullptr) continue;
        for(int64_t c : a->cells) g.props[c] |= 1u;
        for(int64_t c : b->cells) g.props[c] |= 1u;
      }
    }
  }

  // results for callers that embed the solver (tests): macroscopic fields of the last output(), errors of the analytic check
  std::vector<double> vars;
  bool                m_varsFresh = false;
  char*               m_encBuf    = nullptr; // page-locked text of the output fields (lbm_b200_encode_output)
  int64_t             m_encCap    = 0;
  double  maxError = NAN, l2Error = NAN, gre = NAN;
  int64_t stepsRun = 0;
  bool    converged = false;
  double  lastResidual = NAN;

  // `lbm --bench`: the synthetic cube of SURVEY.md section 8d (S3) through the same C ABI: S^3 cells (LBM_BENCH_SIZE, default 256),
  // D3Q19 BGK fp64, omega = 1/0.6, periodic x, bounce-back walls on -y/+y/-z, moving lid on +z; 50 warm-up + 200 timed steps.
  int64_t runBenchmark() {
    const char* env = std::getenv("LBM_BENCH_SIZE");
    const int64_t S = env != nullptr ? std::atoll(env) : 256;
    if(S < 8 || S > 640) TERMM(-1, "LBM_BENCH_SIZE must be in [8, 640]");
    const int64_t shape[3] = {S, S, S};
    const int32_t periodic[3] = {1, 0, 0};
    const int64_t n = lbm_b200_box_ncells(3, shape);
    std::vector<int64_t> nghbr(static_cast<size_t>(n) * 26);
    call(lbm_b200_box_topology(3, shape, periodic, nghbr.data(), 26, nullptr, nullptr));
    lbm_b200_config cfg;
    lbm_b200_default_config(&cfg);
    cfg.ndim = 3;
    cfg.ndist = 19;
    cfg.arithmetic = LBM_B200_FAST;
    cfg.track_vars = 0;
    cfg.omega = 1.0 / 0.6;
    call(lbm_b200_create(&cfg, n, &m_gpu));
    call(lbm_b200_set_topology(m_gpu, nghbr.data(), 26));
    // surfaces in the reference's (lexicographic) order: +y, +z, -y, -z; x is periodic
    const int order[4] = {3, 5, 2, 4};
    for(int dir : order) {
      std::vector<int64_t> cells;
      for(int64_t c = 0; c < n; ++c)
        if(nghbr[static_cast<size_t>(c) * 26 + dir] < 0) cells.push_back(c);
      std::vector<double> normals(cells.size() * 3, 0.0);
      for(size_t k = 0; k < cells.size(); ++k) normals[k * 3 + dir / 2] = dir % 2 ? 1.0 : -1.0;
      if(dir == 5) {
        const double lid[3] = {0.05, 0.0, 0.0};
        call(lbm_b200_add_dirichlet_bb(m_gpu, cells.data(), normals.data(), static_cast<int64_t>(cells.size()), lid));
      } else {
        call(lbm_b200_add_wall_bb(m_gpu, cells.data(), normals.data(), static_cast<int64_t>(cells.size()), 0.0));
      }
    }
    std::vector<int64_t>().swap(nghbr);
    call(lbm_b200_init(m_gpu));
    call(lbm_b200_step(m_gpu, 50));
    float ms_total = 0, ms_main = 0;
    const int64_t steps = 200;
    call(lbm_b200_step_timed(m_gpu, steps, &ms_total, &ms_main));
    lbm_b200_stats st;
    call(lbm_b200_get_stats(m_gpu, &st));
    const double mlups = static_cast<double>(n) * steps / (ms_total * 1e-3) / 1e6;
    const double gbs   = st.bytes_per_cell_alg * static_cast<double>(n) * steps / (ms_main * 1e-3) / 1e9;
    std::cout << "bench: " << S << "^3 D3Q19 BGK fp64, " << n << " cells (" << st.cells_fast << " on the chunk path), " << steps << " steps in "
              << ms_total << " ms\n"
              << "bench: " << mlups << " MLUPS, " << gbs << " GB/s algorithmic (2*Q*8 B per cell update)" << std::endl;
    stepsRun = steps;
    return 0;
  }

  int64_t run() override {
    if(m_benchmark) return runBenchmark();
    loadConfiguration();
    if(m_rank.world > 1) return runPartitioned();
    RunRecord& rec = RunRecord::get();
    rec.createLbmTimers();
    initPostprocess();
    setupGpu();
    vars.assign(static_cast<size_t>(m_grid.g.n) * nvar(), 0.0);
    rec.timers.stop(rec.lbmInit);
    rec.timers.start(rec.lbmMain);
    rec.timers.start(rec.lbmPost);
    executePostprocess(PP_ATSTART);
    rec.timers.stop(rec.lbmPost);
    using clk = std::chrono::steady_clock;
    auto    lastInfo = clk::now();
    int64_t lastStep = 0;
    for(m_timeStep = 0; m_timeStep < m_maxTimeStep && !converged; ++m_timeStep) {
      // writeInfo, solver.cpp:217-230
      if(m_timeStep > 0 && m_timeStep % m_infoInterval == 0) {
        const double dt = std::chrono::duration<double>(clk::now() - lastInfo).count();
        std::cerr << m_timeStep << "/" << m_maxTimeStep << " " << (m_timeStep - lastStep) / dt << "it/s \n";
        lastInfo = clk::now();
        lastStep = m_timeStep;
      }
      rec.timers.start(rec.lbmComp);
      converged = convergenceCondition();
      call(lbm_b200_step(m_gpu, 1));
      const bool last = m_timeStep == m_maxTimeStep - 1 || converged;
      // the launches are asynchronous: before a file is written the device catches up, so that "Computation" holds the device's time
      // and "IO" only the output
      if(outputDue(last)) call(lbm_b200_synchronize(m_gpu));
      rec.timers.stop(rec.lbmComp);
      rec.timers.start(rec.lbmIO);
      output(last);
      rec.timers.stop(rec.lbmIO);
    }
    stepsRun = m_timeStep;
    rec.timers.start(rec.lbmPost);
    fetchVars(); // the final state for the atEnd hooks, the analytic comparison and the callers of vars (output() leaves it on the device)
    executePostprocess(PP_ATEND);
    rec.timers.stop(rec.lbmPost);
    rec.timers.stop(rec.lbmMain);
    if(m_diverged) TERMM(-1, "Solution diverged");
    if(m_cfg.has("analyticalSolution")) compareToAnalyticalResult();
    std::cout << "LBM Solver finished <||" << std::endl;
    finishLog();
    return 0;
  }

  // the end of lbm_log: the closing message and the timer table (main.cpp:289-300 of the reference prints it when the run is over)
  void finishLog() {
    RunRecord& rec = RunRecord::get();
    rec.createLbmTimers();
    rec.timers.stop(rec.lbmTotal);
    rec.lbm_log("LBM Solver finished <||");
    rec.timers.display(rec.lbm_log);
    rec.lbm_log.close();
  }
  // a line on stderr that the reference also writes to its log
  static void say(const std::string& text) {
    std::cerr << text << std::endl;
    RunRecord::get().lbm_log(text);
  }
  bool outputDue(bool forced) const {
    return (m_diverged && !m_wroteDiverged) || (m_timeStep > 0 && m_timeStep % m_solutionInterval == 0) || forced;
  }

 private:
  // ---- one rank of a partitioned run (the reference has no decomposition to mirror) --------------------------
  static int rowsCallback(void* user, const int64_t* ids, int64_t n, int64_t* rows, int64_t* sources) {
    const auto* u = static_cast<const UniformGrid*>(user);
    std::string err;
    if(rows != nullptr && !u->rows(ids, n, rows, u->nn_diag, nullptr, &err)) return 1;
    if(sources != nullptr && !u->sources(ids, n, sources, u->nn_diag, &err)) return 1;
    return 0;
  }

  void setupGpuPartitioned(int device) {
    if(poisson()) TERMM(-1, "the Poisson equation types are not partitioned");
    if(!m_cfg.opt_str("forcing", "").empty()) TERMM(-1, "forcing is not partitioned");
    UniformGrid& u = m_ugrid;
    if(u.n == 0) TERMM(-1, "partitioned run without a grid (transferGrid was not called)");
    u.build_surfaces();
    const Json& boundary = m_cfg.at("boundary");
    struct BcRef { const Json* conf; const Surface* srf; std::vector<double> normals; };
    std::vector<BcRef> order;
    for(const auto& gk : boundary.obj) {
      const size_t nkeys = gk.second.size();
      for(const auto& sk : gk.second.obj) {
        const std::string sname = nkeys > 1 ? gk.first + "_" + sk.first : gk.first;
        const Surface* srf = nullptr;
        for(const Surface& s : u.surfaces)
          if(s.name == sname) srf = &s;
        if(srf == nullptr) TERMM(-1, "Invalid bndryId \"" + sname + "\"");
        if(srf->cells.empty() || !sk.second.opt_bool("generateBndry", true)) continue;
        BcRef r{&sk.second, srf, {}};
        for(int64_t c : srf->cells)
          for(int d = 0; d < m_ndim; ++d) r.normals.push_back(srf->normal.at(c)[d]);
        order.push_back(std::move(r));
      }
    }
    // the pressure surfaces' GLOBAL lists, in application order: the velocity halo of their inward neighbours
    std::vector<const int64_t*> pcells;
    std::vector<const double*>  pnormals;
    std::vector<int64_t>        pcount;
    for(const BcRef& r : order)
      if(r.conf->at("type").as_string() == "pressure") {
        pcells.push_back(r.srf->cells.data());
        pnormals.push_back(r.normals.data());
        pcount.push_back(static_cast<int64_t>(r.srf->cells.size()));
      }
    call(lbm_b200_partition_create(u.n, m_ndim, m_ndist, u.nn_diag, m_rank.rank, m_rank.world, &LBMSolver::rowsCallback, &u,
                                   static_cast<int32_t>(pcells.size()), pcells.data(), pnormals.data(), pcount.data(), &m_part));
    lbm_b200_partition_view v;
    call(lbm_b200_partition_get(m_part, &v));
    m_nOwned = v.n_owned;
    m_nLocal = v.n_owned + v.n_ghost;
    m_lo     = v.lo;
    lbm_b200_config cfg = gpuConfig();
    cfg.device = device;
    call(lbm_b200_create(&cfg, m_nLocal, &m_gpu));
    call(lbm_b200_set_topology(m_gpu, v.nghbr, v.stride));
    for(const BcRef& r : order) {
      const Json&       bc   = *r.conf;
      const std::string type = bc.at("type").as_string();
      const int64_t     n    = static_cast<int64_t>(r.srf->cells.size());
      std::vector<int64_t> local(static_cast<size_t>(n)), index(static_cast<size_t>(n));
      const bool wall_bb = type == "wall" && bc.at("model").as_string() == "bounceback";
      const bool dir_bb  = type == "dirichlet" && bc.at("model").as_string() == "bounceback";
      if(!wall_bb && !dir_bb && type != "pressure") // every rank refuses, whether or not it owns a cell of the surface
        TERMM(-1, "boundary condition type " + type + " is not partitioned (wall/bounceback, dirichlet/bounceback, pressure are)");
      const int64_t m = lbm_b200_partition_restrict(m_part, r.srf->cells.data(), n, local.data(), index.data());
      if(m == 0) continue;
      std::vector<double> normals(static_cast<size_t>(m) * m_ndim);
      for(int64_t k = 0; k < m; ++k)
        for(int d = 0; d < m_ndim; ++d) normals[k * m_ndim + d] = r.normals[index[k] * m_ndim + d];
      if(type == "wall" && bc.at("model").as_string() == "bounceback") {
        call(lbm_b200_add_wall_bb(m_gpu, local.data(), normals.data(), m, bc.opt("tangentialVelocity", 0.0)));
      } else if(type == "dirichlet" && bc.at("model").as_string() == "bounceback") {
        const auto val = bc.at("value").as_doubles();
        if(static_cast<int>(val.size()) < m_ndim) TERMM(-1, "dirichlet value needs one entry per dimension");
        call(lbm_b200_add_dirichlet_bb(m_gpu, local.data(), normals.data(), m, val.data()));
      } else if(type == "pressure") {
        call(lbm_b200_add_pressure(m_gpu, local.data(), normals.data(), m, bc.at("pressure").as_double()));
      } else {
        TERMM(-1, "boundary condition type " + type + " is not partitioned (wall/bounceback, dirichlet/bounceback, pressure are)");
      }
    }
    call(lbm_b200_partition_apply(m_part, m_gpu));
  }

  // NCCL bootstrap without MPI: rank 0 writes the 128-byte id next to a temporary name and renames it, the others wait for the file
  void exchangeId(char* id) {
    const std::string path = m_rank.id_file;
    if(m_rank.rank == 0) {
      call(lbm_b200_comm_unique_id(id));
      const std::string tmp = path + ".tmp";
      std::ofstream o(tmp, std::ios::binary | std::ios::trunc);
      o.write(id, 128);
      o.close();
      if(!o || std::rename(tmp.c_str(), path.c_str()) != 0) TERMM(-1, "cannot write the NCCL id file " + path);
      return;
    }
    for(int tries = 0; tries < 6000; ++tries) { // up to 10 minutes
      struct stat st;
      // a file left behind by a run that died is older than this process (minus a margin for staggered starts): not ours
      if(::stat(path.c_str(), &st) == 0 && st.st_mtime + 120 >= m_startTime) {
        std::ifstream i(path, std::ios::binary);
        if(i && i.read(id, 128) && i.gcount() == 128) return;
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    TERMM(-1, "timed out waiting for the NCCL id file " + path);
  }

  // Peer-to-peer halo (lbm_b200_p2p_export / _import): the mailbox descriptions travel through files next to the NCCL id file, the
  // way the reference would MPI_Allgather them.  Every file starts with the run's NCCL id (unique per launch), so a file left behind by
  // an earlier run is never mistaken for this one's; a rank that cannot take part (velocity halo across a cut) says so in its file and
  // then ALL ranks stay on NCCL.  LBM_B200_HALO=nccl switches the mailboxes off.
  void connectPeerToPeer(const char* id) {
    const char* mode = std::getenv("LBM_B200_HALO");
    if(mode != nullptr && std::string(mode) == "nccl") return;
    std::vector<char> mine(128 + 1 + LBM_B200_P2P_BLOB, 0);
    std::memcpy(mine.data(), id, 128);
    mine[128] = lbm_b200_p2p_export(m_gpu, mine.data() + 129) == LBM_B200_OK ? 1 : 0;
    const std::string base = m_rank.id_file + ".p2p.";
    {
      const std::string path = base + std::to_string(m_rank.rank), tmp = path + ".tmp";
      std::ofstream o(tmp, std::ios::binary | std::ios::trunc);
      o.write(mine.data(), static_cast<std::streamsize>(mine.size()));
      o.close();
      if(!o || std::rename(tmp.c_str(), path.c_str()) != 0) TERMM(-1, "cannot write " + path);
    }
    std::vector<char> blobs(static_cast<size_t>(m_rank.world) * LBM_B200_P2P_BLOB, 0);
    bool all_able = true;
    for(int r = 0; r < m_rank.world; ++r) {
      const std::string path = base + std::to_string(r);
      std::vector<char> buf(mine.size());
      bool got = false;
      for(int tries = 0; tries < 6000 && !got; ++tries) { // up to 10 minutes
        std::ifstream i(path, std::ios::binary);
        if(i && i.read(buf.data(), static_cast<std::streamsize>(buf.size())) && std::memcmp(buf.data(), id, 128) == 0) got = true;
        else std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
      if(!got) TERMM(-1, "timed out waiting for " + path);
      all_able = all_able && buf[128] == 1;
      std::memcpy(blobs.data() + static_cast<size_t>(r) * LBM_B200_P2P_BLOB, buf.data() + 129, LBM_B200_P2P_BLOB);
    }
    if(all_able) call(lbm_b200_p2p_import(m_gpu, m_rank.world, blobs.data()));
    else if(m_rank.rank == 0) std::cerr << "peer-to-peer halo not available for this configuration: NCCL send / receive" << std::endl;
    // every rank has read every file once it has passed the residual's first all-reduce; rank r removes its own file at exit
    m_p2pFile = base + std::to_string(m_rank.rank);
  }

  int64_t runPartitioned() {
    std::cerr << "Rank " << m_rank.rank << " of " << m_rank.world << ": partitioned run" << std::endl;
    setupGpuPartitioned(m_rank.local);
    char id[128];
    exchangeId(id);
    call(lbm_b200_comm_init(m_gpu, id, m_rank.rank, m_rank.world));
    call(lbm_b200_init(m_gpu));
    connectPeerToPeer(id);
    if(m_rank.rank == 0) std::remove(m_rank.id_file.c_str()); // every rank has joined the communicator by now
    vars.assign(static_cast<size_t>(m_nLocal) * nvar(), 0.0);
    for(m_timeStep = 0; m_timeStep < m_maxTimeStep && !converged; ++m_timeStep) {
      if(m_rank.rank == 0 && m_timeStep > 0 && m_timeStep % m_infoInterval == 0) std::cerr << m_timeStep << "/" << m_maxTimeStep << " \n";
      converged = convergenceCondition(); // lbm_b200_residual is collective here: every rank takes the same decision
      call(lbm_b200_step(m_gpu, 1));
      outputPartitioned(m_timeStep == m_maxTimeStep - 1 || converged);
    }
    stepsRun = m_timeStep;
    if(m_diverged) TERMM(-1, "Solution diverged");
    if(m_cfg.has("analyticalSolution") && m_rank.rank == 0)
      std::cerr << "analyticalSolution is not evaluated in a partitioned run (it needs the fields of all ranks)" << std::endl;
    lbm_b200_destroy(m_gpu); // NCCL communicator teardown before the process exits
    m_gpu = nullptr;
    if(!m_p2pFile.empty()) std::remove(m_p2pFile.c_str());
    std::cout << "LBM Solver finished <||" << std::endl;
    finishLog();
    return 0;
  }

  // every rank writes the cells it owns: out/<solution_filename>_<step>_rank<r>.vtp, the reference's file format
  void outputPartitioned(bool forced) {
    if(!((m_timeStep > 0 && m_timeStep % m_solutionInterval == 0) || forced)) return;
    call(lbm_b200_get_moments(m_gpu, vars.data()));
    if(!m_cfg.opt_bool("write_output", true)) return;
    if(m_ownCenter.empty()) {
      std::vector<int64_t> ids(static_cast<size_t>(m_nOwned)), rows(static_cast<size_t>(m_nOwned) * m_ugrid.nn_diag);
      for(int64_t k = 0; k < m_nOwned; ++k) ids[k] = m_lo + k;
      m_ownCenter.resize(static_cast<size_t>(m_nOwned) * m_ndim);
      std::string err;
      if(!m_ugrid.rows(ids.data(), m_nOwned, rows.data(), m_ugrid.nn_diag, m_ownCenter.data(), &err)) TERMM(-1, err);
    }
    ::mkdir(m_outputDir.c_str(), 0755);
    const std::string stem = m_outputDir + m_solutionName + "_" + std::to_string(m_timeStep) + "_rank" + std::to_string(m_rank.rank);
    const int NVAR = nvar();
    static const char* names[4] = {"U", "V", "W", "rho"};
    std::vector<vtk::Column> cols;
    for(int v = 0; v < NVAR; ++v) cols.push_back(vtk::Column{v == m_ndim ? "rho" : names[v], vars.data() + v, NVAR});
    std::cerr << "  Writing " << stem << ".vtp with #" << m_nOwned << " cells" << std::endl;
    if(!vtk::write_points(stem + ".vtp", m_ndim, m_nOwned, m_ownCenter.data(), nullptr, cols))
      TERMM(-1, "Invalid output directory set! (value: " + m_outputDir + ")");
  }

  bool poisson() const { return m_equation == "poisson"; }
  // noVars<LBTYPE>(EQ), src/lbm/variables.h: velocity + density, or the potential alone
  int  nvar() const { return poisson() ? 1 : m_ndim + 1; }

  void call(int rc) {
    if(rc != 0) TERMM(-1, std::string("lbm_b200: ") + lbm_b200_last_error());
  }

  // solver.cpp:71-163
  void loadConfiguration() {
    const std::string method = m_cfg.opt_str("method", "bgk");
    if(method == "bgk") m_collision = LBM_B200_BGK;
    else if(method == "trt") m_collision = LBM_B200_TRT;   // extension
    else if(method == "mrt") m_collision = LBM_B200_MRT;   // extension
    else TERMM(-1, "Invalid equation configuration!");     // constants.h:75
    m_infoInterval = m_cfg.opt_int("info_interval", 10);
    m_convInterval = m_cfg.opt_int("conv_interval", m_infoInterval);
    if(!m_cfg.has("maxSteps")) TERMM(-1, "The required configuration value is missing: maxSteps");
    m_maxTimeStep = m_cfg.at("maxSteps").as_int();
    m_outputDir   = m_cfg.opt_str("output_dir", "out/");
    m_solutionInterval = m_cfg.opt_int("solution_interval", 100);
    m_solutionName     = m_cfg.opt_str("solution_filename", "solution");
    if(m_outputDir.empty()) TERMM(-1, "Invalid output directory set! (value: " + m_outputDir + ")");
    if(m_outputDir.back() != '/') m_outputDir += '/';
    m_refLength = m_cfg.opt("refLength", 1.0);
    // solver.cpp:100: m_finestGridSpacing = 1 / (size^(1/NDIM) - 1)
    const double finestGridSpacing = 1.0 / (std::pow(static_cast<double>(m_grid.g.n), 1.0 / m_ndim) - 1);
    const double lbm_cs = 1.0 / gcem_sqrt(3.0); // constants.h:26
    if(poisson()) { // solver.cpp:128-137
      if(method != "bgk") TERMM(-1, "Invalid equation configuration!");
      if(!m_cfg.has("relaxation")) TERMM(-1, "The required configuration value is missing: relaxation");
      m_relaxTime = m_cfg.at("relaxation").as_double();
      m_re        = 0;
      m_ma        = 1.0 / lbm_cs;
      m_omega     = 1.0 / m_relaxTime;
      m_nu        = (2.0 * m_relaxTime - 1) / 6.0;
      m_dt        = finestGridSpacing * m_ma * lbm_cs / m_refLength;
      // collisionStep, solver.cpp:589-599
      const std::string app = m_cfg.opt_str("equation_application", "debye_huckel");
      if(app == "simple_diff_reaction") {
        if(!m_cfg.has("equation_th")) TERMM(-1, "The required configuration value is missing: equation_th");
        m_poissonRate = m_cfg.at("equation_th").as_double();
      } else if(app == "debye_huckel") {
        m_poissonRate = 27.79;
      } else {
        TERMM(-1, "Invalid poisson equation application");
      }
      std::cerr << "<<<<<<<<<<<<>>>>>>>>>>>>>\nLBM Type " << method << "\nLBM Model " << m_model << "\nEquation poisson\nNo. of variables 1"
                << "\nNo. Leaf cells: " << m_grid.g.n_leaf << "\nMax Mesh Level: " << m_grid.g.max_level << "\nNo. Bnd cells: " << m_grid.g.n_bnd
                << "\nRelaxation Time: " << m_relaxTime << "\nTimestep: " << m_dt << "\nOmega: " << m_omega << "\nReynolds Number: " << m_re
                << "\nViscosity: " << m_nu << "\n+++++++++++++++++++++++++" << std::endl;
      return;
    }
    if(m_cfg.has("reynoldsnumber") && m_cfg.has("relaxation")) TERMM(-1, "Only set either reynoldsnumber or relaxation");
    if(!m_cfg.has("ma")) TERMM(-1, "The required configuration value is missing: ma");
    m_ma = m_cfg.at("ma").as_double();
    if(m_cfg.has("relaxation")) {
      m_relaxTime = m_cfg.at("relaxation").as_double();
      m_omega     = 1.0 / m_relaxTime;
      m_nu        = (2 * m_relaxTime - 1) / 6.0;
      m_re        = m_ma * m_refLength / m_nu;
    } else {
      if(!m_cfg.has("reynoldsnumber")) TERMM(-1, "The required configuration value is missing: reynoldsnumber");
      m_re        = m_cfg.at("reynoldsnumber").as_double();
      m_nu        = m_ma / m_re * m_refLength;
      m_omega     = 2.0 / (1.0 + 2.0 * m_nu * std::pow(2.0, m_grid.g.max_level)); // solver.cpp:119
      m_relaxTime = 1.0 / m_omega;
    }
    std::cerr << "<<<<<<<<<<<<>>>>>>>>>>>>>\nLBM Type " << method << "\nLBM Model " << m_model << "\nNo. of variables " << m_ndim + 1
              << "\nNo. Leaf cells: " << m_grid.g.n_leaf << "\nMax Mesh Level: " << m_grid.g.max_level << "\nNo. Bnd cells: " << m_grid.g.n_bnd
              << "\nRelaxation Time: " << m_relaxTime << "\nOmega: " << m_omega << "\nReynolds Number: " << m_re << "\nViscosity: " << m_nu
              << "\n+++++++++++++++++++++++++" << std::endl;
  }

  lbm_b200_config gpuConfig() const {
    lbm_b200_config cfg;
    lbm_b200_default_config(&cfg);
    cfg.ndim = m_ndim;
    cfg.ndist = m_ndist;
    cfg.collision = m_collision;
    cfg.precision = m_cfg.opt_str("precision", "fp64") == "fp32" ? LBM_B200_FP32 : LBM_B200_FP64;       // extension key
    cfg.arithmetic = m_cfg.opt_str("arithmetic", "strict") == "fast" ? LBM_B200_FAST : LBM_B200_STRICT; // extension key
    cfg.omega = m_omega;
    cfg.omega_minus = m_cfg.opt("omega_minus", m_omega);
    if(m_cfg.has("trt_magic")) { // Lambda = (1/w+ - 1/2)(1/w- - 1/2)
      const double lam = m_cfg.at("trt_magic").as_double();
      cfg.omega_minus  = 1.0 / (lam / (1.0 / m_omega - 0.5) + 0.5);
    }
    for(double& r : cfg.mrt_rates) r = m_omega;
    if(m_cfg.has("mrt_rates")) {
      const auto r = m_cfg.at("mrt_rates").as_doubles();
      for(size_t i = 0; i < r.size() && i < 27; ++i) cfg.mrt_rates[i] = r[i];
    }
    cfg.track_vars = static_cast<int32_t>(m_convInterval > 1 ? m_convInterval : 1);
    return cfg;
  }

  void setupGpu() {
    markBoundaryProperties();
    const SolverGrid& g = m_grid.g;
    lbm_b200_config cfg = gpuConfig();
    call(lbm_b200_create(&cfg, g.n, &m_gpu));
    call(lbm_b200_set_topology(m_gpu, g.nghbr.data(), g.nn_diag));
    call(lbm_b200_set_geometry(m_gpu, g.center.data(), g.bbmin, g.bbmax, g.cell_length));
    if(poisson()) call(lbm_b200_set_poisson(m_gpu, m_dt, m_poissonRate));
    // setupBndryCnds, bnd.h:71-142
    const Json& boundary = m_cfg.at("boundary");
    for(const auto& gk : boundary.obj) {
      const size_t nkeys = gk.second.size();
      for(const auto& sk : gk.second.obj) {
        const std::string sname = nkeys > 1 ? gk.first + "_" + sk.first : gk.first;
        const Surface*    srf   = g.find(sname);
        if(srf == nullptr) TERMM(-1, "Invalid bndryId \"" + sname + "\"");
        if(srf->cells.empty()) continue; // "WARNING: Skipping ... no valid cells!"
        const Json&       bc   = sk.second;
        if(!bc.has("type")) TERMM(-1, "The required configuration value is missing: type");
        const std::string type = bc.at("type").as_string();
        std::vector<double> normals;
        for(int64_t c : srf->cells)
          for(int d = 0; d < m_ndim; ++d) normals.push_back(srf->normal.at(c)[d]);
        const int64_t* cells = srf->cells.data();
        const int64_t  nc    = static_cast<int64_t>(srf->cells.size());
        const bool generate = bc.opt_bool("generateBndry", true);
        if(type == "periodic") {
          const Surface* other = g.find(bc.at("connection").as_string());
          if(other == nullptr) TERMM(-1, "Invalid bndryId \"" + bc.at("connection").as_string() + "\"");
          if(!generate) continue; // LBMBnd_dummy
          call(lbm_b200_add_periodic(m_gpu, cells, normals.data(), nc, other->cells.data(), static_cast<int64_t>(other->cells.size()),
                                     bc.has("pressure") ? bc.at("pressure").as_double() : NAN));
        } else if(type == "wall") {
          if(!generate) continue;
          const std::string model = bc.at("model").as_string();
          if(model == "bounceback") call(lbm_b200_add_wall_bb(m_gpu, cells, normals.data(), nc, bc.opt("tangentialVelocity", 0.0)));
          else if(model == "equilibrium" || model == "neem" || model == "nebb") {
            const int kind = model == "equilibrium" ? LBM_B200_WALL_EQUILIBRIUM : (model == "neem" ? LBM_B200_WALL_NEEM : LBM_B200_WALL_NEBB);
            std::vector<double> vel(3, 0.0);
            const bool has_v = bc.has("velocity");
            if(has_v) {
              const auto v = bc.at("velocity").as_doubles();
              for(int d = 0; d < m_ndim && d < static_cast<int>(v.size()); ++d) vel[d] = v[d];
            }
            call(lbm_b200_add_wall_wetnode(m_gpu, kind, cells, normals.data(), nc, has_v ? 1 : 0, vel.data()));
          }
          else TERMM(-1, "Invalid wall boundary model: " + model);
        } else if(type == "pressure") {
          if(!generate) continue;
          call(lbm_b200_add_pressure(m_gpu, cells, normals.data(), nc, bc.at("pressure").as_double()));
        } else if(type == "outlet" || type == "inlet") {
          TERMM(-1, "Broken"); // bnd.h:176-183
        } else if((type == "dirichlet" || type == "neumann") && bc.at("model").as_string() == "neem") {
          // LBMBnd_DirichletNEEM / LBMBnd_NeumannNEEM (bnd.h:116-137): Poisson equation only (bnd_dirichlet.h:329-331 "FIX ME")
          if(!generate) continue;
          if(!poisson()) TERMM(-1, "FIX ME");
          if(!bc.has("value")) TERMM(-1, "The required configuration value is missing: value");
          const Json& val = bc.at("value");
          std::vector<double> values(static_cast<size_t>(nc));
          if(val.is_string()) { // math expression at the cell centres, bnd_dirichlet.h:268-281
            const Expression e(val.as_string());
            try {
              for(int64_t k = 0; k < nc; ++k) values[k] = e.eval(&g.center[cells[k] * m_ndim], m_ndim);
            } catch(const std::runtime_error& err) {
              TERMM(-1, err.what());
            }
          } else if(val.is_array()) {
            if(!val.arr.empty() && val.arr[0].is_string()) TERMM(-1, "Impl"); // bnd_dirichlet.h:270
            std::fill(values.begin(), values.end(), val.as_doubles().at(0));
          } else {
            std::fill(values.begin(), values.end(), val.as_double());
          }
          if(bc.opt_bool("setAnalyticalValue", false)) TERMM(-1, "setAnalyticalValue is not supported by this host");
          call(lbm_b200_add_poisson_neem(m_gpu, type == "neumann" ? 1 : 0, cells, normals.data(), nc, values.data(), 0.0));
        } else if(type == "dirichlet") {
          if(!generate) continue;
          const std::string model = bc.at("model").as_string();
          // the reference's bounce-back Dirichlet condition for the Poisson equation ends the run in its first apply():
          // TERMM(-1, "this is incorrect!") (bnd_dirichlet.h:98-106, test/poisson/poisson1D_BBDirichlet.json); same outcome here
          if(poisson() && model == "bounceback") TERMM(-1, "this is incorrect!");
          if(poisson()) TERMM(-1, "dirichlet model " + model + " is not available for the Poisson equation on this host");
          if(model != "bounceback") TERMM(-1, "dirichlet model " + model + " is not available on the GPU path yet (SURVEY.md section 8f N1)");
          const auto v = bc.at("value").as_doubles();
          if(static_cast<int>(v.size()) < m_ndim) TERMM(-1, "dirichlet value needs one entry per dimension");
          call(lbm_b200_add_dirichlet_bb(m_gpu, cells, normals.data(), nc, v.data()));
        } else {
          TERMM(-1, "Invalid bndCndType: " + type);
        }
      }
    }
    if(!m_cfg.opt_str("forcing", "").empty()) { // solver.cpp:630-647
      std::cerr << "Using forcing!" << std::endl;
      const Surface *in = g.find("cube_-x"), *out = g.find("cube_+x");
      if(in == nullptr || out == nullptr) TERMM(-1, "Invalid bndryId \"cube_-x\"");
      call(lbm_b200_set_forcing(m_gpu, in->cells.data(), static_cast<int64_t>(in->cells.size()), out->cells.data(),
                                static_cast<int64_t>(out->cells.size()), m_cfg.at("poiseuillePressureGradient").as_double()));
    }
    call(lbm_b200_init(m_gpu));
  }

  // solver.cpp:233-263
  bool convergenceCondition() {
    if(!(m_timeStep > 0 && m_timeStep % m_convInterval == 0)) return false;
    const int NVAR = nvar();
    std::vector<double> conv(NVAR);
    int32_t bad = 0;
    call(lbm_b200_residual(m_gpu, conv.data(), &bad));
    static const char* names3[4] = {"U", "V", "W", "rho"};
    std::ostringstream line;
    line << m_timeStep << ": ";
    for(int v = 0; v < NVAR; ++v) line << "d" << (poisson() ? "P" : (v == m_ndim ? "rho" : names3[v])) << "=" << conv[v] << " ";
    say(line.str());
    double maxConv = conv[0];
    for(double c : conv) maxConv = std::max(maxConv, c);
    lastResidual = maxConv;
    const double crit = m_cfg.opt("convergence", 1E-12);
    if(m_timeStep > 1 && maxConv < crit) {
      std::ostringstream line;
      line << "Reached convergence to: " << maxConv;
      say(line.str());
      return true;
    }
    if(m_timeStep > 1 && (bad || std::isnan(maxConv) || std::isinf(maxConv))) {
      say("Solution diverged!");
      m_diverged = true;
      return true;
    }
    return false;
  }

  // moments of the current fold in the reference's cell order, once per output step
  void fetchVars() {
    if(m_varsFresh) return;
    call(lbm_b200_get_moments(m_gpu, vars.data()));
    m_varsFresh = true;
  }

  // solver.cpp:323-384: the moments of the current fold go to out/<solution_filename>_<step>.vtp in the reference's own binary
  // VTK flavour (vtk_writer.hpp, byte-identical to the reference's file); "output_format": "ascii" is an extension
  void output(bool forced, const std::string& postfix = "") {
    // solver.cpp:327-333: once the solution has diverged, the next call writes one forced file with the postfix "bdiv" and nothing else
    if(m_diverged && !m_wroteDiverged) {
      m_wroteDiverged = true;
      output(true, "bdiv");
      return;
    }
    m_varsFresh = false; // `vars` is fetched only by the paths that read it on the host (fetchVars)
    if(!((m_timeStep > 0 && m_timeStep % m_solutionInterval == 0) || forced)) return;
    if(!m_cfg.opt_bool("write_output", true)) return;
    const std::string stem = m_outputDir + m_solutionName + "_" + std::to_string(m_timeStep) + postfix;
    ::mkdir(m_outputDir.c_str(), 0755);
    const std::string format = m_cfg.opt_str("output_format", "binary");
    if(format == "ascii") {
      fetchVars();
      return writeVtpAscii(stem + ".vtp");
    }
    if(format != "binary") TERMM(-1, "Invalid output_format: " + format);
    const SolverGrid&     g    = m_grid.g;
    const int             NVAR = nvar();
    std::vector<uint8_t>  keep = cellFilter();
    int64_t               nout = 0;
    for(uint8_t k : keep) nout += k;
    say("  Writing " + stem + ".vtp with #" + std::to_string(nout) + " cells"); // IO.h:423
    static const char* names[4] = {"U", "V", "W", "rho"};                                // variables.h: VELSTR, "rho"
    std::vector<vtk::Column> cols;
    if(poisson()) fetchVars();
    if(poisson()) cols.push_back(vtk::Column{"V", vars.data(), 1}); // the electric potential, solver.cpp:368-378
    else for(int v = 0; v < NVAR; ++v) cols.push_back(vtk::Column{v == m_ndim ? "rho" : names[v], vars.data() + v, NVAR});
    // device side of the output path (lbm_b200_encode_output): the cell filter, the 15-decimal rounding and the base64 text of every
    // field are produced on the GPU; the host only assembles the file.  Solver kinds / values the device encoder does not take
    // (LBM_B200_EUNSUP) go through the host writer, which produces the same bytes.  LBM_B200_HOST_OUTPUT=1 forces the host writer.
    if(!poisson() && nout > 0 && std::getenv("LBM_B200_HOST_OUTPUT") == nullptr) {
      // the text lands in page-locked memory kept from one output step to the next (lbm_b200_host_alloc)
      const int64_t need = static_cast<int64_t>(NVAR) * lbm_b200_output_chars(nout);
      if(need > m_encCap) {
        lbm_b200_host_free(m_encBuf);
        m_encBuf = nullptr;
        m_encCap = 0;
        void* ptr = nullptr;
        if(lbm_b200_host_alloc(&ptr, need) == LBM_B200_OK) {
          m_encBuf = static_cast<char*>(ptr);
          m_encCap = need;
        }
      }
      std::vector<int64_t> off(static_cast<size_t>(NVAR) + 1, 0);
      if(m_encBuf != nullptr && lbm_b200_encode_output(m_gpu, keep.data(), m_encBuf, m_encCap, off.data()) == LBM_B200_OK) {
        for(int v = 0; v < NVAR; ++v) {
          cols[v].encoded     = m_encBuf + off[v];
          cols[v].encoded_len = static_cast<size_t>(off[v + 1] - off[v]);
        }
      }
    }
    if(!poisson() && cols[0].encoded == nullptr) fetchVars(); // host writer: the moments come to the host
    if(nout == 0) TERMM(-1, "ERROR: Invalid call to encodeLE() with length = 0"); // base64.h:219-223 (the reference exits there)
    if(!vtk::write_points(stem + ".vtp", m_ndim, g.n, g.center.data(), keep.data(), cols))
      TERMM(-1, "Invalid output directory set! (value: " + m_outputDir + ")");
  }

  // CellFilterManager (cell_filter.h:60-96): leafCells = childless cells, the level filters compare the cell's level
  std::vector<uint8_t> cellFilter() const {
    const SolverGrid&    g = m_grid.g;
    std::vector<uint8_t> keep(static_cast<size_t>(g.n), 1);
    auto leaf_only = [&]() {
      for(int64_t c = 0; c < g.n; ++c) keep[c] = keep[c] && (g.props[c] & (1u << 14)) != 0;
    };
    auto level_only = [&](long long lvl) {
      for(int64_t c = 0; c < g.n; ++c) keep[c] = keep[c] && g.level[c] == lvl;
    };
    if(!m_cfg.has("cellFilter")) { // default {"cellFilter": "leafCells"} (solver.cpp:86)
      leaf_only();
      return keep;
    }
    const Json& fc = m_cfg.at("cellFilter");
    if(!fc.has("cellFilter")) TERMM(-1, "Invalid output configuration missing \"cellFilter\" key"); // cell_filter.h:78
    const Json&              f = fc.at("cellFilter");
    std::vector<std::string> list;
    if(f.is_array()) TERMM(-1, "untested"); // cell_filter.h:74
    list.push_back(f.as_string());
    for(const std::string& name : list) {
      if(name == "highestLvl") { level_only(g.max_level); continue; }
      if(name == "lowestLvl" || name == "partitionLvl") { level_only(g.part_level); continue; }
      if(name == "leafCells") { leaf_only(); continue; }
      if(name == "targetLvl") {
        if(!fc.has("outputLvl")) TERMM(-1, "The required configuration value is missing: outputLvl");
        if(fc.at("outputLvl").as_int() < g.part_level) TERMM(-1, "Outputting a lvl below the partition lvl is not possible!");
        level_only(fc.at("outputLvl").as_int());
        continue;
      }
      TERMM(-1, "Unknown output filter " + name);
    }
    return keep;
  }

  void writeVtpAscii(const std::string& path) {
    const SolverGrid& g = m_grid.g;
    std::ofstream o(path);
    if(!o) TERMM(-1, "Invalid output directory set! (value: " + m_outputDir + ")");
    const int NVAR = nvar();
    std::cerr << "  Writing " << path << " with #" << g.n << " cells" << std::endl;
    o << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PolyData>\n<Piece NumberOfPoints=\""
      << g.n << "\" NumberOfVerts=\"0\" NumberOfLines=\"0\" NumberOfStrips=\"0\" NumberOfPolys=\"0\">\n<Points>\n"
      << "<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
    o << std::setprecision(17);
    for(int64_t c = 0; c < g.n; ++c) {
      for(int d = 0; d < 3; ++d) o << (d < m_ndim ? g.center[c * m_ndim + d] : 0.0) << " ";
      o << "\n";
    }
    o << "</DataArray>\n</Points>\n<PointData>\n";
    static const char* names[4] = {"U", "V", "W", "rho"};
    for(int v = 0; v < NVAR; ++v) {
      o << "<DataArray type=\"Float64\" Name=\"" << (poisson() ? "V" : (v == m_ndim ? "rho" : names[v])) << "\" format=\"ascii\">\n";
      for(int64_t c = 0; c < g.n; ++c) o << vars[c * NVAR + v] << "\n";
      o << "</DataArray>\n";
    }
    o << "</PointData>\n</Piece>\n</PolyData>\n</VTKFile>\n";
  }

  // ---- postprocessing (src/postprocess/postprocessing.h:45-122): functions of type "line" hooked to atStart / beforeTimestep /
  // afterTimestep / atEnd; each execution writes ./line.csv = x,y,u of the leaf cells whose centre is within half a cell of the line
  enum PpHook { PP_ATSTART, PP_BEFORETIMESTEP, PP_AFTERTIMESTEP, PP_ATEND, PP_NUM };
  std::vector<std::vector<int64_t>> m_ppLines[PP_NUM];

  void initPostprocess() {
    if(!m_cfg.has("postprocessing")) return;
    const SolverGrid& g = m_grid.g;
    for(const auto& kv : m_cfg.at("postprocessing").obj) {
      const Json& conf = kv.second;
      if(!conf.is_object()) continue; // getAllObjects, configuration.h:281-292
      if(!conf.has("type")) TERMM(-1, "The required configuration value is missing: type");
      if(!conf.has("execute")) TERMM(-1, "The required configuration value is missing: execute");
      if(conf.at("type").as_string() != "line") TERMM(-1, "Invalid functype"); // postprocessing_func.h:7-12
      const std::string at = conf.at("execute").as_string();
      int hook = -1;
      if(at == "atStart") hook = PP_ATSTART;
      else if(at == "beforeTimestep") hook = PP_BEFORETIMESTEP;
      else if(at == "afterTimestep") hook = PP_AFTERTIMESTEP;
      else if(at == "atEnd") hook = PP_ATEND;
      else TERMM(-1, "Invalid hook");
      if(hook == PP_BEFORETIMESTEP || hook == PP_AFTERTIMESTEP)
        TERMM(-1, "postprocessing at every time step would read m_vars back from the GPU each step; only atStart / atEnd are available");
      // PostprocessCartesianFunctionLine::init (postprocessing_cartesian.h:17-44), Line = Eigen::Hyperplane::Through(A, B)
      // (common/line.h:7-21): unit normal (-dy, dx) / |AB|, offset -A.n, distance |n.p + offset|
      if(m_ndim != 2) TERMM(-1, "postprocessing type \"line\" exists for 2D only");
      if(!conf.has("A")) TERMM(-1, "The required configuration value is missing: A");
      if(!conf.has("B")) TERMM(-1, "The required configuration value is missing: B");
      const auto   A = conf.at("A").as_doubles(), B = conf.at("B").as_doubles();
      const double tx = B[0] - A[0], ty = B[1] - A[1];
      double       nx = -ty, ny = tx;
      const double z  = nx * nx + ny * ny;
      if(z > 0) { const double r = std::sqrt(z); nx /= r; ny /= r; }
      const double offset = -(A[0] * nx + A[1] * ny);
      std::vector<int64_t> cells;
      for(int64_t c = 0; c < g.n; ++c) { // leaf cells within half of their own length (postprocessing_cartesian.h:22-30)
        if(!(g.props[c] & (1u << 14))) continue;
        const double distance = std::abs(nx * g.center[c * 2] + ny * g.center[c * 2 + 1] + offset);
        const double len = g.multi_level ? g.length_on_level[g.level[c]] : g.cell_length;
        if(0.5 * len >= distance) cells.push_back(c);
      }
      // the reference's comparator ("some coordinate is smaller") is not a strict weak ordering; the same std::sort on the same
      // input reproduces its order (ascending y for the axis-parallel lines of the reference's configurations)
      std::sort(cells.begin(), cells.end(), [&g](int64_t a, int64_t b) {
        for(int d = 0; d < 2; ++d)
          if(g.center[a * 2 + d] < g.center[b * 2 + d]) return true;
        return false;
      });
      m_ppLines[hook].push_back(std::move(cells));
    }
  }

  void executePostprocess(int hook) {
    if(m_ppLines[hook].empty()) return;
    static const char* hookName[PP_NUM] = {"atStart", "beforeTimestep", "afterTimestep", "atEnd"};
    std::cerr << "Executing postprocessing at hook:" << hookName[hook] << std::endl;
    const SolverGrid& g    = m_grid.g;
    const int         NVAR = nvar();
    if(hook == PP_ATSTART) m_varsFresh = false;
    fetchVars(); // atEnd: the state of the last, forced output()
    for(const auto& cells : m_ppLines[hook]) {
      std::cerr << "  Writing line.csv" << std::endl;
      writeLineCsv("line.csv", cells, g, vars.data(), NVAR); // relative to the working directory, like the reference
    }
  }

 public:
  // ASCII::writePointsCSV (IO.h:36-81): coordinates at 15 significant digits, values as toStringVector prints them (fixed, 15 decimals)
  static void writeLineCsv(const std::string& path, const std::vector<int64_t>& cells, const SolverGrid& g, const double* v, int nvar) {
    std::ofstream o;
    o << std::setprecision(15);
    o.open(path);
    o << "x,y,u\n";
    for(int64_t c : cells) {
      std::ostringstream t;
      t.precision(15);
      t << std::fixed << v[c * nvar];
      o << g.center[c * 2] << "," << g.center[c * 2 + 1] << "," << t.str() << "\n";
    }
  }
  // tests: the cell lists of the configured "line" functions, hook by hook
  void                                     setupPostprocess() { initPostprocess(); }
  const std::vector<std::vector<int64_t>>& postprocessLines(int hook) const { return m_ppLines[hook]; }

 private:

  // solver.cpp:388-482 with analytical_solutions.h:16-33,51-59
  void compareToAnalyticalResult() {
    const std::string name = m_cfg.at("analyticalSolution").as_string();
    // analytical::getAnalyticalSolution<NDIM> (src/lbm/analytical_solutions.h:101-131): the solution as a vector of NDIM components.
    // The Poisson solutions are evaluated with libm here (the reference uses gcem's series at run time): the pass / fail thresholds
    // are orders of magnitude away from that difference, the printed error may differ in its last digits.
    std::function<void(const double*, double*)> sol;
    if(m_ndim == 2 && name == "couette2D_1_5") {
      const double reynoldsNum = 0.75, relaxTime = 0.9, refL = 1.0;
      const double dynViscosity = (2.0 * relaxTime - 1.0) / 6.0;
      const double refV = reynoldsNum * dynViscosity / refL;
      sol = [refV](const double* x, double* u) { u[0] = refV / 5.0 * x[1]; u[1] = 0; };
    } else if(m_ndim == 2 && name == "poiseuille2D_1") {
      const double dp = m_cfg.at("poiseuillePressureGradient").as_double(), nu = m_nu;
      sol = [dp, nu](const double* x, double* u) { u[0] = dp / (2.0 * nu) * x[1] * (1.0 - 0.0 - x[1]); u[1] = 0; };
    } else if(m_ndim == 2 && name == "poissonCHAI08_2") { // :83-87
      const double mu = std::sqrt(4 + M_PI * M_PI);
      sol = [mu](const double* x, double* u) { u[0] = std::cos(M_PI * x[0]) * std::sinh(mu * (1 - x[1])) / std::sinh(mu); u[1] = 0; };
    } else if(m_ndim == 1 && name == "poissonCHAI08_1") { // :70-78, Debye-Hueckel k = 27.79
      sol = [](const double* x, double* u) {
        const double k = 27.79, ep = std::exp(k), em = std::exp(-k);
        u[0] = (ep - 1.0) / (ep - em) * std::exp(-k * x[0]) + (1.0 - em) / (ep - em) * std::exp(k * x[0]);
      };
    } else if(m_ndim == 1 && name == "poissonSimpleDiffReaction") { // :92-95
      sol = [](const double* x, double* u) { u[0] = std::cosh(1.0 * (1.0 - x[0])) / std::cosh(1.0); };
    } else {
      TERMM(-1, "Invalid analyticalSolution :" + name + " selected!");
    }
    const SolverGrid& g = m_grid.g;
    std::vector<char> excluded(static_cast<size_t>(g.n), 0);
    if(m_cfg.has("analyticalSolutionExcludeSurface"))
      for(const Json& s : m_cfg.at("analyticalSolutionExcludeSurface").arr) {
        const Surface* srf = g.find(s.as_string());
        if(srf == nullptr) TERMM(-1, "Invalid bndryId \"" + s.as_string() + "\"");
        for(int64_t c : srf->cells) excluded[c] = 1;
      }
    const int NVAR = nvar();
    double sumError = 0, sumErrorSq = 0, sumSolution = 0, sumSolutionSq = 0;
    maxError = 0;
    // shiftCenter (solver.cpp:485-501): poissonCHAI08_1 compares at points spaced equidistantly over (0, 1)
    const bool   shift  = name == "poissonCHAI08_1";
    const double extent = shift ? std::abs(g.center[0] - g.center[static_cast<size_t>(g.n - 1)]) : 0.0;
    for(int64_t c = 0; c < g.n; ++c) {
      if(excluded[c]) continue;
      double x[3] = {0, 0, 0}, u[3] = {0, 0, 0}, v[3] = {0, 0, 0};
      for(int d = 0; d < m_ndim; ++d) x[d] = g.center[c * m_ndim + d];
      if(shift) {
        const double pos = std::floor(x[0] / (extent / static_cast<double>(g.n - 1)));
        x[0] = pos * 1.0 / static_cast<double>(g.n - 1);
      }
      sol(x, u);
      if(poisson()) v[0] = vars[c];                                        // the other components of the reference's vector are never set
      else for(int d = 0; d < m_ndim; ++d) v[d] = vars[c * NVAR + d];
      double d2 = 0, s2 = 0;
      for(int d = 0; d < m_ndim; ++d) { d2 += (v[d] - u[d]) * (v[d] - u[d]); s2 += u[d] * u[d]; }
      const double delta = std::sqrt(d2);
      sumError += delta;
      sumErrorSq += delta * delta;
      sumSolution += std::sqrt(s2);
      sumSolutionSq += s2;
      maxError = std::max(delta, maxError);
    }
    l2Error = (gcem_sqrt(sumErrorSq) / gcem_sqrt(sumSolutionSq)) / std::pow(static_cast<double>(g.n), 1.0 / m_ndim);
    gre     = sumError / sumSolution;
    {
      std::ostringstream l1, l2, l3;
      l1 << "max. Error: " << maxError;
      l2 << "avg. L2: " << l2Error;
      l3 << "global relative error: " << gre;
      say("Comparing to analytical result " + name); // one log message per line, like solver.cpp:473-480
      say(l1.str());
      say(l2.str());
      say(l3.str());
    }
    bool failed = false;
    if(m_cfg.opt("errorMax", 1.0) < maxError) {
      failed = true;
      std::cerr << "Error bounds for the maximum error have failed!!! ( < " << m_cfg.opt("errorMax", 1.0) << ")" << std::endl;
    }
    if(m_cfg.opt("errorL2", 1.0) < l2Error || std::isnan(l2Error)) {
      failed = true;
      std::cerr << "Error bounds for the L2 error have failed!!! ( < " << m_cfg.opt("errorL2", 1.0) << ")" << std::endl;
    }
    if(m_cfg.opt("errorGRE", 1.0) < gre || std::isnan(gre)) {
      failed = true;
      std::cerr << "Error bounds for the global relative error have failed!!! ( < " << m_cfg.opt("errorGRE", 1.0) << ")" << std::endl;
    }
    if(failed) TERMM(-1, "Analytical testcase failed");
  }

  Json        m_cfg;
  std::string m_equation = "navierstokes";
  std::string m_configFile, m_model = "D2Q9", m_outputDir = "out/", m_solutionName = "solution";
  int         m_ndim = 2, m_ndist = 9, m_collision = LBM_B200_BGK;
  bool        m_benchmark = false, m_diverged = false, m_wroteDiverged = false;
  std::string m_p2pFile;
  long long   m_infoInterval = 10, m_convInterval = 10, m_solutionInterval = 100, m_maxTimeStep = 0, m_timeStep = 0;
  double      m_dt = 0, m_poissonRate = 27.79; // Poisson equation: m_dt (solver.cpp:136), poisson_D (solver.cpp:589-599)
  double      m_refLength = 1.0, m_ma = 0.01, m_re = 1, m_nu = 0, m_relaxTime = 0.9, m_omega = 1.0 / 0.9;
  LBMGrid     m_grid;
  lbm_b200_solver* m_gpu = nullptr;
  // partitioned runs
  RankInfo    m_rank;
  UniformGrid m_ugrid;
  lbm_b200_partition* m_part = nullptr;
  int64_t     m_nOwned = 0, m_nLocal = 0, m_lo = 0;
  time_t      m_startTime = 0;
  std::vector<double> m_ownCenter;
};

} // namespace lbmhost
