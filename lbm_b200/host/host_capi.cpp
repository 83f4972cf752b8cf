// host_capi.cpp -- C entry points over the host-side mirror (grid pipeline + LBMSolver) for tests and embedding.
#include <cstring>
#include <string>

#include "lbm_solver.hpp"
#include "uniform_grid.hpp"
#include "expr.hpp"

using namespace lbmhost;

namespace {
void set_err(char* err, int n, const std::string& msg) {
  if(err != nullptr && n > 0) {
    std::strncpy(err, msg.c_str(), static_cast<size_t>(n - 1));
    err[n - 1] = 0;
  }
}
struct GridHandle {
  GridGenerator gen;
  LBMSolver     solver;
};
} // namespace

extern "C" {

// gridder run + transferGrid: the tables the LBM solver would stream through (no GPU needed)
void* lbmhost_grid_build(const char* config_path, char* err, int errlen) {
  auto* h = new GridHandle();
  try {
    h->gen.init(0, nullptr, config_path);
    h->gen.run();
    h->solver.init(0, nullptr, config_path);
    h->solver.transferGrid(h->gen.grid());
    h->solver.markBoundaryProperties(); // state after loadConfiguration(), which is where the reference's dump was taken
  } catch(const std::exception& e) {
    set_err(err, errlen, e.what());
    delete h;
    return nullptr;
  }
  return h;
}
void lbmhost_grid_free(void* h) { delete static_cast<GridHandle*>(h); }
int64_t lbmhost_grid_ncells(void* h) { return static_cast<GridHandle*>(h)->solver.solverGrid().n; }
int lbmhost_grid_ndim(void* h) { return static_cast<GridHandle*>(h)->solver.solverGrid().ndim; }
int lbmhost_grid_stride(void* h) { return static_cast<GridHandle*>(h)->solver.solverGrid().nn_diag; }
double lbmhost_grid_cell_length(void* h) { return static_cast<GridHandle*>(h)->solver.solverGrid().cell_length; }
void lbmhost_grid_copy(void* h, int64_t* nghbr, double* center, uint16_t* props) {
  const SolverGrid& g = static_cast<GridHandle*>(h)->solver.solverGrid();
  if(nghbr) std::memcpy(nghbr, g.nghbr.data(), g.nghbr.size() * sizeof(int64_t));
  if(center) std::memcpy(center, g.center.data(), g.center.size() * sizeof(double));
  if(props) std::memcpy(props, g.props.data(), g.props.size() * sizeof(uint16_t));
}
int lbmhost_grid_nsurfaces(void* h) { return static_cast<int>(static_cast<GridHandle*>(h)->solver.solverGrid().surfaces.size()); }
const char* lbmhost_grid_surface_name(void* h, int k) { return static_cast<GridHandle*>(h)->solver.solverGrid().surfaces[k].name.c_str(); }
int64_t lbmhost_grid_surface_size(void* h, int k) {
  return static_cast<int64_t>(static_cast<GridHandle*>(h)->solver.solverGrid().surfaces[k].cells.size());
}
void lbmhost_grid_surface_copy(void* h, int k, int64_t* cells, double* normals) {
  const SolverGrid& g = static_cast<GridHandle*>(h)->solver.solverGrid();
  const Surface&    s = g.surfaces[k];
  for(size_t i = 0; i < s.cells.size(); ++i) {
    cells[i] = s.cells[i];
    for(int d = 0; d < g.ndim; ++d) normals[i * g.ndim + d] = s.normal.at(s.cells[i])[d];
  }
}

// full pipeline like main(): gridder -> transferGrid -> LBMSolver::run. Returns 0 or the TERMM code.
// out[0..5] = max error, L2 error, global relative error, steps run, converged flag, last residual; vars (optional) = n*NVAR
int lbmhost_run(const char* config_path, double* out, double* vars, int64_t nvars, char* err, int errlen) {
  try {
    GridGenerator gen;
    gen.init(0, nullptr, config_path);
    gen.run();
    LBMSolver solver;
    solver.init(0, nullptr, config_path);
    solver.transferGrid(gen.grid());
    int rc = 0;
    try {
      rc = static_cast<int>(solver.run());
    } catch(const TermError& e) {
      set_err(err, errlen, e.what());
      rc = e.code;
    }
    if(out != nullptr) {
      out[0] = solver.maxError;
      out[1] = solver.l2Error;
      out[2] = solver.gre;
      out[3] = static_cast<double>(solver.stepsRun);
      out[4] = solver.converged ? 1.0 : 0.0;
      out[5] = solver.lastResidual;
    }
    if(vars != nullptr) {
      const int64_t n = std::min<int64_t>(nvars, static_cast<int64_t>(solver.vars.size()));
      std::memcpy(vars, solver.vars.data(), static_cast<size_t>(n) * sizeof(double));
    }
    return rc;
  } catch(const TermError& e) {
    set_err(err, errlen, e.what());
    return e.code;
  } catch(const std::exception& e) {
    set_err(err, errlen, e.what());
    return -1;
  }
}

// the reference's solution file (vtk_writer.hpp) from caller-supplied fields: vars = [n][nvar] as m_vars, names = nvar C strings.
// Returns 0, or -1 if the file cannot be written.  No GPU involved: this is the host-side formatting only.
int lbmhost_write_points(const char* path, int ndim, int64_t n, const double* center, const uint8_t* keep, int nvar, const double* vars,
                         const char* const* names) {
  std::vector<vtk::Column> cols;
  for(int v = 0; v < nvar; ++v) cols.push_back(vtk::Column{names[v], vars + v, nvar});
  return vtk::write_points(path, ndim, n, center, keep, cols) ? 0 : -1;
}
// postprocessing type "line" of a configuration on caller-supplied m_vars ([n][ndim+1]): writes what the reference's atEnd hook
// writes to ./line.csv into `out_path`.  Returns the number of cells on the (last configured) line, -1 on error.
int64_t lbmhost_postprocess_line(void* h, const double* vars, const char* out_path, char* err, int errlen) {
  auto* gh = static_cast<GridHandle*>(h);
  try {
    if(gh->solver.postprocessLines(3).empty()) gh->solver.setupPostprocess();
    const auto& lines = gh->solver.postprocessLines(3);
    if(lines.empty()) return 0;
    const SolverGrid& g = gh->solver.solverGrid();
    for(const auto& cells : lines) LBMSolver::writeLineCsv(out_path, cells, g, vars, gh->solver.noVars());
    return static_cast<int64_t>(lines.back().size());
  } catch(const std::exception& e) {
    set_err(err, errlen, e.what());
    return -1;
  }
}
// One rank's set-up of a partitioned run (LBMSolver::setupGpuPartitioned: on-demand grid rows, native partition, restricted boundary
// conditions, halo lists) on an inspection-only handle (device -1): returns the lbm_b200_solver* for lbm_b200_debug_plan, NULL on error.
// The caller destroys it with lbm_b200_destroy.  No GPU, no NCCL: this is the part of the multi-GPU host that can be checked on a CPU.
void* lbmhost_partitioned_handle(const char* config_path, int rank, int world, char* err, int errlen) {
  try {
    GridGenerator gen;
    gen.init(0, nullptr, config_path);
    LBMSolver solver;
    solver.init(0, nullptr, config_path);
    solver.setRank(rank, world);
    GridGen g; // transferGrid only needs the dimensionality in this mode
    GeneratedGrid gg;
    gg.gen.configure(Json::parse_file(config_path));
    solver.transferGrid(gg);
    return solver.buildPartitioned(-1);
  } catch(const std::exception& e) {
    set_err(err, errlen, e.what());
    return nullptr;
  }
}

// ---- single-level grid, rows on demand (uniform_grid.hpp): what one rank of a partitioned run asks the grid pipeline
void* lbmhost_ugrid_build(const char* config_path, char* err, int errlen) {
  auto* u = new UniformGrid();
  try {
    u->configure(Json::parse_file(config_path));
  } catch(const std::exception& e) {
    set_err(err, errlen, e.what());
    delete u;
    return nullptr;
  }
  return u;
}
void lbmhost_ugrid_free(void* h) { delete static_cast<UniformGrid*>(h); }
int64_t lbmhost_ugrid_ncells(void* h) { return static_cast<UniformGrid*>(h)->n; }
int lbmhost_ugrid_ndim(void* h) { return static_cast<UniformGrid*>(h)->ndim; }
int lbmhost_ugrid_stride(void* h) { return static_cast<UniformGrid*>(h)->nn_diag; }
int lbmhost_ugrid_level(void* h) { return static_cast<UniformGrid*>(h)->level; }
double lbmhost_ugrid_cell_length(void* h) { return static_cast<UniformGrid*>(h)->cell_length(); }
void lbmhost_ugrid_bbox(void* h, double* lo, double* hi) {
  const auto* u = static_cast<UniformGrid*>(h);
  for(int d = 0; d < u->ndim; ++d) { lo[d] = u->bbmin[d]; hi[d] = u->bbmax[d]; }
}
int lbmhost_ugrid_rows(void* h, const int64_t* ids, int64_t count, int64_t* nghbr, int stride, double* center, char* err, int errlen) {
  std::string e;
  if(static_cast<UniformGrid*>(h)->rows(ids, count, nghbr, stride, center, &e)) return 0;
  set_err(err, errlen, e);
  return -1;
}
int lbmhost_ugrid_sources(void* h, const int64_t* ids, int64_t count, int64_t* src, int stride, char* err, int errlen) {
  std::string e;
  if(static_cast<UniformGrid*>(h)->sources(ids, count, src, stride, &e)) return 0;
  set_err(err, errlen, e);
  return -1;
}
int lbmhost_ugrid_nsurfaces(void* h, char* err, int errlen) {
  auto* u = static_cast<UniformGrid*>(h);
  try {
    if(u->surfaces.empty()) u->build_surfaces();
  } catch(const std::exception& e) {
    set_err(err, errlen, e.what());
    return -1;
  }
  return static_cast<int>(u->surfaces.size());
}
const char* lbmhost_ugrid_surface_name(void* h, int k) { return static_cast<UniformGrid*>(h)->surfaces[k].name.c_str(); }
int64_t lbmhost_ugrid_surface_size(void* h, int k) { return static_cast<int64_t>(static_cast<UniformGrid*>(h)->surfaces[k].cells.size()); }
void lbmhost_ugrid_surface_copy(void* h, int k, int64_t* cells, double* normals) {
  const auto*    u = static_cast<UniformGrid*>(h);
  const Surface& s = u->surfaces[k];
  for(size_t i = 0; i < s.cells.size(); ++i) {
    cells[i] = s.cells[i];
    for(int d = 0; d < u->ndim; ++d) normals[i * u->ndim + d] = s.normal.at(s.cells[i])[d];
  }
}

// a boundary-value expression ("value": "cos(pi*x)", expr.hpp) at n points: points = [n][ndim]; returns 0, or -1 with a message
int lbmhost_eval_expression(const char* text, const double* points, int64_t n, int ndim, double* out, char* err, int errlen) {
  try {
    const Expression e(text);
    for(int64_t k = 0; k < n; ++k) out[k] = e.eval(points + k * ndim, ndim);
  } catch(const std::exception& ex) {
    set_err(err, errlen, ex.what());
    return -1;
  }
  return 0;
}

void lbmhost_round15(const double* in, double* out, int64_t n) {
  for(int64_t i = 0; i < n; ++i) out[i] = vtk::round15(in[i]);
}

} // extern "C"
