/* lbm_b200.h -- C ABI of the B200-native lattice-Boltzmann time step.
 *
 * This is the drop-in boundary for the hot path of SFCMM/LBM (reference: /root/reference, v0.0.2):
 * everything LBMSolver::run does between allocateMemory() and the end of the step loop
 * (src/lbm/solver.cpp:176-197), i.e. the 8-pass timeStep (src/lbm/solver.cpp:307-320), the boundary
 * kernels (src/lbm/bnd/), the forcing (src/lbm/solver.cpp:626-696) and the residual reduction
 * (src/lbm/solver.cpp:233-263,809-815).  The reference keeps its grid generator, JSON configuration and
 * output; it hands this library the tables it already owns and asks for steps.  INTEGRATION.md shows the
 * binding a maintainer adds to src/lbm/solver.cpp.
 *
 * Conventions
 *   - plain C types; the caller owns every host buffer it passes, the library owns all device memory;
 *   - every function returns 0 on success and a negative LBM_B200_E* code on failure; the message is
 *     available from lbm_b200_last_error() (thread-local).  The reference-side wrapper turns a non-zero
 *     return into TERMM(-1, msg) (src/common/term.h:6-35);
 *   - one host thread per handle, one CUDA device per handle (reference: single caller thread,
 *     SURVEY.md section 8b);
 *   - cell ids, direction indices and variable slots are the reference's: populations i = 0..Q-1 in the
 *     order of LBMethod<>::m_dirs (src/lbm/constants.h:296-422, rest population last), variables
 *     u,v[,w],rho (src/lbm/variables.h:12-79), host arrays in the reference's array-of-structures layout
 *     f[cell*Q + i], vars[cell*NVAR + v] (src/lbm/solver.h:156-172);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     LBM_B200_ECUDA.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_B200_ABI_VERSION 1

typedef struct lbm_b200_solver lbm_b200_solver;

enum {
  LBM_B200_OK      = 0,
  LBM_B200_EINVAL  = -1, /* bad argument / unsupported configuration */
  LBM_B200_ESTATE  = -2, /* call order violated (e.g. step before init) */
  LBM_B200_ECUDA   = -3, /* CUDA runtime error or no device */
  LBM_B200_ENOMEM  = -4,
  LBM_B200_EUNSUP  = -5  /* semantics of the reference that the device plan cannot reproduce exactly */
};

enum { LBM_B200_FP64 = 0, LBM_B200_FP32 = 1 };
/* collision operator: the reference parses only "bgk" (src/lbm/constants.h:70-76); TRT / MRT are extensions */
enum { LBM_B200_BGK = 0, LBM_B200_TRT = 1, LBM_B200_MRT = 2 };
/* LBM_B200_STRICT: the reference's operation order, divisions kept, no FMA contraction -> bit-identical to
 * the reference in fp64.  LBM_B200_FAST: algebraically equal, reciprocal multiplies + FMA (<= 1e-12 rel.). */
enum { LBM_B200_STRICT = 0, LBM_B200_FAST = 1 };

typedef struct {
  int32_t abi_version;  /* LBM_B200_ABI_VERSION */
  int32_t ndim;         /* 2 or 3 (1 for the Poisson lattice D1Q3) */
  int32_t ndist;        /* 9 (D2Q9), 19 (D3Q19), 27 (D3Q27) -- src/lbm/constants.h:145-162; 3 (D1Q3), 5 (D2Q5): Poisson equation only */
  int32_t precision;    /* LBM_B200_FP64 / LBM_B200_FP32 */
  int32_t collision;    /* LBM_B200_BGK / TRT / MRT */
  int32_t arithmetic;   /* LBM_B200_STRICT / LBM_B200_FAST */
  int32_t device;       /* CUDA device ordinal */
  int32_t track_vars;   /* 0: m_vars/m_varsold not kept (residual unavailable, get_moments still works);
                           1: kept every step; k>1: kept for the two steps before every multiple of k
                           (k = solver.conv_interval, src/lbm/solver.cpp:76-77,233-234) */
  double  omega;        /* m_omega, src/lbm/solver.cpp:108-123 */
  double  omega_minus;  /* TRT: relaxation rate of the antisymmetric part */
  double  mrt_rates[27]; /* MRT: per-direction-pair rates, see DESIGN.md */
} lbm_b200_config;

/* Fills *cfg with defaults (D2Q9, fp64, BGK, strict, device 0, track_vars 1, omega 1). */
void lbm_b200_default_config(lbm_b200_config* cfg);

/* Replaces allocateMemory() (src/lbm/solver.cpp:166-173). ncells = grid().totalSize(). */
int lbm_b200_create(const lbm_b200_config* cfg, int64_t ncells, lbm_b200_solver** out);
void lbm_b200_destroy(lbm_b200_solver* s);

/* The push table the reference streams through: nghbr[cell*stride + dir] = CartesianGrid::neighbor(cell,dir)
 * (src/cartesiangrid.h:111-124), -1 = INVALID_CELLID.  stride = cartesian::maxNoNghbrsDiag<NDIM>() (8 / 26);
 * only the first Q-1 columns are read.  The table need not be symmetric and is used as is. */
int lbm_b200_set_topology(lbm_b200_solver* s, const int64_t* nghbr, int32_t stride);

/* Cell centres (grid().center(cell,dir)), the grid bounding box and the cell length on the finest level.
 * Needed only by the periodic boundary condition and the forcing, which match cells by coordinates
 * (src/lbm/bnd/bnd_periodic.h:31-98, src/lbm/solver.cpp:651-693). */
int lbm_b200_set_geometry(lbm_b200_solver* s, const double* center, const double* bbmin, const double* bbmax,
                          double cell_length);

/* Boundary conditions, to be added in the reference's application order (LBMBndManager::m_bndrys,
 * src/lbm/bnd/bnd.h:71-142).  cells[k] / normals[k*ndim..] = Surface::getCellList() / normal_p(cell)
 * (src/common/surface.h:37,69); duplicates in the list are allowed and behave as in the reference. */
/* LBMBnd_wallBB<..., TANGENTIALVELO> (src/lbm/bnd/bnd_wall.h:9-95); tangential == 0 -> no slip. */
int lbm_b200_add_wall_bb(lbm_b200_solver* s, const int64_t* cells, const double* normals, int64_t n, double tangential);
/* LBMBnd_DirichletBB (src/lbm/bnd/bnd_dirichlet.h:24-121); value[ndim] = wall velocity. */
int lbm_b200_add_dirichlet_bb(lbm_b200_solver* s, const int64_t* cells, const double* normals, int64_t n,
                              const double* value);
/* Wet-node walls: LBMBnd_wallEq / LBMBnd_wallNEEM / LBMBnd_wallNEBB (src/lbm/bnd/bnd_wall.h:101-479; NEBB is D2Q9 only).
 * has_velocity = the configuration has a "velocity" key (velocity[ndim]); without it the no-slip variants are used.
 * A solver that contains one of these runs the reference's passes in the reference's order on the GPU. */
enum { LBM_B200_WALL_EQUILIBRIUM = 0, LBM_B200_WALL_NEEM = 1, LBM_B200_WALL_NEBB = 2 };
int lbm_b200_add_wall_wetnode(lbm_b200_solver* s, int32_t model, const int64_t* cells, const double* normals, int64_t n,
                              int32_t has_velocity, const double* velocity);
/* Poisson equation types (solver.equation = "poisson": LBEquationType::Poisson, src/lbm/solver.cpp:283-293,540-543,562-566,589-612;
 * lattices D1Q3 / D2Q5 / D2Q9).  One variable per cell, the potential: lbm_b200_get_vars / get_moments / residual then use
 * NVAR = 1.  dt = m_dt (solver.cpp:100,136), rate = poisson_D (equation_th for "simple_diff_reaction", 27.79 otherwise, :589-599).
 * Call before lbm_b200_init; the only boundary conditions of this equation are the two NEEM conditions below.  A solver in this mode
 * runs the reference's passes in the reference's order on the GPU (lbm_b200/csrc/poisson.cuh); fp64, BGK, single GPU. */
int lbm_b200_set_poisson(lbm_b200_solver* s, double dt, double rate);
/* LBMBnd_DirichletNEEM (neumann == 0, src/lbm/bnd/bnd_dirichlet.h:250-368) / LBMBnd_NeumannNEEM (neumann != 0,
 * src/lbm/bnd/bnd_neumann.h:15-66).  values[k] = m_value of entry k: the constant of the configuration, or the configuration's math
 * expression evaluated at the cell centre by the caller (the reference uses exprtk, bnd_dirichlet.h:268-281); gradient = m_gradValue. */
int lbm_b200_add_poisson_neem(lbm_b200_solver* s, int32_t neumann, const int64_t* cells, const double* normals, int64_t n,
                              const double* values, double gradient);
/* LBMBnd_Pressure, anti-bounce-back (src/lbm/bnd/bnd_pressure.h:12-113). */
int lbm_b200_add_pressure(lbm_b200_solver* s, const int64_t* cells, const double* normals, int64_t n, double pressure);
/* LBMBnd_Periodic (src/lbm/bnd/bnd_periodic.h:175-215); connected = cell list of the connected surface;
 * pressure = NAN for the plain variant. Needs lbm_b200_set_geometry. */
int lbm_b200_add_periodic(lbm_b200_solver* s, const int64_t* cells, const double* normals, int64_t n,
                          const int64_t* connected, int64_t nconnected, double pressure);
/* LBMSolver::forcing (src/lbm/solver.cpp:626-696): inlet = surface "cube_-x", outlet = "cube_+x",
 * gradient = solver.poiseuillePressureGradient.  Needs lbm_b200_set_geometry. */
int lbm_b200_set_forcing(lbm_b200_solver* s, const int64_t* inlet, int64_t ninlet, const int64_t* outlet,
                         int64_t noutlet, double gradient);

/* ---- multi-GPU (one process per GPU). The reference has no domain decomposition (the MPI calls carry
 * no simulation data); this is the B200-native replacement for the exchange its CHANGELOG only plans.
 * The SFC-ordered cell list is cut into contiguous ranges, one per rank. A rank passes to lbm_b200_create / set_topology its
 * owned cells followed by `nghost` ghost cells (copies of the remote cells its owned cells push to or pull from), with
 * neighbour ids local to that list. After every step the library sends exactly the post-collision populations
 * (cell, direction) that a peer's pull needs and receives the ones its own ghosts must hold: ncclSend/ncclRecv in one
 * group on the solver's stream over NVLink. Lists are per peer, concatenated in peer order; both sides must use the
 * same order (lbm_b200/partition.py derives them deterministically from the tables, without communication). */
int lbm_b200_set_ghosts(lbm_b200_solver* s, int64_t nghost);
int lbm_b200_set_halo(lbm_b200_solver* s, int32_t npeers, const int32_t* peers, const int64_t* send_count,
                      const int64_t* send_cell, const int32_t* send_dir, const int64_t* recv_count,
                      const int64_t* recv_cell, const int32_t* recv_dir);
/* Velocity halo of the pressure boundary condition.  LBMBnd_Pressure::apply extrapolates the boundary velocity from m_vars of the
 * two inward neighbours n1 = N(c, inside), n2 = N(n1, inside) (src/lbm/bnd/bnd_pressure.h:68-84).  Where a partition cut separates
 * a pressure cell from n1 or n2, the owner of the neighbour sends its velocity with every halo exchange.  Per peer of
 * lbm_b200_set_halo (same order, call it first): send_cell = owned cells whose velocity that peer needs, recv_cell = ghost
 * cells whose velocity arrives from it; the k-th item a rank sends is the k-th item the peer receives.  The ghost rows of the
 * table must contain the links c -> n1 -> n2 (lbm_b200/partition.py writes them). */
int lbm_b200_set_vars_halo(lbm_b200_solver* s, const int64_t* send_count, const int64_t* send_cell, const int64_t* recv_count,
                           const int64_t* recv_cell);
/* NCCL bootstrap: rank 0 creates the 128-byte id, the host program broadcasts it (torch.distributed / MPI), every rank
 * calls lbm_b200_comm_init before lbm_b200_init. */
int lbm_b200_comm_unique_id(char* out128);
int lbm_b200_comm_init(lbm_b200_solver* s, const char* id128, int32_t rank, int32_t nranks);
/* Rows of the synthetic box's table for an arbitrary list of global cell ids (what one rank of a partitioned run needs). */
int lbm_b200_box_rows(int32_t ndim, const int64_t* shape, const int32_t* periodic, const int64_t* cells, int64_t ncells,
                      int64_t* nghbr, int32_t stride, double* center);

/* ---- partition helper (host code, no CUDA): the local problem of one rank, built from table rows the caller supplies through a
 * callback -- a full table, the synthetic box, or an on-demand provider such as lbm_b200/host/uniform_grid.hpp -- so that no rank needs
 * the whole table.  Equal-count contiguous ranges of the SFC-ordered list (uniform weights: the reference's only WeightMethod,
 * src/loadbalancing_weights.h:19-29); ghosts = remote cells an owned cell pushes to or pulls from; halo lists in (global id,
 * direction) order, identical on both sides without communication; pressure = the GLOBAL cell lists of all pressure boundary
 * conditions in application order (velocity halo of their inward neighbours, see lbm_b200_set_vars_halo).
 * rows_fn(user, ids, n, rows, sources): rows[r*stride + j] = N(ids[r], j), sources[r*stride + j] = the cell whose push in direction j
 * lands in ids[r] (-1: none); either output pointer may be NULL; returns 0 on success. */
typedef struct lbm_b200_partition lbm_b200_partition;
typedef int (*lbm_b200_rows_fn)(void* user, const int64_t* ids, int64_t n, int64_t* rows, int64_t* sources);
int lbm_b200_partition_create(int64_t ncells_global, int32_t ndim, int32_t ndist, int32_t stride, int32_t rank, int32_t world,
                              lbm_b200_rows_fn rows_fn, void* user, int32_t npressure, const int64_t* const* pressure_cells,
                              const double* const* pressure_normals, const int64_t* pressure_count, lbm_b200_partition** out);
typedef struct {
  int64_t lo, hi;              /* this rank owns the global cells [lo, hi) */
  int64_t n_owned, n_ghost;    /* local list = owned cells, then ghosts */
  const int64_t* ghosts;       /* [n_ghost] global ids, ascending */
  const int64_t* nghbr;        /* [(n_owned + n_ghost) * stride] local ids: the table for lbm_b200_set_topology */
  int32_t stride, npeers;
  const int32_t* peers;
  const int64_t *send_count, *recv_count, *send_cell, *recv_cell;   /* as lbm_b200_set_halo takes them */
  const int32_t *send_dir, *recv_dir;
  const int64_t *vsend_count, *vrecv_count, *vsend_cell, *vrecv_cell; /* as lbm_b200_set_vars_halo takes them */
} lbm_b200_partition_view;
int lbm_b200_partition_get(const lbm_b200_partition* p, lbm_b200_partition_view* out);
/* Entries of a boundary-condition cell list that this rank owns, order kept: local_cells[k] = local id, index[k] = position in the
 * input list (to pick the matching normals / values).  Returns the count. */
int64_t lbm_b200_partition_restrict(const lbm_b200_partition* p, const int64_t* cells, int64_t n, int64_t* local_cells, int64_t* index);
/* lbm_b200_set_ghosts + lbm_b200_set_halo (+ lbm_b200_set_vars_halo when there is one) on a solver created with n_owned + n_ghost cells. */
int lbm_b200_partition_apply(const lbm_b200_partition* p, lbm_b200_solver* s);
void lbm_b200_partition_destroy(lbm_b200_partition* p);

/* Optional: run the kernels on this cudaStream_t (default: the legacy default stream). */
int lbm_b200_set_stream(lbm_b200_solver* s, void* cuda_stream);

/* Replaces initialCondition() (src/lbm/solver.cpp:267-304): builds the device plan, uploads it and sets
 * rho = 1, u = boundary preset, f = fold = feq.  After this call the topology and BCs are frozen. */
int lbm_b200_init(lbm_b200_solver* s);

/* Replaces `nsteps` iterations of timeStep() (src/lbm/solver.cpp:307-320). Asynchronous on the stream. */
int lbm_b200_step(lbm_b200_solver* s, int64_t nsteps);
int lbm_b200_synchronize(lbm_b200_solver* s);

/* Replaces sumAbsDiff for every variable (src/lbm/solver.cpp:809-815) and the NaN/Inf test
 * (src/lbm/solver.cpp:254-260): out[v] = sum over cells |vars - varsold|.  Requires track_vars. */
/* In a partitioned run (lbm_b200_comm_init) the sums cover the whole domain: one ncclAllReduce of NVAR + 1 doubles over the
 * ranks' owned cells; the call is then collective (every rank calls it at the same step). */
int lbm_b200_residual(lbm_b200_solver* s, double* out, int32_t* diverged);

/* State read-back in the reference's layout.  Any pointer may be NULL.
 *   f, fold      : m_f (post-collision) and m_fold (after streaming + boundary conditions) [ncells*Q]
 *   vars, varsold: m_vars / m_varsold as convergenceCondition() would see them now [ncells*NVAR] (track_vars)
 *   moments      : what output() recomputes from the current fold (src/lbm/solver.cpp:336) [ncells*NVAR] */
int lbm_b200_get_populations(lbm_b200_solver* s, double* f, double* fold);
int lbm_b200_get_vars(lbm_b200_solver* s, double* vars, double* varsold);
int lbm_b200_get_moments(lbm_b200_solver* s, double* moments);
/* Overwrites m_fold and, if f != NULL, m_f (restart / tests).  Both [ncells*Q], reference layout.  m_fold alone is the input of the
 * next time step: collisionStep overwrites all of m_f before anything reads it (solver.cpp:601-613), so f may be NULL, in which
 * case lbm_b200_get_populations returns the previous m_f until a step has run. */
int lbm_b200_set_populations(lbm_b200_solver* s, const double* f, const double* fold);

int64_t lbm_b200_steps_done(const lbm_b200_solver* s);

/* Plan statistics for DESIGN.md / bench.py. */
typedef struct {
  int64_t ncells;          /* cells owned */
  int64_t cells_fast;      /* cells updated by the template-indexed chunk path (no index traffic) */
  int64_t cells_generic;   /* cells updated through per-slot link codes */
  int64_t chunk_cells;     /* cells per SFC chunk */
  int64_t slots_bc;        /* (cell,dir) slots written by a boundary condition */
  int64_t slots_stale;     /* slots nothing ever writes (keep their initial value) */
  int64_t device_bytes;    /* device memory held */
  int64_t launches;        /* kernels launched since init */
  int64_t launches_main;   /* of which: fused stream+collide launches */
  double  bytes_per_cell_alg; /* 2*Q*sizeof(real) */
  int64_t h2d_bytes;       /* bytes copied host->device / device->host by set_/get_ calls since init */
  int64_t d2h_bytes;
  int64_t cells_ghost;     /* copies of cells owned by other ranks */
  int64_t halo_bytes;      /* bytes sent + received by the halo exchange since init */
} lbm_b200_stats;
int lbm_b200_get_stats(const lbm_b200_solver* s, lbm_b200_stats* out);

/* Inspection of the device plan (host-side layout planning of lbm_b200/csrc/plan.hpp; no arithmetic, no CUDA call).  Works on
 * a normal handle and on an inspection-only handle created with config.device = -1 (which refuses init / step).  The tests
 * use it to check on a machine without a GPU that the plan resolves the reference's preApply -> push -> apply order, the chunk
 * templates, wall descriptors, ghost blocks and halo lists correctly.  Pointers stay valid until lbm_b200_destroy. */
typedef struct {
  int64_t n, n_owned, npad, chunk, nsel, n_fast_chunks, n_fast_outer, gen_begin, n_gen, n_gen_outer, gen_stride, ghost_begin,
      n_ghost_blocks, n_values_static;
  const int32_t*  ref2dev;   /* [n] reference cell -> device slot */
  const uint16_t* tmpl;      /* [(Q-1)*chunk] sel << 10 | offset (device in-chunk order) */
  const int32_t*  chunk_nb;  /* [n_fast_chunks*(nsel+1)] neighbour-chunk bases (-1 wall), wall descriptor id */
  const int32_t*  codes;     /* [(Q-1)*gen_stride] link codes of the generic range */
  const int32_t*  copytab;   /* [n_copy*2] cell, dir */
  int64_t         n_copy;
  const double*   addtab;    /* [n_add*4] v0, v1, v2, count */
  int64_t         n_add;
  const double*   wall_desc; /* [n_wall*4] per wall descriptor nsel * (Q-1) entries [selector][direction] of v0, v1, v2, count (count -1:
                                anti-bounce-back slot) */
  int64_t         n_wall;
  const double*   abb_p;     /* [n_abb] pressure of every anti-bounce-back entry */
  const int32_t*  abb_cells; /* [n_abb*3] device cell, inward neighbours n1, n2 (< 0: -(slot+1) of the received velocity halo) */
  int64_t         n_abb;
  const double*   values;    /* [n_values] stored slot values (static part filled at init) */
  int64_t         n_values;
  const int64_t*  stale_ref; /* [n_stale] device cell * Q + dir of every slot nothing writes */
  int64_t         n_stale;
  const int64_t*  send_index; /* flat device indices dir * npad + cell, wire order */
  int64_t         n_send;
  const int64_t*  recv_index;
  int64_t         n_recv;
  const int32_t*  vsend_cells; /* [n_vsend] device cells whose velocity is sent, wire order */
  int64_t         n_vsend;
  int64_t         n_vrecv;     /* received velocity items */
  const int32_t*  chunk_abb_base; /* [n_fast_chunks] row of chunk_abb (chunks on a pressure face), -1 otherwise */
  const int32_t*  chunk_abb;      /* [n_chunk_abb_rows*chunk] anti-bounce-back entry of the cell at that device offset, -1 elsewhere */
  int64_t         n_chunk_abb_rows;
} lbm_b200_plan_view;
int lbm_b200_debug_plan(lbm_b200_solver* s, lbm_b200_plan_view* out);

/* Time `nsteps` steps with CUDA events on the solver's stream; *ms_total covers all kernels of those steps,
 * *ms_main only the fused stream+collide kernel (events around each launch). Synchronous. */
int lbm_b200_step_timed(lbm_b200_solver* s, int64_t nsteps, float* ms_total, float* ms_main);

/* Synthetic benchmark grid.  The reference's `--bench` is unimplemented for the LBM solver (initBenchmark:
 * TERMM("Not implemented!"), src/lbm/solver.cpp:39-46) and its grid generator's bench set-up aborts
 * (src/gridgenerator/gridGenerator.cpp:43-54); this builds what that path would hand to the solver for a uniform box of
 * shape[0] x shape[1] (x shape[2]) cells: cells in the order of the reference's space-filling curve
 * (include/common/math/hilbert.h:16-48), nghbr[cell*stride + dir] with axis neighbours (grid-level periodic links
 * where periodic[d] != 0, src/cartesiangrid.h:608-706) and diagonal neighbours composed from axis steps in x, y, z
 * order (src/cartesiangrid.h:451-493, extended to 3D in LBMethod<D3Q27>::m_dirs order).  stride >= 8 (2D) / 26 (3D).
 * center[cell*ndim + d] = (coord + 0.5) / max(shape) and coords[cell*ndim + d] (integer) are optional. */
int64_t lbm_b200_box_ncells(int32_t ndim, const int64_t* shape);
/* hilbert::index<NDIM>(x, level) of the reference (include/common/math/hilbert.h:16-48): x in unit-cube coordinates,
 * 1 <= ndim <= 4.  Returns -1 on bad arguments. */
int64_t lbm_b200_sfc_index(int32_t ndim, const double* x, int32_t level);
int lbm_b200_box_topology(int32_t ndim, const int64_t* shape, const int32_t* periodic, int64_t* nghbr, int32_t stride,
                          double* center, int64_t* coords);

const char* lbm_b200_last_error(void);
int         lbm_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
