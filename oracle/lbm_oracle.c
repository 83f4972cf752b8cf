/* lbm_oracle.c -- CPU restatement of the SFCMM/LBM time step (the parity oracle).
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; it is the checker, never the product.  The product path is the CUDA library
 * built from lbm_b200/csrc and fails loudly when that library is missing.
 *
 * Every function follows the reference (/root/reference, SFCMM/LBM v0.0.2) pass by pass, with the same
 * operation order, the same array-of-structures storage and the same ascending-index sums, so that on the
 * reference's own configurations (D2Q9, BGK, fp64) it reproduces the reference's raw m_f / m_fold / m_vars /
 * m_varsold arrays bit for bit.  That claim is pinned by tests/test_oracle_golden.py against dumps of the
 * reference binary (oracle/_ref/lbm_ref, built by oracle/ref_build/Makefile) committed under tests/golden/.
 *
 * Pinned by the reference (file:line relative to /root/reference):
 *   storage                       src/lbm/solver.h:156-172,206-210   f[c*Q+i], vars[c*NVAR+v], slots u,v[,w],rho
 *   time step (8 passes)          src/lbm/solver.cpp:307-320
 *   currToOldVars                 src/lbm/solver.cpp:505-510
 *   updateMacroscopicValues       src/lbm/solver.cpp:513-553
 *   calcEquilibriumMoments        src/lbm/solver.cpp:556-571, src/lbm/equilibrium_func.h:52-54,68-84
 *   collisionStep (BGK)           src/lbm/solver.cpp:601-613
 *   forcing                       src/lbm/solver.cpp:626-696
 *   propagationStep (push)        src/lbm/solver.cpp:715-740
 *   bounce-back / Dirichlet BB    src/lbm/bnd/bnd_dirichlet.h:79-121, src/lbm/bnd/bnd_wall.h:31-88
 *   anti-bounce-back pressure     src/lbm/bnd/bnd_pressure.h:32-106
 *   periodic boundary condition   src/lbm/bnd/bnd_periodic.h:31-119,175-215
 *   residual                      src/lbm/solver.cpp:233-263,809-815
 *   initial condition             src/lbm/solver.cpp:267-304
 *   lattice tables                src/lbm/constants.h:296-422
 *   wet-node walls                src/lbm/bnd/bnd_wetnode.h:12-72 (limited distribution sets), src/lbm/bnd/bnd_dirichlet.h:134-248
 *                                 (equilibrium), src/lbm/bnd/bnd_wall.h:101-306 (NEEM), :313-479 (NEBB, D2Q9 only),
 *                                 src/lbm/moments.h:120-154 (density from a limited set)
 *
 *   Poisson equation src/lbm/solver.cpp:283-293 (initial condition), :540-543 (potential), :562-566 with
 *                                 src/lbm/equilibrium_func.h:149-163 (equilibrium), :589-612 (collision + source term), lattices
 *                                 D1Q3 / D2Q5 src/lbm/constants.h:248-292, Dirichlet NEEM src/lbm/bnd/bnd_dirichlet.h:250-368,
 *                                 Neumann NEEM src/lbm/bnd/bnd_neumann.h:15-66, potential from a cell's populations
 *                                 src/lbm/moments.h:70-91 -- pinned on the five Poisson cases of the reference's test/run.sh
 *
 * NOT pinned by the reference ("parity unpinned", new behaviour):
 *   - D3Q19 / D3Q27 runs: the reference compiles these templates but cannot reach them
 *     (src/lbm/solverExe.h:37-90); the dimension-generic source text above is followed literally.
 *     The pressure BC writes all NDIM velocity components (the reference writes only u and v,
 *     src/lbm/bnd/bnd_pressure.h:92-93).
 *   - TRT and MRT collision (literature definitions; equal rates reduce to the BGK path).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXQ 27
#include "mrt_tables.h"
/* MRT moment bases, padded to ORC_MAXQ columns so that one pointer type serves the three lattices */
static const int mrt_m9[9][ORC_MAXQ]   = LBM_MRT_D2Q9_M;
static const int mrt_m19[19][ORC_MAXQ] = LBM_MRT_D3Q19_M;
static const int mrt_m27[27][ORC_MAXQ] = LBM_MRT_D3Q27_M;
static const int mrt_n9[9] = LBM_MRT_D2Q9_NORM, mrt_n19[19] = LBM_MRT_D3Q19_NORM, mrt_n27[27] = LBM_MRT_D3Q27_NORM;
#define ORC_MAXD 3

enum { ORC_BGK = 0, ORC_TRT = 1, ORC_MRT = 2 };
enum { ORC_BC_BB = 1, ORC_BC_BB_TANGENTIAL = 2, ORC_BC_DIRICHLET_BB = 3, ORC_BC_PRESSURE = 4, ORC_BC_PERIODIC = 5,
       ORC_BC_WALL_EQ = 6, ORC_BC_WALL_NEEM = 7, ORC_BC_WALL_NEBB = 8, ORC_BC_DIRICHLET_NEEM = 9, ORC_BC_NEUMANN_NEEM = 10 };
enum { ORC_EQ_NAVIER_STOKES = 0, ORC_EQ_POISSON = 1 };

static const double kEps = 2.220446049250313e-16; /* GDoubleEps, include/common/sfcmm_types.h:50 */

/* src/lbm/constants.h:27 -- the reference writes the speed of sound squared as the double 1.0/3.0 */
static const double kCssq = 1.0 / 3.0;

typedef struct {
  int    ndim, ndist;
  double c[ORC_MAXQ][ORC_MAXD];
  int    opp[ORC_MAXQ];
  double w[ORC_MAXQ];
  double poisson_alpha, poisson_w[ORC_MAXQ]; /* CHAI08 eq. 2.3, constants.h:262-265,283-286,305-307 */
} OrcLattice;

/* src/lbm/constants.h:296-319 */
static const int kD2Q9c[9][2]   = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {1, 1}, {1, -1}, {-1, -1}, {-1, 1}, {0, 0}};
static const int kD2Q9opp[9]    = {1, 0, 3, 2, 6, 7, 4, 5, 8};
/* src/lbm/constants.h:322-422 (D3Q19 is the 18-direction prefix of D3Q27 plus the rest population) */
static const int kD3c[26][3]    = {{-1, 0, 0},  {1, 0, 0},   {0, -1, 0},  {0, 1, 0},  {0, 0, -1}, {0, 0, 1},  {-1, -1, 0},
                                   {-1, 1, 0},  {1, -1, 0},  {1, 1, 0},   {-1, 0, -1}, {-1, 0, 1}, {1, 0, -1}, {1, 0, 1},
                                   {0, -1, -1}, {0, -1, 1},  {0, 1, -1},  {0, 1, 1},  {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1},
                                   {-1, 1, 1},  {1, -1, -1}, {1, -1, 1},  {1, 1, -1}, {1, 1, 1}};
static const int kD3Q19opp[19]  = {1, 0, 3, 2, 5, 4, 9, 8, 7, 6, 13, 12, 11, 10, 17, 16, 15, 14, 18};
static const int kD3Q27opp[27]  = {1,  0,  3,  2,  5,  4,  9,  8,  7,  6,  13, 12, 11, 10,
                                   17, 16, 15, 14, 25, 24, 23, 22, 21, 20, 19, 18, 26};

static int orc_lattice(OrcLattice* L, int ndim, int ndist) {
  memset(L, 0, sizeof(*L));
  L->ndim  = ndim;
  L->ndist = ndist;
  if(ndim == 1 && ndist == 3) { /* constants.h:248-270 */
    L->c[0][0] = -1; L->c[1][0] = 1; L->c[2][0] = 0;
    L->opp[0] = 1; L->opp[1] = 0; L->opp[2] = 2;
    L->w[0] = 1.0 / 6.0; L->w[1] = 1.0 / 6.0; L->w[2] = 2.0 / 3.0;
    L->poisson_alpha = 1.0 / 3.0;
    L->poisson_w[0] = 0.5; L->poisson_w[1] = 0.5; L->poisson_w[2] = 0;
    return 0;
  }
  if(ndim == 2 && ndist == 5) { /* constants.h:272-292 */
    static const int c5[5][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {0, 0}};
    static const int o5[5]    = {1, 0, 3, 2, 4};
    for(int i = 0; i < 5; ++i) {
      L->c[i][0] = c5[i][0];
      L->c[i][1] = c5[i][1];
      L->opp[i]  = o5[i];
      L->w[i]    = i < 4 ? 1.0 / 6.0 : 1.0 / 3.0;
      L->poisson_w[i] = i < 4 ? 0.25 : 0.0;
    }
    L->poisson_alpha = 1.0 / 2.0;
    return 0;
  }
  if(ndim == 2 && ndist == 9) {
    for(int i = 0; i < 9; ++i) {
      L->c[i][0] = kD2Q9c[i][0];
      L->c[i][1] = kD2Q9c[i][1];
      L->opp[i]  = kD2Q9opp[i];
      L->w[i]    = i < 4 ? 1.0 / 9.0 : (i < 8 ? 1.0 / 36.0 : 4.0 / 9.0);
      L->poisson_w[i] = i < 8 ? 1.0 / 8.0 : 0.0;
    }
    L->poisson_alpha = 1.0 / 3.0;
    return 0;
  }
  if(ndim == 3 && (ndist == 19 || ndist == 27)) {
    for(int i = 0; i < ndist - 1; ++i) {
      for(int d = 0; d < 3; ++d) L->c[i][d] = kD3c[i][d];
    }
    for(int i = 0; i < ndist; ++i) L->opp[i] = ndist == 19 ? kD3Q19opp[i] : kD3Q27opp[i];
    if(ndist == 19) {
      for(int i = 0; i < 19; ++i) L->w[i] = i < 6 ? 1.0 / 18.0 : (i < 18 ? 1.0 / 36.0 : 1.0 / 3.0);
    } else {
      for(int i = 0; i < 27; ++i) L->w[i] = i < 6 ? 2.0 / 27.0 : (i < 18 ? 1.0 / 54.0 : (i < 26 ? 1.0 / 216.0 : 8.0 / 27.0));
    }
    return 0;
  }
  return -1;
}

typedef struct {
  int      kind;
  int64_t  n;
  int64_t* cells;
  double*  normals; /* n*ndim; Surface::normal_p(cell) i.e. the LAST normal stored for that cell (surface.h:52-55) */
  double   value[ORC_MAXD];
  double   tangential;
  double   pressure;
  double*  wallval;  /* ORC_BC_BB_TANGENTIAL: n*ndist, bnd_wall.h:31-72 */
  int64_t* link;     /* ORC_BC_PERIODIC: n*ndist, bnd_periodic.h:59-98 */
  int*     linkdist; /* n*ndist */
  int*     nset;     /* n */
  /* wet-node walls */
  int      has_velocity;      /* configuration key "velocity" present */
  int*     lim_n;             /* n: size of the limited set (0 marks a corner), bnd_wetnode.h:54-61 */
  int*     lim_dist;          /* n*ndist: the set, ascending (std::set) */
  double*  lim_const;         /* n*ndist */
  int64_t* cell2bnd;          /* n: index used for entry k = LAST entry with the same cell (unordered_map, bnd_dirichlet.h:176-181) */
  int64_t* ext;               /* NEEM: extrapolation cell per entry, bnd_wall.h:212-246 */
  /* Poisson: Dirichlet / Neumann NEEM, bnd_dirichlet.h:250-368, bnd_neumann.h:15-66 */
  double*  values;            /* n: m_value[bndId][0] (Neumann: rewritten every step) */
  int*     extdir;            /* n: direction from the boundary cell to its extrapolation cell */
  double   grad;              /* m_gradValue */
} OrcBc;

typedef struct {
  OrcLattice L;
  int        nvar, stride, model, omp_collide;
  int        push_conflict; /* two cells push into one slot (multi-level grids with grid-level periodic links) */
  int        equation;      /* ORC_EQ_* */
  double     poisson_dt, poisson_rate; /* m_dt (solver.cpp:136) and poisson_D (solver.cpp:589-599) */
  int64_t    n;
  int64_t*   nghbr; /* n*stride push table, -1 = no neighbour (cartesiangrid.h:111-124) */
  double*    center; /* n*ndim or NULL */
  double     bbmin[ORC_MAXD], bbmax[ORC_MAXD], cell_length;
  double     omega;
  double     omega_minus;     /* TRT: odd-moment rate */
  double     mrt_rates[ORC_MAXQ];
  double *   f, *fold, *feq, *vars, *varsold;
  unsigned char* periodic; /* CellProperties::periodic, set when a periodic boundary condition is added (bnd.h:217-221) */
  OrcBc*     bc;
  int        nbc;
  /* forcing, solver.cpp:626-696 */
  int      forcing;
  int64_t *inlet, *outlet;
  int64_t  ninlet, noutlet;
  double   p_in, p_out;
  int64_t  step;
} Orc;

/* ------------------------------------------------------------------------------------------------ helpers */

static int in_direction(const OrcLattice* L, const double* normal, int dist) {
  /* constants.h:83-86: normal . c_dist >= eps */
  double dot = 0;
  for(int d = 0; d < L->ndim; ++d) dot += normal[d] * L->c[dist][d];
  return dot >= kEps;
}

/* equilibrium_func.h:52-54 -- kept with its divisions by constants */
static inline double default_eq(double w, double rho, double cu, double vsq) {
  return w * rho * (1.0 + cu / kCssq + cu * cu / (2.0 * kCssq * kCssq) - vsq / (2.0 * kCssq));
}
/* equilibrium_func.h:109-111 */
static inline double symm_eq(double w, double rho, double cu, double vsq) {
  return w * rho * (1.0 + cu * cu / (2.0 * kCssq * kCssq) - vsq / (2.0 * kCssq));
}

static inline double eq_dist(const OrcLattice* L, int dist, double rho, const double* u, int symm) {
  double vsq = 0;
  for(int d = 0; d < L->ndim; ++d) vsq += u[d] * u[d];
  double cu = 0;
  for(int d = 0; d < L->ndim; ++d) cu += u[d] * L->c[dist][d];
  return symm ? symm_eq(L->w[dist], rho, cu, vsq) : default_eq(L->w[dist], rho, cu, vsq);
}

/* equilibrium_func.h:68-84 */
static void eq_all(const OrcLattice* L, double* feq, double rho, const double* u) {
  double vsq = 0;
  for(int d = 0; d < L->ndim; ++d) vsq += u[d] * u[d];
  for(int i = 0; i < L->ndist; ++i) {
    double cu = 0;
    for(int d = 0; d < L->ndim; ++d) cu += u[d] * L->c[i][d];
    feq[i] = default_eq(L->w[i], rho, cu, vsq);
  }
}

/* ------------------------------------------------------------------------------------------------ lifecycle */

Orc* orc_create(int ndim, int ndist, int64_t ncells, const int64_t* nghbr, int stride, double omega) {
  Orc* o = (Orc*)calloc(1, sizeof(Orc));
  if(orc_lattice(&o->L, ndim, ndist) != 0) {
    free(o);
    return NULL;
  }
  o->nvar   = ndim + 1;
  o->stride = stride;
  o->n      = ncells;
  o->omega  = omega;
  o->model  = ORC_BGK;
  o->nghbr  = (int64_t*)malloc(sizeof(int64_t) * (size_t)ncells * (size_t)stride);
  memcpy(o->nghbr, nghbr, sizeof(int64_t) * (size_t)ncells * (size_t)stride);
  size_t nq = (size_t)ncells * (size_t)ndist, nv = (size_t)ncells * (size_t)o->nvar;
  o->f       = (double*)calloc(nq, sizeof(double));
  o->fold    = (double*)calloc(nq, sizeof(double));
  o->feq     = (double*)calloc(nq, sizeof(double));
  o->vars    = (double*)calloc(nv, sizeof(double));
  o->varsold = (double*)calloc(nv, sizeof(double));
  o->periodic = (unsigned char*)calloc((size_t)ncells, 1);
  /* The reference's propagation is an OpenMP loop (solver.cpp:728): where two cells push into the same slot it races; run
   * serially there, so that the highest source wins as in the reference's single-thread order (the fixtures of such cases are
   * generated with one thread, tests/golden/make_golden.py). */
  {
    unsigned char* seen = (unsigned char*)calloc((size_t)ncells, 1);
    for(int i = 0; i < ndist - 1 && !o->push_conflict; ++i) {
      memset(seen, 0, (size_t)ncells);
      for(int64_t c = 0; c < ncells; ++c) {
        const int64_t nb = nghbr[c * stride + i];
        if(nb < 0) continue;
        if(seen[nb]) { o->push_conflict = 1; break; }
        seen[nb] = 1;
      }
    }
    free(seen);
  }
  return o;
}

void orc_set_geometry(Orc* o, const double* center, const double* bbmin, const double* bbmax, double cell_length) {
  size_t nb = sizeof(double) * (size_t)o->n * (size_t)o->L.ndim;
  o->center = (double*)malloc(nb);
  memcpy(o->center, center, nb);
  for(int d = 0; d < o->L.ndim; ++d) {
    o->bbmin[d] = bbmin[d];
    o->bbmax[d] = bbmax[d];
  }
  o->cell_length = cell_length;
}

/* New behaviour (not in the reference): two-relaxation-time and multiple-relaxation-time collision.
 * TRT: f_i' = f_i - omega (f+_i - feq+_i) - omega_minus (f-_i - feq-_i) with the symmetric / antisymmetric
 * split over opposite pairs.
 * MRT: moment space, f' = f - M^-1 S M (f - feq) with the orthogonal integer basis of oracle/mrt_tables.h (generated by
 * tools/gen_mrt_tables.py: d'Humieres-type, reference direction order), S = diag(rates), M^-1 = M^T diag(1/|row|^2).  Conserved rows
 * (density, momentum) are never relaxed; all rates equal = BGK up to rounding.  PARITY UNPINNED (literature form, no reference run). */
void orc_set_collision(Orc* o, int model, double omega_minus, const double* rates) {
  o->model       = model;
  o->omega_minus = omega_minus;
  if(rates != NULL) memcpy(o->mrt_rates, rates, sizeof(double) * (size_t)o->L.ndist);
}

void orc_set_omp_collide(Orc* o, int on) { o->omp_collide = on; }

static OrcBc* new_bc(Orc* o, int kind, const int64_t* cells, const double* normals, int64_t n) {
  o->bc     = (OrcBc*)realloc(o->bc, sizeof(OrcBc) * (size_t)(o->nbc + 1));
  OrcBc* b  = &o->bc[o->nbc++];
  memset(b, 0, sizeof(*b));
  b->kind    = kind;
  b->n       = n;
  b->cells   = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  b->normals = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * (size_t)o->L.ndim);
  memcpy(b->cells, cells, sizeof(int64_t) * (size_t)n);
  memcpy(b->normals, normals, sizeof(double) * (size_t)n * (size_t)o->L.ndim);
  b->pressure = NAN;
  return b;
}

/* wall bounce-back; tangential == 0 selects the no-slip instantiation (bnd.h:232-240) */
int orc_add_bc_wall_bb(Orc* o, const int64_t* cells, const double* normals, int64_t n, double tangential) {
  const OrcLattice* L = &o->L;
  if(fabs(tangential) > kEps) {
    if(L->ndim != 2) return -1; /* bnd_wall.h:52-54: TERMM("Not implemented") */
    OrcBc* b      = new_bc(o, ORC_BC_BB_TANGENTIAL, cells, normals, n);
    b->tangential = tangential;
    b->wallval    = (double*)calloc((size_t)n * (size_t)L->ndist, sizeof(double));
    for(int64_t k = 0; k < n; ++k) {
      const double* nrm = &b->normals[k * L->ndim];
      for(int id = 0; id < L->ndist; ++id) {
        if(in_direction(L, nrm, id)) {
          const int    inside = L->opp[id];
          const double t[2]   = {nrm[1], nrm[0]}; /* bnd_wall.h:48-51 */
          const double tdot   = t[0] * L->c[inside][0] + t[1] * L->c[inside][1];
          /* bnd_wall.h:57-66: directions anti-parallel to the normal are skipped and stay 0; for axis
           * normals their tangential projection is 0 as well, so the stored value is 0 either way. */
          const double ndot = nrm[0] * L->c[inside][0] + nrm[1] * L->c[inside][1];
          const double nn   = sqrt(L->c[inside][0] * L->c[inside][0] + L->c[inside][1] * L->c[inside][1]);
          const int parallel = fabs(acos(ndot / nn) - 3.14159265358979323846) < 10 * kEps;
          if(!parallel) b->wallval[k * L->ndist + inside] = tangential * tdot;
        }
      }
    }
  } else {
    new_bc(o, ORC_BC_BB, cells, normals, n);
  }
  return 0;
}

int orc_add_bc_dirichlet_bb(Orc* o, const int64_t* cells, const double* normals, int64_t n, const double* value) {
  OrcBc* b = new_bc(o, ORC_BC_DIRICHLET_BB, cells, normals, n);
  for(int d = 0; d < o->L.ndim; ++d) b->value[d] = value[d];
  return 0;
}

int orc_add_bc_pressure(Orc* o, const int64_t* cells, const double* normals, int64_t n, double pressure) {
  OrcBc* b    = new_bc(o, ORC_BC_PRESSURE, cells, normals, n);
  b->pressure = pressure;
  return 0;
}

/* bnd_periodic.h:31-98.  `pressure` NAN selects the plain copy variant. Needs orc_set_geometry. */
int orc_add_bc_periodic(Orc* o, const int64_t* cells, const double* normals, int64_t n, const int64_t* conn,
                        int64_t nconn, double pressure) {
  const OrcLattice* L = &o->L;
  if(o->center == NULL) return -1;
  OrcBc* b    = new_bc(o, ORC_BC_PERIODIC, cells, normals, n);
  b->pressure = pressure;
  for(int64_t k = 0; k < n; ++k) o->periodic[cells[k]] = 1; /* surfA.setProperty(periodic), bnd.h:219 */
  for(int64_t k = 0; k < nconn; ++k) o->periodic[conn[k]] = 1; /* surfB.setProperty(periodic), bnd.h:220 */
  b->link     = (int64_t*)malloc(sizeof(int64_t) * (size_t)n * (size_t)L->ndist);
  b->linkdist = (int*)malloc(sizeof(int) * (size_t)n * (size_t)L->ndist);
  b->nset     = (int*)calloc((size_t)n, sizeof(int));
  const double maxMatch = 10 * kEps;
  for(int64_t k = 0; k < n; ++k) {
    const int64_t c   = cells[k];
    const double* nrm = &b->normals[k * L->ndim];
    const double* ctr = &o->center[c * L->ndim];
    int           ns  = 0;
    for(int dist = 0; dist < L->ndist; ++dist) {
      if(in_direction(L, nrm, dist)) {
        double coord[ORC_MAXD];
        int    inside = 1;
        for(int d = 0; d < L->ndim; ++d) {
          coord[d] = fabs(nrm[d]) > 0 ? o->bbmin[d] : ctr[d] + L->c[dist][d] * o->cell_length;
          if(coord[d] < o->bbmin[d] || coord[d] > o->bbmax[d]) inside = 0;
        }
        if(inside) b->linkdist[k * L->ndist + ns++] = dist;
      }
    }
    b->nset[k] = ns;
    for(int id = 0; id < ns; ++id) {
      const int dist = b->linkdist[k * L->ndist + id];
      double    ca[ORC_MAXD];
      for(int d = 0; d < L->ndim; ++d) ca[d] = fabs(nrm[d]) > 0 ? ctr[d] : ctr[d] + L->c[dist][d] * o->cell_length;
      int64_t link = -1;
      for(int64_t j = 0; j < nconn && link < 0; ++j) {
        const double* cb = &o->center[conn[j] * L->ndim];
        for(int d = 0; d < L->ndim; ++d) {
          if(fabs(ca[d] - cb[d]) <= maxMatch) { /* first match in ANY coordinate, bnd_periodic.h:75-91 */
            link = conn[j];
            break;
          }
        }
      }
      b->link[k * L->ndist + id] = link;
    }
  }
  return 0;
}

/* Wet-node wall family: kind = ORC_BC_WALL_EQ / _NEEM / _NEBB; has_velocity = the "velocity" key is present.
 * LBMBnd_wallWetnode constructor, bnd_wetnode.h:24-63. */
static int orthogonal(const OrcLattice* L, const double* normal, int dist) {
  double dot = 0; /* constants.h:88-91 */
  for(int d = 0; d < L->ndim; ++d) dot += normal[d] * L->c[dist][d];
  return fabs(dot) <= kEps;
}

int orc_add_bc_wall_wetnode(Orc* o, int kind, const int64_t* cells, const double* normals, int64_t n, int has_velocity,
                            const double* velocity) {
  const OrcLattice* L = &o->L;
  const int Q = L->ndist, D = L->ndim;
  if(kind == ORC_BC_WALL_NEBB && !(D == 2 && Q == 9)) return -1; /* bnd_wall.h:325-328 */
  OrcBc* b        = new_bc(o, kind, cells, normals, n);
  b->has_velocity = has_velocity;
  for(int d = 0; d < D; ++d) b->value[d] = has_velocity ? velocity[d] : 0.0;
  b->lim_n     = (int*)calloc((size_t)n, sizeof(int));
  b->lim_dist  = (int*)calloc((size_t)n * (size_t)Q, sizeof(int));
  b->lim_const = (double*)calloc((size_t)n * (size_t)Q, sizeof(double));
  b->cell2bnd  = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  for(int64_t k = 0; k < n; ++k) {
    const int64_t c   = cells[k];
    const double* nrm = &b->normals[k * D];
    double*       cst = &b->lim_const[k * Q];
    int           m   = 0;
    for(int dir = 0; dir < Q - 1; ++dir) {
      const int     op     = L->opp[dir];
      const int     per    = o->periodic[c];
      const int64_t nb_opp = o->nghbr[c * o->stride + op];
      if(in_direction(L, nrm, dir) && (per || nb_opp != -1)) {
        b->lim_dist[k * Q + m++] = dir;
        cst[dir]                 = 2;
      } else if(orthogonal(L, nrm, dir) && (per || nb_opp != -1)) {
        b->lim_dist[k * Q + m++] = dir;
        cst[dir] = (o->nghbr[c * o->stride + dir] == -1 && !per) ? 2 : 1; /* corner, bnd_wetnode.h:43-50 */
      }
    }
    int sumC = 0; /* std::accumulate(..., 0): integer accumulation */
    for(int i = 0; i < Q; ++i) sumC = (int)(sumC + cst[i]);
    if(sumC != Q - 1) {
      m = 0; /* "we clear to mark a corner" */
    } else {
      b->lim_dist[k * Q + m++] = Q - 1;
      cst[Q - 1]               = 1;
    }
    b->lim_n[k] = m;
  }
  for(int64_t k = 0; k < n; ++k) { /* cell2Bnd: the last entry of a cell wins */
    int64_t idx = k;
    for(int64_t j = n - 1; j > k; --j)
      if(cells[j] == cells[k]) { idx = j; break; }
    b->cell2bnd[k] = idx;
  }
  if(kind == ORC_BC_WALL_NEEM) {
    b->ext = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for(int64_t k = 0; k < n; ++k) {
      const double* nrm = &b->normals[k * D];
      int           ex  = -1;
      for(int d = 0; d < D && ex < 0; ++d) {
        if(nrm[d] < 0) ex = 2 * d + 1;
        else if(nrm[d] > 0) ex = 2 * d;
      }
      b->ext[k] = ex < 0 ? -1 : o->nghbr[cells[k] * o->stride + ex];
      if(b->ext[k] == -1) return -2; /* TERMM("No valid extrapolation cellId"), bnd_wall.h:243-246 */
    }
  }
  return 0;
}

/* solver.cpp:640-647: inlet = surface cube_-x, outlet = cube_+x, p_out = 1.0, p_in = 1.0 + gradient */
/* Poisson equation (solver.cpp EQ == LBEquationType::Poisson): one variable (the potential) instead of velocity + density.
 * dt = m_dt (solver.cpp:100,136), rate = poisson_D (equation_th for "simple_diff_reaction", else the Debye-Hueckel constant 27.79,
 * solver.cpp:589-599).  Call before adding boundary conditions. */
void orc_set_poisson(Orc* o, double dt, double rate) {
  o->equation     = ORC_EQ_POISSON;
  o->poisson_dt   = dt;
  o->poisson_rate = rate;
  o->nvar         = 1;
  free(o->vars);
  free(o->varsold);
  o->vars    = (double*)calloc((size_t)o->n, sizeof(double));
  o->varsold = (double*)calloc((size_t)o->n, sizeof(double));
}

/* LBMBnd_DirichletNEEM / LBMBnd_NeumannNEEM constructor (bnd_dirichlet.h:265-332): per entry the value and the extrapolation cell
 * = the neighbour opposite to the first missing axis neighbour, or the diagonal one at a 2D corner.  Returns -2 where the reference
 * aborts ("No valid extrapolation cellId"). */
int orc_add_bc_poisson_neem(Orc* o, int neumann, const int64_t* cells, const double* normals, int64_t n, const double* values, double grad) {
  OrcBc* bc = new_bc(o, neumann ? ORC_BC_NEUMANN_NEEM : ORC_BC_DIRICHLET_NEEM, cells, normals, n);
  const int D = o->L.ndim;
  bc->values = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  bc->ext    = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  bc->extdir = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  bc->grad   = grad;
  static const int opp8[8] = {1, 0, 3, 2, 6, 7, 4, 5}; /* cartesian::oppositeDir incl. the 2D diagonals */
  for(int64_t k = 0; k < n; ++k) {
    bc->values[k] = values[k];
    int ed = -1;
    for(int dist = 0; dist < 2 * D; ++dist) {
      if(o->nghbr[cells[k] * o->stride + dist] != -1) continue;
      if(ed < 0) ed = dist;
      else {
        if(ed == 0 && dist == 2) ed = 6;
        if(ed == 0 && dist == 3) ed = 7;
        if(ed == 1 && dist == 3) ed = 4;
        if(ed == 1 && dist == 2) ed = 5;
      }
    }
    if(ed < 0) return -2;
    bc->extdir[k] = D == 3 ? (ed ^ 1) : opp8[ed];
    bc->ext[k]    = o->nghbr[cells[k] * o->stride + bc->extdir[k]];
    if(bc->ext[k] < 0) return -2;
  }
  return 0;
}

int orc_set_forcing(Orc* o, const int64_t* inlet, int64_t ninlet, const int64_t* outlet, int64_t noutlet, double gradient) {
  if(o->center == NULL) return -1;
  o->forcing = 1;
  o->inlet   = (int64_t*)malloc(sizeof(int64_t) * (size_t)ninlet);
  o->outlet  = (int64_t*)malloc(sizeof(int64_t) * (size_t)noutlet);
  memcpy(o->inlet, inlet, sizeof(int64_t) * (size_t)ninlet);
  memcpy(o->outlet, outlet, sizeof(int64_t) * (size_t)noutlet);
  o->ninlet  = ninlet;
  o->noutlet = noutlet;
  o->p_out   = 1.0;
  o->p_in    = o->p_out + gradient;
  return 0;
}

void orc_destroy(Orc* o) {
  if(o == NULL) return;
  for(int i = 0; i < o->nbc; ++i) {
    free(o->bc[i].cells);
    free(o->bc[i].normals);
    free(o->bc[i].wallval);
    free(o->bc[i].link);
    free(o->bc[i].linkdist);
    free(o->bc[i].nset);
    free(o->bc[i].lim_n);
    free(o->bc[i].lim_dist);
    free(o->bc[i].lim_const);
    free(o->bc[i].cell2bnd);
    free(o->bc[i].ext);
    free(o->bc[i].values);
    free(o->bc[i].extdir);
  }
  free(o->periodic);
  free(o->bc);
  free(o->nghbr);
  free(o->center);
  free(o->f);
  free(o->fold);
  free(o->feq);
  free(o->vars);
  free(o->varsold);
  free(o->inlet);
  free(o->outlet);
  free(o);
}

/* ------------------------------------------------------------------------------------------------ init */

/* solver.cpp:267-304 with bnd_dirichlet.h:44-50 and bnd_pressure.h:32-38 */
void orc_init(Orc* o) {
  const OrcLattice* L = &o->L;
  const int Q = L->ndist, NV = o->nvar, D = L->ndim;
  memset(o->vars, 0, sizeof(double) * (size_t)o->n * (size_t)NV);
  memset(o->varsold, 0, sizeof(double) * (size_t)o->n * (size_t)NV);
  if(o->equation == ORC_EQ_POISSON) { /* solver.cpp:283-293; initCnd of the Dirichlet NEEM condition bnd_dirichlet.h:335-341 */
    for(int b = 0; b < o->nbc; ++b) {
      const OrcBc* bc = &o->bc[b];
      if(bc->kind == ORC_BC_DIRICHLET_NEEM)
        for(int64_t k = 0; k < bc->n; ++k) o->vars[bc->cells[k]] = bc->values[k];
    }
    for(int64_t c = 0; c < o->n; ++c) {
      const double phi = o->vars[c];
      for(int i = 0; i < Q - 1; ++i) {
        const double v = L->w[i] * phi;
        o->feq[c * Q + i] = o->f[c * Q + i] = o->fold[c * Q + i] = v;
      }
      const double v0 = (L->w[Q - 1] - 1.0) * phi;
      o->feq[c * Q + Q - 1] = o->f[c * Q + Q - 1] = o->fold[c * Q + Q - 1] = v0;
    }
    o->step = 0;
    return;
  }
  for(int b = 0; b < o->nbc; ++b) {
    const OrcBc* bc = &o->bc[b];
    if(bc->kind == ORC_BC_DIRICHLET_BB) {
      for(int64_t k = 0; k < bc->n; ++k)
        for(int d = 0; d < D; ++d) o->vars[bc->cells[k] * NV + d] = bc->value[d];
    } else if(bc->kind == ORC_BC_PRESSURE) {
      for(int64_t k = 0; k < bc->n; ++k) o->vars[bc->cells[k] * NV + D] = bc->pressure;
    }
  }
  for(int64_t c = 0; c < o->n; ++c) {
    o->vars[c * NV + D] = 1.0;
    eq_all(L, &o->feq[c * Q], o->vars[c * NV + D], &o->vars[c * NV]);
    for(int i = 0; i < Q; ++i) {
      o->f[c * Q + i]    = o->feq[c * Q + i];
      o->fold[c * Q + i] = o->feq[c * Q + i];
    }
  }
  o->step = 0;
}

/* ------------------------------------------------------------------------------------------------ passes */

/* solver.cpp:513-553 */
/* moments.h:83-91: potential = 1/(1 - w_rest) * sum of the moving populations, ascending from 0 */
static inline double potential_of(const OrcLattice* L, const double* fo) {
  double acc = 0;
  for(int i = 0; i < L->ndist - 1; ++i) acc += fo[i];
  return 1.0 / (1.0 - L->w[L->ndist - 1]) * acc;
}

static void pass_moments(Orc* o) {
  const OrcLattice* L = &o->L;
  const int Q = L->ndist, NV = o->nvar, D = L->ndim;
  if(o->equation == ORC_EQ_POISSON) { /* solver.cpp:540-543 */
#pragma omp parallel for schedule(static)
    for(int64_t c = 0; c < o->n; ++c) o->vars[c] = potential_of(L, &o->fold[c * Q]);
    return;
  }
#pragma omp parallel for schedule(static)
  for(int64_t c = 0; c < o->n; ++c) {
    const double* fo  = &o->fold[c * Q];
    double        rho = 0.0; /* std::accumulate(..., 0.0), ascending */
    for(int i = 0; i < Q; ++i) rho += fo[i];
    o->vars[c * NV + D] = rho;
    for(int d = 0; d < D; ++d) {
      double v = 0;
      for(int i = 0; i < Q - 1; ++i) v += L->c[i][d] * fo[i];
      o->vars[c * NV + d] = v / rho;
    }
  }
}

/* solver.cpp:556-571 */
static void pass_equilibrium(Orc* o) {
  const int Q = o->L.ndist, NV = o->nvar, D = o->L.ndim;
  if(o->equation == ORC_EQ_POISSON) { /* equilibrium_func.h:149-163 */
#pragma omp parallel for schedule(static)
    for(int64_t c = 0; c < o->n; ++c) {
      for(int i = 0; i < Q - 1; ++i) o->feq[c * Q + i] = o->L.w[i] * o->vars[c];
      o->feq[c * Q + Q - 1] = (o->L.w[Q - 1] - 1.0) * o->vars[c];
    }
    return;
  }
#pragma omp parallel for schedule(static)
  for(int64_t c = 0; c < o->n; ++c) eq_all(&o->L, &o->feq[c * Q], o->vars[c * NV + D], &o->vars[c * NV]);
}

/* solver.cpp:601-613 (BGK); TRT / MRT are new behaviour */
static void collide_cell(const Orc* o, int64_t c) {
  const OrcLattice* L = &o->L;
  const int         Q = L->ndist;
  double*           f = &o->f[c * Q];
  const double*     fo = &o->fold[c * Q];
  const double*     fe = &o->feq[c * Q];
  if(o->equation == ORC_EQ_POISSON) { /* solver.cpp:601-612 */
    const double relax_time  = 1.0 / o->omega; /* omega = 1.0 / m_relaxTime, solver.cpp:132-134 */
    const double diffusivity = L->poisson_alpha * 1.0 * (0.5 - relax_time) * o->poisson_dt; /* gcem::pow(m_latticeVelocity = 1, 2) */
    const double rhs         = o->poisson_rate * o->poisson_rate * o->vars[c];
    for(int i = 0; i < Q; ++i) {
      f[i] = (1 - o->omega) * fo[i] + o->omega * fe[i];
      if(i != Q - 1) f[i] += o->poisson_dt * diffusivity * L->poisson_w[i] * rhs;
    }
    return;
  }
  if(o->model == ORC_BGK) {
    for(int i = 0; i < Q; ++i) f[i] = (1 - o->omega) * fo[i] + o->omega * fe[i];
  } else if(o->model == ORC_MRT) {
    const int (*M)[ORC_MAXQ] = Q == 9 ? mrt_m9 : (Q == 19 ? mrt_m19 : mrt_m27);
    const int* norm          = Q == 9 ? mrt_n9 : (Q == 19 ? mrt_n19 : mrt_n27);
    /* base rate s0 = the rate most non-conserved moments share (ties: first in basis order); evaluated as
     * f' = f - s0 (f - feq) - sum_k (s_k - s0)/|M_k|^2 M_k^T M_k (f - feq), rows with s_k == s0 skipped */
    double s0 = o->mrt_rates[L->ndim + 1];
    int    best_n = 0;
    for(int k = L->ndim + 1; k < Q; ++k) {
      int n = 0;
      for(int j = L->ndim + 1; j < Q; ++j) n += o->mrt_rates[j] == o->mrt_rates[k] ? 1 : 0;
      if(n > best_n) { best_n = n; s0 = o->mrt_rates[k]; }
    }
    double     fneq[ORC_MAXQ];
    for(int i = 0; i < Q; ++i) {
      fneq[i] = fo[i] - fe[i];
      f[i]    = fo[i] - s0 * fneq[i];
    }
    for(int k = L->ndim + 1; k < Q; ++k) { /* rows 0 .. ndim are density and momentum */
      const double sk = (o->mrt_rates[k] - s0) / (double)norm[k];
      if(sk == 0) continue;
      double m     = 0;
      int    first = 1;
      for(int i = 0; i < Q; ++i) {
        if(M[k][i] == 0) continue;
        const double t = (double)M[k][i] * fneq[i];
        m              = first ? t : m + t;
        first          = 0;
      }
      const double d = sk * m;
      for(int i = 0; i < Q; ++i)
        if(M[k][i] != 0) f[i] = f[i] - (double)M[k][i] * d;
    }
  } else {
    for(int i = 0; i < Q; ++i) {
      const int    j  = L->opp[i];
      const double wp = o->omega, wm = o->omega_minus;
      const double fp  = 0.5 * (fo[i] + fo[j]);
      const double fm  = 0.5 * (fo[i] - fo[j]);
      const double fep = 0.5 * (fe[i] + fe[j]);
      const double fem = 0.5 * (fe[i] - fe[j]);
      f[i]             = fo[i] - wp * (fp - fep) - wm * (fm - fem);
    }
  }
}

static void pass_collision(Orc* o) {
  if(o->omp_collide) {
#pragma omp parallel for schedule(static)
    for(int64_t c = 0; c < o->n; ++c) collide_cell(o, c);
  } else {
    for(int64_t c = 0; c < o->n; ++c) collide_cell(o, c); /* the reference has no OpenMP pragma here */
  }
}

/* solver.cpp:626-696 */
static void pass_forcing(Orc* o) {
  if(!o->forcing) return;
  const OrcLattice* L = &o->L;
  const int Q = L->ndist, NV = o->nvar, D = L->ndim;
  for(int64_t a = 0; a < o->ninlet; ++a) {
    const int64_t val = o->nghbr[o->inlet[a] * o->stride + 1];
    const double* ci  = &o->center[val * D];
    for(int64_t b = 0; b < o->noutlet; ++b) {
      const int64_t out = o->outlet[b];
      if(fabs(ci[1] - o->center[out * D + 1]) < kEps) {
        double vsq = 0;
        for(int d = 0; d < D; ++d) vsq += o->vars[val * NV + d] * o->vars[val * NV + d];
        for(int i = 0; i < Q; ++i) {
          const double cu = o->vars[val * NV + 0] * L->c[i][0];
          o->f[out * Q + i] = default_eq(L->w[i], o->p_out, cu, vsq) + o->f[val * Q + i] - o->feq[val * Q + i];
        }
      }
    }
  }
  for(int64_t b = 0; b < o->noutlet; ++b) {
    const int64_t val = o->nghbr[o->outlet[b] * o->stride + 0];
    const double* co  = &o->center[val * D];
    for(int64_t a = 0; a < o->ninlet; ++a) {
      const int64_t in = o->inlet[a];
      if(fabs(o->center[in * D + 1] - co[1]) < kEps) {
        double vsq = 0;
        for(int d = 0; d < D; ++d) vsq += o->vars[val * NV + d] * o->vars[val * NV + d];
        for(int i = 0; i < Q; ++i) {
          const double cu = o->vars[val * NV + 0] * L->c[i][0];
          o->f[in * Q + i] = default_eq(L->w[i], o->p_in, cu, vsq) + o->f[val * Q + i] - o->feq[val * Q + i];
        }
      }
    }
  }
}

/* solver.cpp:699-712 -> bnd.h:48-53 */
static void pass_pre_apply(Orc* o) {
  const OrcLattice* L = &o->L;
  const int Q = L->ndist, NV = o->nvar, D = L->ndim;
  for(int b = 0; b < o->nbc; ++b) {
    const OrcBc* bc = &o->bc[b];
    if(bc->kind == ORC_BC_PRESSURE) {
      for(int64_t k = 0; k < bc->n; ++k) o->vars[bc->cells[k] * NV + D] = bc->pressure;
    } else if(bc->kind == ORC_BC_PERIODIC) {
      for(int64_t k = 0; k < bc->n; ++k) {
        const int64_t c = bc->cells[k];
        if(!isnan(bc->pressure)) {
          const int64_t l0 = bc->link[k * Q + 0];
          for(int i = 0; i < Q; ++i)
            o->fold[l0 * Q + i] = eq_dist(L, i, bc->pressure, &o->vars[c * NV], 0) + o->f[c * Q + i] - o->feq[c * Q + i];
          o->vars[l0 * NV + D] = bc->pressure;
        } else {
          for(int id = 0; id < bc->nset[k]; ++id) {
            const int dist                          = bc->linkdist[k * Q + id];
            o->fold[bc->link[k * Q + id] * Q + dist] = o->f[c * Q + dist];
          }
          o->vars[bc->link[k * Q + 0] * NV + D] = 1.0;
        }
      }
    }
  }
}

/* solver.cpp:715-740 */
static void pass_propagation(Orc* o) {
  const int Q = o->L.ndist;
#pragma omp parallel for schedule(static) if(!o->push_conflict)
  for(int64_t c = 0; c < o->n; ++c) {
    for(int i = 0; i < Q - 1; ++i) {
      const int64_t nb = o->nghbr[c * o->stride + i];
      if(nb != -1) o->fold[nb * Q + i] = o->f[c * Q + i];
    }
    o->fold[c * Q + Q - 1] = o->f[c * Q + Q - 1];
  }
}

/* bnd_dirichlet.h:79-121 */
static void bb_cell(Orc* o, int64_t c, const double* nrm, int mode, const double* values) {
  const OrcLattice* L = &o->L;
  const int         Q = L->ndist;
  for(int i = 0; i < Q - 1; ++i) {
    if(o->nghbr[c * o->stride + i] == -1 && in_direction(L, nrm, i)) {
      const int op         = L->opp[i];
      o->fold[c * Q + op] = o->f[c * Q + i];
      const double density = 1.0;
      if(mode == 1) { /* SCALAR, bnd_dirichlet.h:111-112 */
        o->fold[c * Q + op] += density * 2.0 / kCssq * L->w[op] * values[op];
      } else if(mode == 2) { /* vector, bnd_dirichlet.h:113-117 */
        for(int d = 0; d < L->ndim; ++d) o->fold[c * Q + op] += density * 2.0 / kCssq * L->w[op] * L->c[op][d] * values[d];
      }
    }
  }
}

/* moments.h:120-154 */
static void density_limited(Orc* o, int64_t c, const OrcBc* bc, int64_t idx, int noslip) {
  const int Q = o->L.ndist, NV = o->nvar, D = o->L.ndim;
  if(bc->lim_n[idx] == 0) return; /* corner: the density of the moments pass stays */
  double rho = 0;
  for(int m = 0; m < bc->lim_n[idx]; ++m) {
    const int dist = bc->lim_dist[idx * Q + m];
    rho += bc->lim_const[idx * Q + dist] * o->fold[c * Q + dist];
  }
  if(!noslip) {
    const double* nrm = &bc->normals[idx * D];
    for(int d = 0; d < D; ++d) {
      if(nrm[d] > kEps) rho *= 1.0 / (1.0 + o->vars[c * NV + d]);
      else if(nrm[d] < 0) rho *= 1.0 / (1.0 - o->vars[c * NV + d]);
    }
  }
  o->vars[c * NV + D] = rho;
}

/* LBMBnd_DirichletEQ::apply<VALZERO>, bnd_dirichlet.h:211-241 */
static void wall_eq_cell(Orc* o, const OrcBc* bc, int64_t k) {
  const OrcLattice* L = &o->L;
  const int Q = L->ndist, NV = o->nvar, D = L->ndim;
  const int64_t c   = bc->cells[k];
  const int64_t idx = bc->cell2bnd[k];
  const int     valzero = !bc->has_velocity;
  for(int d = 0; d < D; ++d) o->vars[c * NV + d] = bc->value[d];
  density_limited(o, c, bc, idx, valzero);
  if(valzero) {
    for(int i = 0; i < Q; ++i) o->fold[c * Q + i] = default_eq(L->w[i], o->vars[c * NV + D], 0, 0);
  } else {
    eq_all(L, &o->fold[c * Q], o->vars[c * NV + D], &o->vars[c * NV]);
  }
}

/* LBMBnd_wallNEBB, bnd_wall.h:366-466 (D2Q9 slots: 0 -x, 1 +x, 2 -y, 3 +y, 4 ++, 5 +-, 6 --, 7 -+) */
static void wall_nebb(Orc* o, const OrcBc* bc) {
  const int Q = 9, NV = 3, D = 2;
  double*   fo = o->fold;
  if(!bc->has_velocity) {
    for(int64_t k = 0; k < bc->n; ++k)
      for(int d = 0; d < D; ++d) o->vars[bc->cells[k] * NV + d] = 0;
    for(int64_t k = 0; k < bc->n; ++k) density_limited(o, bc->cells[k], bc, k, 1);
    for(int64_t k = 0; k < bc->n; ++k) {
      const int64_t c = bc->cells[k];
      for(int dist = 0; dist < 4; ++dist)
        if(o->nghbr[c * o->stride + dist] == -1) fo[c * Q + (dist ^ 1)] = fo[c * Q + dist];
    }
    /* bnd_wall.h:393-414: the entry index is never advanced inside this loop -> the normal of entry 0 is used for all cells */
    const double* nrm = &bc->normals[0];
    for(int64_t k = 0; k < bc->n; ++k) {
      double* f = &fo[bc->cells[k] * Q];
      if(nrm[0] < 0) {
        f[6] = f[4] + 0.5 * (f[3] - f[2]);
        f[7] = f[5] - 0.5 * (f[3] - f[2]);
      } else if(nrm[0] > 0) {
        f[4] = f[6] - 0.5 * (f[3] - f[2]);
        f[5] = f[7] + 0.5 * (f[3] - f[2]);
      } else if(nrm[1] < 0) {
        f[5] = f[7] - 0.5 * (f[1] - f[0]);
        f[6] = f[4] + 0.5 * (f[1] - f[0]);
      } else if(nrm[1] > 0) {
        f[7] = f[5] + 0.5 * (f[1] - f[0]);
        f[4] = f[6] - 0.5 * (f[1] - f[0]);
      }
    }
  } else {
    const double* wv = bc->value;
    for(int64_t k = 0; k < bc->n; ++k) {
      const int64_t c = bc->cells[k];
      density_limited(o, c, bc, k, 0);
      for(int d = 0; d < D; ++d) o->vars[c * NV + d] = wv[d];
    }
    for(int64_t k = 0; k < bc->n; ++k) {
      const int64_t c   = bc->cells[k];
      const double* nrm = &bc->normals[k * D];
      for(int dist = 0; dist < 4; ++dist) {
        const int dir = dist / 2;
        if(o->nghbr[c * o->stride + dist] == -1 && fabs(nrm[dir]) > 0) {
          fo[c * Q + (dist ^ 1)] = fo[c * Q + dist];
          if(nrm[dir] < 0) fo[c * Q + (dist ^ 1)] -= 2.0 / 3.0 * o->vars[c * NV + D] * wv[dir];
          else fo[c * Q + (dist ^ 1)] += 2.0 / 3.0 * o->vars[c * NV + D] * wv[dir];
        }
      }
    }
    for(int64_t k = 0; k < bc->n; ++k) {
      const int64_t c   = bc->cells[k];
      const double  rho = o->vars[c * NV + D];
      const double* nrm = &bc->normals[k * D];
      double*       f   = &fo[c * Q];
      if(nrm[0] > 0) {
        f[6] = f[4] + 0.5 * (f[3] - f[2]) - 0.5 * rho * wv[1] - 1.0 / 6.0 * rho * wv[0];
        f[7] = f[5] - 0.5 * (f[3] - f[2]) + 0.5 * rho * wv[1] - 1.0 / 6.0 * rho * wv[0];
      } else if(nrm[0] < 0) {
        f[4] = f[6] - 0.5 * (f[3] - f[2]) + 0.5 * rho * wv[1] + 1.0 / 6.0 * rho * wv[0];
        f[5] = f[7] + 0.5 * (f[3] - f[2]) - 0.5 * rho * wv[1] + 1.0 / 6.0 * rho * wv[0];
      } else if(nrm[1] > 0) {
        f[5] = f[7] - 0.5 * (f[1] - f[0]) + 0.5 * rho * wv[0] - 1.0 / 6.0 * rho * wv[1];
        f[6] = f[4] + 0.5 * (f[1] - f[0]) - 0.5 * rho * wv[0] - 1.0 / 6.0 * rho * wv[1];
      } else if(nrm[1] < 0) {
        f[7] = f[5] + 0.5 * (f[1] - f[0]) - 0.5 * rho * wv[0] + 1.0 / 6.0 * rho * wv[1];
        f[4] = f[6] - 0.5 * (f[1] - f[0]) + 0.5 * rho * wv[0] + 1.0 / 6.0 * rho * wv[1];
      }
    }
  }
}

/* solver.cpp:743-755 -> bnd.h:60-65 */
static void pass_apply(Orc* o) {
  const OrcLattice* L = &o->L;
  const int Q = L->ndist, NV = o->nvar, D = L->ndim;
  for(int b = 0; b < o->nbc; ++b) {
    const OrcBc* bc = &o->bc[b];
    switch(bc->kind) {
      case ORC_BC_BB:
        for(int64_t k = 0; k < bc->n; ++k) bb_cell(o, bc->cells[k], &bc->normals[k * D], 0, NULL);
        break;
      case ORC_BC_BB_TANGENTIAL:
        for(int64_t k = 0; k < bc->n; ++k) bb_cell(o, bc->cells[k], &bc->normals[k * D], 1, &bc->wallval[k * Q]);
        break;
      case ORC_BC_DIRICHLET_BB:
        for(int64_t k = 0; k < bc->n; ++k) bb_cell(o, bc->cells[k], &bc->normals[k * D], 2, bc->value);
        break;
      case ORC_BC_PRESSURE: /* bnd_pressure.h:55-106 */
        for(int64_t k = 0; k < bc->n; ++k) {
          const int64_t c   = bc->cells[k];
          const double* nrm = &bc->normals[k * D];
          int           ins = -1;
          for(int d = 0; d < D && ins < 0; ++d) {
            if(nrm[d] < 0) ins = 2 * d + 1;
            else if(nrm[d] > 0) ins = 2 * d;
          }
          const int64_t n1 = o->nghbr[c * o->stride + ins];
          const int64_t n2 = o->nghbr[n1 * o->stride + ins];
          double        ue[ORC_MAXD];
          for(int d = 0; d < D; ++d) ue[d] = 1.5 * o->vars[n1 * NV + d] - 0.5 * o->vars[n2 * NV + d];
          o->vars[c * NV + D] = bc->pressure;
          for(int d = 0; d < D; ++d) o->vars[c * NV + d] = ue[d];
          for(int i = 0; i < Q - 1; ++i) {
            if(o->nghbr[c * o->stride + i] == -1 && in_direction(L, nrm, i)) {
              const int op         = L->opp[i];
              o->fold[c * Q + op] = -o->f[c * Q + i] + 2 * eq_dist(L, i, bc->pressure, ue, 1);
            }
          }
        }
        break;
      case ORC_BC_WALL_EQ:
        for(int64_t k = 0; k < bc->n; ++k) wall_eq_cell(o, bc, k);
        break;
      case ORC_BC_WALL_NEEM: { /* bnd_wall.h:262-298 */
        for(int64_t k = 0; k < bc->n; ++k) wall_eq_cell(o, bc, k);
        for(int64_t k = 0; k < bc->n; ++k) { /* calcDensity, moments.h:68-76 */
          const int64_t e = bc->ext[k];
          double rho = o->fold[e * Q];
          for(int i = 1; i < Q; ++i) rho += o->fold[e * Q + i];
          o->vars[e * NV + D] = rho;
        }
        for(int64_t k = 0; k < bc->n; ++k) { /* calcVelocity, moments.h:17-35 */
          const int64_t e = bc->ext[k];
          for(int d = 0; d < D; ++d) {
            double v = 0;
            for(int i = 0; i < Q - 1; ++i) v += L->c[i][d] * o->fold[e * Q + i];
            o->vars[e * NV + d] = v / o->vars[e * NV + D];
          }
        }
        for(int64_t k = 0; k < bc->n; ++k) {
          const int64_t c = bc->cells[k], e = bc->ext[k];
          for(int i = 0; i < Q; ++i)
            o->fold[c * Q + i] += o->fold[e * Q + i] - eq_dist(L, i, o->vars[e * NV + D], &o->vars[e * NV], 0);
        }
        break;
      }
      case ORC_BC_WALL_NEBB: wall_nebb(o, bc); break;
      case ORC_BC_NEUMANN_NEEM: /* bnd_neumann.h:47-60: the boundary value follows from the normal gradient, then Dirichlet */
        for(int64_t k = 0; k < bc->n; ++k) {
          const int64_t e  = bc->ext[k];
          const int64_t e2 = o->nghbr[e * o->stride + bc->extdir[k]];
          o->vars[e2]      = potential_of(L, &o->fold[e2 * Q]);
          bc->values[k]    = (4.0 * o->vars[e] - o->vars[e2] + bc->grad) / 3.0;
        }
        /* fall through */
      case ORC_BC_DIRICHLET_NEEM: /* bnd_dirichlet.h:347-365 */
        for(int64_t k = 0; k < bc->n; ++k) o->vars[bc->ext[k]] = potential_of(L, &o->fold[bc->ext[k] * Q]);
        for(int64_t k = 0; k < bc->n; ++k) {
          const int64_t c = bc->cells[k], e = bc->ext[k];
          o->vars[c] = bc->values[k];
          for(int i = 0; i < Q - 1; ++i) o->fold[c * Q + i] = L->w[i] * bc->values[k] + o->fold[e * Q + i] - L->w[i] * o->vars[e];
          o->fold[c * Q + Q - 1] = (L->w[Q - 1] - 1.0) * bc->values[k] + o->fold[e * Q + Q - 1] - (L->w[Q - 1] - 1.0) * o->vars[e];
        }
        break;
      default: break; /* periodic: apply is empty (bnd_periodic.h:208-209) */
    }
  }
}

/* solver.cpp:307-320 */
void orc_step(Orc* o, int64_t nsteps) {
  for(int64_t s = 0; s < nsteps; ++s) {
    memcpy(o->varsold, o->vars, sizeof(double) * (size_t)o->n * (size_t)o->nvar); /* currToOldVars :505-510 */
    pass_moments(o);
    pass_equilibrium(o);
    pass_collision(o);
    pass_forcing(o);
    pass_pre_apply(o);
    pass_propagation(o);
    pass_apply(o);
    ++o->step;
  }
}

/* The same step cut in two, so that a test can emulate a domain-decomposed run: everything before the propagation
 * (which produces m_f), then -- after the test has copied the m_f entries that cross a partition cut into its ghost
 * cells -- the propagation and the boundary conditions. */
void orc_step_collide(Orc* o) {
  memcpy(o->varsold, o->vars, sizeof(double) * (size_t)o->n * (size_t)o->nvar);
  pass_moments(o);
  pass_equilibrium(o);
  pass_collision(o);
  pass_forcing(o);
  pass_pre_apply(o);
}
void orc_step_stream(Orc* o) {
  pass_propagation(o);
  pass_apply(o);
  ++o->step;
}

/* solver.cpp:336 -- output() recomputes the moments of the current fold */
void orc_update_moments(Orc* o) { pass_moments(o); }

/* solver.cpp:809-815; returns 1 if NaN/Inf (solver.cpp:254-260) */
int orc_residual(const Orc* o, double* out) {
  int bad = 0;
  for(int v = 0; v < o->nvar; ++v) {
    double conv = 0.0;
    for(int64_t c = 0; c < o->n; ++c) conv += fabs(o->vars[c * o->nvar + v] - o->varsold[c * o->nvar + v]);
    out[v] = conv;
    if(isnan(conv) || isinf(conv)) bad = 1;
  }
  return bad;
}

double*  orc_f(Orc* o) { return o->f; }
double*  orc_fold(Orc* o) { return o->fold; }
double*  orc_feq(Orc* o) { return o->feq; }
double*  orc_vars(Orc* o) { return o->vars; }
double*  orc_varsold(Orc* o) { return o->varsold; }
int64_t  orc_steps_done(const Orc* o) { return o->step; }
int      orc_nvar(const Orc* o) { return o->nvar; }
int      orc_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
