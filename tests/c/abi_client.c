/* A plain C99 client of include/lbm_b200.h: proves that the drop-in boundary is a C ABI (no C++ / torch types in the signatures) and
 * exercises the entry points that need no GPU -- the ones a reference-side binding would call first.  Built and run by
 * tests/test_abi_c_client.py with gcc -std=c99 -pedantic. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lbm_b200.h"

static const int64_t SHAPE[2] = {8, 4};
static int64_t*      g_table  = NULL; /* full table of the 8 x 4 box */

static int rows_from_table(void* user, const int64_t* ids, int64_t n, int64_t* rows, int64_t* sources) {
  const int64_t N = 32;
  (void)user;
  for(int64_t r = 0; r < n; ++r) {
    for(int j = 0; j < 8; ++j) {
      if(rows != NULL) rows[r * 8 + j] = g_table[ids[r] * 8 + j];
      if(sources != NULL) {
        int64_t src = -1;
        for(int64_t c = 0; c < N; ++c)
          if(g_table[c * 8 + j] == ids[r]) src = c; /* the highest source wins */
        sources[r * 8 + j] = src;
      }
    }
  }
  return 0;
}

#define CHECK(cond)                                                        \
  do {                                                                     \
    if(!(cond)) {                                                          \
      fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, lbm_b200_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while(0)

int main(void) {
  const int32_t periodic[2] = {1, 0};
  lbm_b200_config          cfg;
  lbm_b200_solver*         s = NULL;
  lbm_b200_partition*      part = NULL;
  lbm_b200_partition_view  v;
  lbm_b200_plan_view       plan;
  double                   x[2] = {0.75, 0.25};
  int64_t                  n;

  CHECK(lbm_b200_abi_version() == LBM_B200_ABI_VERSION);
  n = lbm_b200_box_ncells(2, SHAPE);
  CHECK(n == 32);
  g_table = (int64_t*)malloc(sizeof(int64_t) * (size_t)n * 8);
  CHECK(lbm_b200_box_topology(2, SHAPE, periodic, g_table, 8, NULL, NULL) == LBM_B200_OK);
  CHECK(lbm_b200_sfc_index(2, x, 1) == 3); /* UnitTest/test_hilbert.cpp: the quadrant x >= 0.5, y < 0.5 */

  /* the local problem of rank 1 of 2, then an inspection-only solver (device -1) set up from it: everything up to the device plan */
  CHECK(lbm_b200_partition_create(n, 2, 9, 8, 1, 2, rows_from_table, NULL, 0, NULL, NULL, NULL, &part) == LBM_B200_OK);
  CHECK(lbm_b200_partition_get(part, &v) == LBM_B200_OK);
  CHECK(v.n_owned == 16 && v.n_ghost > 0 && v.npeers == 1 && v.peers[0] == 0);
  lbm_b200_default_config(&cfg);
  cfg.device = -1;
  cfg.omega  = 1.2;
  CHECK(lbm_b200_create(&cfg, v.n_owned + v.n_ghost, &s) == LBM_B200_OK);
  CHECK(lbm_b200_set_topology(s, v.nghbr, v.stride) == LBM_B200_OK);
  CHECK(lbm_b200_partition_apply(part, s) == LBM_B200_OK);
  CHECK(lbm_b200_debug_plan(s, &plan) == LBM_B200_OK);
  CHECK(plan.n_owned == 16 && plan.n_send > 0 && plan.n_send == plan.n_recv);
  /* no CPU compute path: an inspection handle refuses to initialise */
  CHECK(lbm_b200_init(s) == LBM_B200_ECUDA);
  CHECK(strstr(lbm_b200_last_error(), "no CPU compute path") != NULL);
  lbm_b200_destroy(s);
  lbm_b200_partition_destroy(part);
  free(g_table);
  printf("abi_client ok\n");
  return 0;
}
