/* Single-rank stand-in for <mpi.h>, used ONLY to compile the unmodified reference
 * (SFCMM/LBM, /root/reference) into oracle/_ref/lbm_ref on a machine without an MPI
 * installation.  TEST INFRASTRUCTURE: nothing in the product links or includes this.
 *
 * The reference carries no simulation data over MPI: it only
 * initialises MPI, asks for rank/size, synchronises and averages timers.  Every call
 * below is therefore the trivial one-rank answer.
 */
#ifndef LBM_B200_ORACLE_MPI_SHIM_H
#define LBM_B200_ORACLE_MPI_SHIM_H
#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Info;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Errhandler;

#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_MAX_INFO_KEY 256
#define MPI_MAX_INFO_VAL 1024
#define MPI_DOUBLE 1
#define MPI_SUM 1
#define MPI_THREAD_FUNNELED 1
#define MPI_ERRORS_RETURN 0
#define MPI_SUCCESS 0

inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
inline int MPI_Init_thread(int*, char***, int required, int* provided) {
  *provided = required;
  return MPI_SUCCESS;
}
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* rank) {
  *rank = 0;
  return MPI_SUCCESS;
}
inline int MPI_Comm_size(MPI_Comm, int* size) {
  *size = 1;
  return MPI_SUCCESS;
}
inline int MPI_Comm_set_errhandler(MPI_Comm, MPI_Errhandler) { return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Abort(MPI_Comm, int code) {
  std::exit(code);
  return MPI_SUCCESS;
}
inline int MPI_Reduce(const void* send, void* recv, int count, MPI_Datatype, MPI_Op, int, MPI_Comm) {
  std::memcpy(recv, send, sizeof(double) * static_cast<size_t>(count));
  return MPI_SUCCESS;
}
inline double MPI_Wtime() {
  using clk = std::chrono::steady_clock;
  return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}
inline int MPI_Info_create(MPI_Info* info) {
  *info = 0;
  return MPI_SUCCESS;
}
inline int MPI_Info_set(MPI_Info, const char*, const char*) { return MPI_SUCCESS; }
inline int MPI_Info_get_nkeys(MPI_Info, int* n) {
  *n = 0;
  return MPI_SUCCESS;
}
inline int MPI_Info_get_nthkey(MPI_Info, int, char* key) {
  key[0] = 0;
  return MPI_SUCCESS;
}
inline int MPI_Info_get_valuelen(MPI_Info, const char*, int* len, int* flag) {
  *len  = 0;
  *flag = 0;
  return MPI_SUCCESS;
}
inline int MPI_Info_get(MPI_Info, const char*, int, char* value, int* flag) {
  value[0] = 0;
  *flag    = 0;
  return MPI_SUCCESS;
}
#endif
