// output.cuh -- device side of the reference's solution output.
//
// LBMSolver::output (/root/reference/src/lbm/solver.cpp:323-384) recomputes the macroscopic fields, filters the cells
// (cell_filter.h:84-96), turns every value into a 15-decimal string and back (string_helper.h:93-107, IO.h:479), and writes each
// field as base64( uint64 header || little-endian doubles ) with '=' padding (base64.h:216-270, IO.h:395-399).  On the host that
// costs 0.48 s per 256^3 output on 16 cores even with lbm_b200/host/vtk_writer.hpp's fast rounding (3 s on two); here the filter gather,
// the rounding and the base64 text are produced on the device, and what crosses PCIe is the text the file stores (16 ms into page-locked
// memory).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace lbm {
namespace out {

// The double that std::stod returns for the text `std::fixed << std::setprecision(15) << x` prints (vtk_writer.hpp: round15, same
// algorithm): |x| = m 2^e exactly; N = round-half-even(m 10^15 2^e); result = N / 10^15 (one IEEE division of two exactly
// representable integers) as long as N < 2^53.  *slow is set when the value needs the host's slow path (|x| >= 9.007..., NaN, Inf).
__device__ __forceinline__ double round15(double x, int* slow) {
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(x));
  const unsigned long long frac = bits & ((1ull << 52) - 1);
  const int                bexp = static_cast<int>((bits >> 52) & 0x7FF);
  if(bexp == 0x7FF) { *slow = 1; return x; }
  const unsigned long long m = bexp == 0 ? frac : (frac | (1ull << 52));
  const int                e = (bexp == 0 ? 1 : bexp) - 1075; // |x| = m * 2^e
  if(e >= 0) { *slow = 1; return x; }
  const int k = -e;
  unsigned long long q = 0;
  if(k < 104) { // P = m * 10^15 < 2^103: for k >= 104 the value is below one half and rounds to zero
    const unsigned long long T  = 1000000000000000ull;
    const unsigned long long lo = m * T, hi = __umul64hi(m, T); // P = hi : lo
    // q = P >> k, rem = P mod 2^k, half = 2^(k-1)
    unsigned long long qhi, qlo, rhi, rlo, hhi, hlo;
    if(k >= 64) {
      qhi = 0; qlo = hi >> (k - 64);
      rhi = (k - 64) == 0 ? 0 : (hi & ((1ull << (k - 64)) - 1)); rlo = lo;
      hhi = (k - 1) >= 64 ? (1ull << (k - 1 - 64)) : 0; hlo = (k - 1) >= 64 ? 0 : (1ull << (k - 1));
    } else {
      qhi = hi >> k; qlo = (lo >> k) | (hi << (64 - k)); // k in 1..63
      rhi = 0; rlo = lo & ((1ull << k) - 1);
      hhi = 0; hlo = 1ull << (k - 1);
    }
    const bool gt = rhi > hhi || (rhi == hhi && rlo > hlo);
    const bool eq = rhi == hhi && rlo == hlo;
    if(gt || (eq && (qlo & 1))) { ++qlo; if(qlo == 0) ++qhi; }
    if(qhi != 0 || qlo >= (1ull << 53)) { *slow = 1; return x; }
    q = qlo;
  }
  const double r = __ddiv_rn(static_cast<double>(q), 1e15);
  return (bits >> 63) ? -r : r;
}

// column v of the output: the kept cells' values of variable v, rounded; src = per-cell SoA array [nvar][stride] in device cell order,
// sel[k] = device cell of the k-th kept cell (reference order)
template <class Real>
__global__ void k_output_column(const Real* __restrict__ src, int64_t stride, int var, const int32_t* __restrict__ sel, int64_t n,
                                double* __restrict__ col, int* __restrict__ slow) {
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(k >= n) return;
  int s = 0;
  col[k] = round15(static_cast<double>(src[static_cast<size_t>(var) * stride + sel[k]]), &s);
  if(s) *slow = 1;
}

// base64( header || col[0..n) ), padded: 4 characters per 3-byte group, one group per thread
static __global__ void k_base64_field(const double* __restrict__ col, int64_t n, unsigned long long header, char* __restrict__ text, int64_t ngroups) {
  const int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if(g >= ngroups) return;
  const int64_t nbytes = 8 + 8 * n;
  const unsigned char* data = reinterpret_cast<const unsigned char*>(col);
  unsigned b[3];
  int      have = 0;
#pragma unroll
  for(int t = 0; t < 3; ++t) {
    const int64_t i = 3 * g + t;
    if(i < nbytes) {
      b[t] = i < 8 ? static_cast<unsigned>((header >> (8 * i)) & 0xFF) : data[i - 8];
      ++have;
    } else b[t] = 0;
  }
  const unsigned v = (b[0] << 16) | (b[1] << 8) | b[2];
  const char* T = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  char c0 = T[v >> 18], c1 = T[(v >> 12) & 63], c2 = T[(v >> 6) & 63], c3 = T[v & 63];
  if(have == 1) { c2 = '='; c3 = '='; }
  if(have == 2) c3 = '=';
  // one 32-bit store per group (text and every field's offset in it are multiples of four characters)
  reinterpret_cast<uint32_t*>(text)[g] = static_cast<uint32_t>(static_cast<unsigned char>(c0)) | (static_cast<uint32_t>(static_cast<unsigned char>(c1)) << 8)
                                         | (static_cast<uint32_t>(static_cast<unsigned char>(c2)) << 16) | (static_cast<uint32_t>(static_cast<unsigned char>(c3)) << 24);
}

static inline int64_t base64_chars(int64_t n_values) { return (8 + 8 * n_values + 2) / 3 * 4; }

} // namespace out
} // namespace lbm
