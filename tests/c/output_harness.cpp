// output_harness.cpp -- TEST INFRASTRUCTURE: the device code of lbm_b200/csrc/output.cuh (15-decimal rounding, base64 of a field) executed
// on the CPU against the host writer lbm_b200/host/vtk_writer.hpp, which tests/test_vtk_writer.py pins byte for byte against files written
// by the reference binary.
#include <cstring>
#include <string>
#include <vector>

#include "../../lbm_b200/csrc/output.cuh"
#include "../../lbm_b200/host/vtk_writer.hpp"

// k_output_column over a fake launch: column `var` of the cells sel[] of a [nvar][stride] array, rounded; returns the number of entries
// that differ from the host's round15 of the same value (entries the device flags as slow are compared as passed through unchanged)
template <class Real>
static int64_t check_column(const Real* src, int64_t stride, int var, const int32_t* sel, int64_t n, int* slow_out) {
  std::vector<double> col(static_cast<size_t>(n), -1.0);
  int slow = 0;
  blockDim.x = 128;
  gridDim.x  = static_cast<unsigned>((n + 127) / 128);
  for(unsigned b = 0; b < gridDim.x; ++b)
    for(unsigned t = 0; t < 128; ++t) {
      blockIdx.x = b;
      threadIdx.x = t;
      lbm::out::k_output_column<Real>(src, stride, var, sel, n, col.data(), &slow);
    }
  int64_t bad = 0;
  for(int64_t k = 0; k < n; ++k) {
    const double x = static_cast<double>(src[static_cast<size_t>(var) * stride + sel[k]]);
    int s = 0;
    lbm::out::round15(x, &s);
    const double want = s ? x : lbmhost::vtk::round15(x);
    bad += std::memcmp(&col[k], &want, 8) != 0 ? 1 : 0;
  }
  *slow_out = slow;
  return bad;
}

extern "C" {

// number of values whose device rounding differs from the host's (values flagged "slow" by the device must be exactly those the host sends
// through its strtod path: they are skipped in the comparison and counted in *n_slow)
int64_t oh_check_round15(const double* x, int64_t n, int64_t* n_slow) {
  int64_t bad = 0;
  *n_slow = 0;
  for(int64_t i = 0; i < n; ++i) {
    int slow = 0;
    const double d = lbm::out::round15(x[i], &slow);
    if(slow) { ++*n_slow; continue; }
    const double h = lbmhost::vtk::round15(x[i]);
    bad += std::memcmp(&d, &h, 8) != 0 ? 1 : 0;
  }
  return bad;
}

// base64 payload of one field from k_base64_field (one "thread" per group) == vtk::append_array's text ?
int oh_check_base64(const double* col, int64_t n) {
  const int64_t chars = lbm::out::base64_chars(n), ngroups = chars / 4;
  std::vector<char> text(static_cast<size_t>(chars));
  blockDim.x = 256;
  gridDim.x  = static_cast<unsigned>((ngroups + 255) / 256);
  for(unsigned b = 0; b < gridDim.x; ++b)
    for(unsigned t = 0; t < 256; ++t) {
      blockIdx.x = b;
      threadIdx.x = t;
      lbm::out::k_base64_field(col, n, static_cast<unsigned long long>(n) * 8ull, text.data(), ngroups);
    }
  std::string want;
  lbmhost::vtk::append_array(want, col, n);
  return want.size() == text.size() && std::memcmp(want.data(), text.data(), text.size()) == 0 ? 0 : 1;
}

int64_t oh_check_column_f64(const double* src, int64_t stride, int var, const int32_t* sel, int64_t n, int* slow) { return check_column(src, stride, var, sel, n, slow); }
int64_t oh_check_column_f32(const float* src, int64_t stride, int var, const int32_t* sel, int64_t n, int* slow) { return check_column(src, stride, var, sel, n, slow); }

// the host writer's work for the fields of one file (vtk_writer.hpp, points_stream): per variable, the strided column of the kept cells
// through round15, then base64.  Returns the number of characters produced; timed by tools/bench_output.py beside the device encoder.
int64_t oh_host_encode(const double* vars, int64_t n, int nvar) {
  int64_t             chars = 0;
  std::vector<double> col(static_cast<size_t>(n));
  for(int v = 0; v < nvar; ++v) {
#pragma omp parallel for schedule(static) if(n > (1 << 14))
    for(int64_t k = 0; k < n; ++k) col[k] = lbmhost::vtk::round15(vars[k * nvar + v]);
    std::string t;
    lbmhost::vtk::append_array(t, col.data(), n);
    chars += static_cast<int64_t>(t.size());
  }
  return chars;
}

} // extern "C"
