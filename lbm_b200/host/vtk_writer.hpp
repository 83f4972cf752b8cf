// vtk_writer.hpp -- the solution file of the reference, byte for byte (host code).
//
// The reference writes `out/<solution_filename>_<step>.vtp` through VTK::BINARY::writePoints (/root/reference/src/common/IO.h:414-507),
// fed by LBMSolver::output (/root/reference/src/lbm/solver.cpp:323-384).  What ends up in the file, restated from the file itself and
// from those two functions:
//   * every value goes through a decimal string first (toStringVector: fixed notation, 15 decimals,
//     /root/reference/include/common/util/string_helper.h:93-107) and back through std::stod (IO.h:479) -- i.e. it is rounded to a
//     multiple of 1e-15 and a small negative value becomes -0.0;
//   * each array is base64( uint64 header || little-endian data ) (base64.h:216-270), the header being 8 x the NUMBER OF ELEMENTS
//     (binary::BYTE_SIZE * length, base64.h:236) -- the byte count only when the element is 8 bytes wide, twice the byte count for
//     the Float32 point coordinates; '=' padding is 4 - (chars mod 4) characters (IO.h:395-399);
//   * point coordinates are narrowed to float and padded to three components, connectivity is 0..n-1 as int64.
// The reference's decimal round trip costs two locale-aware stream operations per value (16-35 % of its main loop, SURVEY.md
// section 6); round15() below gets the same double with one 128-bit multiply, a shift and one division, and the base64 text is
// produced in parallel.  `tests/test_vtk_writer.py` compares the bytes against files written by the reference binary.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace lbmhost {
namespace vtk {

// The double that `std::stod` returns for the text `std::fixed << std::setprecision(15) << x` prints.
// |x| = m * 2^e exactly; N = round-half-even(m * 10^15 * 2^e) is what the 15 printed decimals spell (glibc prints the exactly
// rounded decimal expansion); the text parses to the correctly rounded quotient N / 10^15, which is what an IEEE division of
// the two exactly representable integers yields as long as N < 2^53.  Larger magnitudes (|x| > 9.007) take the slow way.
inline double round15(double x) {
  if(!std::isfinite(x)) return x;
  uint64_t bits;
  std::memcpy(&bits, &x, sizeof bits);
  const uint64_t frac = bits & ((uint64_t(1) << 52) - 1);
  const int      bexp = static_cast<int>((bits >> 52) & 0x7FF);
  const uint64_t m    = bexp == 0 ? frac : (frac | (uint64_t(1) << 52));
  const int      e    = (bexp == 0 ? 1 : bexp) - 1075; // |x| = m * 2^e
  if(e < 0) {
    const int k = -e;
    unsigned __int128 q = 0;
    if(k < 104) { // m * 10^15 < 2^103, so for k >= 104 the value is below one half and rounds to zero
      const unsigned __int128 P    = static_cast<unsigned __int128>(m) * static_cast<unsigned __int128>(1000000000000000ULL);
      const unsigned __int128 half = static_cast<unsigned __int128>(1) << (k - 1);
      const unsigned __int128 rem  = P & ((half << 1) - 1);
      q                            = P >> k;
      if(rem > half || (rem == half && (q & 1))) ++q;
    }
    if(q < (static_cast<unsigned __int128>(1) << 53)) return std::copysign(static_cast<double>(static_cast<uint64_t>(q)) / 1e15, x);
  }
  char buf[400];
  std::snprintf(buf, sizeof buf, "%.15f", x);
  return std::strtod(buf, nullptr);
}

// standard base64 of `n` bytes into out[0 .. 4*ceil(n/3)) WITHOUT padding characters; returns the number of characters
inline size_t base64_into(const uint8_t* in, size_t n, char* out) {
  static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  const int64_t full = static_cast<int64_t>(n / 3);
#pragma omp parallel for schedule(static) if(full > (1 << 16))
  for(int64_t g = 0; g < full; ++g) {
    const uint8_t* p = in + 3 * g;
    const uint32_t v = (uint32_t(p[0]) << 16) | (uint32_t(p[1]) << 8) | p[2];
    char*          o = out + 4 * g;
    o[0] = T[v >> 18];
    o[1] = T[(v >> 12) & 63];
    o[2] = T[(v >> 6) & 63];
    o[3] = T[v & 63];
  }
  size_t         nc  = static_cast<size_t>(full) * 4;
  const size_t   rem = n - static_cast<size_t>(full) * 3;
  const uint8_t* p   = in + 3 * full;
  if(rem == 1) {
    out[nc++] = T[p[0] >> 2];
    out[nc++] = T[(p[0] & 3) << 4];
  } else if(rem == 2) {
    out[nc++] = T[p[0] >> 2];
    out[nc++] = T[((p[0] & 3) << 4) | (p[1] >> 4)];
    out[nc++] = T[(p[1] & 15) << 2];
  }
  return nc;
}

// one <DataArray> payload: header (8 * element count) + data, base64, the reference's padding rule; appended to `text`.
// The 8 header bytes and the first data byte form three complete groups, so the data is encoded in place from its second byte.
template <class T>
inline void append_array(std::string& text, const T* data, int64_t length) {
  const size_t   nbytes = sizeof(T) * static_cast<size_t>(length);
  const uint64_t header = static_cast<uint64_t>(length) * 8; // base64.h:236: length * BYTE_SIZE, whatever sizeof(T) is
  const uint8_t* bytes  = reinterpret_cast<const uint8_t*>(data);
  uint8_t        head[9];
  std::memcpy(head, &header, 8);
  head[8]         = bytes[0];
  const size_t at = text.size();
  text.resize(at + (8 + nbytes + 2) / 3 * 4 + 4);
  size_t nchars = base64_into(head, 9, &text[at]);
  nchars += base64_into(bytes + 1, nbytes - 1, &text[at + nchars]);
  // IO.h:395-399: ceil(bytes * 8 / 6) characters, then 4 - (chars mod 4) pad characters.  For chars mod 4 == 0 the reference
  // indexes one past its 4-entry padding table (undefined behaviour); correct base64 needs no padding there and none is written.
  const size_t pad = nchars % 4 == 0 ? 0 : 4 - nchars % 4;
  text.resize(at + nchars);
  text.append(pad, '=');
}

struct Column {
  std::string   name;
  const double* values; // [n_all * stride]
  int64_t       stride; // distance between consecutive cells
  // payload already encoded on the device (lbm_b200_encode_output: filter, rounding and base64 done there): written as is
  const char*   encoded     = nullptr;
  size_t        encoded_len = 0;
};

// The file, handed to `sink(const std::string&)` piece by piece (one piece per array, the buffer is reused).
// `keep` (may be null = all) is the cell filter (cell_filter.h:84-96) evaluated per cell.
template <class Sink>
inline void points_stream(Sink&& sink, int ndim, int64_t n_all, const double* center, const uint8_t* keep, const std::vector<Column>& columns) {
  std::vector<int64_t> ids; // kept cells; empty when nothing is filtered
  int64_t              n = n_all;
  if(keep != nullptr) {
    for(int64_t c = 0; c < n_all; ++c)
      if(keep[c]) ids.push_back(c);
    n = static_cast<int64_t>(ids.size());
  }
  auto id = [&](int64_t k) { return keep != nullptr ? ids[k] : k; };
  std::string t;
  t.reserve(static_cast<size_t>(n) * 16 + 4096);
  t += "<VTKFile type=\"PolyData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n  <PolyData>\n";
  t += "<Piece NumberOfPoints=\"" + std::to_string(n) + "\" NumberOfVerts=\"1\" NumberOfLines=\"0\" NumberOfStrips=\"0\" NumberOfPolys=\"0\" > \n";
  t += "<Points>\n<DataArray type=\"Float32\" Name=\"Points\" NumberOfComponents=\"3\" format=\"binary\"> \n";
  {
    std::vector<float> xyz(static_cast<size_t>(n) * 3, 0.0f);
#pragma omp parallel for schedule(static) if(n > (1 << 16))
    for(int64_t k = 0; k < n; ++k)
      for(int d = 0; d < ndim; ++d) xyz[static_cast<size_t>(k) * 3 + d] = static_cast<float>(center[id(k) * ndim + d]);
    append_array(t, xyz.data(), n * 3);
  }
  t += "\n        </DataArray>\n      </Points>\n      <Verts>\n        <DataArray type=\"Int64\" Name=\"connectivity\" format=\"binary\"> \n";
  sink(t);
  t.clear();
  std::vector<double> col(static_cast<size_t>(n)); // also the connectivity buffer (same width)
  {
    int64_t* conn = reinterpret_cast<int64_t*>(col.data());
    for(int64_t k = 0; k < n; ++k) conn[k] = k;
    append_array(t, conn, n);
  }
  t += "\n        </DataArray> \n        <DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\"> \n" + std::to_string(n)
       + "\n        </DataArray> \n        </Verts> \n      <PointData> \n";
  sink(t);
  for(const Column& c : columns) {
    t.clear();
    t += "<DataArray type=\"Float64\" Name=\"" + c.name + "\" format=\"binary\">\n";
    if(c.encoded != nullptr) {
      t.append(c.encoded, c.encoded_len);
    } else {
#pragma omp parallel for schedule(static) if(n > (1 << 14))
      for(int64_t k = 0; k < n; ++k) col[k] = round15(c.values[id(k) * c.stride]);
      append_array(t, col.data(), n);
    }
    t += "\n        </DataArray> \n";
    sink(t);
  }
  sink(std::string("      </PointData> \n    </Piece>\n  </PolyData>\n</VTKFile> \n"));
}

inline bool write_points(const std::string& path, int ndim, int64_t n_all, const double* center, const uint8_t* keep,
                         const std::vector<Column>& columns) {
  if(n_all <= 0) return false; // the reference exits in encodeLE_header for an empty array (base64.h:219-223)
  if(keep != nullptr) {
    bool any = false;
    for(int64_t c = 0; c < n_all && !any; ++c) any = keep[c] != 0;
    if(!any) return false;
  }
  std::FILE* f = std::fopen(path.c_str(), "wb");
  if(f == nullptr) return false;
  bool ok = true;
  points_stream([&](const std::string& piece) { ok = ok && std::fwrite(piece.data(), 1, piece.size(), f) == piece.size(); }, ndim, n_all, center,
                keep, columns);
  return (std::fclose(f) == 0) && ok;
}

} // namespace vtk
} // namespace lbmhost
