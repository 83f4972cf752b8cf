// geometry.hpp -- analytic geometries and the inside / cut tests the grid pipeline asks (host code).
//
// Mirrors the reference's GeometryManager for analytic objects (/root/reference/src/geometry.h): GeomBox (:700-850),
// GeomSphere (:640-698), GeomCube (:852-915), body / subtract logic of pointIsInside (:1011-1048), cutWithCell over all
// objects (:1055-1063) and over one named object (:1065-1075), bounding box (:1086-1104).  STL geometries are out of scope
// (no reference configuration uses one).
// 3D cut tests for box and sphere do not exist in the reference (TERMM("impl") :796-798, TERMM("not implemented") :690-692);
// the 3D forms here extend the 2D source text dimension by dimension and are "parity unpinned".
#pragma once
#include <cmath>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "json.hpp"

namespace lbmhost {

static constexpr double kEps = std::numeric_limits<double>::epsilon(); // GDoubleEps

struct Geom {
  std::string name, body = "unique", type;
  bool        subtract = false;
  int         ndim     = 2;
  virtual ~Geom()                                                  = default;
  virtual bool point_inside(const double* x) const                  = 0;
  virtual bool cut_with_cell(const double* c, double len) const     = 0;
  virtual void bbox(double* lo, double* hi) const                   = 0;
};

struct GeomBox final : Geom {
  double A[3] = {0, 0, 0}, B[3] = {0, 0, 0};
  bool point_inside(const double* x) const override { // geometry.h:752-760
    for(int d = 0; d < ndim; ++d)
      if(A[d] > x[d] || B[d] < x[d]) return false;
    return true;
  }
  bool cut_with_cell(const double* c, double len) const override { // geometry.h:764-809
    const double h = 0.5 * len;
    if(ndim == 1) return std::abs(A[0] - c[0]) <= h || std::abs(B[0] - c[0]) <= h;
    // for every axis a: the cell lies within the box extended by h in all OTHER axes and touches face A_a or B_a.
    // In 2D this is exactly the reference's two blocks (x-range -> y faces :779-786, y-range -> x faces :787-794); the 3D
    // form ("range in all axes but a, faces of axis a") is the extension the reference leaves as TERMM("impl").
    for(int a = ndim - 1; a >= 0; --a) { // 2D order of the reference: y faces first (:779-786), then x faces (:787-794)
      bool in_range = true;
      for(int r = 0; r < ndim; ++r) {
        if(r == a) continue;
        if(!(A[r] - h <= c[r] && B[r] + h >= c[r])) in_range = false;
      }
      if(in_range && (std::abs(c[a] - A[a]) <= h || std::abs(c[a] - B[a]) <= h)) return true;
    }
    return false;
  }
  void bbox(double* lo, double* hi) const override {
    for(int d = 0; d < ndim; ++d) { lo[d] = A[d]; hi[d] = B[d]; }
  }
};

struct GeomSphere final : Geom {
  double C[3] = {0, 0, 0}, R = 0;
  double dist(const double* x) const {
    double s = 0; // Eigen's norm(): sqrt of the sum of squares, summed in index order
    for(int d = 0; d < ndim; ++d) s += (x[d] - C[d]) * (x[d] - C[d]);
    return std::sqrt(s);
  }
  bool point_inside(const double* x) const override { return dist(x) < R + kEps; } // geometry.h:660-662
  bool cut_with_cell(const double* c, double len) const override {                   // geometry.h:664-698
    const double distance = dist(c);
    const double h        = 0.5 * len;
    const double rr       = std::sqrt(ndim == 3 ? 3.0 : 2.0) * h; // reference (2D): gcem::sqrt(2) * halfCellLength
    if(distance <= R + rr && distance >= R - rr) {
      int inside = 0;
      const int nvert = 1 << ndim;
      for(int v = 0; v < nvert; ++v) {
        // 2D loop order of the reference: dirX outer, dirY inner; vertex = origin + len*dirX*ex + len*dirY*ey
        double p[3];
        int    bits[3];
        if(ndim == 2) { bits[0] = (v >> 1) & 1; bits[1] = v & 1; }
        else { bits[0] = (v >> 2) & 1; bits[1] = (v >> 1) & 1; bits[2] = v & 1; }
        for(int d = 0; d < ndim; ++d) p[d] = (c[d] - h) + len * bits[d];
        if(dist(p) <= R) ++inside;
      }
      if(inside < nvert) return true;
    }
    return false;
  }
  void bbox(double* lo, double* hi) const override {
    for(int d = 0; d < ndim; ++d) { lo[d] = C[d] - R; hi[d] = C[d] + R; }
  }
};

struct GeomCube final : Geom {
  double C[3] = {0, 0, 0}, length = 0; // `length` acts as a half-width (geometry.h:866-873)
  bool point_inside(const double* x) const override {
    for(int d = 0; d < ndim; ++d)
      if(std::abs(x[d] - C[d]) > length) return false;
    return true;
  }
  bool cut_with_cell(const double* c, double len) const override { // geometry.h:875-882
    for(int d = 0; d < ndim; ++d)
      if(std::abs(c[d] - C[d]) > length + len) return false;
    return true;
  }
  void bbox(double* lo, double* hi) const override { // geometry.h:884-893
    const double r = std::sqrt(static_cast<double>(ndim)) * length;
    for(int d = 0; d < ndim; ++d) { lo[d] = C[d] - r; hi[d] = C[d] + r; }
  }
};

class GeometryManager {
 public:
  int ndim = 2;
  std::vector<std::unique_ptr<Geom>> objs; // json key order (sorted), geometry.h:929-966

  void setup(const Json& geometry, int dim) {
    ndim = dim;
    for(const auto& kv : geometry.obj) {
      const Json& g = kv.second;
      if(!g.has("type")) throw std::runtime_error("Malformed json: No \"type\" given!");
      const std::string type = g.at("type").as_string();
      std::unique_ptr<Geom> o;
      if(type == "box") {
        auto b = std::make_unique<GeomBox>();
        const auto A = g.at("A").as_doubles(), B = g.at("B").as_doubles();
        if(static_cast<int>(A.size()) != dim || static_cast<int>(B.size()) != dim)
          throw std::runtime_error("Invalid dimensionality given for box corner points");
        for(int d = 0; d < dim; ++d) {
          b->A[d] = A[d];
          b->B[d] = B[d];
          if(A[d] > B[d]) throw std::runtime_error("ERROR: The specification of the box is invalid");
        }
        o = std::move(b);
      } else if(type == "sphere") {
        auto s = std::make_unique<GeomSphere>();
        const auto C = g.at("center").as_doubles();
        for(int d = 0; d < dim && d < static_cast<int>(C.size()); ++d) s->C[d] = C[d];
        s->R = g.at("radius").as_double();
        o = std::move(s);
      } else if(type == "cube") {
        auto c = std::make_unique<GeomCube>();
        const auto C = g.at("center").as_doubles();
        for(int d = 0; d < dim && d < static_cast<int>(C.size()); ++d) c->C[d] = C[d];
        c->length = g.at("length").as_double();
        o = std::move(c);
      } else if(type == "stl") {
        throw std::runtime_error("STL geometries are not supported by this host (out of scope, SURVEY.md section 2 row 8)");
      } else {
        continue; // "Unknown geometry type": the reference logs and skips (geometry.h:957-961)
      }
      o->name     = kv.first;
      o->type     = type;
      o->ndim     = dim;
      o->body     = g.opt_str("body", "unique");
      o->subtract = g.opt_bool("subtract", false);
      if(o->body == "unique") o->body = o->name;
      objs.push_back(std::move(o));
    }
    // bodies in first-appearance order (the reference iterates an unordered_map; the order only matters when several
    // bodies with subtraction overlap, which no reference configuration has)
    for(const auto& o : objs) {
      bool known = false;
      for(const auto& b : bodies) known = known || b == o->body;
      if(!known) bodies.push_back(o->body);
    }
  }

  bool point_inside(const double* x) const { // geometry.h:1011-1048
    for(const std::string& body : bodies) {
      bool has_sub = false;
      for(const auto& o : objs)
        if(o->body == body && o->subtract) has_sub = true;
      for(const auto& o : objs) {
        if(o->body != body) continue;
        if(o->point_inside(x)) {
          if(!has_sub) return true;
          if(o->subtract) return false;
          for(const auto& s : objs)
            if(s->body == body && s->subtract && s->point_inside(x)) return false;
          return true;
        }
      }
    }
    return false;
  }
  bool cut_with_cell(const double* c, double len) const {
    for(const auto& o : objs)
      if(o->cut_with_cell(c, len)) return true;
    return false;
  }
  bool cut_with_cell(const std::string& name, const double* c, double len) const {
    for(const auto& o : objs)
      if(o->name == name && o->cut_with_cell(c, len)) return true;
    return false;
  }
  void bbox(double* lo, double* hi) const { // geometry.h:1086-1104
    for(size_t k = 0; k < objs.size(); ++k) {
      double a[3], b[3];
      objs[k]->bbox(a, b);
      for(int d = 0; d < ndim; ++d) {
        if(k == 0 || lo[d] > a[d]) lo[d] = a[d];
        if(k == 0 || hi[d] < b[d]) hi[d] = b[d];
      }
    }
  }
  size_t size() const { return objs.size(); }

 private:
  std::vector<std::string> bodies;
};

} // namespace lbmhost
