// json.hpp -- small JSON reader for the reference's configuration files (host code).
//
// The reference reads its configuration with nlohmann::json 3.9.1, whose objects are std::map: iterating
// `solver.boundary` visits geometry names and surface keys in byte-lexicographic order, and that order is the boundary
// condition application order (/root/reference/src/lbm/bnd/bnd.h:71-142).  Objects here are
// std::map as well, so iteration order is the same by construction.
#pragma once
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace lbmhost {

class Json {
 public:
  enum Type { Null, Bool, Number, String, Array, Object };
  Type                         type = Null;
  bool                         b    = false;
  double                       num  = 0;
  bool                         is_int = false;
  std::string                  str;
  std::vector<Json>            arr;
  std::map<std::string, Json>  obj;

  bool is_object() const { return type == Object; }
  bool is_array() const { return type == Array; }
  bool is_number() const { return type == Number; }
  bool is_string() const { return type == String; }
  bool has(const std::string& k) const { return type == Object && obj.count(k) > 0; }
  size_t size() const { return type == Object ? obj.size() : (type == Array ? arr.size() : 0); }

  const Json& at(const std::string& k) const {
    auto it = obj.find(k);
    if(type != Object || it == obj.end()) throw std::runtime_error("The required configuration value is missing: " + k);
    return it->second;
  }
  double as_double() const {
    if(type != Number) throw std::runtime_error("configuration value is not a number");
    return num;
  }
  long long as_int() const {
    if(type != Number) throw std::runtime_error("configuration value is not a number");
    return static_cast<long long>(num);
  }
  bool as_bool() const {
    if(type != Bool) throw std::runtime_error("configuration value is not a boolean");
    return b;
  }
  const std::string& as_string() const {
    if(type != String) throw std::runtime_error("configuration value is not a string");
    return str;
  }
  std::vector<double> as_doubles() const {
    if(type != Array) throw std::runtime_error("configuration value is not an array");
    std::vector<double> v;
    for(const Json& e : arr) v.push_back(e.as_double());
    return v;
  }
  // optional accessors in the style of config::opt_config_value (src/common/configuration.h:15-68)
  double opt(const std::string& k, double dflt) const { return has(k) ? obj.at(k).as_double() : dflt; }
  long long opt_int(const std::string& k, long long dflt) const { return has(k) ? obj.at(k).as_int() : dflt; }
  bool opt_bool(const std::string& k, bool dflt) const { return has(k) ? obj.at(k).as_bool() : dflt; }
  std::string opt_str(const std::string& k, const std::string& dflt) const { return has(k) ? obj.at(k).as_string() : dflt; }

  static Json parse(const std::string& text) {
    size_t pos = 0;
    Json   v   = parse_value(text, pos);
    skip_ws(text, pos);
    if(pos != text.size()) throw std::runtime_error("JSON: trailing characters");
    return v;
  }
  static Json parse_file(const std::string& path) {
    std::ifstream f(path);
    if(!f) throw std::runtime_error("Unable to open configuration file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
  }

 private:
  static void skip_ws(const std::string& s, size_t& p) {
    while(p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) ++p;
  }
  static std::string parse_string(const std::string& s, size_t& p) {
    if(s[p] != '"') throw std::runtime_error("JSON: expected string");
    ++p;
    std::string out;
    while(p < s.size() && s[p] != '"') {
      if(s[p] == '\\') {
        ++p;
        if(p >= s.size()) break;
        switch(s[p]) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            unsigned code = static_cast<unsigned>(std::strtoul(s.substr(p + 1, 4).c_str(), nullptr, 16));
            p += 4;
            if(code < 0x80) out += static_cast<char>(code);
            else out += '?';
            break;
          }
          default: out += s[p];
        }
        ++p;
      } else {
        out += s[p++];
      }
    }
    if(p >= s.size()) throw std::runtime_error("JSON: unterminated string");
    ++p;
    return out;
  }
  static Json parse_value(const std::string& s, size_t& p) {
    skip_ws(s, p);
    if(p >= s.size()) throw std::runtime_error("JSON: unexpected end");
    Json v;
    const char c = s[p];
    if(c == '{') {
      v.type = Object;
      ++p;
      skip_ws(s, p);
      if(p < s.size() && s[p] == '}') { ++p; return v; }
      while(true) {
        skip_ws(s, p);
        std::string k = parse_string(s, p);
        skip_ws(s, p);
        if(p >= s.size() || s[p] != ':') throw std::runtime_error("JSON: expected ':'");
        ++p;
        v.obj[k] = parse_value(s, p);
        skip_ws(s, p);
        if(p < s.size() && s[p] == ',') { ++p; continue; }
        if(p < s.size() && s[p] == '}') { ++p; break; }
        throw std::runtime_error("JSON: expected ',' or '}'");
      }
    } else if(c == '[') {
      v.type = Array;
      ++p;
      skip_ws(s, p);
      if(p < s.size() && s[p] == ']') { ++p; return v; }
      while(true) {
        v.arr.push_back(parse_value(s, p));
        skip_ws(s, p);
        if(p < s.size() && s[p] == ',') { ++p; continue; }
        if(p < s.size() && s[p] == ']') { ++p; break; }
        throw std::runtime_error("JSON: expected ',' or ']'");
      }
    } else if(c == '"') {
      v.type = String;
      v.str  = parse_string(s, p);
    } else if(s.compare(p, 4, "true") == 0) {
      v.type = Bool; v.b = true; p += 4;
    } else if(s.compare(p, 5, "false") == 0) {
      v.type = Bool; v.b = false; p += 5;
    } else if(s.compare(p, 4, "null") == 0) {
      p += 4;
    } else {
      const char* begin = s.c_str() + p;
      char*       end   = nullptr;
      v.num             = std::strtod(begin, &end); // correctly rounded, like nlohmann's number parser
      if(end == begin) throw std::runtime_error("JSON: bad token");
      v.type   = Number;
      v.is_int = true;
      for(const char* q = begin; q < end; ++q)
        if(*q == '.' || *q == 'e' || *q == 'E') v.is_int = false;
      p += static_cast<size_t>(end - begin);
    }
    return v;
  }
};

} // namespace lbmhost
