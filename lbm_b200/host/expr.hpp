// expr.hpp -- math expressions of boundary values ("value": "cos(pi*x)") for the host mirror.
//
// The reference evaluates such strings with exprtk at the centre of every boundary cell (MathExpression<NDIM>,
// /root/reference/src/common/configuration.h + external exprtk; call site src/lbm/bnd/bnd_dirichlet.h:268-281) with the variables
// x, y, z and the constant pi.  exprtk is a third-party dependency that is not restated here; this is a small recursive-descent
// evaluator for the grammar the reference's configurations use:
//     expr   := term (('+' | '-') term)*
//     term   := unary (('*' | '/') unary)*
//     unary  := ('-' | '+') unary | power
//     power  := atom ('^' unary)?                       (right associative, binds tighter than unary minus on its left)
//     atom   := number | 'pi' | 'x' | 'y' | 'z' | func '(' expr ')' | '(' expr ')'
//     func   := sin cos tan sinh cosh tanh asin acos atan exp log sqrt abs
// Arithmetic: IEEE double, libm functions, an integer power as repeated multiplication (what exprtk's optimiser emits for
// `a^2`).  tests/test_host_expr.py compares the values with the ones the reference wrote into m_vars for
// test/poisson/poisson2D_helmholtz.json.
#pragma once
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace lbmhost {

class Expression {
 public:
  explicit Expression(std::string text) : m_text(std::move(text)) {}

  double eval(const double* xyz, int ndim) const {
    Parser p{m_text, 0, xyz, ndim};
    const double v = p.expr();
    p.skip();
    if(p.pos != m_text.size()) throw std::runtime_error("Invalid math expression: unexpected '" + m_text.substr(p.pos) + "' in \"" + m_text + "\"");
    return v;
  }

 private:
  std::string m_text;

  struct Parser {
    const std::string& s;
    size_t             pos;
    const double*      xyz;
    int                ndim;

    void skip() {
      while(pos < s.size() && (s[pos] == ' ' || s[pos] == '\t')) ++pos;
    }
    bool take(char c) {
      skip();
      if(pos < s.size() && s[pos] == c) { ++pos; return true; }
      return false;
    }
    double expr() {
      double v = term();
      for(;;) {
        if(take('+')) v = v + term();
        else if(take('-')) v = v - term();
        else return v;
      }
    }
    double term() {
      double v = unary();
      for(;;) {
        if(take('*')) v = v * unary();
        else if(take('/')) v = v / unary();
        else return v;
      }
    }
    double unary() {
      if(take('-')) return -unary();
      if(take('+')) return unary();
      return power();
    }
    double power() {
      const double base = atom();
      if(!take('^')) return base;
      const double e = unary();
      if(e == std::floor(e) && std::abs(e) <= 16) { // integer exponent: repeated multiplication
        double r = 1.0;
        for(int i = 0; i < static_cast<int>(std::abs(e)); ++i) r = r * base;
        return e < 0 ? 1.0 / r : r;
      }
      return std::pow(base, e);
    }
    double atom() {
      skip();
      if(pos >= s.size()) throw std::runtime_error("Invalid math expression: unexpected end of \"" + s + "\"");
      if(take('(')) {
        const double v = expr();
        if(!take(')')) throw std::runtime_error("Invalid math expression: missing ')' in \"" + s + "\"");
        return v;
      }
      const char c = s[pos];
      if((c >= '0' && c <= '9') || c == '.') {
        char*        end = nullptr;
        const double v   = std::strtod(s.c_str() + pos, &end);
        pos              = static_cast<size_t>(end - s.c_str());
        return v;
      }
      size_t b = pos;
      while(pos < s.size() && ((s[pos] >= 'a' && s[pos] <= 'z') || (s[pos] >= 'A' && s[pos] <= 'Z') || s[pos] == '_' || (pos > b && s[pos] >= '0' && s[pos] <= '9'))) ++pos;
      const std::string name = s.substr(b, pos - b);
      if(name.empty()) throw std::runtime_error(std::string("Invalid math expression: unexpected '") + c + "' in \"" + s + "\"");
      if(name == "pi") return 3.14159265358979323846264338327950288419716939937510;
      if(name == "x") return xyz[0];
      if(name == "y") return ndim > 1 ? xyz[1] : 0.0;
      if(name == "z") return ndim > 2 ? xyz[2] : 0.0;
      if(!take('(')) throw std::runtime_error("Invalid math expression: unknown symbol '" + name + "' in \"" + s + "\"");
      const double a = expr();
      if(!take(')')) throw std::runtime_error("Invalid math expression: missing ')' in \"" + s + "\"");
      if(name == "sin") return std::sin(a);
      if(name == "cos") return std::cos(a);
      if(name == "tan") return std::tan(a);
      if(name == "sinh") return std::sinh(a);
      if(name == "cosh") return std::cosh(a);
      if(name == "tanh") return std::tanh(a);
      if(name == "asin") return std::asin(a);
      if(name == "acos") return std::acos(a);
      if(name == "atan") return std::atan(a);
      if(name == "exp") return std::exp(a);
      if(name == "log") return std::log(a);
      if(name == "sqrt") return std::sqrt(a);
      if(name == "abs") return std::abs(a);
      throw std::runtime_error("Invalid math expression: unknown function '" + name + "' in \"" + s + "\"");
    }
  };
};

} // namespace lbmhost
