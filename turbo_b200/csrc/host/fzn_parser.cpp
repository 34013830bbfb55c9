// fzn_parser.cpp — hand-written recursive-descent FlatZinc parser (see fzn_parser.hpp).
#include "fzn_parser.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>

namespace fzn {
namespace {

enum Tok { T_END, T_IDENT, T_INT, T_FLOAT, T_STRING, T_SYM };

struct Lexer {
  const char* p;
  const char* end;
  int line = 1;
  Tok tok = T_END;
  std::string text;     // identifier / string / symbol
  int64_t ival = 0;

  Lexer(const char* t, size_t n) : p(t), end(t + n) { next(); }

  [[noreturn]] void fail(const std::string& m) const { throw ParseError{m, line}; }

  void skip_ws() {
    for (;;) {
      while (p < end && isspace((unsigned char)*p)) { if (*p == '\n') ++line; ++p; }
      if (p < end && *p == '%') { while (p < end && *p != '\n') ++p; continue; }
      break;
    }
  }

  void next() {
    skip_ws();
    text.clear();
    if (p >= end) { tok = T_END; return; }
    char c = *p;
    if (isalpha((unsigned char)c) || c == '_') {
      const char* s = p;
      while (p < end && (isalnum((unsigned char)*p) || *p == '_')) ++p;
      text.assign(s, p);
      tok = T_IDENT;
      return;
    }
    if (isdigit((unsigned char)c) || ((c == '-' || c == '+') && p + 1 < end && isdigit((unsigned char)p[1]))) {
      const char* s = p;
      ++p;
      while (p < end && isdigit((unsigned char)*p)) ++p;
      bool is_float = false;
      if (p + 1 < end && *p == '.' && isdigit((unsigned char)p[1])) {      // 1.5 but not 1..5
        is_float = true; ++p;
        while (p < end && isdigit((unsigned char)*p)) ++p;
      }
      if (p < end && (*p == 'e' || *p == 'E') && is_float) {
        ++p; if (p < end && (*p == '-' || *p == '+')) ++p;
        while (p < end && isdigit((unsigned char)*p)) ++p;
      }
      text.assign(s, p);
      if (is_float) { tok = T_FLOAT; return; }
      tok = T_INT;
      errno = 0;
      ival = strtoll(text.c_str(), nullptr, 10);
      return;
    }
    if (c == '"') {
      ++p;
      const char* s = p;
      while (p < end && *p != '"') { if (*p == '\n') ++line; ++p; }
      text.assign(s, p);
      if (p < end) ++p;
      tok = T_STRING;
      return;
    }
    tok = T_SYM;
    if (c == ':' && p + 1 < end && p[1] == ':') { text = "::"; p += 2; return; }
    if (c == '.' && p + 1 < end && p[1] == '.') { text = ".."; p += 2; return; }
    text.assign(1, c);
    ++p;
  }

  bool is_sym(const char* s) const { return tok == T_SYM && text == s; }
  bool is_ident(const char* s) const { return tok == T_IDENT && text == s; }
  void expect_sym(const char* s) { if (!is_sym(s)) fail(std::string("expected '") + s + "' but found '" + text + "'"); next(); }
  bool accept_sym(const char* s) { if (is_sym(s)) { next(); return true; } return false; }
  std::string expect_ident() { if (tok != T_IDENT) fail("expected an identifier but found '" + text + "'"); std::string t = text; next(); return t; }
  int64_t expect_int() { if (tok != T_INT) fail("expected an integer but found '" + text + "'"); int64_t v = ival; next(); return v; }
};

struct TypeSpec {
  bool is_bool = false, is_float = false, is_set = false;
  bool has_bounds = false;
  int64_t lb = 0, ub = 0;
  std::vector<int64_t> values;   // set-literal domain {a,b,c}
};

struct Parser {
  Lexer lx;
  std::unique_ptr<Model> m;

  Parser(const char* t, size_t n) : lx(t, n), m(new Model()) {}

  static void normalize_set(std::vector<int64_t>& vals, Expr& e) {
    std::sort(vals.begin(), vals.end());
    vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
    e.kind = Expr::SET;
    e.ranges.clear();
    for (size_t i = 0; i < vals.size();) {
      size_t j = i;
      while (j + 1 < vals.size() && vals[j + 1] == vals[j] + 1) ++j;
      e.ranges.push_back({vals[i], vals[j]});
      i = j + 1;
    }
  }

  TypeSpec parse_type() {
    TypeSpec t;
    if (lx.is_ident("set")) {
      lx.next();
      if (!lx.is_ident("of")) lx.fail("expected 'of' after 'set'");
      lx.next();
      t = parse_type();
      t.is_set = true;
      return t;
    }
    if (lx.is_ident("int")) { lx.next(); return t; }
    if (lx.is_ident("bool")) { lx.next(); t.is_bool = true; t.has_bounds = true; t.lb = 0; t.ub = 1; return t; }
    if (lx.is_ident("float")) { lx.next(); t.is_float = true; return t; }
    if (lx.tok == T_INT) {
      t.lb = lx.expect_int();
      lx.expect_sym("..");
      t.ub = lx.expect_int();
      t.has_bounds = true;
      return t;
    }
    if (lx.tok == T_FLOAT) {
      lx.next(); lx.expect_sym(".."); lx.next();
      t.is_float = true;
      return t;
    }
    if (lx.is_sym("{")) {
      lx.next();
      while (!lx.is_sym("}")) {
        t.values.push_back(lx.expect_int());
        if (!lx.accept_sym(",")) break;
      }
      lx.expect_sym("}");
      std::sort(t.values.begin(), t.values.end());
      t.values.erase(std::unique(t.values.begin(), t.values.end()), t.values.end());
      t.has_bounds = true;
      if (t.values.empty()) { t.lb = 1; t.ub = 0; }
      else { t.lb = t.values.front(); t.ub = t.values.back(); }
      return t;
    }
    lx.fail("unsupported type '" + lx.text + "'");
  }

  // Generic expression; identifiers are resolved against declared names when possible.
  Expr parse_expr() {
    Expr e;
    if (lx.tok == T_INT) {
      int64_t v = lx.expect_int();
      if (lx.is_sym("..")) {
        lx.next();
        int64_t w = lx.expect_int();
        e.kind = Expr::SET;
        if (v <= w) e.ranges.push_back({v, w});
        return e;
      }
      e.kind = Expr::INT; e.value = v;
      return e;
    }
    if (lx.tok == T_FLOAT) lx.fail("floating-point values are not supported");
    if (lx.tok == T_STRING) { e.kind = Expr::STRING; e.name = lx.text; lx.next(); return e; }
    if (lx.is_sym("[")) {
      lx.next();
      e.kind = Expr::ARRAY;
      while (!lx.is_sym("]")) {
        e.elems.push_back(parse_expr());
        if (!lx.accept_sym(",")) break;
      }
      lx.expect_sym("]");
      return e;
    }
    if (lx.is_sym("{")) {
      lx.next();
      std::vector<int64_t> vals;
      while (!lx.is_sym("}")) {
        vals.push_back(lx.expect_int());
        if (!lx.accept_sym(",")) break;
      }
      lx.expect_sym("}");
      normalize_set(vals, e);
      return e;
    }
    if (lx.tok == T_IDENT) {
      std::string id = lx.text;
      lx.next();
      if (id == "true" || id == "false") { e.kind = Expr::BOOL; e.value = id == "true"; return e; }
      if (lx.is_sym("(")) {
        lx.next();
        e.kind = Expr::CALL; e.name = id;
        while (!lx.is_sym(")")) {
          e.elems.push_back(parse_expr());
          if (!lx.accept_sym(",")) break;
        }
        lx.expect_sym(")");
        return e;
      }
      if (lx.is_sym("[")) {
        lx.next();
        int64_t idx = lx.expect_int();
        lx.expect_sym("]");
        auto it = m->names.find(id);
        if (it == m->names.end() || it->second.kind != Expr::ARRAY) lx.fail("'" + id + "' is not an array");
        if (idx < 1 || idx > (int64_t)it->second.elems.size()) lx.fail("index out of bounds in '" + id + "'");
        return it->second.elems[(size_t)idx - 1];
      }
      auto vi = m->var_index.find(id);
      if (vi != m->var_index.end()) { e.kind = Expr::VAR; e.var = vi->second; return e; }
      auto ni = m->names.find(id);
      if (ni != m->names.end()) return ni->second;
      e.kind = Expr::IDENT; e.name = id;
      return e;
    }
    lx.fail("unexpected token '" + lx.text + "'");
  }

  std::vector<Expr> parse_annotations() {
    std::vector<Expr> anns;
    while (lx.is_sym("::")) {
      lx.next();
      anns.push_back(parse_expr());
    }
    return anns;
  }

  static const Expr* find_ann(const std::vector<Expr>& anns, const char* name) {
    for (const Expr& a : anns)
      if ((a.kind == Expr::CALL || a.kind == Expr::IDENT) && a.name == name) return &a;
    return nullptr;
  }

  int declare_var(const std::string& name, const TypeSpec& t, const std::vector<Expr>& anns) {
    if (t.is_float) lx.fail("float variables are not supported ('" + name + "')");
    if (t.is_set) lx.fail("arrays of set variables are not supported ('" + name + "')");
    Var v;
    v.name = name;
    v.is_bool = t.is_bool;
    if (t.has_bounds) {
      v.has_lb = v.has_ub = true; v.lb = t.lb; v.ub = t.ub;
      if (!t.values.empty()) {
        size_t k = 0;
        for (int64_t x = t.lb; x <= t.ub; ++x) {
          while (k < t.values.size() && t.values[k] < x) ++k;
          if (k >= t.values.size() || t.values[k] != x) v.holes.push_back(x);
        }
      }
    }
    v.output = find_ann(anns, "output_var") != nullptr;
    v.introduced = find_ann(anns, "var_is_introduced") != nullptr;
    int idx = (int)m->vars.size();
    m->vars.push_back(v);
    m->var_index[name] = idx;
    if (v.output) m->output_order.push_back("v:" + name);
    return idx;
  }

  // `var set of <universe>: name`: one Boolean membership variable per universe value, bound to `name` as a SETVAR
  void declare_set_var(const std::string& name, const TypeSpec& t, const std::vector<Expr>& anns) {
    if (!t.has_bounds) lx.fail("set variable '" + name + "' needs a finite universe");
    if (find_ann(anns, "output_var")) lx.fail("output of set variables is not supported ('" + name + "')");
    std::vector<int64_t> universe = t.values;
    if (universe.empty()) for (int64_t v = t.lb; v <= t.ub; ++v) universe.push_back(v);
    if (universe.size() > 4096) lx.fail("the universe of set variable '" + name + "' is too large");
    Expr s;
    normalize_set(universe, s);
    s.kind = Expr::SETVAR;
    TypeSpec b;
    b.is_bool = true; b.has_bounds = true; b.lb = 0; b.ub = 1;
    std::vector<Expr> intro;
    Expr tag; tag.kind = Expr::IDENT; tag.name = "var_is_introduced";
    intro.push_back(tag);
    for (int64_t v : universe) {
      Expr e; e.kind = Expr::VAR;
      e.var = declare_var(name + "{" + std::to_string(v) + "}", b, intro);
      s.elems.push_back(e);
    }
    m->names[name] = s;
  }

  void parse_var_decl() {            // after 'var'
    TypeSpec t = parse_type();
    lx.expect_sym(":");
    std::string name = lx.expect_ident();
    std::vector<Expr> anns = parse_annotations();
    if (t.is_set) {
      declare_set_var(name, t, anns);
      if (lx.is_sym("=")) lx.fail("initialised set variables are not supported ('" + name + "')");
      lx.expect_sym(";");
      return;
    }
    int idx = declare_var(name, t, anns);
    if (lx.accept_sym("=")) {
      Expr e = parse_expr();
      if (e.kind == Expr::VAR) m->vars[idx].alias_var = e.var;
      else if (e.kind == Expr::INT || e.kind == Expr::BOOL) { m->vars[idx].has_alias_const = true; m->vars[idx].alias_const = e.value; }
      else lx.fail("unsupported initialiser for variable '" + name + "'");
    }
    lx.expect_sym(";");
  }

  void parse_array_decl() {          // after 'array'
    lx.expect_sym("[");
    int64_t lo = lx.expect_int();
    lx.expect_sym("..");
    int64_t hi = lx.expect_int();
    lx.expect_sym("]");
    if (!lx.is_ident("of")) lx.fail("expected 'of'");
    lx.next();
    bool is_var = false;
    if (lx.is_ident("var")) { is_var = true; lx.next(); }
    TypeSpec t = parse_type();
    lx.expect_sym(":");
    std::string name = lx.expect_ident();
    std::vector<Expr> anns = parse_annotations();
    Expr value;
    value.kind = Expr::ARRAY;
    if (lx.accept_sym("=")) {
      value = parse_expr();
      if (value.kind != Expr::ARRAY) lx.fail("array '" + name + "' must be initialised with an array literal");
    } else if (is_var) {
      // an array of fresh variables (legal FlatZinc, rarely emitted)
      for (int64_t i = lo; i <= hi; ++i) {
        Expr e; e.kind = Expr::VAR;
        e.var = declare_var(name + "[" + std::to_string(i) + "]", t, {});
        value.elems.push_back(e);
      }
    }
    lx.expect_sym(";");
    if (t.is_float) return;           // float parameter arrays are parsed and ignored
    if ((int64_t)value.elems.size() != std::max<int64_t>(0, hi - lo + 1)) lx.fail("array '" + name + "' has the wrong number of elements");
    if (is_var) {
      for (const Expr& e : value.elems)
        if (e.kind != Expr::VAR && e.kind != Expr::INT && e.kind != Expr::BOOL) lx.fail("unsupported element in variable array '" + name + "'");
      if (const Expr* oa = find_ann(anns, "output_array")) {
        OutputArray out;
        out.name = name; out.elems = value.elems; out.is_bool = t.is_bool;
        if (oa->elems.size() == 1 && oa->elems[0].kind == Expr::ARRAY)
          for (const Expr& d : oa->elems[0].elems) {
            if (d.kind == Expr::SET && d.ranges.size() == 1) out.dims.push_back(d.ranges[0]);
            else if (d.kind == Expr::SET && d.ranges.empty()) out.dims.push_back({1, 0});
            else lx.fail("unsupported output_array dimension in '" + name + "'");
          }
        m->output_arrays.push_back(out);
        m->output_order.push_back("a:" + name);
      }
    }
    m->names[name] = value;
  }

  void parse_param_decl() {          // scalar parameter: type ':' ident '=' expr ';'
    TypeSpec t = parse_type();
    lx.expect_sym(":");
    std::string name = lx.expect_ident();
    parse_annotations();
    lx.expect_sym("=");
    if (t.is_float) { while (!lx.is_sym(";") && lx.tok != T_END) lx.next(); lx.expect_sym(";"); return; }
    Expr e = parse_expr();
    lx.expect_sym(";");
    m->names[name] = e;
  }

  void collect_search(const Expr& a) {
    if (a.kind != Expr::CALL) return;
    if (a.name == "seq_search") {
      if (a.elems.size() == 1 && a.elems[0].kind == Expr::ARRAY)
        for (const Expr& s : a.elems[0].elems) collect_search(s);
      return;
    }
    if (a.name == "int_search" || a.name == "bool_search") {
      if (a.elems.size() < 3) lx.fail("malformed search annotation");
      SearchAnn s;
      const Expr& vs = a.elems[0];
      if (vs.kind == Expr::ARRAY) s.vars = vs.elems;
      else if (vs.kind == Expr::VAR) s.vars.push_back(vs);
      else lx.fail("the first argument of a search annotation must be an array of variables");
      s.var_sel = a.elems[1].name;
      s.val_sel = a.elems[2].name;
      m->search.push_back(s);
    }
    // other annotations (set_search, float_search, restart_*, ...) are ignored
  }

  void parse_solve() {               // after 'solve'
    std::vector<Expr> anns = parse_annotations();
    for (const Expr& a : anns) collect_search(a);
    if (lx.is_ident("satisfy")) { lx.next(); m->solve = Model::SATISFY; }
    else if (lx.is_ident("minimize")) { lx.next(); m->solve = Model::MINIMIZE; m->objective = parse_expr(); }
    else if (lx.is_ident("maximize")) { lx.next(); m->solve = Model::MAXIMIZE; m->objective = parse_expr(); }
    else lx.fail("expected satisfy, minimize or maximize");
    lx.expect_sym(";");
  }

  void parse_constraint() {          // after 'constraint'
    Constraint c;
    c.line = lx.line;
    Expr e = parse_expr();
    if (e.kind == Expr::BOOL) {       // `constraint true;` / `constraint false;`
      c.name = "bool_eq";
      Expr t; t.kind = Expr::BOOL; t.value = 1;
      c.args = {e, t};
    } else if (e.kind == Expr::CALL) {
      c.name = e.name;
      c.args = std::move(e.elems);
    } else lx.fail("expected a predicate call after 'constraint'");
    parse_annotations();
    lx.expect_sym(";");
    m->constraints.push_back(std::move(c));
  }

  std::unique_ptr<Model> run() {
    bool seen_solve = false;
    while (lx.tok != T_END) {
      if (lx.is_ident("predicate")) { while (!lx.is_sym(";") && lx.tok != T_END) lx.next(); lx.expect_sym(";"); }
      else if (lx.is_ident("var")) { lx.next(); parse_var_decl(); }
      else if (lx.is_ident("array")) { lx.next(); parse_array_decl(); }
      else if (lx.is_ident("constraint")) { lx.next(); parse_constraint(); }
      else if (lx.is_ident("solve")) { lx.next(); parse_solve(); seen_solve = true; }
      else if (lx.is_ident("int") || lx.is_ident("bool") || lx.is_ident("float") || lx.is_ident("set") || lx.tok == T_INT || lx.is_sym("{")) parse_param_decl();
      else lx.fail("unexpected token '" + lx.text + "' at the start of an item");
    }
    if (!seen_solve) lx.fail("missing solve item");
    return std::move(m);
  }
};

}  // namespace

std::unique_ptr<Model> parse(const char* text, size_t len) {
  Parser p(text, len);
  return p.run();
}

}  // namespace fzn
