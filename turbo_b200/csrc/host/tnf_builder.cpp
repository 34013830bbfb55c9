// tnf_builder.cpp — ternarisation of a FlatZinc model (see tnf_builder.hpp) and the FlatZinc-level
// solution checker required by north_star ("every reported solution re-checked against the FlatZinc
// constraints").
#include "tnf_builder.hpp"

#include <algorithm>
#include <stdexcept>

using fzn::Expr;

namespace {

constexpr int64_t NINF = TB_NEG_INF, PINF = TB_POS_INF;

int64_t sat(int64_t v) { return v <= NINF ? NINF : (v >= PINF ? PINF : v); }
// product of two bounds that may be infinite (0 * oo = 0: the other factor is exactly 0)
int64_t sat_mul(int64_t a, int64_t b) {
  if (a == 0 || b == 0) return 0;
  const bool neg = (a < 0) != (b < 0);
  if (a == NINF || a == PINF || b == NINF || b == PINF) return neg ? NINF : PINF;
  const __int128 p = (__int128)a * (__int128)b;
  return p <= (__int128)NINF ? NINF : (p >= (__int128)PINF ? PINF : (int64_t)p);
}
bool is_inf(int64_t v) { return v == NINF || v == PINF; }
int64_t fdiv(int64_t a, int64_t b) { int64_t q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }
int64_t cdiv(int64_t a, int64_t b) { int64_t q = a / b, r = a % b; return (r != 0 && ((r < 0) == (b < 0))) ? q + 1 : q; }

struct Term { int64_t c; int v; };

struct Builder {
  tb_model& out;
  const fzn::Model& m;
  std::vector<int64_t> lb, ub;
  std::vector<tb_prop> props;
  std::map<int64_t, int> consts;
  bool failed = false;

  Builder(tb_model& o, const fzn::Model& mm) : out(o), m(mm) {}

  int fresh(int64_t l, int64_t u) {
    lb.push_back(sat(l)); ub.push_back(sat(u));
    return (int)lb.size() - 1;
  }
  int cst(int64_t k) {
    if (k <= NINF || k >= PINF) throw std::runtime_error("integer constant out of the 32-bit range: " + std::to_string(k));
    auto it = consts.find(k);
    if (it != consts.end()) return it->second;
    int v = fresh(k, k);
    consts[k] = v;
    return v;
  }
  int fresh_bool() { return fresh(0, 1); }
  void prop(int op, int x, int y, int z) { props.push_back(tb_prop{op, x, y, z}); }
  bool is_const(int v) const { return lb[v] == ub[v] && !is_inf(lb[v]); }
  void restrict_lb(int v, int64_t l) { l = sat(l); if (l > lb[v]) lb[v] = l; if (lb[v] > ub[v]) failed = true; }
  void restrict_ub(int v, int64_t u) { u = sat(u); if (u < ub[v]) ub[v] = u; if (lb[v] > ub[v]) failed = true; }

  // ---- interval arithmetic for the initial domains of auxiliary variables -------------------------
  static int64_t add_lo(int64_t a, int64_t b) { return (a == NINF || b == NINF) ? NINF : sat(a + b); }
  static int64_t add_hi(int64_t a, int64_t b) { return (a == PINF || b == PINF) ? PINF : sat(a + b); }
  void mul_bounds(int64_t c, int v, int64_t& l, int64_t& u) const {
    if (c == 0) { l = u = 0; return; }
    int64_t a = is_inf(lb[v]) ? lb[v] : sat(c * lb[v]), b = is_inf(ub[v]) ? ub[v] : sat(c * ub[v]);
    if (c > 0) { l = a; u = b; }
    else { l = is_inf(ub[v]) ? NINF : sat(c * ub[v]); u = is_inf(lb[v]) ? PINF : sat(c * lb[v]); }
  }

  // ---- operands ------------------------------------------------------------------------------------------
  int operand(const Expr& e) {
    switch (e.kind) {
      case Expr::INT: case Expr::BOOL: return cst(e.value);
      case Expr::VAR: return out.var_of_model[(size_t)e.var];
      case Expr::CALL: {          // a predicate used as a term: reify it (test_data/bug1.fzn)
        int r = fresh_bool();
        std::vector<Expr> args = e.elems;
        Expr re; re.kind = Expr::VAR; re.var = -1 - r;      // negative: already a TNF variable
        args.push_back(re);
        post(e.name + "_reif", args);
        return r;
      }
      default: throw std::runtime_error("unsupported term in a constraint");
    }
  }
  int tnf_var(const Expr& e) {
    if (e.kind == Expr::VAR && e.var < 0) return -1 - e.var;
    return operand(e);
  }
  std::vector<int> operands(const Expr& e) {
    if (e.kind != Expr::ARRAY) throw std::runtime_error("expected an array argument");
    std::vector<int> r;
    for (const Expr& x : e.elems) r.push_back(tnf_var(x));
    return r;
  }
  static std::vector<int64_t> ints(const Expr& e) {
    if (e.kind != Expr::ARRAY) throw std::runtime_error("expected an array of integers");
    std::vector<int64_t> r;
    for (const Expr& x : e.elems) {
      if (x.kind != Expr::INT && x.kind != Expr::BOOL) throw std::runtime_error("expected integer literals");
      r.push_back(x.value);
    }
    return r;
  }
  static int64_t int_of(const Expr& e) {
    if (e.kind != Expr::INT && e.kind != Expr::BOOL) throw std::runtime_error("expected an integer literal");
    return e.value;
  }

  // not(r) as a variable: r + nr = 1
  int negation(int r) {
    if (is_const(r)) return cst(1 - lb[r]);
    int nr = fresh_bool();
    prop(TB_OP_ADD, cst(1), r, nr);
    return nr;
  }

  // ---- linear expressions ------------------------------------------------------------------------------------
  // c * v as a variable
  int scaled(int64_t c, int v) {
    if (c == 1) return v;
    if (is_const(v)) return cst(c * lb[v]);
    int64_t l, u; mul_bounds(c, v, l, u);
    int mvar = fresh(l, u);
    if (c == -1) prop(TB_OP_ADD, cst(0), mvar, v);
    else prop(TB_OP_MUL, mvar, cst(c), v);
    return mvar;
  }

  // Normalise: fold constants, merge duplicates, positive coefficients first. Returns the constant part.
  int64_t normalise(std::vector<Term>& ts) {
    int64_t k = 0;
    std::vector<Term> r;
    for (const Term& t : ts) {
      if (t.c == 0) continue;
      if (is_const(t.v)) { k += t.c * lb[t.v]; continue; }
      bool merged = false;
      for (Term& q : r) if (q.v == t.v) { q.c += t.c; merged = true; break; }
      if (!merged) r.push_back(t);
    }
    r.erase(std::remove_if(r.begin(), r.end(), [](const Term& t) { return t.c == 0; }), r.end());
    std::stable_partition(r.begin(), r.end(), [](const Term& t) { return t.c > 0; });
    ts.swap(r);
    return k;
  }

  // Builds  sum(ts) = target  (target < 0: a fresh variable is created and returned).
  // A left-nested chain of ADD propagators; a subtraction  acc' = acc - m  is stored as  acc = acc' + m.
  int sum(std::vector<Term> ts, int target = -1) {
    int64_t k = normalise(ts);
    if (k != 0) ts.push_back(Term{1, cst(k)});       // constants are variables
    if (ts.empty()) {
      if (target < 0) return cst(0);
      prop(TB_OP_EQ, cst(1), target, cst(0));
      return target;
    }
    // move a positive term to the front so that the chain starts with a plain variable when possible
    for (size_t i = 0; i < ts.size(); ++i) if (ts[i].c > 0) { std::swap(ts[0], ts[i]); break; }
    if (ts.size() == 1) {
      if (target < 0) return scaled(ts[0].c, ts[0].v);
      if (ts[0].c == 1) prop(TB_OP_EQ, cst(1), ts[0].v, target);
      else if (ts[0].c == -1) prop(TB_OP_ADD, cst(0), target, ts[0].v);
      else prop(TB_OP_MUL, target, cst(ts[0].c), ts[0].v);
      return target;
    }
    int acc = scaled(ts[0].c, ts[0].v);
    int64_t al = lb[acc], au = ub[acc];
    for (size_t i = 1; i < ts.size(); ++i) {
      const bool last = i + 1 == ts.size();
      const bool neg = ts[i].c < 0;
      int mvar = scaled(neg ? -ts[i].c : ts[i].c, ts[i].v);
      int64_t nl, nu;
      if (!neg) { nl = add_lo(al, lb[mvar]); nu = add_hi(au, ub[mvar]); }
      else { nl = (al == NINF || ub[mvar] == PINF) ? NINF : sat(al - ub[mvar]); nu = (au == PINF || lb[mvar] == NINF) ? PINF : sat(au - lb[mvar]); }
      int nacc = (last && target >= 0) ? target : fresh(nl, nu);
      if (!neg) prop(TB_OP_ADD, nacc, acc, mvar);
      else prop(TB_OP_ADD, acc, nacc, mvar);
      acc = nacc; al = nl; au = nu;
    }
    return acc;
  }

  std::vector<Term> terms(const Expr& cs, const Expr& xs) {
    std::vector<int64_t> c = ints(cs);
    std::vector<int> x = operands(xs);
    if (c.size() != x.size()) throw std::runtime_error("linear constraint with mismatched arrays");
    std::vector<Term> ts;
    for (size_t i = 0; i < c.size(); ++i) ts.push_back(Term{c[i], x[i]});
    return ts;
  }

  // r = (sum(ts) <= c); r is a TNF variable (cst(1) for a hard constraint)
  void lin_le(std::vector<Term> ts, int64_t c, int r) {
    int64_t k = normalise(ts);
    c -= k;
    if (ts.empty()) { post_truth(r, 0 <= c); return; }
    if (ts.size() == 1) {
      const Term t = ts[0];
      if (t.c > 0) { int64_t bound = fdiv(c, t.c); if (r == cst(1)) restrict_ub(t.v, bound); else prop(TB_OP_LEQ, r, t.v, cst_clamped(bound)); }
      else { int64_t bound = cdiv(c, t.c); if (r == cst(1)) restrict_lb(t.v, bound); else prop(TB_OP_LEQ, r, cst_clamped(bound), t.v); }
      return;
    }
    if (ts.size() == 2 && ts[0].c == 1 && ts[1].c == -1) {       // x - y <= c
      int y = ts[1].v;
      if (c != 0) { int t = fresh(add_lo(lb[y], c), add_hi(ub[y], c)); prop(TB_OP_ADD, t, y, cst(c)); y = t; }
      prop(TB_OP_LEQ, r, ts[0].v, y);
      return;
    }
    int s = sum(ts);
    prop(TB_OP_LEQ, r, s, cst(c));
  }
  int cst_clamped(int64_t k) { return cst(std::max<int64_t>(NINF + 1, std::min<int64_t>(PINF - 1, k))); }

  // r = (sum(ts) == c)
  void lin_eq(std::vector<Term> ts, int64_t c, int r) {
    int64_t k = normalise(ts);
    c -= k;
    if (ts.empty()) { post_truth(r, c == 0); return; }
    const bool hard = r == cst(1);
    if (ts.size() == 1) {
      const Term t = ts[0];
      if (c % t.c != 0) { post_truth(r, false); return; }
      int64_t val = c / t.c;
      if (hard) { restrict_lb(t.v, val); restrict_ub(t.v, val); }
      else prop(TB_OP_EQ, r, t.v, cst_clamped(val));
      return;
    }
    if (ts.size() == 2 && ts[0].c == 1 && ts[1].c == -1) {       // x - y == c
      int y = ts[1].v;
      if (c != 0) { int t = fresh(add_lo(lb[y], c), add_hi(ub[y], c)); prop(TB_OP_ADD, t, y, cst(c)); y = t; }
      prop(TB_OP_EQ, r, ts[0].v, y);
      return;
    }
    if (hard) { sum(ts, cst(c)); return; }
    int s = sum(ts);
    prop(TB_OP_EQ, r, s, cst(c));
  }

  // force the boolean variable r to a truth value
  void post_truth(int r, bool truth) {
    if (truth) restrict_lb(r, 1); else restrict_ub(r, 0);
  }

  // ---- element ---------------------------------------------------------------------------------------------------
  // (i = k) => (v = a_k) for every k  
  void element(int idx, const std::vector<int>& arr, int v) {
    restrict_lb(idx, 1); restrict_ub(idx, (int64_t)arr.size());
    int64_t lo = PINF, hi = NINF;
    for (int a : arr) { lo = std::min(lo, lb[a]); hi = std::max(hi, ub[a]); }
    if (!arr.empty()) { restrict_lb(v, lo); restrict_ub(v, hi); }
    for (size_t k = 0; k < arr.size(); ++k) {
      int b = fresh_bool(), c = fresh_bool();
      prop(TB_OP_EQ, b, idx, cst((int64_t)k + 1));
      prop(TB_OP_EQ, c, v, arr[k]);
      prop(TB_OP_LEQ, cst(1), b, c);
    }
  }

  // r = (x in S)
  void set_in(int x, const Expr& s, int r) {
    if (s.kind != Expr::SET) throw std::runtime_error("expected a set literal");
    const bool hard = r == cst(1);
    if (s.ranges.empty()) { post_truth(r, false); return; }
    if (hard) {
      restrict_lb(x, s.ranges.front().first); restrict_ub(x, s.ranges.back().second);
      for (size_t i = 0; i + 1 < s.ranges.size(); ++i)
        for (int64_t h = s.ranges[i].second + 1; h < s.ranges[i + 1].first; ++h) prop(TB_OP_EQ, cst(0), x, cst(h));
      return;
    }
    std::vector<int> members;
    for (const auto& rg : s.ranges) {
      int b;
      if (rg.first == rg.second) { b = fresh_bool(); prop(TB_OP_EQ, b, x, cst(rg.first)); }
      else {
        int b1 = fresh_bool(), b2 = fresh_bool();
        prop(TB_OP_LEQ, b1, cst(rg.first), x);
        prop(TB_OP_LEQ, b2, x, cst(rg.second));
        b = fresh_bool();
        prop(TB_OP_MIN, b, b1, b2);
      }
      members.push_back(b);
    }
    chain(TB_OP_MAX, members, r);
  }

  // ---- set variables: one membership Boolean per universe value (fzn_parser.cpp, declare_set_var) -----------------
  static std::vector<int64_t> universe_of(const Expr& s) {
    std::vector<int64_t> u;
    for (const auto& rg : s.ranges) for (int64_t v = rg.first; v <= rg.second; ++v) u.push_back(v);
    return u;
  }
  static bool set_has(const Expr& s, int64_t v) {
    for (const auto& rg : s.ranges) if (v >= rg.first && v <= rg.second) return true;
    return false;
  }

  // r = (x in S), S a set variable: (x = v) => (r = member_v) for every universe value, and r => x in universe
  void set_in_var(int x, const Expr& S, int r) {
    const std::vector<int64_t> u = universe_of(S);
    if (u.empty()) { post_truth(r, false); return; }
    for (size_t k = 0; k < u.size(); ++k) {
      if (u[k] < lb[x] || u[k] > ub[x]) continue;             // x can never take this value
      const int mv = tnf_var(S.elems[k]);
      int e = fresh_bool(), c = fresh_bool();
      prop(TB_OP_EQ, e, x, cst(u[k]));
      prop(TB_OP_EQ, c, r, mv);
      prop(TB_OP_LEQ, cst(1), e, c);
    }
    Expr uni; uni.kind = Expr::SET; uni.ranges = S.ranges;
    int inside = fresh_bool();
    set_in(x, uni, inside);
    prop(TB_OP_LEQ, cst(1), r, inside);
  }

  // S = arr[idx], arr an array of constant sets: member_v = (idx in {k : v in arr[k]}); a set that leaves the universe
  // cannot be selected
  void set_element(int idx, const Expr& arr, const Expr& S) {
    if (arr.kind != Expr::ARRAY || S.kind != Expr::SETVAR) throw std::runtime_error("array_set_element expects an array of constant sets and a set variable");
    const std::vector<int64_t> u = universe_of(S);
    restrict_lb(idx, 1); restrict_ub(idx, (int64_t)arr.elems.size());
    for (size_t k = 0; k < arr.elems.size(); ++k) {
      const Expr& sk = arr.elems[k];
      if (sk.kind != Expr::SET) throw std::runtime_error("array_set_element: variable sets inside the array are not supported");
      bool inside = true;
      for (const auto& rg : sk.ranges) for (int64_t v = rg.first; v <= rg.second; ++v) inside = inside && set_has(S, v);
      if (!inside) prop(TB_OP_EQ, cst(0), idx, cst((int64_t)k + 1));
    }
    for (size_t j = 0; j < u.size(); ++j) {
      std::vector<int64_t> ks;
      for (size_t k = 0; k < arr.elems.size(); ++k) if (set_has(arr.elems[k], u[j])) ks.push_back((int64_t)k + 1);
      Expr K; K.kind = Expr::SET;
      for (size_t i = 0; i < ks.size();) {
        size_t e = i;
        while (e + 1 < ks.size() && ks[e + 1] == ks[e] + 1) ++e;
        K.ranges.push_back({ks[i], ks[e]});
        i = e + 1;
      }
      set_in(idx, K, tnf_var(S.elems[j]));
    }
  }

  // r = op-fold(vs) for MIN (conjunction) / MAX (disjunction) over 0..1 variables
  void chain(int op, const std::vector<int>& vs, int r) {
    if (vs.empty()) { post_truth(r, op == TB_OP_MIN); return; }
    if (vs.size() == 1) { prop(TB_OP_EQ, cst(1), r, vs[0]); return; }
    int acc = vs[0];
    for (size_t i = 1; i < vs.size(); ++i) {
      int n = (i + 1 == vs.size()) ? r : fresh_bool();
      prop(op, n, acc, vs[i]);
      acc = n;
    }
  }

  // ---- constraints ------------------------------------------------------------------------------------------------
  void post(const std::string& name, const std::vector<Expr>& a) {
    auto need = [&](size_t n) { if (a.size() != n) throw std::runtime_error("constraint " + name + " expects " + std::to_string(n) + " arguments"); };
    const int ONE = cst(1), ZERO = cst(0);
    // half reification  r -> C  (the _imp predicates of recent MiniZinc): r <= r' with r' <-> C
    if (name.size() > 4 && name.compare(name.size() - 4, 4, "_imp") == 0 && !a.empty()) {
      const int r = tnf_var(a.back());
      const int full = fresh_bool();
      std::vector<Expr> args(a.begin(), a.end() - 1);
      Expr re; re.kind = Expr::VAR; re.var = -1 - full;      // negative: already a TNF variable
      args.push_back(re);
      post(name.substr(0, name.size() - 4) + "_reif", args);
      prop(TB_OP_LEQ, ONE, r, full);
      return;
    }
    if (name == "int_pow") {          // z = x ^ k for a constant exponent k >= 0: a chain of multiplications
      need(3);
      if (a[1].kind != Expr::INT || a[1].value < 0 || a[1].value > 30) throw std::runtime_error("int_pow needs a constant exponent in 0..30");
      const int x = tnf_var(a[0]), z = tnf_var(a[2]);
      const int64_t k = a[1].value;
      if (k == 0) { prop(TB_OP_EQ, ONE, z, ONE); return; }
      if (k == 1) { prop(TB_OP_EQ, ONE, z, x); return; }
      int acc = x;
      for (int64_t i = 2; i <= k; ++i) {
        int64_t l, u; 
        {   // bounds of acc * x from the corner products (saturating)
          const int64_t c[4] = {sat_mul(lb[acc], lb[x]), sat_mul(lb[acc], ub[x]), sat_mul(ub[acc], lb[x]), sat_mul(ub[acc], ub[x])};
          l = *std::min_element(c, c + 4); u = *std::max_element(c, c + 4);
        }
        const int nx = i == k ? z : fresh(l, u);
        prop(TB_OP_MUL, nx, acc, x);
        acc = nx;
      }
      return;
    }
    if (name == "bool_lin_eq" || name == "bool_lin_le") { post(name == "bool_lin_eq" ? "int_lin_eq" : "int_lin_le", a); return; }
    if (name == "array_int_minimum" || name == "array_int_maximum") {
      need(2);
      std::vector<int> vs = operands(a[1]);
      if (vs.empty()) throw std::runtime_error(name + " over an empty array");
      const int op = name == "array_int_minimum" ? TB_OP_MIN : TB_OP_MAX;
      const int m = tnf_var(a[0]);
      if (vs.size() == 1) { prop(TB_OP_EQ, ONE, m, vs[0]); return; }
      int acc = vs[0];
      for (size_t i = 1; i < vs.size(); ++i) {
        int64_t l = op == TB_OP_MIN ? std::min(lb[acc], lb[vs[i]]) : std::max(lb[acc], lb[vs[i]]);
        int64_t u = op == TB_OP_MIN ? std::min(ub[acc], ub[vs[i]]) : std::max(ub[acc], ub[vs[i]]);
        const int nx = i + 1 == vs.size() ? m : fresh(l, u);
        prop(op, nx, acc, vs[i]);
        acc = nx;
      }
      return;
    }
    if (name == "int_lin_le") { need(3); lin_le(terms(a[0], a[1]), int_of(a[2]), ONE); }
    else if (name == "int_lin_le_reif") { need(4); lin_le(terms(a[0], a[1]), int_of(a[2]), tnf_var(a[3])); }
    else if (name == "int_lin_eq") { need(3); lin_eq(terms(a[0], a[1]), int_of(a[2]), ONE); }
    else if (name == "int_lin_eq_reif") { need(4); lin_eq(terms(a[0], a[1]), int_of(a[2]), tnf_var(a[3])); }
    else if (name == "int_lin_ne") { need(3); lin_eq(terms(a[0], a[1]), int_of(a[2]), ZERO); }
    else if (name == "int_lin_ne_reif") { need(4); lin_eq(terms(a[0], a[1]), int_of(a[2]), negation(tnf_var(a[3]))); }
    else if (name == "int_lin_lt" || name == "int_lin_lt_reif") {
      lin_le(terms(a[0], a[1]), int_of(a[2]) - 1, a.size() == 4 ? tnf_var(a[3]) : ONE);
    }
    else if (name == "int_lin_ge" || name == "int_lin_ge_reif" || name == "int_lin_gt" || name == "int_lin_gt_reif") {
      std::vector<Term> ts = terms(a[0], a[1]);
      for (Term& t : ts) t.c = -t.c;
      const bool strict = name.find("_gt") != std::string::npos;
      lin_le(ts, -int_of(a[2]) - (strict ? 1 : 0), a.size() == 4 ? tnf_var(a[3]) : ONE);
    }
    else if (name == "int_eq" || name == "bool_eq" || name == "bool2int") { need(2); lin_eq({{1, tnf_var(a[0])}, {-1, tnf_var(a[1])}}, 0, ONE); }
    else if (name == "int_ne") { need(2); lin_eq({{1, tnf_var(a[0])}, {-1, tnf_var(a[1])}}, 0, ZERO); }
    else if (name == "int_le" || name == "bool_le") { need(2); lin_le({{1, tnf_var(a[0])}, {-1, tnf_var(a[1])}}, 0, ONE); }
    else if (name == "int_lt" || name == "bool_lt") { need(2); lin_le({{1, tnf_var(a[0])}, {-1, tnf_var(a[1])}}, -1, ONE); }
    else if (name == "int_eq_reif" || name == "bool_eq_reif") { need(3); lin_eq({{1, tnf_var(a[0])}, {-1, tnf_var(a[1])}}, 0, tnf_var(a[2])); }
    else if (name == "int_ne_reif" || name == "bool_xor") {
      if (a.size() == 2) { prop(TB_OP_ADD, ONE, tnf_var(a[0]), tnf_var(a[1])); }     // bool_xor(a,b): a + b = 1
      else { need(3); lin_eq({{1, tnf_var(a[0])}, {-1, tnf_var(a[1])}}, 0, negation(tnf_var(a[2]))); }
    }
    else if (name == "int_le_reif" || name == "bool_le_reif") { need(3); lin_le({{1, tnf_var(a[0])}, {-1, tnf_var(a[1])}}, 0, tnf_var(a[2])); }
    else if (name == "int_lt_reif" || name == "bool_lt_reif") { need(3); lin_le({{1, tnf_var(a[0])}, {-1, tnf_var(a[1])}}, -1, tnf_var(a[2])); }
    else if (name == "int_plus") { need(3); prop(TB_OP_ADD, tnf_var(a[2]), tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "int_minus") { need(3); prop(TB_OP_ADD, tnf_var(a[0]), tnf_var(a[2]), tnf_var(a[1])); }
    else if (name == "int_times") { need(3); prop(TB_OP_MUL, tnf_var(a[2]), tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "int_div") { need(3); prop(TB_OP_TDIV, tnf_var(a[2]), tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "int_mod") { need(3); prop(TB_OP_TMOD, tnf_var(a[2]), tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "int_min") { need(3); prop(TB_OP_MIN, tnf_var(a[2]), tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "int_max") { need(3); prop(TB_OP_MAX, tnf_var(a[2]), tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "int_abs") {
      need(2);
      int x = tnf_var(a[0]), y = tnf_var(a[1]);
      int n = scaled(-1, x);
      prop(TB_OP_MAX, y, x, n);
    }
    else if (name == "int_negate") { need(2); prop(TB_OP_ADD, ZERO, tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "bool_and") { need(3); prop(TB_OP_MIN, tnf_var(a[2]), tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "bool_or") { need(3); prop(TB_OP_MAX, tnf_var(a[2]), tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "bool_not") { need(2); prop(TB_OP_ADD, ONE, tnf_var(a[0]), tnf_var(a[1])); }
    else if (name == "array_bool_and") { need(2); chain(TB_OP_MIN, operands(a[0]), tnf_var(a[1])); }
    else if (name == "array_bool_or") { need(2); chain(TB_OP_MAX, operands(a[0]), tnf_var(a[1])); }
    else if (name == "array_bool_xor") {          // odd number of true literals: sum = 2k + 1
      need(1);
      std::vector<Term> ts;
      std::vector<int> vs = operands(a[0]);
      for (int v : vs) ts.push_back(Term{1, v});
      int k = fresh(0, (int64_t)vs.size() / 2);
      ts.push_back(Term{-2, k});
      lin_eq(ts, 1, ONE);
    }
    else if (name == "bool_clause" || name == "bool_clause_reif") {
      // sum(pos) - sum(neg) >= 1 - |neg|   <=>   -sum(pos) + sum(neg) <= |neg| - 1
      std::vector<int> pos = operands(a[0]), neg = operands(a[1]);
      std::vector<Term> ts;
      for (int v : pos) ts.push_back(Term{-1, v});
      for (int v : neg) ts.push_back(Term{1, v});
      lin_le(ts, (int64_t)neg.size() - 1, a.size() == 3 ? tnf_var(a[2]) : ONE);
    }
    else if (name == "array_int_element" || name == "array_bool_element" || name == "array_var_int_element" ||
             name == "array_var_bool_element") {
      need(3);
      element(tnf_var(a[0]), operands(a[1]), tnf_var(a[2]));
    }
    else if (name == "set_in") { need(2); if (a[1].kind == Expr::SETVAR) set_in_var(tnf_var(a[0]), a[1], ONE); else set_in(tnf_var(a[0]), a[1], ONE); }
    else if (name == "set_in_reif") { need(3); if (a[1].kind == Expr::SETVAR) set_in_var(tnf_var(a[0]), a[1], tnf_var(a[2])); else set_in(tnf_var(a[0]), a[1], tnf_var(a[2])); }
    else if (name == "array_set_element") { need(3); set_element(tnf_var(a[0]), a[1], a[2]); }
    else throw std::runtime_error("unsupported FlatZinc constraint '" + name + "'");
  }
};

int var_order_of(const std::string& s) {
  if (s == "input_order") return TB_VAR_INPUT_ORDER;
  if (s == "first_fail") return TB_VAR_FIRST_FAIL;
  if (s == "anti_first_fail") return TB_VAR_ANTI_FIRST_FAIL;
  if (s == "smallest") return TB_VAR_SMALLEST;
  if (s == "largest") return TB_VAR_LARGEST;
  return TB_VAR_FIRST_FAIL;      // occurrence, most_constrained, max_regret, dom_w_deg: not available
}
int val_order_of(const std::string& s) {
  if (s == "indomain_min" || s == "indomain") return TB_VAL_MIN;
  if (s == "indomain_max") return TB_VAL_MAX;
  if (s == "indomain_split") return TB_VAL_SPLIT;
  if (s == "indomain_reverse_split") return TB_VAL_REVERSE_SPLIT;
  return TB_VAL_MIN;             // median / middle / random are not expressible on intervals (barebones :385)
}

}  // namespace

void tb_model::finalize() {
  strategies.clear();
  for (size_t i = 0; i < strat_vars.size(); ++i) {
    tb_strategy s;
    s.var_order = strat_orders[i].first; s.val_order = strat_orders[i].second;
    s.n = (int32_t)strat_vars[i].size();
    s.vars = strat_vars[i].empty() ? nullptr : strat_vars[i].data();
    strategies.push_back(s);
  }
  problem.nvars = (int32_t)lb.size();
  problem.nprops = (int32_t)props.size();
  problem.lb = lb.data(); problem.ub = ub.data();
  problem.props = props.data();
  problem.nstrategies = (int32_t)strategies.size();
  problem.strategies = strategies.data();
  problem.has_eps_strategy = has_eps_strategy ? 1 : 0;
  problem.obj_var = obj_var;
}

void tb_model::push_eps_strategy(int var_order, int val_order) {
  if (has_eps_strategy) { strat_orders[0] = {var_order, val_order}; finalize(); return; }
  strat_vars.insert(strat_vars.begin(), std::vector<int32_t>{});
  strat_orders.insert(strat_orders.begin(), std::make_pair(var_order, val_order));
  has_eps_strategy = true;
  finalize();
}

std::unique_ptr<tb_model> build_tnf(std::unique_ptr<fzn::Model> src) {
  std::unique_ptr<tb_model> out(new tb_model());
  const fzn::Model& m = *src;
  Builder b(*out, m);
  // constants 0, 1, 2 are variables 0, 1, 2 (common_solving.hpp:521)
  for (int k = 0; k < 3; ++k) b.cst(k);
  out->var_of_model.resize(m.vars.size());
  for (size_t i = 0; i < m.vars.size(); ++i) {
    const fzn::Var& v = m.vars[i];
    out->var_of_model[i] = b.fresh(v.has_lb ? v.lb : NINF, v.has_ub ? v.ub : PINF);
    if (v.has_lb && v.has_ub && v.lb > v.ub) b.failed = true;
  }
  for (size_t i = 0; i < m.vars.size(); ++i) {
    const fzn::Var& v = m.vars[i];
    const int x = out->var_of_model[i];
    for (int64_t h : v.holes) b.prop(TB_OP_EQ, b.cst(0), x, b.cst(h));
    if (v.alias_var >= 0) b.prop(TB_OP_EQ, b.cst(1), x, out->var_of_model[(size_t)v.alias_var]);
    if (v.has_alias_const) { b.restrict_lb(x, v.alias_const); b.restrict_ub(x, v.alias_const); }
  }
  for (const fzn::Constraint& c : m.constraints) {
    try { b.post(c.name, c.args); }
    catch (const std::runtime_error& e) { throw std::runtime_error(std::string(e.what()) + " (line " + std::to_string(c.line) + ")"); }
  }
  // objective: always minimise; maximize x adds __MINIMIZE_OBJ = -x (common_solving.hpp:489-510)
  if (m.solve != fzn::Model::SATISFY) {
    int x = b.operand(m.objective);
    out->user_obj_var = x;
    if (m.solve == fzn::Model::MINIMIZE) { out->objective_kind = 0; out->obj_var = x; }
    else { out->objective_kind = 1; out->obj_var = b.scaled(-1, x); }
  }
  // search strategies, then the default first_fail / indomain_min over every variable (:640-650)
  for (const fzn::SearchAnn& s : m.search) {
    std::vector<int32_t> vs;
    for (const Expr& e : s.vars) {
      if (e.kind == Expr::VAR) vs.push_back(out->var_of_model[(size_t)e.var]);
      else if (e.kind == Expr::INT || e.kind == Expr::BOOL) vs.push_back(b.cst(e.value));
    }
    if (vs.empty()) continue;       // an empty list would mean "all variables" in the ABI
    out->strat_vars.push_back(std::move(vs));
    out->strat_orders.push_back({var_order_of(s.var_sel), val_order_of(s.val_sel)});
  }
  out->strat_vars.push_back({});
  out->strat_orders.push_back({TB_VAR_FIRST_FAIL, TB_VAL_MIN});

  out->lb.resize(b.lb.size()); out->ub.resize(b.ub.size());
  for (size_t i = 0; i < b.lb.size(); ++i) { out->lb[i] = (int32_t)b.lb[i]; out->ub[i] = (int32_t)b.ub[i]; }
  out->props = std::move(b.props);
  out->root_failed = b.failed;
  out->parsed_variables = (int)m.vars.size();
  out->parsed_constraints = (int)m.constraints.size();
  out->src = std::move(src);
  out->finalize();
  return out;
}

// ---- FlatZinc-level checker -----------------------------------------------------------------------------------
namespace {

struct Checker {
  const fzn::Model& m;
  const std::vector<int64_t>& val;

  static int64_t tdiv(int64_t a, int64_t b) { return a / b; }

  int64_t ev(const Expr& e) const {
    switch (e.kind) {
      case Expr::INT: case Expr::BOOL: return e.value;
      case Expr::VAR: return val[(size_t)e.var];
      case Expr::CALL: return holds(e.name, e.elems) ? 1 : 0;
      default: throw std::runtime_error("checker: unsupported term");
    }
  }
  std::vector<int64_t> evs(const Expr& e) const {
    std::vector<int64_t> r;
    for (const Expr& x : e.elems) r.push_back(ev(x));
    return r;
  }
  int64_t dot(const Expr& cs, const Expr& xs) const {
    int64_t s = 0;
    for (size_t i = 0; i < cs.elems.size(); ++i) s += cs.elems[i].value * ev(xs.elems[i]);
    return s;
  }
  bool in_set_var(int64_t x, const Expr& S) const {           // membership in a set variable at the point
    size_t j = 0;
    for (const auto& rg : S.ranges) {
      if (x >= rg.first && x <= rg.second) return ev(S.elems[j + (size_t)(x - rg.first)]) != 0;
      j += (size_t)(rg.second - rg.first + 1);
    }
    return false;
  }
  static bool in_set(int64_t x, const Expr& s) {
    for (const auto& r : s.ranges) if (r.first <= x && x <= r.second) return true;
    return false;
  }

  bool holds(const std::string& name, const std::vector<Expr>& a) const {
    const size_t n = a.size();
    if (name.size() > 4 && name.compare(name.size() - 4, 4, "_imp") == 0) {       // r -> C
      std::vector<Expr> base(a.begin(), a.end() - 1);
      return ev(a[n - 1]) == 0 || holds(name.substr(0, name.size() - 4), base);
    }
    if (name == "int_pow") {
      int64_t r = 1;
      for (int64_t i = 0; i < ev(a[1]); ++i) r *= ev(a[0]);
      return ev(a[1]) >= 0 && r == ev(a[2]);
    }
    if (name == "bool_lin_eq") return dot(a[0], a[1]) == a[2].value;
    if (name == "bool_lin_le") return dot(a[0], a[1]) <= a[2].value;
    if (name == "array_int_minimum" || name == "array_int_maximum") {
      std::vector<int64_t> vs = evs(a[1]);
      if (vs.empty()) return false;
      return ev(a[0]) == (name == "array_int_minimum" ? *std::min_element(vs.begin(), vs.end()) : *std::max_element(vs.begin(), vs.end()));
    }
    if (name.size() > 5 && name.compare(name.size() - 5, 5, "_reif") == 0) {
      std::vector<Expr> base(a.begin(), a.end() - 1);
      return holds(name.substr(0, name.size() - 5), base) == (ev(a[n - 1]) != 0);
    }
    if (name == "int_lin_le") return dot(a[0], a[1]) <= a[2].value;
    if (name == "int_lin_lt") return dot(a[0], a[1]) < a[2].value;
    if (name == "int_lin_ge") return dot(a[0], a[1]) >= a[2].value;
    if (name == "int_lin_gt") return dot(a[0], a[1]) > a[2].value;
    if (name == "int_lin_eq") return dot(a[0], a[1]) == a[2].value;
    if (name == "int_lin_ne") return dot(a[0], a[1]) != a[2].value;
    if (name == "int_eq" || name == "bool_eq" || name == "bool2int") return ev(a[0]) == ev(a[1]);
    if (name == "int_ne") return ev(a[0]) != ev(a[1]);
    if (name == "int_le" || name == "bool_le") return ev(a[0]) <= ev(a[1]);
    if (name == "int_lt" || name == "bool_lt") return ev(a[0]) < ev(a[1]);
    if (name == "bool_xor") return n == 2 ? ev(a[0]) != ev(a[1]) : (ev(a[0]) != ev(a[1])) == (ev(a[2]) != 0);
    if (name == "int_plus") return ev(a[0]) + ev(a[1]) == ev(a[2]);
    if (name == "int_minus") return ev(a[0]) - ev(a[1]) == ev(a[2]);
    if (name == "int_times") return ev(a[0]) * ev(a[1]) == ev(a[2]);
    if (name == "int_div") return ev(a[1]) != 0 && tdiv(ev(a[0]), ev(a[1])) == ev(a[2]);
    if (name == "int_mod") return ev(a[1]) != 0 && ev(a[0]) % ev(a[1]) == ev(a[2]);
    if (name == "int_min") return std::min(ev(a[0]), ev(a[1])) == ev(a[2]);
    if (name == "int_max") return std::max(ev(a[0]), ev(a[1])) == ev(a[2]);
    if (name == "int_abs") return std::llabs(ev(a[0])) == ev(a[1]);
    if (name == "int_negate") return -ev(a[0]) == ev(a[1]);
    if (name == "bool_and") return ((ev(a[0]) != 0) && (ev(a[1]) != 0)) == (ev(a[2]) != 0);
    if (name == "bool_or") return ((ev(a[0]) != 0) || (ev(a[1]) != 0)) == (ev(a[2]) != 0);
    if (name == "bool_not") return (ev(a[0]) != 0) != (ev(a[1]) != 0);
    if (name == "array_bool_and") { bool all = true; for (int64_t v : evs(a[0])) all = all && v != 0; return all == (ev(a[1]) != 0); }
    if (name == "array_bool_or") { bool any = false; for (int64_t v : evs(a[0])) any = any || v != 0; return any == (ev(a[1]) != 0); }
    if (name == "array_bool_xor") { int c = 0; for (int64_t v : evs(a[0])) c += v != 0; return c % 2 == 1; }
    if (name == "bool_clause") {
      for (int64_t v : evs(a[0])) if (v != 0) return true;
      for (int64_t v : evs(a[1])) if (v == 0) return true;
      return false;
    }
    if (name == "array_int_element" || name == "array_bool_element" || name == "array_var_int_element" || name == "array_var_bool_element") {
      int64_t i = ev(a[0]);
      if (i < 1 || i > (int64_t)a[1].elems.size()) return false;
      return ev(a[1].elems[(size_t)i - 1]) == ev(a[2]);
    }
    if (name == "set_in") return a[1].kind == Expr::SETVAR ? in_set_var(ev(a[0]), a[1]) : in_set(ev(a[0]), a[1]);
    if (name == "array_set_element") {
      const int64_t i = ev(a[0]);
      if (a[1].kind != Expr::ARRAY || a[2].kind != Expr::SETVAR || i < 1 || i > (int64_t)a[1].elems.size()) return false;
      const Expr& sk = a[1].elems[(size_t)i - 1];
      for (const auto& rg : sk.ranges) for (int64_t v = rg.first; v <= rg.second; ++v) if (!in_set_var(v, a[2])) return false;   // sk is a subset of S
      size_t j = 0;
      for (const auto& rg : a[2].ranges) for (int64_t v = rg.first; v <= rg.second; ++v, ++j) if ((ev(a[2].elems[j]) != 0) != in_set(v, sk)) return false;
      return true;
    }
    throw std::runtime_error("checker: unsupported constraint '" + name + "'");
  }
};

}  // namespace

int check_flatzinc(const fzn::Model& m, const std::vector<int64_t>& value, std::string* first_violation) {
  int bad = 0;
  auto report = [&](const std::string& s) { if (bad == 0 && first_violation) *first_violation = s; ++bad; };
  for (size_t i = 0; i < m.vars.size(); ++i) {
    const fzn::Var& v = m.vars[i];
    const int64_t x = value[i];
    if ((v.has_lb && x < v.lb) || (v.has_ub && x > v.ub) || std::binary_search(v.holes.begin(), v.holes.end(), x))
      report("domain of " + v.name);
    if (v.alias_var >= 0 && value[(size_t)v.alias_var] != x) report("alias of " + v.name);
    if (v.has_alias_const && v.alias_const != x) report("value of " + v.name);
  }
  Checker c{m, value};
  for (const fzn::Constraint& k : m.constraints)
    if (!c.holds(k.name, k.args)) report(k.name + " at line " + std::to_string(k.line));
  return bad;
}
