// tnf_simplify.cpp — TNF simplifier (host side, C++17): SURVEY.md §8(f).1.
//
// Replaces lala-core's Simplifier as it is driven by CP::preprocess_tcn
// (reference include/common_solving.hpp:538-565; the class itself is un-vendored): a loop of
//   root fixpoint -> meet equivalence classes -> algebraic simplification -> elimination of entailed
//   constraints -> ICSE (common sub-expression elimination on x = y op z) -> useless-variable elimination
// until nothing changes, then a renumbering of what is left.  The mapping back (SimplifierStats /
// `print_variable` through the equivalence classes, :526-531) is kept in the model, so solutions of the
// reduced network expand to points of the full one and are printed / re-checked there.
//
// The root fixpoint is NOT computed here: the caller passes the engine's fixpoint (tb_propagate on the
// GPU in the `turbo` driver), so interval narrowing has exactly one implementation in the product.  What
// runs on the host is symbolic: union-find over variables, the `ask` table of DESIGN.md (entailment on
// bounds) and hashing of (op, y, z).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <unordered_map>

#include "tnf_builder.hpp"

void tb_set_error_internal(const char* s);   // engine.cu

namespace {

struct Simplifier {
  std::vector<int32_t> lb, ub;        // full index space; authoritative on class representatives only
  std::vector<int32_t> parent;
  std::vector<tb_prop> ps;            // operands are always representatives
  bool failed = false;
  tb_simplify_stats st{};
  std::vector<tb_prop> defs;          // propagators removed by eliminate_functional, in elimination order

  int find(int v) {
    while (parent[(size_t)v] != v) { parent[(size_t)v] = parent[(size_t)parent[(size_t)v]]; v = parent[(size_t)v]; }
    return v;
  }
  bool fixed(int v) const { return lb[(size_t)v] == ub[(size_t)v]; }
  bool fixed_to(int v, int32_t k) const { return lb[(size_t)v] == k && ub[(size_t)v] == k; }

  // x := x meet [l, u]; returns true when the domain moved
  bool meet(int v, int32_t l, int32_t u) {
    bool moved = false;
    if (l > lb[(size_t)v]) { lb[(size_t)v] = l; moved = true; }
    if (u < ub[(size_t)v]) { ub[(size_t)v] = u; moved = true; }
    if (lb[(size_t)v] > ub[(size_t)v]) failed = true;
    return moved;
  }

  // Merge the classes of a and b (lowest index stays the representative: the constants 0, 1, 2 keep their
  // slots and FlatZinc variables win over the auxiliaries introduced after them).
  bool unite(int a, int b) {
    a = find(a); b = find(b);
    if (a == b) return false;
    if (b < a) std::swap(a, b);
    parent[(size_t)b] = a;
    meet(a, lb[(size_t)b], ub[(size_t)b]);
    ++st.merged_variables;
    return true;
  }

  void substitute() {
    for (tb_prop& p : ps) { p.x = find(p.x); p.y = find(p.y); p.z = find(p.z); }
  }

  static bool commutative(int op) { return op == TB_OP_ADD || op == TB_OP_MUL || op == TB_OP_MIN || op == TB_OP_MAX || op == TB_OP_EQ; }

  // `ask` on bounds (frozen operator spec): -1 = the propagator is violated on singletons.
  int entailed(const tb_prop& p) const {
    const int64_t xl = lb[(size_t)p.x], xu = ub[(size_t)p.x], yl = lb[(size_t)p.y], yu = ub[(size_t)p.y], zl = lb[(size_t)p.z], zu = ub[(size_t)p.z];
    switch (p.op) {
      case TB_OP_EQ:
        if (xl >= 1) return yl == yu && zl == zu && yl == zl;
        if (xu <= 0) return yu < zl || zu < yl;
        return 0;
      case TB_OP_LEQ:
        if (xl >= 1) return yu <= zl;
        if (xu <= 0) return yl > zu;
        return 0;
      default: break;
    }
    if (xl != xu || yl != yu || zl != zu) return 0;
    bool holds = false;
    switch (p.op) {
      case TB_OP_ADD: holds = xl == yl + zl; break;
      case TB_OP_MUL: holds = xl == yl * zl; break;
      case TB_OP_TDIV: holds = zl != 0 && xl == yl / zl; break;
      case TB_OP_TMOD: holds = zl != 0 && xl == yl % zl; break;
      case TB_OP_MIN: holds = xl == std::min(yl, zl); break;
      case TB_OP_MAX: holds = xl == std::max(yl, zl); break;
      default: break;
    }
    return holds ? 1 : -1;
  }

  // All singleton classes with the same value become one variable ("no constant in a TCN" keeps constants as
  // variables, common_solving.hpp:725-727; one per value is enough).
  bool merge_constants() {
    bool changed = false;
    std::unordered_map<int32_t, int> first;
    for (int v = 0; v < (int)parent.size(); ++v) {
      if (parent[(size_t)v] != v || !fixed(v)) continue;
      auto it = first.find(lb[(size_t)v]);
      if (it == first.end()) first.emplace(lb[(size_t)v], v);
      else changed |= unite(it->second, v);
    }
    return changed;
  }

  // Algebraic simplification + elimination of entailed constraints: propagators that are equalities in
  // disguise merge their operands and disappear; entailed ones disappear.
  bool algebraic_and_entailed() {
    bool changed = false;
    std::vector<tb_prop> keep;
    keep.reserve(ps.size());
    for (tb_prop p : ps) {
      p.x = find(p.x); p.y = find(p.y); p.z = find(p.z);
      bool drop = false;
      switch (p.op) {
        case TB_OP_EQ:
          if (p.y == p.z) { changed |= meet(p.x, 1, 1); drop = true; }
          else if (lb[(size_t)p.x] >= 1) { unite(p.y, p.z); changed = true; drop = true; }
          break;
        case TB_OP_LEQ:
          if (p.y == p.z) { changed |= meet(p.x, 1, 1); drop = true; }
          break;
        case TB_OP_ADD:
          if (fixed_to(p.z, 0)) { changed |= unite(p.x, p.y); drop = true; }
          else if (fixed_to(p.y, 0)) { changed |= unite(p.x, p.z); drop = true; }
          else if (p.x == p.y && lb[(size_t)p.x] != INT32_MIN && ub[(size_t)p.x] != INT32_MAX) { changed |= meet(p.z, 0, 0); drop = true; }   // x = x + z
          else if (p.x == p.z && lb[(size_t)p.x] != INT32_MIN && ub[(size_t)p.x] != INT32_MAX) { changed |= meet(p.y, 0, 0); drop = true; }
          break;
        case TB_OP_MUL:
          if (fixed_to(p.z, 1)) { changed |= unite(p.x, p.y); drop = true; }
          else if (fixed_to(p.y, 1)) { changed |= unite(p.x, p.z); drop = true; }
          else if (fixed_to(p.y, 0) || fixed_to(p.z, 0)) { changed |= meet(p.x, 0, 0); drop = true; }
          break;
        case TB_OP_MIN:
          // one operand never exceeds the other: the minimum IS that operand (conjunctions with a constant true, ...)
          if (p.y == p.z || ub[(size_t)p.y] <= lb[(size_t)p.z]) { changed |= unite(p.x, p.y); drop = true; }
          else if (ub[(size_t)p.z] <= lb[(size_t)p.y]) { changed |= unite(p.x, p.z); drop = true; }
          break;
        case TB_OP_MAX:
          if (p.y == p.z || lb[(size_t)p.y] >= ub[(size_t)p.z]) { changed |= unite(p.x, p.y); drop = true; }
          else if (lb[(size_t)p.z] >= ub[(size_t)p.y]) { changed |= unite(p.x, p.z); drop = true; }
          break;
        default: break;
      }
      if (failed) return changed;
      if (drop) { ++st.eliminated_equalities; continue; }
      const int e = entailed(p);
      if (e < 0) { failed = true; return changed; }
      if (e > 0) { ++st.eliminated_entailed; continue; }
      keep.push_back(p);
    }
    if (keep.size() != ps.size()) changed = true;
    ps.swap(keep);
    return changed;
  }

  // ICSE: x1 = y op z and x2 = y op z  =>  x1 == x2, one propagator is enough.
  bool icse() {
    bool changed = false, again = true;
    while (again && !failed) {
      again = false;
      substitute();
      std::unordered_map<uint64_t, std::vector<size_t>> seen;     // hash -> indices of kept propagators
      std::vector<tb_prop> keep;
      keep.reserve(ps.size());
      for (tb_prop p : ps) {
        if (commutative(p.op) && p.z < p.y) std::swap(p.y, p.z);
        const uint64_t h = ((uint64_t)(uint32_t)p.op * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)p.y << 32 | (uint32_t)p.z);
        bool dup = false;
        for (size_t k : seen[h]) {
          const tb_prop& q = keep[k];
          if (q.op == p.op && q.y == p.y && q.z == p.z) {
            if (find(q.x) != find(p.x)) { unite(q.x, p.x); again = true; }
            dup = true;
            break;
          }
        }
        if (dup) { ++st.eliminated_icse; changed = true; continue; }
        seen[h].push_back(keep.size());
        keep.push_back(p);
      }
      ps.swap(keep);
    }
    return changed;
  }

  // Does dom(x) contain every value y op z can take on the current box?  Then the propagator never prunes y or z.
  bool covers(const tb_prop& p) const {
    const int64_t yl = lb[(size_t)p.y], yu = ub[(size_t)p.y], zl = lb[(size_t)p.z], zu = ub[(size_t)p.z];
    const bool finite = yl != INT32_MIN && yu != INT32_MAX && zl != INT32_MIN && zu != INT32_MAX;
    int64_t lo, hi;
    switch (p.op) {
      case TB_OP_ADD: if (!finite) return false; lo = yl + zl; hi = yu + zu; break;
      case TB_OP_MUL: {
        if (!finite) return false;
        const int64_t c[4] = {yl * zl, yl * zu, yu * zl, yu * zu};
        lo = *std::min_element(c, c + 4); hi = *std::max_element(c, c + 4);
        break;
      }
      case TB_OP_MIN: lo = std::min(yl, zl); hi = std::min(yu, zu); break;
      case TB_OP_MAX: lo = std::max(yl, zl); hi = std::max(yu, zu); break;
      case TB_OP_EQ: case TB_OP_LEQ: lo = 0; hi = 1; break;
      default: return false;
    }
    return (int64_t)lb[(size_t)p.x] <= lo && (int64_t)ub[(size_t)p.x] >= hi;
  }

  // Functionally defined variables: x occurs in one propagator only, as its result, and its domain covers the
  // range of y op z.  The propagator can then never prune anything: drop it and compute x when a solution is
  // expanded.  Dropping it may leave y or z in the same situation (cascade).  `keep` marks variables the engine
  // needs (objective, the constants 0, 1, 2).
  void eliminate_functional(const std::vector<char>& keep) {
    std::vector<int> occ(parent.size(), 0);
    for (const tb_prop& p : ps) { ++occ[(size_t)p.x]; ++occ[(size_t)p.y]; ++occ[(size_t)p.z]; }
    std::vector<char> dead(ps.size(), 0);
    bool again = true;
    while (again) {
      again = false;
      for (size_t i = 0; i < ps.size(); ++i) {
        if (dead[i]) continue;
        const tb_prop& p = ps[i];
        if (occ[(size_t)p.x] != 1 || keep[(size_t)p.x] || fixed(p.x) || !covers(p)) continue;
        dead[i] = 1;
        defs.push_back(p);
        --occ[(size_t)p.x]; --occ[(size_t)p.y]; --occ[(size_t)p.z];
        ++st.eliminated_functional;
        again = true;
      }
    }
    std::vector<tb_prop> keep_ps;
    keep_ps.reserve(ps.size());
    for (size_t i = 0; i < ps.size(); ++i) if (!dead[i]) keep_ps.push_back(ps[i]);
    ps.swap(keep_ps);
  }
};

int64_t eval_op(int op, int64_t y, int64_t z) {
  switch (op) {
    case TB_OP_ADD: return y + z;
    case TB_OP_MUL: return y * z;
    case TB_OP_MIN: return std::min(y, z);
    case TB_OP_MAX: return std::max(y, z);
    case TB_OP_EQ: return y == z ? 1 : 0;
    case TB_OP_LEQ: return y <= z ? 1 : 0;
    default: return 0;
  }
}

}  // namespace

extern "C" tb_status tb_model_simplify(tb_model* m, tb_fixpoint_fn fixpoint, void* ctx, tb_simplify_stats* out_stats) {
  if (!m || !fixpoint) { tb_set_error_internal("tb_model_simplify: null argument"); return TB_ERR_INVALID; }
  if (m->simplified) { if (out_stats) *out_stats = m->simplify_stats; return TB_OK; }
  const int V = (int)m->lb.size();
  Simplifier s;
  s.lb = m->lb; s.ub = m->ub; s.ps = m->props;
  s.parent.resize((size_t)V);
  std::iota(s.parent.begin(), s.parent.end(), 0);
  s.st.vars_before = V; s.st.props_before = (int32_t)m->props.size();
  s.failed = m->root_failed;

  // compaction of the current state: representatives that still occur (or must be kept) get dense indices
  std::vector<int32_t> dense((size_t)V, -1);
  std::vector<int32_t> clb, cub, members;
  std::vector<tb_prop> cps;
  auto compact = [&](bool drop_useless) {
    std::vector<char> used((size_t)V, drop_useless ? 0 : 1);
    if (drop_useless) {
      for (const tb_prop& p : s.ps) used[(size_t)p.x] = used[(size_t)p.y] = used[(size_t)p.z] = 1;
      for (int k = 0; k < 3 && k < V; ++k) used[(size_t)s.find(k)] = 1;              // the constants 0, 1, 2 stay variables
      if (m->obj_var >= 0) used[(size_t)s.find(m->obj_var)] = 1;
      if (m->user_obj_var >= 0) used[(size_t)s.find(m->user_obj_var)] = 1;
    }
    std::fill(dense.begin(), dense.end(), -1);
    clb.clear(); cub.clear(); members.clear();
    for (int v = 0; v < V; ++v)
      if (s.parent[(size_t)v] == v && used[(size_t)v]) { dense[(size_t)v] = (int32_t)clb.size(); clb.push_back(s.lb[(size_t)v]); cub.push_back(s.ub[(size_t)v]); members.push_back(v); }
    cps = s.ps;
    for (tb_prop& p : cps) { p.x = dense[(size_t)p.x]; p.y = dense[(size_t)p.y]; p.z = dense[(size_t)p.z]; }
  };

  bool changed = true;
  while (changed && !s.failed) {
    changed = false;
    ++s.st.iterations;
    // 1. root fixpoint of what is left, by the engine
    s.substitute();
    compact(false);
    tb_problem pb;
    memset(&pb, 0, sizeof(pb));
    tb_strategy all{TB_VAR_INPUT_ORDER, TB_VAL_MIN, 0, nullptr};
    pb.nvars = (int32_t)clb.size(); pb.nprops = (int32_t)cps.size();
    pb.lb = clb.data(); pb.ub = cub.data(); pb.props = cps.data();
    pb.nstrategies = 1; pb.strategies = &all; pb.obj_var = -1;
    std::vector<int32_t> nlb = clb, nub = cub;
    int32_t f = 0;
    if (!cps.empty()) {
      tb_status rc = fixpoint(ctx, &pb, nlb.data(), nub.data(), &f);
      if (rc != TB_OK) return rc;
    }
    if (f) { s.failed = true; break; }
    for (size_t i = 0; i < members.size(); ++i) changed |= s.meet(members[i], nlb[i], nub[i]);
    if (s.failed) break;
    // 2. symbolic passes
    changed |= s.merge_constants();
    changed |= s.algebraic_and_entailed();
    if (s.failed) break;
    changed |= s.icse();
  }

  // ---- install the reduced network -----------------------------------------------------------------------
  m->full_lb = m->lb; m->full_ub = m->ub; m->full_props = m->props;
  m->root_failed = m->root_failed || s.failed;
  s.substitute();
  if (!m->root_failed) {
    std::vector<char> keep((size_t)V, 0);
    for (int k = 0; k < 3 && k < V; ++k) keep[(size_t)s.find(k)] = 1;
    if (m->obj_var >= 0) keep[(size_t)s.find(m->obj_var)] = 1;
    if (m->user_obj_var >= 0) keep[(size_t)s.find(m->user_obj_var)] = 1;
    s.eliminate_functional(keep);
  }
  compact(true);
  m->rep_of_full.resize((size_t)V);
  for (int v = 0; v < V; ++v) m->rep_of_full[(size_t)v] = s.find(v);
  m->defs = s.defs;
  m->red_of_full.assign((size_t)V, -1);
  for (int v = 0; v < V; ++v) {
    const int r = s.find(v);
    m->red_of_full[(size_t)v] = dense[(size_t)r];
    m->full_lb[(size_t)v] = s.lb[(size_t)r]; m->full_ub[(size_t)v] = s.ub[(size_t)r];
  }
  s.st.eliminated_variables = 0;
  for (int v = 0; v < V; ++v) if (s.parent[(size_t)v] == v && dense[(size_t)v] < 0) ++s.st.eliminated_variables;
  m->lb = clb; m->ub = cub; m->props = cps;
  // strategies: representatives, first occurrence only; eliminated and root-assigned variables never branch
  std::vector<std::vector<int32_t>> nsv;
  std::vector<std::pair<int, int>> nso;
  for (size_t i = 0; i < m->strat_vars.size(); ++i) {
    const bool was_all = m->strat_vars[i].empty();
    std::vector<int32_t> vs;
    std::vector<char> seen(clb.size(), 0);
    for (int32_t v : m->strat_vars[i]) {
      const int32_t d = m->red_of_full[(size_t)v];
      if (d < 0 || seen[(size_t)d] || clb[(size_t)d] == cub[(size_t)d]) continue;
      seen[(size_t)d] = 1;
      vs.push_back(d);
    }
    if (!was_all && vs.empty()) continue;          // an empty list would mean "all variables" in the ABI
    nsv.push_back(std::move(vs));
    nso.push_back(m->strat_orders[i]);
  }
  m->strat_vars.swap(nsv); m->strat_orders.swap(nso);
  if (m->obj_var >= 0) m->obj_var = m->red_of_full[(size_t)m->obj_var];
  if (m->user_obj_var >= 0) m->user_obj_var = m->red_of_full[(size_t)m->user_obj_var];
  s.st.vars_after = (int32_t)clb.size(); s.st.props_after = (int32_t)cps.size();
  s.st.root_failed = m->root_failed ? 1 : 0;
  m->simplified = true;
  m->simplify_stats = s.st;
  m->finalize();
  if (out_stats) *out_stats = s.st;
  return TB_OK;
}

// Point of the full network for a store of the reduced one: members of a class take the value of their
// representative, eliminated variables (no constraint left on them) the lower bound of their root domain.
void tb_model_expand_internal(const tb_model* m, const int32_t* lb, const int32_t* ub, std::vector<int32_t>& flb, std::vector<int32_t>& fub) {
  const size_t V = m->red_of_full.size();
  flb.resize(V); fub.resize(V);
  // values of the class representatives: from the store, or the root domain when nothing constrains the class
  for (size_t v = 0; v < V; ++v) {
    if ((size_t)m->rep_of_full[v] != v) continue;
    const int32_t d = m->red_of_full[v];
    if (d >= 0) { flb[v] = lb[d]; fub[v] = ub ? ub[d] : lb[d]; }
    else { flb[v] = m->full_lb[v]; fub[v] = m->full_ub[v]; }
  }
  // functionally defined variables, innermost definition last: evaluate in reverse elimination order at the point lb
  for (size_t i = m->defs.size(); i-- > 0;) {
    const tb_prop& p = m->defs[i];
    const int32_t x = (int32_t)eval_op(p.op, flb[(size_t)p.y], flb[(size_t)p.z]);
    flb[(size_t)p.x] = x; fub[(size_t)p.x] = x;
  }
  for (size_t v = 0; v < V; ++v) {
    const size_t r = (size_t)m->rep_of_full[v];
    if (r != v) { flb[v] = flb[r]; fub[v] = fub[r]; }
  }
}

extern "C" int32_t tb_model_num_full_variables(const tb_model* m) {
  if (!m) return 0;
  return (int32_t)(m->simplified ? m->red_of_full.size() : m->lb.size());
}

extern "C" tb_status tb_model_expand_solution(const tb_model* m, const int32_t* lb, const int32_t* ub, int32_t* full_lb, int32_t* full_ub) {
  if (!m || !lb || !full_lb) { tb_set_error_internal("tb_model_expand_solution: null argument"); return TB_ERR_INVALID; }
  if (!m->simplified) {
    memcpy(full_lb, lb, m->lb.size() * sizeof(int32_t));
    if (full_ub) memcpy(full_ub, ub ? ub : lb, m->lb.size() * sizeof(int32_t));
    return TB_OK;
  }
  std::vector<int32_t> a, b;
  tb_model_expand_internal(m, lb, ub, a, b);
  memcpy(full_lb, a.data(), a.size() * sizeof(int32_t));
  if (full_ub) memcpy(full_ub, b.data(), b.size() * sizeof(int32_t));
  return TB_OK;
}
