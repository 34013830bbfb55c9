// model_capi.cpp — C surface of the host front-end (tb_model_* in include/turbo_b200.h):
// FlatZinc loading, the synthetic TNF generator of BASELINE config 5, the binary .tnf format used
// by tests/golden, the solution printer and the solution checkers.
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#include "tnf_builder.hpp"

void tb_set_error_internal(const char* s);   // engine.cu

namespace {

tb_status fail(tb_status rc, const std::string& msg) { tb_set_error_internal(msg.c_str()); return rc; }

tb_status parse_text(tb_model** out, const char* text, size_t len, uint32_t flags) {
  (void)flags;
  if (!out) return fail(TB_ERR_INVALID, "null argument");
  *out = nullptr;
  try {
    std::unique_ptr<fzn::Model> src = fzn::parse(text, len);
    std::unique_ptr<tb_model> m = build_tnf(std::move(src));
    *out = m.release();
    return TB_OK;
  } catch (const fzn::ParseError& e) {
    return fail(TB_ERR_PARSE, "line " + std::to_string(e.line) + ": " + e.message);
  } catch (const std::exception& e) {
    return fail(TB_ERR_UNSUPPORTED, e.what());
  }
}

// ---- truncated-division helpers for tb_model_check_tnf (independent of the kernels and the oracle) --
bool prop_holds(const tb_prop& p, const int32_t* v) {
  const int64_t x = v[p.x], y = v[p.y], z = v[p.z];
  switch (p.op) {
    case TB_OP_ADD: return x == y + z;
    case TB_OP_MUL: return x == y * z;
    case TB_OP_TDIV: return z != 0 && x == y / z;
    case TB_OP_TMOD: return z != 0 && x == y % z;
    case TB_OP_MIN: return x == std::min(y, z);
    case TB_OP_MAX: return x == std::max(y, z);
    case TB_OP_EQ: return (x == 0 || x == 1) && x == (y == z ? 1 : 0);
    case TB_OP_LEQ: return (x == 0 || x == 1) && x == (y <= z ? 1 : 0);
    default: return false;
  }
}

}  // namespace

extern "C" {

tb_status tb_model_parse_fzn(tb_model** out, const char* text, size_t len, uint32_t flags) {
  if (!text) return fail(TB_ERR_INVALID, "null text");
  return parse_text(out, text, len, flags);
}

tb_status tb_model_load_fzn(tb_model** out, const char* path, uint32_t flags) {
  if (!path) return fail(TB_ERR_INVALID, "null path");
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(TB_ERR_IO, std::string("cannot open ") + path);
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  return parse_text(out, text.data(), text.size(), flags);
}

// BASELINE config 5: planted solution, std::mt19937_64, constants 0/1/2 as variables.
tb_status tb_model_synthetic(tb_model** out, int32_t nvars, int32_t nprops, uint64_t seed) {
  if (!out || nvars < 8 || nprops < 0) return fail(TB_ERR_INVALID, "tb_model_synthetic: need nvars >= 8");
  std::mt19937_64 rng(seed);
  auto uni = [&](int64_t lo, int64_t hi) { return (int64_t)(lo + (int64_t)(rng() % (uint64_t)(hi - lo + 1))); };
  std::unique_ptr<tb_model> m(new tb_model());
  std::vector<int32_t> s((size_t)nvars);
  for (int v = 0; v < nvars; ++v) s[v] = (int32_t)uni(-500, 500);
  for (int v = 3; v < nvars; ++v) if (rng() % 10 == 0) s[v] = (int32_t)(rng() & 1);     // reified results
  s[0] = 0; s[1] = 1; s[2] = 2;
  std::unordered_map<int32_t, std::vector<int32_t>> by_value;
  for (int v = 0; v < nvars; ++v) by_value[s[v]].push_back(v);
  m->props.reserve((size_t)nprops);
  while ((int)m->props.size() < nprops) {
    const unsigned pick = (unsigned)(rng() % 100);
    int op = pick < 40 ? TB_OP_ADD : pick < 65 ? TB_OP_LEQ : pick < 80 ? TB_OP_EQ : pick < 85 ? TB_OP_MIN : pick < 90 ? TB_OP_MAX : TB_OP_MUL;
    const int y = (int)(rng() % (uint64_t)nvars), z = (int)(rng() % (uint64_t)nvars);
    int64_t r;
    switch (op) {
      case TB_OP_ADD: r = (int64_t)s[y] + s[z]; break;
      case TB_OP_LEQ: r = s[y] <= s[z]; break;
      case TB_OP_EQ: r = s[y] == s[z]; break;
      case TB_OP_MIN: r = std::min(s[y], s[z]); break;
      case TB_OP_MAX: r = std::max(s[y], s[z]); break;
      default:
        if (std::abs(s[y]) > 20 || std::abs(s[z]) > 20) continue;
        r = (int64_t)s[y] * s[z];
        break;
    }
    auto it = by_value.find((int32_t)r);
    if (r < -1000000 || r > 1000000 || it == by_value.end()) continue;
    const int x = it->second[(size_t)(rng() % it->second.size())];
    m->props.push_back(tb_prop{op, x, y, z});
  }
  m->lb.resize((size_t)nvars); m->ub.resize((size_t)nvars);
  for (int v = 0; v < nvars; ++v) {
    if (rng() % 10 == 0) { m->lb[v] = m->ub[v] = s[v]; continue; }
    m->lb[v] = s[v] - (int32_t)uni(0, 32);
    m->ub[v] = s[v] + (int32_t)uni(0, 32);
  }
  for (int k = 0; k < 3; ++k) m->lb[k] = m->ub[k] = k;
  // the result of a reified comparison is a 0..1 variable (precondition of TB_OP_EQ / TB_OP_LEQ)
  for (const tb_prop& p : m->props)
    if (p.op == TB_OP_EQ || p.op == TB_OP_LEQ) { m->lb[p.x] = std::max(m->lb[p.x], 0); m->ub[p.x] = std::min(m->ub[p.x], 1); }
  m->strat_vars.push_back({});
  m->strat_orders.push_back({TB_VAR_FIRST_FAIL, TB_VAL_MIN});
  m->obj_var = -1; m->objective_kind = -1;
  m->finalize();
  *out = m.release();
  return TB_OK;
}

// ---- binary TNF files: "TNF1", header ints, lb[], ub[], props[], strategies -------------------------
tb_status tb_model_save_tnf(const tb_model* m, const char* path) {
  if (!m || !path) return fail(TB_ERR_INVALID, "null argument");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(TB_ERR_IO, std::string("cannot write ") + path);
  const int32_t hdr[8] = {0x31464e54, (int32_t)m->lb.size(), (int32_t)m->props.size(), (int32_t)m->strat_vars.size(),
                          m->obj_var, m->has_eps_strategy ? 1 : 0, m->objective_kind, m->user_obj_var};
  bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1;
  if (!m->lb.empty()) { ok = ok && fwrite(m->lb.data(), 4, m->lb.size(), f) == m->lb.size(); ok = ok && fwrite(m->ub.data(), 4, m->ub.size(), f) == m->ub.size(); }
  if (!m->props.empty()) ok = ok && fwrite(m->props.data(), sizeof(tb_prop), m->props.size(), f) == m->props.size();
  for (size_t i = 0; i < m->strat_vars.size(); ++i) {
    const int32_t sh[3] = {m->strat_orders[i].first, m->strat_orders[i].second, (int32_t)m->strat_vars[i].size()};
    ok = ok && fwrite(sh, sizeof(sh), 1, f) == 1;
    if (sh[2]) ok = ok && fwrite(m->strat_vars[i].data(), 4, (size_t)sh[2], f) == (size_t)sh[2];
  }
  fclose(f);
  return ok ? TB_OK : fail(TB_ERR_IO, "short write");
}

tb_status tb_model_load_tnf(tb_model** out, const char* path) {
  if (!out || !path) return fail(TB_ERR_INVALID, "null argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return fail(TB_ERR_IO, std::string("cannot open ") + path);
  int32_t hdr[8];
  std::unique_ptr<tb_model> m(new tb_model());
  bool ok = fread(hdr, sizeof(hdr), 1, f) == 1 && hdr[0] == 0x31464e54 && hdr[1] >= 0 && hdr[2] >= 0 && hdr[3] >= 0;
  if (ok) {
    m->lb.resize((size_t)hdr[1]); m->ub.resize((size_t)hdr[1]); m->props.resize((size_t)hdr[2]);
    if (hdr[1]) { ok = ok && fread(m->lb.data(), 4, m->lb.size(), f) == m->lb.size(); ok = ok && fread(m->ub.data(), 4, m->ub.size(), f) == m->ub.size(); }
    if (hdr[2]) ok = ok && fread(m->props.data(), sizeof(tb_prop), m->props.size(), f) == m->props.size();
    for (int i = 0; ok && i < hdr[3]; ++i) {
      int32_t sh[3];
      ok = fread(sh, sizeof(sh), 1, f) == 1 && sh[2] >= 0;
      if (!ok) break;
      std::vector<int32_t> vs((size_t)sh[2]);
      if (sh[2]) ok = fread(vs.data(), 4, vs.size(), f) == vs.size();
      m->strat_vars.push_back(std::move(vs));
      m->strat_orders.push_back({sh[0], sh[1]});
    }
    m->obj_var = hdr[4]; m->has_eps_strategy = hdr[5] != 0; m->objective_kind = hdr[6]; m->user_obj_var = hdr[7];
  }
  fclose(f);
  if (!ok) return fail(TB_ERR_PARSE, std::string("malformed TNF file ") + path);
  m->finalize();
  *out = m.release();
  return TB_OK;
}

const tb_problem* tb_model_problem(const tb_model* m) { return m ? &m->problem : nullptr; }
int32_t tb_model_objective_kind(const tb_model* m) { return m ? m->objective_kind : -1; }
int32_t tb_model_user_objective_var(const tb_model* m) { return m ? m->user_obj_var : -1; }
int32_t tb_model_num_parsed_variables(const tb_model* m) { return m ? m->parsed_variables : 0; }
int32_t tb_model_num_parsed_constraints(const tb_model* m) { return m ? m->parsed_constraints : 0; }
int32_t tb_model_root_failed(const tb_model* m) { return m && m->root_failed ? 1 : 0; }
void tb_model_destroy(tb_model* m) { delete m; }

tb_status tb_model_push_eps_strategy(tb_model* m, int32_t var_order, int32_t val_order) {
  if (!m) return fail(TB_ERR_INVALID, "null model");
  m->push_eps_strategy(var_order, val_order);
  return TB_OK;
}

int32_t tb_model_check_tnf(const tb_model* m, const int32_t* lb) {
  if (!m || !lb) return -1;
  int bad = 0;
  if (m->simplified) {             // expand to the full network and check every original propagator there
    std::vector<int32_t> flb, fub;
    tb_model_expand_internal(m, lb, nullptr, flb, fub);
    for (size_t v = 0; v < flb.size(); ++v) bad += (flb[v] < m->full_lb[v] || flb[v] > m->full_ub[v]) ? 1 : 0;
    for (const tb_prop& p : m->full_props) bad += prop_holds(p, flb.data()) ? 0 : 1;
    return bad;
  }
  for (size_t v = 0; v < m->lb.size(); ++v) bad += (lb[v] < m->lb[v] || lb[v] > m->ub[v]) ? 1 : 0;
  for (const tb_prop& p : m->props) bad += prop_holds(p, lb) ? 0 : 1;
  return bad;
}

int32_t tb_model_check_solution(const tb_model* m, const int32_t* lb, const int32_t* ub) {
  (void)ub;
  if (!m || !lb || !m->src) return -1;
  std::vector<int32_t> flb, fub;
  if (m->simplified) { tb_model_expand_internal(m, lb, nullptr, flb, fub); lb = flb.data(); }
  std::vector<int64_t> value(m->src->vars.size());
  for (size_t i = 0; i < value.size(); ++i) value[i] = lb[m->var_of_model[i]];
  try {
    std::string first;
    int bad = check_flatzinc(*m->src, value, &first);
    if (bad) tb_set_error_internal(("violated: " + first).c_str());
    return bad;
  } catch (const std::exception& e) {
    tb_set_error_internal(e.what());
    return -1;
  }
}

// SolverOutput::print_solution (lala-parsing; called at common_solving.hpp:849): `name = v;` lines
// for output_var items and `name = arrayNd(dims..., [..]);` for output_array items, in declaration order.
size_t tb_model_format_solution(const tb_model* m, const int32_t* lb, const int32_t* ub, char* buf, size_t cap) {
  (void)ub;
  if (!m || !lb || !m->src) return 0;
  std::vector<int32_t> flb, fub;
  if (m->simplified) { tb_model_expand_internal(m, lb, nullptr, flb, fub); lb = flb.data(); }
  const fzn::Model& src = *m->src;
  std::string s;
  auto value_of = [&](const fzn::Expr& e, bool is_bool) -> std::string {
    int64_t v = e.kind == fzn::Expr::VAR ? lb[m->var_of_model[(size_t)e.var]] : e.value;
    const bool b = is_bool || e.kind == fzn::Expr::BOOL || (e.kind == fzn::Expr::VAR && src.vars[(size_t)e.var].is_bool);
    if (b) return v ? "true" : "false";
    return std::to_string(v);
  };
  for (const std::string& item : src.output_order) {
    const std::string name = item.substr(2);
    if (item[0] == 'v') {
      fzn::Expr e; e.kind = fzn::Expr::VAR; e.var = src.var_index.at(name);
      s += name + " = " + value_of(e, false) + ";\n";
    } else {
      for (const fzn::OutputArray& a : src.output_arrays) {
        if (a.name != name) continue;
        s += name + " = array" + std::to_string(a.dims.size()) + "d(";
        for (const auto& d : a.dims) s += std::to_string(d.first) + ".." + std::to_string(d.second) + ", ";
        s += "[";
        for (size_t i = 0; i < a.elems.size(); ++i) { if (i) s += ", "; s += value_of(a.elems[i], a.is_bool); }
        s += "]);\n";
      }
    }
  }
  if (buf && cap) {
    size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}

}  // extern "C"
