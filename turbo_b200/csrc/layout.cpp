// layout.cpp — see layout.h.
#include "layout.h"

#include <algorithm>
#include <numeric>

namespace {

inline bool is_small(const int32_t* lb, const int32_t* ub, int v) {
  return lb[v] >= -TBC_SMALL_LIMIT && ub[v] <= TBC_SMALL_LIMIT;
}
inline bool is_fixed(const int32_t* lb, const int32_t* ub, int v) { return lb[v] == ub[v]; }
// a root constant whose value fits a 21-bit field
inline bool is_fieldk(const int32_t* lb, const int32_t* ub, int v) {
  return lb[v] == ub[v] && lb[v] >= -TBC_CONST_LIMIT && lb[v] < TBC_CONST_LIMIT;
}

struct Item { int cls; int x, y, z; };   // x/z may hold a constant depending on the class

inline bool loads_x(int c) { return !(c == TBC_ADD_XK || c == TBC_EQ_T || c == TBC_EQ_F || c == TBC_LEQ_T || c == TBC_LEQ_F); }
inline bool loads_z(int c) { return !(c == TBC_ADD_ZK || c == TBC_EQ_ZK || c == TBC_LEQ_ZK); }

}  // namespace

const char* tb_class_name(int cls) {
  static const char* names[TBC_NUM] = {"add_s", "add_xk", "add_zk", "add_g", "mul", "tdiv", "tmod", "min", "max",
                                       "eq_s", "eq_t", "eq_f", "eq_zk", "eq_g", "leq_s", "leq_t", "leq_f", "leq_zk", "leq_g"};
  return cls >= 0 && cls < TBC_NUM ? names[cls] : "?";
}

int tb_classify(const tb_prop& p, const int32_t* lb, const int32_t* ub, bool* swap_yz) {
  *swap_yz = false;
  const bool sx = is_small(lb, ub, p.x), sy = is_small(lb, ub, p.y), sz = is_small(lb, ub, p.z);
  switch (p.op) {
    case TB_OP_ADD:
      if (!(sx && sy && sz)) return TBC_ADD_G;
      if (is_fieldk(lb, ub, p.x)) return TBC_ADD_XK;
      if (is_fieldk(lb, ub, p.z)) return TBC_ADD_ZK;
      if (is_fieldk(lb, ub, p.y)) { *swap_yz = true; return TBC_ADD_ZK; }
      return TBC_ADD_S;
    case TB_OP_MUL: return TBC_MUL;
    case TB_OP_TDIV: return TBC_TDIV;
    case TB_OP_TMOD: return TBC_TMOD;
    case TB_OP_MIN: return TBC_MIN;
    case TB_OP_MAX: return TBC_MAX;
    case TB_OP_EQ:
      if (!(sy && sz)) return TBC_EQ_G;
      if (is_fixed(lb, ub, p.x) && lb[p.x] == 1) return TBC_EQ_T;
      if (is_fixed(lb, ub, p.x) && lb[p.x] == 0) return TBC_EQ_F;
      if (is_fieldk(lb, ub, p.z)) return TBC_EQ_ZK;
      if (is_fieldk(lb, ub, p.y)) { *swap_yz = true; return TBC_EQ_ZK; }
      return TBC_EQ_S;
    default:   // TB_OP_LEQ
      if (!(sy && sz)) return TBC_LEQ_G;
      if (is_fixed(lb, ub, p.x) && lb[p.x] == 1) return TBC_LEQ_T;
      if (is_fixed(lb, ub, p.x) && lb[p.x] == 0) return TBC_LEQ_F;
      if (is_fieldk(lb, ub, p.z)) return TBC_LEQ_ZK;
      return TBC_LEQ_S;
  }
}

tb_status tb_build_layout(const tb_problem* pb, const TnfLayoutOptions& opt, TnfLayout* out, std::string* err) {
  const int V = pb->nvars, P = pb->nprops;
  if (V > TBC_MAX_VARS) {
    if (err) *err = "more than 2^21 variables: the device propagator word has 21-bit fields";
    return TB_ERR_UNSUPPORTED;
  }
  TnfLayout& L = *out;
  L = TnfLayout();
  L.nvars = V;
  int align = std::max(1, opt.slot_align);
  if (opt.nbanks > 0) align = std::max(align, opt.nbanks);
  L.nslots = std::max(align, (V + align - 1) / align * align);
  L.referenced.assign((size_t)V, 0);

  // ---- classify, sort by class (stable: the ternariser's order carries the model's locality) ---------------
  std::vector<Item> items((size_t)P);
  for (int i = 0; i < P; ++i) {
    const tb_prop& p = pb->props[i];
    bool swap = false;
    Item it;
    it.cls = tb_classify(p, pb->lb, pb->ub, &swap);
    it.x = p.x; it.y = swap ? p.z : p.y; it.z = swap ? p.y : p.z;
    items[(size_t)i] = it;
    L.referenced[p.x] = L.referenced[p.y] = L.referenced[p.z] = 1;
  }
  std::vector<int> order((size_t)P);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return items[a].cls < items[b].cls; });

  // ---- STORE_CLUSTER: which CTA a variable lives in ------------------------------------------------------------
  // part[v] in 0..C-1, every part at most nslots / C variables. Breadth-first order over the propagator hypergraph
  // (variables that occur together are visited together), cut into C consecutive runs.
  const int C = opt.cluster > 1 ? opt.cluster : 0;
  std::vector<int> part;
  if (C) {
    const int cap = L.nslots / C;
    part.assign((size_t)V, -1);
    // variable -> propagators (CSR), loaded operands only
    std::vector<int> deg((size_t)V, 0);
    auto each_operand = [&](const Item& it, auto&& f) { if (loads_x(it.cls)) f(it.x); f(it.y); if (loads_z(it.cls)) f(it.z); };
    for (const Item& it : items) each_operand(it, [&](int v) { ++deg[v]; });
    std::vector<size_t> off((size_t)V + 1, 0);
    for (int v = 0; v < V; ++v) off[v + 1] = off[v] + (size_t)deg[v];
    std::vector<int> inc(off[V]);
    {
      std::vector<size_t> fill(off.begin(), off.end() - 1);
      for (int i = 0; i < P; ++i) each_operand(items[(size_t)i], [&](int v) { inc[fill[v]++] = i; });
    }
    std::vector<int> bfs;
    bfs.reserve((size_t)V);
    std::vector<char> seen((size_t)V, 0), pseen((size_t)P, 0);
    for (int root = 0; root < V; ++root) {
      if (seen[root] || !deg[root]) continue;
      size_t head = bfs.size();
      bfs.push_back(root); seen[root] = 1;
      while (head < bfs.size()) {
        const int v = bfs[head++];
        for (size_t k = off[v]; k < off[v + 1]; ++k) {
          const int i = inc[k];
          if (pseen[i]) continue;
          pseen[i] = 1;
          each_operand(items[(size_t)i], [&](int u) { if (!seen[u]) { seen[u] = 1; bfs.push_back(u); } });
        }
      }
    }
    for (int v = 0; v < V; ++v) if (!seen[v]) bfs.push_back(v);       // variables no propagator loads
    for (size_t k = 0; k < bfs.size(); ++k) part[(size_t)bfs[k]] = std::min(C - 1, (int)(k / (size_t)cap));
    // Refinement: label propagation under the capacity. A variable wants the part most of its co-operands live in;
    // wishes a -> b are granted in pairs with wishes b -> a (largest gains first), so every part keeps its size.
    {
      std::vector<int> want((size_t)V), gain((size_t)V);
      for (int pass = 0; pass < 12; ++pass) {
        for (int v = 0; v < V; ++v) {
          want[(size_t)v] = part[(size_t)v]; gain[(size_t)v] = 0;
          if (!deg[v] || deg[v] > 4096) continue;            // (hubs stay where they are: everybody's neighbour)
          int cnt[16] = {0};
          for (size_t k = off[v]; k < off[v + 1]; ++k)
            each_operand(items[(size_t)inc[k]], [&](int u) { if (u != v) ++cnt[part[(size_t)u] & 15]; });
          int best = part[(size_t)v];
          for (int q = 0; q < C; ++q) if (cnt[q] > cnt[best]) best = q;
          want[(size_t)v] = best; gain[(size_t)v] = cnt[best] - cnt[part[(size_t)v]];
        }
        std::vector<std::vector<int>> wish((size_t)C * C);
        for (int v = 0; v < V; ++v) if (want[(size_t)v] != part[(size_t)v]) wish[(size_t)part[(size_t)v] * C + want[(size_t)v]].push_back(v);
        size_t moved = 0;
        for (int p = 0; p < C; ++p)
          for (int q = p + 1; q < C; ++q) {
            auto& ab = wish[(size_t)p * C + q];
            auto& ba = wish[(size_t)q * C + p];
            auto by_gain = [&](int x, int y) { return gain[(size_t)x] > gain[(size_t)y]; };
            std::stable_sort(ab.begin(), ab.end(), by_gain);
            std::stable_sort(ba.begin(), ba.end(), by_gain);
            const size_t n = std::min(ab.size(), ba.size());
            for (size_t k = 0; k < n; ++k) { part[(size_t)ab[k]] = q; part[(size_t)ba[k]] = p; }
            moved += 2 * n;
          }
        if (moved * 200 < (size_t)V) break;                  // under half a percent moved: settled
      }
    }
    // ---- and which CTA a propagator is evaluated in: the one most of its operands live in. Inside a class the chunks
    // are filled so that the chunk a CTA's warps visit (ch mod (C * warps) / warps) holds that CTA's propagators.
    const int W = std::max(1, opt.cluster_warps);
    auto home = [&](const Item& it) {
      int cnt[16] = {0};
      each_operand(it, [&](int v) { ++cnt[part[v] & 15]; });
      int best = part[it.y];
      for (int q = 0; q < C; ++q) if (cnt[q] > cnt[best]) best = q;
      return best;
    };
    std::vector<int> reordered;
    reordered.reserve(order.size());
    size_t i = 0, chunk = 0;
    const size_t CHK = (size_t)32 * TBC_U;
    while (i < order.size()) {
      const int c = items[(size_t)order[i]].cls;
      std::vector<std::vector<int>> bucket((size_t)C);
      size_t n = 0;
      for (; i < order.size() && items[(size_t)order[i]].cls == c; ++i, ++n) bucket[(size_t)home(items[(size_t)order[i]])].push_back(order[i]);
      std::vector<size_t> taken((size_t)C, 0);
      size_t left = n;
      while (left) {
        const int cta = (int)((chunk % ((size_t)C * W)) / (size_t)W);
        size_t want = std::min(CHK, left);
        int q = cta;
        while (want) {
          if (taken[(size_t)q] == bucket[(size_t)q].size()) {          // this CTA's propagators are used up: take from the fullest bucket
            q = 0;
            for (int r = 1; r < C; ++r) if (bucket[(size_t)r].size() - taken[(size_t)r] > bucket[(size_t)q].size() - taken[(size_t)q]) q = r;
          }
          reordered.push_back(bucket[(size_t)q][taken[(size_t)q]++]);
          --want; --left;
        }
        ++chunk;
      }
    }
    order.swap(reordered);
  }

  // row table: rows of 32 lanes holding indices into `items`, -1 = padding (filled with a copy of the class's
  // last propagator). A chunk is TBC_U consecutive rows of one class: lane l of a warp evaluates lane l of each.
  const int CH = 32 * TBC_U;
  std::vector<int> lanes;
  lanes.reserve((size_t)P + (size_t)CH * TBC_NUM);
  {
    size_t i = 0;
    for (int c = 0; c < TBC_NUM; ++c) {
      L.cls_begin[c] = (int)(lanes.size() / CH);
      size_t n = 0;
      while (i < order.size() && items[order[i]].cls == c) { lanes.push_back(order[i]); ++i; ++n; }
      L.cls_last[c] = n ? (int)((n - 1) % CH + 1) : 0;
      while (lanes.size() % CH) lanes.push_back(-1);
    }
    L.cls_begin[TBC_NUM] = (int)(lanes.size() / CH);
  }
  const int nrows = (int)(lanes.size() / 32);

  // ---- variable placement --------------------------------------------------------------------------------
  L.slot_of.resize((size_t)V);
  std::iota(L.slot_of.begin(), L.slot_of.end(), 0);
  L.identity = true;
  // role sets: for every chunk and every loaded operand position, the distinct variables the 32 lanes read
  std::vector<std::vector<int>> sets;
  if (nrows) sets.reserve((size_t)nrows * 6);
  // (a 64-bit shared-memory load is served one half-warp at a time: lanes_per_set = 16)
  const int LPS = opt.lanes_per_set > 0 ? opt.lanes_per_set : 32;
  for (int row = 0; row < nrows; ++row) {
    int first = -1, real = 0;
    for (int l = 0; l < 32; ++l) if (lanes[(size_t)row * 32 + l] >= 0) { if (first < 0) first = lanes[(size_t)row * 32 + l]; ++real; }
    if (first < 0) continue;                      // a row of padding only
    const int cls = items[first].cls;
    for (int role = 0; role < 3; ++role) {
      if ((role == 0 && !loads_x(cls)) || (role == 2 && !loads_z(cls))) continue;
      L.loads_per_sweep += (uint64_t)real;
      for (int l0 = 0; l0 < 32; l0 += LPS) {
        std::vector<int> vs;
        for (int l = l0; l < l0 + LPS; ++l) {
          const int id = lanes[(size_t)row * 32 + l];
          if (id < 0) continue;
          const Item& it = items[id];
          vs.push_back(role == 0 ? it.x : (role == 1 ? it.y : it.z));
        }
        if (vs.empty()) continue;
        std::sort(vs.begin(), vs.end());
        vs.erase(std::unique(vs.begin(), vs.end()), vs.end());
        sets.push_back(std::move(vs));
      }
    }
  }
  if (C) {
    // slot = local index * C + CTA (the device finds slot s in CTA s mod C)
    std::vector<int> next((size_t)C);
    std::iota(next.begin(), next.end(), 0);
    for (int v = 0; v < V; ++v) { L.slot_of[v] = next[(size_t)part[v]]; next[(size_t)part[v]] += C; }
    L.identity = true;
    for (int v = 0; v < V; ++v) if (L.slot_of[v] != v) { L.identity = false; break; }
    // how local the sweep is: operand loads that stay in the CTA whose warps visit the chunk
    const int W = std::max(1, opt.cluster_warps);
    unsigned long long local = 0, total = 0;
    for (size_t k = 0; k < lanes.size(); ++k) {
      if (lanes[k] < 0) continue;
      const Item& it = items[(size_t)lanes[k]];
      const int cta = (int)(((k / 32 / TBC_U) % ((size_t)C * W)) / (size_t)W);
      auto f = [&](int v) { ++total; local += (L.slot_of[v] % C) == cta; };
      if (loads_x(it.cls)) f(it.x);
      f(it.y);
      if (loads_z(it.cls)) f(it.z);
    }
    L.cluster_local_fraction = total ? (double)local / (double)total : 0.0;
  }
  if (opt.nbanks > 0 && V > 0) {
    const int NB = opt.nbanks;
    // membership lists
    std::vector<int> deg((size_t)V, 0);
    for (const auto& s : sets) for (int v : s) ++deg[v];
    std::vector<size_t> off((size_t)V + 1, 0);
    for (int v = 0; v < V; ++v) off[v + 1] = off[v] + (size_t)deg[v];
    std::vector<int> memb(off[V]);
    {
      std::vector<size_t> fill(off.begin(), off.end() - 1);
      for (size_t s = 0; s < sets.size(); ++s) for (int v : sets[s]) memb[fill[v]++] = (int)s;
    }
    std::vector<uint16_t> cnt(sets.size() * (size_t)NB, 0);
    std::vector<uint16_t> mx(sets.size(), 0);
    std::vector<int> pop((size_t)NB, 0);
    const int cap = L.nslots / NB;
    std::vector<int> bank((size_t)V, 0);
    std::vector<int> vorder((size_t)V);
    std::iota(vorder.begin(), vorder.end(), 0);
    std::stable_sort(vorder.begin(), vorder.end(), [&](int a, int b) { return deg[a] > deg[b]; });
    std::vector<long long> score((size_t)NB);
    for (int v : vorder) {
      std::fill(score.begin(), score.end(), 0);
      for (size_t k = off[v]; k < off[v + 1]; ++k) {
        const uint16_t* c = &cnt[(size_t)memb[k] * NB];
        const uint16_t m = mx[memb[k]];
        for (int b = 0; b < NB; ++b) score[b] += (c[b] + 1 > m ? 4096 : 0) + c[b];
      }
      int best = -1;
      for (int b = 0; b < NB; ++b) {
        if (pop[b] >= cap) continue;
        if (best < 0 || score[b] < score[best] || (score[b] == score[best] && pop[b] < pop[best])) best = b;
      }
      bank[v] = best;
      ++pop[best];
      for (size_t k = off[v]; k < off[v + 1]; ++k) {
        uint16_t& c = cnt[(size_t)memb[k] * NB + best];
        ++c;
        if (c > mx[memb[k]]) mx[memb[k]] = c;
      }
    }
    // slots: bank b owns b, b + NB, b + 2 NB, ...
    std::vector<int> next((size_t)NB);
    std::iota(next.begin(), next.end(), 0);
    for (int v = 0; v < V; ++v) { L.slot_of[v] = next[bank[v]]; next[bank[v]] += NB; }
    L.identity = true;
    for (int v = 0; v < V; ++v) if (L.slot_of[v] != v) { L.identity = false; break; }
  }
  // bank model: wavefronts per set (one set = the lanes served together)
  {
    const int NB = opt.nbanks > 0 ? opt.nbanks : (LPS == 16 ? 16 : 32);
    unsigned long long wf = 0;
    std::vector<int> c((size_t)NB);
    for (const auto& s : sets) {
      std::fill(c.begin(), c.end(), 0);
      int m = 0;
      for (int v : s) m = std::max(m, ++c[L.slot_of[v] % NB]);
      wf += (unsigned long long)m;
    }
    L.wavefronts_per_load = sets.empty() ? 0.0 : (double)wf / (double)sets.size();
  }

  // ---- device words ------------------------------------------------------------------------------------------
  // lane l of chunk ch reads its TBC_U words (one per row) with one vector load: they are adjacent
  L.words.assign(lanes.size(), 0);
  uint64_t last_word = 0;
  for (size_t i = 0; i < lanes.size(); ++i) {
    const size_t row = i / 32, lane = i % 32, ch = row / TBC_U, u = row % TBC_U;
    const size_t dst = (ch * 32 + lane) * TBC_U + u;
    if (lanes[i] < 0) { L.words[dst] = last_word; continue; }
    const Item& it = items[lanes[i]];
    uint64_t f0, f1, f2;
    if (loads_x(it.cls)) f0 = (uint64_t)L.slot_of[it.x];
    else if (it.cls == TBC_ADD_XK) f0 = (uint64_t)((uint32_t)pb->lb[it.x] & TBC_FIELD_MASK);
    else f0 = 0;
    f1 = (uint64_t)L.slot_of[it.y];
    if (loads_z(it.cls)) f2 = (uint64_t)L.slot_of[it.z];
    else f2 = (uint64_t)((uint32_t)pb->lb[it.z] & TBC_FIELD_MASK);
    last_word = f0 | (f1 << TBC_FIELD_BITS) | (f2 << (2 * TBC_FIELD_BITS));
    L.words[dst] = last_word;
  }

  // ---- watch lists: which chunks load a slot (the active-set fixpoint re-evaluates exactly those when it moves) ----
  {
    std::vector<std::vector<int>> w((size_t)L.nslots);
    L.chunk_of_prop.assign((size_t)P, -1);
    for (size_t i = 0; i < lanes.size(); ++i) {
      if (lanes[i] < 0) continue;
      const int ch = (int)(i / 32 / TBC_U);
      L.chunk_of_prop[(size_t)lanes[i]] = ch;
      const Item& it = items[lanes[i]];
      auto add = [&](int var) { std::vector<int>& l = w[(size_t)L.slot_of[var]]; if (l.empty() || l.back() != ch) l.push_back(ch); };
      if (loads_x(it.cls)) add(it.x);
      add(it.y);
      if (loads_z(it.cls)) add(it.z);
    }
    L.watch_off.assign((size_t)L.nslots + 1, 0);
    L.watch_list.clear();
    for (int sl = 0; sl < L.nslots; ++sl) {
      std::vector<int>& l = w[(size_t)sl];
      std::sort(l.begin(), l.end());
      l.erase(std::unique(l.begin(), l.end()), l.end());
      L.watch_off[(size_t)sl] = (int)L.watch_list.size();
      L.watch_list.insert(L.watch_list.end(), l.begin(), l.end());
    }
    L.watch_off[(size_t)L.nslots] = (int)L.watch_list.size();
    // one 64-bit word per slot answers most marks with a single load: three watchers inline, a flag for "more"
    L.watch_inline.assign((size_t)L.nslots, ~0ull);
    for (int sl = 0; sl < L.nslots; ++sl) {
      const int b = L.watch_off[(size_t)sl], e = L.watch_off[(size_t)sl + 1];
      uint64_t wd = 0;
      for (int k = 0; k < 3; ++k) wd |= (uint64_t)(b + k < e ? (uint32_t)L.watch_list[(size_t)(b + k)] & 0xFFFFu : 0xFFFFu) << (16 * k);
      wd |= (uint64_t)(e - b > 3 ? 0xFFFEu : 0xFFFFu) << 48;
      L.watch_inline[(size_t)sl] = wd;
    }
  }
  return TB_OK;
}

extern "C" tb_status tb_layout_watch_lists(const tb_problem* pb, int32_t nbanks, int32_t* nslots, int32_t* nentries,
                                            int32_t* off, int32_t* list, int32_t* slot_of, int32_t* chunk_of_prop) {
  if (!pb || nbanks < 0) return TB_ERR_INVALID;
  TnfLayoutOptions lo;
  lo.nbanks = nbanks;
  lo.lanes_per_set = 16;
  lo.slot_align = nbanks > 0 ? nbanks : 4;
  TnfLayout L;
  std::string err;
  tb_status rc = tb_build_layout(pb, lo, &L, &err);
  if (rc != TB_OK) return rc;
  if (nslots) *nslots = L.nslots;
  if (nentries) *nentries = (int32_t)L.watch_list.size();
  if (off) std::copy(L.watch_off.begin(), L.watch_off.end(), off);
  if (list) std::copy(L.watch_list.begin(), L.watch_list.end(), list);
  if (slot_of) std::copy(L.slot_of.begin(), L.slot_of.end(), slot_of);
  if (chunk_of_prop) std::copy(L.chunk_of_prop.begin(), L.chunk_of_prop.end(), chunk_of_prop);
  return TB_OK;
}

extern "C" tb_status tb_layout_describe(const tb_problem* pb, int32_t nbanks, tb_layout_info* info, int32_t* slot_of) {
  if (!pb || !info || nbanks < 0) return TB_ERR_INVALID;
  TnfLayoutOptions lo;
  lo.nbanks = nbanks;
  lo.lanes_per_set = 16;
  lo.slot_align = nbanks > 0 ? nbanks : 4;
  TnfLayout L;
  std::string err;
  tb_status rc = tb_build_layout(pb, lo, &L, &err);
  if (rc != TB_OK) return rc;
  *info = tb_layout_info();
  info->nclasses = TBC_NUM;
  info->nchunks = L.cls_begin[TBC_NUM];
  info->nslots = L.nslots;
  info->identity = L.identity ? 1 : 0;
  for (int c = 0; c < TBC_NUM; ++c) {
    const int n = L.cls_begin[c + 1] - L.cls_begin[c];
    info->class_count[c] = n ? (n - 1) * 32 * TBC_U + L.cls_last[c] : 0;
  }
  info->loads_per_sweep = L.loads_per_sweep;
  info->wavefronts_per_load = L.wavefronts_per_load;
  if (slot_of) std::copy(L.slot_of.begin(), L.slot_of.end(), slot_of);
  return TB_OK;
}

extern "C" tb_status tb_layout_cluster_locality(const tb_problem* pb, int32_t cluster, int32_t warps_per_cta, double* placed, double* striped) {
  if (!pb || cluster < 2 || cluster > 16 || warps_per_cta < 1) return TB_ERR_INVALID;
  std::string err;
  for (int pass = 0; pass < 2; ++pass) {
    TnfLayoutOptions lo;
    lo.slot_align = 4 * cluster;
    TnfLayout L;
    double frac = 0.0;
    if (pass == 0) {
      lo.cluster = cluster; lo.cluster_warps = warps_per_cta;
      tb_status rc = tb_build_layout(pb, lo, &L, &err);
      if (rc != TB_OK) return rc;
      frac = L.cluster_local_fraction;
      if (placed) *placed = frac;
    } else {
      // the striped placement (slot = variable index, slot s in CTA s mod C, chunks in the ternariser's order)
      tb_status rc = tb_build_layout(pb, lo, &L, &err);
      if (rc != TB_OK) return rc;
      unsigned long long local = 0, total = 0;
      const size_t period = (size_t)cluster * (size_t)warps_per_cta;
      for (int i = 0; i < pb->nprops; ++i) {
        const int ch = L.chunk_of_prop[(size_t)i];
        if (ch < 0) continue;
        const int cta = (int)(((size_t)ch % period) / (size_t)warps_per_cta);
        bool sw = false;
        const int cls = tb_classify(pb->props[i], pb->lb, pb->ub, &sw);
        const int y = sw ? pb->props[i].z : pb->props[i].y, z = sw ? pb->props[i].y : pb->props[i].z;
        auto f = [&](int v) { ++total; local += (L.slot_of[(size_t)v] % cluster) == cta; };
        if (loads_x(cls)) f(pb->props[i].x);
        f(y);
        if (loads_z(cls)) f(z);
      }
      if (striped) *striped = total ? (double)local / (double)total : 0.0;
    }
  }
  return TB_OK;
}

extern "C" const char* tb_layout_class_name(int32_t cls) { return tb_class_name(cls); }
