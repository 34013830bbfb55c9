/* tnf_oracle.c — CPU oracle (plain C) for interval propagation over TNF propagators and the
 * dive-and-solve search.  TEST INFRASTRUCTURE ONLY — see tnf_oracle.h for who may use it and for
 * the "parity unpinned" statement.
 *
 * Written independently of the CUDA kernels on purpose (different arithmetic style: everything in
 * int64 with extended infinities, sequential Gauss-Seidel updates) so that agreement between the
 * two is evidence, not a tautology.
 */
#include "tnf_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef long long i64;

#define NINF TB_NEG_INF
#define PINF TB_POS_INF
#define BIG (1LL << 40)

/* ---- extended integers ------------------------------------------------------------------------ */

static i64 ext(int32_t v) { return v == NINF ? -BIG : (v == PINF ? BIG : (i64)v); }
static int32_t clamp32(i64 r) { return r <= (i64)NINF ? NINF : (r >= (i64)PINF ? PINF : (int32_t)r); }
static int finite(int32_t l, int32_t u) { return l != NINF && u != PINF && l != PINF && u != NINF; }
static i64 min64(i64 a, i64 b) { return a < b ? a : b; }
static i64 max64(i64 a, i64 b) { return a > b ? a : b; }
static i64 floordiv(i64 a, i64 b) { i64 q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }
static i64 ceildiv(i64 a, i64 b) { i64 q = a / b, r = a % b; return (r != 0 && ((r < 0) == (b < 0))) ? q + 1 : q; }
static i64 abs64(i64 a) { return a < 0 ? -a : a; }

/* VStore::embed restricted to one bound: monotone meet, reports change, flags emptiness. */
static int tell_lb(int32_t* lb, const int32_t* ub, int v, i64 nl, int32_t* failed) {
  int32_t n = clamp32(nl);
  if (n > lb[v]) { lb[v] = n; if (n > ub[v]) *failed = 1; return 1; }
  return 0;
}
static int tell_ub(const int32_t* lb, int32_t* ub, int v, i64 nu, int32_t* failed) {
  int32_t n = clamp32(nu);
  if (n < ub[v]) { ub[v] = n; if (n < lb[v]) *failed = 1; return 1; }
  return 0;
}

/* ---- operators (our frozen spec of PIR::deduce) ------------------------- */

static int deduce_add(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = 0;
  c |= tell_lb(lb, ub, x, ext(lb[y]) + ext(lb[z]), f);
  c |= tell_ub(lb, ub, x, ext(ub[y]) + ext(ub[z]), f);
  c |= tell_lb(lb, ub, y, ext(lb[x]) - ext(ub[z]), f);
  c |= tell_ub(lb, ub, y, ext(ub[x]) - ext(lb[z]), f);
  c |= tell_lb(lb, ub, z, ext(lb[x]) - ext(ub[y]), f);
  c |= tell_ub(lb, ub, z, ext(ub[x]) - ext(lb[y]), f);
  return c;
}

/* y <- y ∩ (x / z) for x = y * z when 0 is not in z (real hull of the quotient, rounded inward). */
static int mul_back(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  if (!finite(lb[x], ub[x]) || !finite(lb[z], ub[z])) return 0;
  if (lb[x] > ub[x] || lb[z] > ub[z]) return 0;        /* already failed: nothing canonical to do */
  if (!(lb[z] > 0 || ub[z] < 0)) return 0;
  i64 xs[2] = { lb[x], ub[x] }, zs[2] = { lb[z], ub[z] };
  i64 lo = BIG, hi = -BIG;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
    lo = min64(lo, ceildiv(xs[i], zs[j]));
    hi = max64(hi, floordiv(xs[i], zs[j]));
  }
  int c = 0;
  c |= tell_lb(lb, ub, y, lo, f);
  c |= tell_ub(lb, ub, y, hi, f);
  return c;
}

static int off_zero(int v, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = 0;
  if (lb[v] == 0) c |= tell_lb(lb, ub, v, 1, f);
  if (ub[v] == 0) c |= tell_ub(lb, ub, v, -1, f);
  return c;
}

static int deduce_mul(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = 0;
  if (finite(lb[y], ub[y]) && finite(lb[z], ub[z])) {
    i64 p[4] = { (i64)lb[y] * lb[z], (i64)lb[y] * ub[z], (i64)ub[y] * lb[z], (i64)ub[y] * ub[z] };
    i64 lo = p[0], hi = p[0];
    for (int i = 1; i < 4; ++i) { lo = min64(lo, p[i]); hi = max64(hi, p[i]); }
    c |= tell_lb(lb, ub, x, lo, f);
    c |= tell_ub(lb, ub, x, hi, f);
  }
  if (lb[x] > 0 || ub[x] < 0) {       /* product is non-zero: neither factor is zero */
    c |= off_zero(y, lb, ub, f);
    c |= off_zero(z, lb, ub, f);
  }
  c |= mul_back(x, y, z, lb, ub, f);
  c |= mul_back(x, z, y, lb, ub, f);
  return c;
}

static i64 tdiv(i64 a, i64 b) { return a / b; }   /* C division truncates */

static int deduce_tdiv(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = off_zero(z, lb, ub, f);                 /* z != 0 */
  if (finite(lb[y], ub[y]) && finite(lb[z], ub[z])) {
    i64 zs[4]; int nz = 0;
    if (lb[z] < 0) { zs[nz++] = lb[z]; zs[nz++] = min64(ub[z], -1); }
    if (ub[z] > 0) { zs[nz++] = max64(lb[z], 1); zs[nz++] = ub[z]; }
    if (nz > 0) {
      i64 lo = BIG, hi = -BIG, ys[2] = { lb[y], ub[y] };
      for (int i = 0; i < 2; ++i) for (int j = 0; j < nz; ++j) {
        i64 q = tdiv(ys[i], zs[j]); lo = min64(lo, q); hi = max64(hi, q);
      }
      c |= tell_lb(lb, ub, x, lo, f);
      c |= tell_ub(lb, ub, x, hi, f);
    }
  }
  if (finite(lb[x], ub[x]) && finite(lb[z], ub[z])) {   /* y = x*z + r, |r| <= |z|-1 */
    i64 m = max64(abs64(lb[z]), abs64(ub[z])) - 1;
    if (m < 0) m = 0;
    i64 p[4] = { (i64)lb[x] * lb[z], (i64)lb[x] * ub[z], (i64)ub[x] * lb[z], (i64)ub[x] * ub[z] };
    i64 lo = p[0], hi = p[0];
    for (int i = 1; i < 4; ++i) { lo = min64(lo, p[i]); hi = max64(hi, p[i]); }
    c |= tell_lb(lb, ub, y, lo - m, f);
    c |= tell_ub(lb, ub, y, hi + m, f);
  }
  return c;
}

static int deduce_tmod(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = off_zero(z, lb, ub, f);
  if (finite(lb[z], ub[z])) {
    i64 m = max64(abs64(lb[z]), abs64(ub[z])) - 1;
    if (m < 0) m = 0;
    c |= tell_lb(lb, ub, x, -m, f);
    c |= tell_ub(lb, ub, x, m, f);
  }
  if (lb[y] >= 0) { c |= tell_lb(lb, ub, x, 0, f); c |= tell_ub(lb, ub, x, ext(ub[y]), f); }
  if (ub[y] <= 0) { c |= tell_ub(lb, ub, x, 0, f); c |= tell_lb(lb, ub, x, ext(lb[y]), f); }
  if (finite(lb[y], ub[y]) && finite(lb[z], ub[z]) && lb[y] == ub[y] && lb[z] == ub[z] && lb[z] != 0) {
    i64 r = (i64)lb[y] % (i64)lb[z];
    c |= tell_lb(lb, ub, x, r, f);
    c |= tell_ub(lb, ub, x, r, f);
  }
  return c;
}

static int deduce_min(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = 0;
  c |= tell_lb(lb, ub, x, lb[y] < lb[z] ? lb[y] : lb[z], f);
  c |= tell_ub(lb, ub, x, ub[y] < ub[z] ? ub[y] : ub[z], f);
  c |= tell_lb(lb, ub, y, lb[x], f);
  c |= tell_lb(lb, ub, z, lb[x], f);
  if (lb[y] > ub[x]) c |= tell_ub(lb, ub, z, ub[x], f);
  if (lb[z] > ub[x]) c |= tell_ub(lb, ub, y, ub[x], f);
  return c;
}

static int deduce_max(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = 0;
  c |= tell_lb(lb, ub, x, lb[y] > lb[z] ? lb[y] : lb[z], f);
  c |= tell_ub(lb, ub, x, ub[y] > ub[z] ? ub[y] : ub[z], f);
  c |= tell_ub(lb, ub, y, ub[x], f);
  c |= tell_ub(lb, ub, z, ub[x], f);
  if (ub[y] < lb[x]) c |= tell_lb(lb, ub, z, lb[x], f);
  if (ub[z] < lb[x]) c |= tell_lb(lb, ub, y, lb[x], f);
  return c;
}

/* remove value k from the bounds of v (k finite) */
static int not_value(int v, int32_t k, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = 0;
  if (lb[v] == k) c |= tell_lb(lb, ub, v, (i64)k + 1, f);
  if (ub[v] == k) c |= tell_ub(lb, ub, v, (i64)k - 1, f);
  return c;
}

static int deduce_eq(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = 0;
  if (lb[x] >= 1) {                      /* y = z */
    c |= tell_lb(lb, ub, y, lb[z], f);
    c |= tell_ub(lb, ub, y, ub[z], f);
    c |= tell_lb(lb, ub, z, lb[y], f);
    c |= tell_ub(lb, ub, z, ub[y], f);
  }
  else if (ub[x] <= 0) {                 /* y != z */
    if (lb[y] == ub[y] && finite(lb[y], ub[y])) c |= not_value(z, lb[y], lb, ub, f);
    if (lb[z] == ub[z] && finite(lb[z], ub[z])) c |= not_value(y, lb[z], lb, ub, f);
  }
  else {
    if (ub[y] < lb[z] || ub[z] < lb[y]) c |= tell_ub(lb, ub, x, 0, f);
    else if (lb[y] == ub[y] && lb[z] == ub[z] && lb[y] == lb[z]) c |= tell_lb(lb, ub, x, 1, f);
  }
  return c;
}

static int deduce_leq(int x, int y, int z, int32_t* lb, int32_t* ub, int32_t* f) {
  int c = 0;
  if (lb[x] >= 1) {                      /* y <= z */
    c |= tell_ub(lb, ub, y, ub[z], f);
    c |= tell_lb(lb, ub, z, lb[y], f);
  }
  else if (ub[x] <= 0) {                 /* y > z */
    c |= tell_lb(lb, ub, y, ext(lb[z]) + 1, f);
    c |= tell_ub(lb, ub, z, ext(ub[y]) - 1, f);
  }
  else {
    if (ub[y] <= lb[z]) c |= tell_lb(lb, ub, x, 1, f);
    else if (lb[y] > ub[z]) c |= tell_ub(lb, ub, x, 0, f);
  }
  return c;
}

int tbo_deduce(const tb_prop* p, int32_t* lb, int32_t* ub, int32_t* failed) {
  int x = p->x, y = p->y, z = p->z;
  if (lb[x] > ub[x] || lb[y] > ub[y] || lb[z] > ub[z]) { *failed = 1; return 0; }
  switch (p->op) {
    case TB_OP_ADD:  return deduce_add(x, y, z, lb, ub, failed);
    case TB_OP_MUL:  return deduce_mul(x, y, z, lb, ub, failed);
    case TB_OP_TDIV: return deduce_tdiv(x, y, z, lb, ub, failed);
    case TB_OP_TMOD: return deduce_tmod(x, y, z, lb, ub, failed);
    case TB_OP_MIN:  return deduce_min(x, y, z, lb, ub, failed);
    case TB_OP_MAX:  return deduce_max(x, y, z, lb, ub, failed);
    case TB_OP_EQ:   return deduce_eq(x, y, z, lb, ub, failed);
    case TB_OP_LEQ:  return deduce_leq(x, y, z, lb, ub, failed);
    default: return 0;
  }
}

int tbo_ask(const tb_prop* p, const int32_t* lb, const int32_t* ub) {
  int x = p->x, y = p->y, z = p->z;
  switch (p->op) {
    case TB_OP_EQ:
      if (lb[x] >= 1) return lb[y] == ub[y] && lb[z] == ub[z] && lb[y] == lb[z];
      if (ub[x] <= 0) return ub[y] < lb[z] || ub[z] < lb[y];
      return 0;
    case TB_OP_LEQ:
      if (lb[x] >= 1) return ub[y] <= lb[z];
      if (ub[x] <= 0) return lb[y] > ub[z];
      return 0;
    default:
      return lb[x] == ub[x] && lb[y] == ub[y] && lb[z] == ub[z];
  }
}

int64_t tbo_fixpoint(const tb_problem* pb, int32_t* lb, int32_t* ub, int32_t* failed,
                     const int32_t* order, uint64_t* num_deductions) {
  int64_t sweeps = 0;
  int changed = 1;
  *failed = 0;
  while (changed && !*failed) {
    changed = 0;
    for (int i = 0; i < pb->nprops; ++i) {
      const tb_prop* p = &pb->props[order ? order[i] : i];
      changed |= tbo_deduce(p, lb, ub, failed);
    }
    ++sweeps;
  }
  if (num_deductions) *num_deductions += (uint64_t)sweeps * (uint64_t)pb->nprops;
  return sweeps;
}

/* ---- search ------------------------------------------------------------------------------------ */

typedef struct {            /* LightBranch<Itv>, barebones_dive_and_solve.hpp:135,358-393 */
  int32_t var;
  int32_t clb[2], cub[2];
  int32_t ropes[2];
  int32_t current_idx;
} decision_t;

typedef struct {
  _Atomic uint64_t next_subproblem;   /* GridData::next_subproblem (:418) */
  _Atomic int32_t appx_best_bound;    /* GridData::appx_best_bound (:426) */
  _Atomic int32_t stop;               /* UnifiedData::stop (:64) */
  volatile int32_t* user_stop;
  struct timespec t0;
  uint64_t timeout_ms;
  uint64_t rank, world;               /* this worker group's shard: idx = k * world + rank */
} shared_t;

typedef struct {
  const tb_problem* pb;
  shared_t* sh;
  int32_t depth_power;
  uint64_t cutnodes;
  int32_t *lb, *ub, *root_lb, *root_ub, *best_lb, *best_ub;
  decision_t* dec;
  int32_t max_depth;
  int32_t depth, cur_strategy, next_unassigned, snap_strategy, snap_next_unassigned;
  int32_t best_bound;
  int stop, leaf, failed_node, overflow;
  tb_stats st;
  uint64_t first_idx;
  int64_t best_time_ns;
} block_t;

static int64_t ns_since(const struct timespec* t0) {
  struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
  return (int64_t)(t.tv_sec - t0->tv_sec) * 1000000000LL + (t.tv_nsec - t0->tv_nsec);
}

static int embed(block_t* b, int v, int32_t l, int32_t u) {
  int c = 0;
  if (l > b->lb[v]) { b->lb[v] = l; c = 1; }
  if (u < b->ub[v]) { b->ub[v] = u; c = 1; }
  return c;
}

static int splittable(const block_t* b, int v) {
  return b->lb[v] != b->ub[v] && b->lb[v] != NINF && b->ub[v] != PINF;
}

/* BlockData::push_decision, barebones_dive_and_solve.hpp:355-405 */
static void push_decision(block_t* b, int val_order, int var) {
  if (b->depth >= b->max_depth) { b->overflow = 1; return; }
  decision_t* d = &b->dec[b->depth];
  int32_t l = b->lb[var], u = b->ub[var];
  d->var = var; d->current_idx = -1;
  int32_t mid = (int32_t)((i64)l + ((i64)u - (i64)l) / 2);
  switch (val_order) {
    case TB_VAL_MIN:   d->clb[0] = l; d->cub[0] = l; d->clb[1] = l + 1; d->cub[1] = u; break;
    case TB_VAL_MAX:   d->clb[0] = u; d->cub[0] = u; d->clb[1] = l; d->cub[1] = u - 1; break;
    case TB_VAL_SPLIT: d->clb[0] = l; d->cub[0] = mid; d->clb[1] = mid + 1; d->cub[1] = u; break;
    default:           d->clb[0] = mid + 1; d->cub[0] = u; d->clb[1] = l; d->cub[1] = mid; break;
  }
  d->ropes[0] = b->depth + 1;
  d->ropes[1] = b->depth > 0 ? b->dec[b->depth - 1].ropes[b->dec[b->depth - 1].current_idx] : -1;
  ++b->depth;
}

/* BlockData::split + input_order_split + lattice_smallest_split (:187-349), sequential reading:
 * scan the strategy's variables from the cursor, lowest index wins ties. Returns 1 if pushed. */
static int split(block_t* b) {
  const tb_problem* pb = b->pb;
  for (int s = b->cur_strategy; s < pb->nstrategies; ++s) {
    const tb_strategy* st = &pb->strategies[s];
    int in_store = st->n == 0;
    int n = in_store ? pb->nvars : st->n;
    int first = -1, best = -1;
    i64 best_val = 0;
    for (int i = b->next_unassigned; i < n; ++i) {
      int v = in_store ? i : st->vars[i];
      if (!splittable(b, v)) continue;
      if (first < 0) first = i;
      if (st->var_order == TB_VAR_INPUT_ORDER) { best = i; break; }
      i64 val;
      switch (st->var_order) {
        case TB_VAR_FIRST_FAIL:      val =  (i64)(uint32_t)((uint32_t)b->ub[v] - (uint32_t)b->lb[v]); break;
        case TB_VAR_ANTI_FIRST_FAIL: val = -(i64)(uint32_t)((uint32_t)b->ub[v] - (uint32_t)b->lb[v]); break;
        case TB_VAR_SMALLEST:        val =  (i64)b->lb[v]; break;
        default:                     val = -(i64)b->ub[v]; break;   /* LARGEST */
      }
      if (best < 0 || val < best_val) { best = i; best_val = val; }
    }
    if (best >= 0) {
      b->next_unassigned = first;
      push_decision(b, st->val_order, in_store ? best : st->vars[best]);
      return !b->overflow;
    }
    b->cur_strategy = s + 1;
    b->next_unassigned = 0;
  }
  return 0;
}

static void check_stop(block_t* b) {
  shared_t* sh = b->sh;
  if (atomic_load(&sh->stop)) { b->stop = 1; return; }
  int must = 0;
  if (sh->user_stop && *sh->user_stop) must = 1;
  if (sh->timeout_ms && (b->st.nodes & 63) == 0 && ns_since(&sh->t0) / 1000000 >= (int64_t)sh->timeout_ms) must = 1;
  if (must) { atomic_store(&sh->stop, 1); b->stop = 1; }
}

/* propagate(), barebones_dive_and_solve.hpp:903-1031 (AC1 accounting). */
static void propagate(block_t* b) {
  const tb_problem* pb = b->pb;
  int32_t failed = 0;
  b->leaf = 0;
  /* The only embed outside deduce that can empty an interval is the objective-bound injection
   * (VStore::embed sets the sticky bot flag there, :761-764); such a node fails with 0 sweeps. */
  int64_t it = 0;
  if (pb->obj_var >= 0 && b->lb[pb->obj_var] > b->ub[pb->obj_var]) failed = 1;
  else it = tbo_fixpoint(pb, b->lb, b->ub, &failed, NULL, &b->st.num_deductions);
  if (!failed) {
    int all = 1;
    for (int i = 0; i < pb->nprops && all; ++i) all = tbo_ask(&pb->props[i], b->lb, b->ub);
    if (all) {
      b->leaf = 1;
      if (pb->obj_var >= 0) {
        if (b->best_bound > b->lb[pb->obj_var]) {
          b->best_bound = b->lb[pb->obj_var];
          int32_t cur = atomic_load(&b->sh->appx_best_bound);
          while (b->best_bound < cur && !atomic_compare_exchange_weak(&b->sh->appx_best_bound, &cur, b->best_bound)) {}
          b->best_time_ns = ns_since(&b->sh->t0);
          memcpy(b->best_lb, b->lb, sizeof(int32_t) * pb->nvars);
          memcpy(b->best_ub, b->ub, sizeof(int32_t) * pb->nvars);
          b->st.solutions++;
        }
      }
      else {                       /* satisfaction: first solution wins, then everybody stops */
        if (b->st.solutions == 0) {
          b->best_time_ns = ns_since(&b->sh->t0);
          memcpy(b->best_lb, b->lb, sizeof(int32_t) * pb->nvars);
          memcpy(b->best_ub, b->ub, sizeof(int32_t) * pb->nvars);
        }
        b->st.solutions++;
        b->st.exhaustive = 0;
        atomic_store(&b->sh->stop, 1);
        b->stop = 1;
      }
    }
  }
  else b->leaf = 1;
  b->failed_node = failed;
  b->st.fixpoint_iterations += (uint64_t)it;
  b->st.nodes++;
  b->st.fails += failed ? 1 : 0;
  if (b->depth > b->st.depth_max) b->st.depth_max = b->depth;
  if (b->cutnodes && b->st.nodes >= b->cutnodes) { b->st.exhaustive = 0; b->stop = 1; }
  check_stop(b);
  if (b->stop && atomic_load(&b->sh->stop)) b->st.exhaustive = 0;
}

static void copy_store(int32_t* dl, int32_t* du, const int32_t* sl, const int32_t* su, int n) {
  memcpy(dl, sl, sizeof(int32_t) * n); memcpy(du, su, sizeof(int32_t) * n);
}

/* steps C-D of gpu_barebones_solve (:663-714). Returns remaining depth. */
static int dive(block_t* b, uint64_t idx) {
  const tb_problem* pb = b->pb;
  b->cur_strategy = 0; b->next_unassigned = 0; b->depth = 0;
  copy_store(b->lb, b->ub, pb->lb, pb->ub, pb->nvars);
  int remaining = b->depth_power;
  b->leaf = 0;
  while (remaining > 0 && !b->leaf && !b->stop) {
    propagate(b);
    if (!b->leaf) {
      if (!split(b)) { b->leaf = 1; b->st.exhaustive = 0; }
      else {
        --remaining; --b->depth;
        int bit = (int)((idx >> remaining) & 1ULL);
        embed(b, b->dec[0].var, b->dec[0].clb[bit], b->dec[0].cub[bit]);
      }
    }
  }
  return remaining;
}

/* step F of gpu_barebones_solve (:742-871). */
static void solve_subproblem(block_t* b) {
  const tb_problem* pb = b->pb;
  shared_t* sh = b->sh;
  if (pb->has_eps_strategy) { if (b->cur_strategy < 1) b->cur_strategy = 1; b->next_unassigned = 0; }
  while (!b->stop) {
    if (pb->obj_var >= 0) {
      int32_t appx = atomic_load(&sh->appx_best_bound);
      if (appx != PINF) {
        embed(b, pb->obj_var, NINF, clamp32(ext(appx) - 1));
        embed(b, pb->obj_var, NINF, clamp32(ext(b->best_bound) - 1));
      }
      if (appx == NINF) { b->stop = 1; atomic_store(&sh->stop, 1); break; }
    }
    propagate(b);
    if (!b->leaf) {
      if (b->depth == 0) {
        copy_store(b->root_lb, b->root_ub, b->lb, b->ub, pb->nvars);
        b->snap_strategy = b->cur_strategy; b->snap_next_unassigned = b->next_unassigned;
      }
      if (!split(b)) { b->leaf = 1; b->st.exhaustive = 0; }
      else {
        decision_t* d = &b->dec[b->depth - 1];
        ++d->current_idx;
        embed(b, d->var, d->clb[d->current_idx], d->cub[d->current_idx]);
      }
    }
    if (b->leaf) {
      if (b->depth == 0) break;
      b->depth = b->dec[b->depth - 1].ropes[b->dec[b->depth - 1].current_idx];
      if (b->depth == -1) break;
      copy_store(b->lb, b->ub, b->root_lb, b->root_ub, pb->nvars);
      for (int i = 0; i < b->depth - 1; ++i)
        embed(b, b->dec[i].var, b->dec[i].clb[b->dec[i].current_idx], b->dec[i].cub[b->dec[i].current_idx]);
      decision_t* d = &b->dec[b->depth - 1];
      ++d->current_idx;
      embed(b, d->var, d->clb[d->current_idx], d->cub[d->current_idx]);
      b->cur_strategy = b->snap_strategy; b->next_unassigned = b->snap_next_unassigned;
    }
  }
}

static int block_init(block_t* b, const tb_problem* pb, shared_t* sh, int32_t depth_power, uint64_t cutnodes) {
  memset(b, 0, sizeof(*b));
  b->pb = pb; b->sh = sh; b->depth_power = depth_power; b->cutnodes = cutnodes;
  size_t n = (size_t)(pb->nvars > 0 ? pb->nvars : 1) * sizeof(int32_t);
  b->lb = malloc(n); b->ub = malloc(n); b->root_lb = malloc(n); b->root_ub = malloc(n);
  b->best_lb = malloc(n); b->best_ub = malloc(n);
  b->max_depth = 10000;     /* MAX_SEARCH_DEPTH, barebones_dive_and_solve.hpp:14 */
  b->dec = malloc(sizeof(decision_t) * (size_t)b->max_depth);
  b->best_bound = PINF;
  b->st.exhaustive = 1;
  return b->lb && b->ub && b->root_lb && b->root_ub && b->best_lb && b->best_ub && b->dec;
}

static void block_free(block_t* b) {
  free(b->lb); free(b->ub); free(b->root_lb); free(b->root_ub); free(b->best_lb); free(b->best_ub); free(b->dec);
}

/* main loop B-G of gpu_barebones_solve (:656-886) for one worker */
static void* block_run(void* arg) {
  block_t* b = (block_t*)arg;
  shared_t* sh = b->sh;
  uint64_t nsub = 1ULL << b->depth_power;
  uint64_t k = b->first_idx;
  for (;;) {
    uint64_t idx = k * sh->world + sh->rank;
    if (idx >= nsub || b->stop) break;
    int remaining = dive(b, idx);
    if (b->leaf && !b->stop) {
      uint64_t next = ((idx >> remaining) + 1ULL) << remaining;
      uint64_t next_k = next <= sh->rank ? 0 : (next - sh->rank + sh->world - 1) / sh->world;
      uint64_t cur = atomic_load(&sh->next_subproblem);
      while (cur < next_k && !atomic_compare_exchange_weak(&sh->next_subproblem, &cur, next_k)) {}
      if ((idx & ((1ULL << remaining) - 1ULL)) == 0) b->st.eps_skipped_subproblems += next - idx;
    }
    else if (!b->stop) {
      solve_subproblem(b);
      if (!(b->cutnodes && b->st.nodes >= b->cutnodes) && !atomic_load(&sh->stop)) b->st.eps_solved_subproblems++;
    }
    if (b->overflow) { b->st.exhaustive = 0; b->stop = 1; }
    if (!b->stop) k = atomic_fetch_add(&sh->next_subproblem, 1);
  }
  if (!(b->cutnodes && b->st.nodes >= b->cutnodes) && !atomic_load(&sh->stop)) b->st.num_blocks_done = 1;
  b->st.cumulative_time_block_ns = ns_since(&sh->t0);
  return NULL;
}

int tbo_dive(const tb_problem* pb, uint64_t idx, int32_t depth,
             int32_t* lb_out, int32_t* ub_out, int32_t* remaining_depth, int32_t* leaf_kind) {
  shared_t sh; memset(&sh, 0, sizeof(sh));
  atomic_store(&sh.appx_best_bound, PINF);
  sh.world = 1;
  clock_gettime(CLOCK_MONOTONIC, &sh.t0);
  block_t b;
  if (!block_init(&b, pb, &sh, depth, 0)) return TB_ERR_NOMEM;
  int remaining = dive(&b, idx);
  copy_store(lb_out, ub_out, b.lb, b.ub, pb->nvars);
  *remaining_depth = remaining;
  *leaf_kind = b.leaf ? (b.failed_node ? 1 : 2) : 0;
  block_free(&b);
  return TB_OK;
}

int tbo_solve(const tb_problem* pb, int32_t depth, uint64_t cutnodes, uint64_t timeout_ms,
              int32_t nthreads, volatile int32_t* stop_flag,
              int32_t* best_lb, int32_t* best_ub, int32_t* has_solution, int32_t* exhaustive,
              tb_stats* stats) {
  return tbo_solve_shard(pb, depth, cutnodes, timeout_ms, nthreads, 0, 1, TB_POS_INF, stop_flag,
                         best_lb, best_ub, has_solution, exhaustive, stats);
}

int tbo_solve_shard(const tb_problem* pb, int32_t depth, uint64_t cutnodes, uint64_t timeout_ms,
                    int32_t nthreads, int32_t rank, int32_t world, int32_t initial_bound,
                    volatile int32_t* stop_flag,
                    int32_t* best_lb, int32_t* best_ub, int32_t* has_solution, int32_t* exhaustive,
                    tb_stats* stats) {
  if (nthreads < 1) nthreads = 1;
  if (world < 1 || rank < 0 || rank >= world) return TB_ERR_INVALID;
  if (depth < 0) depth = 0;
  shared_t sh; memset(&sh, 0, sizeof(sh));
  atomic_store(&sh.appx_best_bound, initial_bound);
  atomic_store(&sh.next_subproblem, (uint64_t)nthreads);
  sh.rank = (uint64_t)rank; sh.world = (uint64_t)world;
  sh.user_stop = stop_flag; sh.timeout_ms = timeout_ms;
  clock_gettime(CLOCK_MONOTONIC, &sh.t0);
  block_t* blocks = calloc((size_t)nthreads, sizeof(block_t));
  pthread_t* th = calloc((size_t)nthreads, sizeof(pthread_t));
  if (!blocks || !th) return TB_ERR_NOMEM;
  for (int i = 0; i < nthreads; ++i) {
    if (!block_init(&blocks[i], pb, &sh, depth, cutnodes)) return TB_ERR_NOMEM;
    blocks[i].first_idx = (uint64_t)i;
  }
  if (nthreads == 1) block_run(&blocks[0]);
  else {
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, block_run, &blocks[i]);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
  }
  /* reduce_blocks, barebones_dive_and_solve.hpp:1033-1067 */
  tb_stats total; memset(&total, 0, sizeof(total));
  total.exhaustive = 1; total.num_blocks = nthreads; total.threads_per_block = 1;
  total.subproblems_power = depth; total.eps_num_subproblems = 1ULL << depth;
  int best_block = -1; int32_t best = PINF; int64_t best_time = 0;
  for (int i = 0; i < nthreads; ++i) {
    const tb_stats* s = &blocks[i].st;
    total.nodes += s->nodes; total.fails += s->fails; total.solutions += s->solutions;
    if (s->depth_max > total.depth_max) total.depth_max = s->depth_max;
    total.exhaustive = total.exhaustive && s->exhaustive;
    total.eps_solved_subproblems += s->eps_solved_subproblems;
    total.eps_skipped_subproblems += s->eps_skipped_subproblems;
    total.num_blocks_done += s->num_blocks_done;
    total.fixpoint_iterations += s->fixpoint_iterations;
    total.num_deductions += s->num_deductions;
    total.cumulative_time_block_ns += s->cumulative_time_block_ns;
    if (s->solutions > 0) {
      if (pb->obj_var < 0) { if (best_block < 0) { best_block = i; best_time = blocks[i].best_time_ns; } }
      else if (blocks[i].best_bound < best || (blocks[i].best_bound == best && blocks[i].best_time_ns <= best_time)) {
        best = blocks[i].best_bound; best_block = i; best_time = blocks[i].best_time_ns;
      }
    }
  }
  total.timers_ns[TB_TIMER_OVERALL] = ns_since(&sh.t0);
  total.timers_ns[TB_TIMER_LATEST_BEST_OBJ_FOUND] = best_block >= 0 ? best_time : 0;
  *has_solution = best_block >= 0;
  if (best_block >= 0) copy_store(best_lb, best_ub, blocks[best_block].best_lb, blocks[best_block].best_ub, pb->nvars);
  *exhaustive = total.exhaustive;
  if (stats) *stats = total;
  for (int i = 0; i < nthreads; ++i) block_free(&blocks[i]);
  free(blocks); free(th);
  return TB_OK;
}
