// tnf_device.cuh — device-side interval narrowing for one TNF propagator  x = y op z.
//
// B200-native replacement of PIR::deduce / PIR::ask (called at reference
// include/barebones_dive_and_solve.hpp:931,944,977; bodies live in lala-pc, not in the tree).
//
// The host layout pass (layout.cpp) sorts the propagator table into *classes* (tnf_classes.h): one
// class = one operator + which operands are root constants + whether 32-bit arithmetic is exact.
// A warp always evaluates 32 propagators of the same class, so the operator is a compile-time
// constant here: no dispatch, no divergence, and operands that are constants cost no load.
//
// One thread evaluates one propagator on the bounds it loaded (Jacobi style: every candidate is
// computed from the loaded snapshot) and publishes only the bounds that moved with shared-memory
// atomicMax/atomicMin.  All rules are monotone and contracting, so the greatest fixpoint is schedule
// independent and agrees bit for bit with the sequential CPU oracle.
//
// Failure detection: see `emptied`.
#pragma once
#include <stdint.h>
#include "../../include/turbo_b200.h"
#include "tnf_classes.h"

#define TBD_NINF INT32_MIN
#define TBD_PINF INT32_MAX

#define F_CHANGED 1
#define F_FAILED 2
#define F_NOT_ENTAILED 4

namespace tbd {

struct Cand {          // candidate bounds for x, y, z (initialised to "no information")
  int xl, xu, yl, yu, zl, zu;
};
struct Snap {          // the bounds one evaluation worked on (constants appear as singletons)
  int xl, xu, yl, yu, zl, zu;
};

__device__ __forceinline__ long long ext64(int v) {
  long long e = v;
  if (v == TBD_NINF) e = -(1LL << 40);
  if (v == TBD_PINF) e = (1LL << 40);
  return e;
}
__device__ __forceinline__ int clamp64(long long r) {
  return r <= (long long)TBD_NINF ? TBD_NINF : (r >= (long long)TBD_PINF ? TBD_PINF : (int)r);
}
__device__ __forceinline__ bool fin(int l, int u) {
  return l != TBD_NINF && u != TBD_PINF && l != TBD_PINF && u != TBD_NINF;
}
// v + 1 / v - 1 on extended integers (infinities absorb)
__device__ __forceinline__ int succ(int v) { return v + (int)(v != TBD_NINF && v != TBD_PINF); }
__device__ __forceinline__ int pred(int v) { return v - (int)(v != TBD_NINF && v != TBD_PINF); }

__device__ __forceinline__ long long fdiv64(long long a, long long b) {
  long long q = a / b, r = a - q * b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
__device__ __forceinline__ long long cdiv64(long long a, long long b) {
  long long q = a / b, r = a - q * b;
  return (r != 0 && ((r < 0) == (b < 0))) ? q + 1 : q;
}
__device__ __forceinline__ long long min4(long long a, long long b, long long c, long long d) {
  return min(min(a, b), min(c, d));
}
__device__ __forceinline__ long long max4(long long a, long long b, long long c, long long d) {
  return max(max(a, b), max(c, d));
}

// ---- x = y + z on extended integers --------------------------------------------------------------
__device__ __forceinline__ void add_ext(int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
  long long exl = ext64(xl), exu = ext64(xu), eyl = ext64(yl), eyu = ext64(yu), ezl = ext64(zl), ezu = ext64(zu);
  c.xl = clamp64(eyl + ezl); c.xu = clamp64(eyu + ezu);
  c.yl = clamp64(exl - ezu); c.yu = clamp64(exu - ezl);
  c.zl = clamp64(exl - eyu); c.zu = clamp64(exu - eyl);
}

// quotient hull for  f = p / d  (d without 0, everything finite), rounded inward
__device__ __forceinline__ void quot(int pl, int pu, int dl, int du, int& ql, int& qu) {
  long long a = pl, b = pu, c = dl, d = du;
  ql = clamp64(min4(cdiv64(a, c), cdiv64(a, d), cdiv64(b, c), cdiv64(b, d)));
  qu = clamp64(max4(fdiv64(a, c), fdiv64(a, d), fdiv64(b, c), fdiv64(b, d)));
}

// ---- x = y * z -------------------------------------------------------------------------------------
__device__ __forceinline__ void mul(int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
  if (fin(yl, yu) && fin(zl, zu)) {
    long long a = (long long)yl * zl, b = (long long)yl * zu, d = (long long)yu * zl, e = (long long)yu * zu;
    c.xl = clamp64(min4(a, b, d, e));
    c.xu = clamp64(max4(a, b, d, e));
  }
  // The sequential oracle narrows x first; a product that is non-zero forbids zero factors.
  int nxl = max(xl, c.xl), nxu = min(xu, c.xu);
  if (nxl > nxu) return;                          // failed on x: the caller flags it
  if (nxl > 0 || nxu < 0) {
    if (yl == 0) c.yl = 1;
    if (yu == 0) c.yu = -1;
    if (zl == 0) c.zl = 1;
    if (zu == 0) c.zu = -1;
  }
  int nyl = max(yl, c.yl), nyu = min(yu, c.yu), nzl = max(zl, c.zl), nzu = min(zu, c.zu);
  if (nyl > nyu || nzl > nzu) return;
  if (fin(nxl, nxu) && fin(nzl, nzu) && (nzl > 0 || nzu < 0)) {
    int ql, qu; quot(nxl, nxu, nzl, nzu, ql, qu);
    c.yl = max(c.yl, ql); c.yu = min(c.yu, qu);
    nyl = max(nyl, ql); nyu = min(nyu, qu);
    if (nyl > nyu) return;
  }
  if (fin(nxl, nxu) && fin(nyl, nyu) && (nyl > 0 || nyu < 0)) {
    int ql, qu; quot(nxl, nxu, nyl, nyu, ql, qu);
    c.zl = max(c.zl, ql); c.zu = min(c.zu, qu);
  }
}

// ---- x = y tdiv z, z != 0 ----------------------------------------------------------------------------
__device__ __forceinline__ void tdiv(int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
  if (zl == 0) { c.zl = 1; zl = 1; }
  if (zu == 0) { c.zu = -1; zu = -1; }
  if (zl > zu) return;
  if (fin(yl, yu) && fin(zl, zu)) {
    long long lo = (1LL << 40), hi = -(1LL << 40);
    long long ys0 = yl, ys1 = yu;
    if (zl < 0) {
      long long d0 = zl, d1 = min(zu, -1);
      lo = min(lo, min4(ys0 / d0, ys0 / d1, ys1 / d0, ys1 / d1));
      hi = max(hi, max4(ys0 / d0, ys0 / d1, ys1 / d0, ys1 / d1));
    }
    if (zu > 0) {
      long long d0 = max(zl, 1), d1 = zu;
      lo = min(lo, min4(ys0 / d0, ys0 / d1, ys1 / d0, ys1 / d1));
      hi = max(hi, max4(ys0 / d0, ys0 / d1, ys1 / d0, ys1 / d1));
    }
    c.xl = clamp64(lo); c.xu = clamp64(hi);
  }
  int nxl = max(xl, c.xl), nxu = min(xu, c.xu);
  if (nxl > nxu) return;
  if (fin(nxl, nxu) && fin(zl, zu)) {
    long long m = max(llabs((long long)zl), llabs((long long)zu)) - 1;
    if (m < 0) m = 0;
    long long a = (long long)nxl * zl, b = (long long)nxl * zu, d = (long long)nxu * zl, e = (long long)nxu * zu;
    c.yl = clamp64(min4(a, b, d, e) - m);
    c.yu = clamp64(max4(a, b, d, e) + m);
  }
}

// ---- x = y tmod z, z != 0 ----------------------------------------------------------------------------
__device__ __forceinline__ void tmod(int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
  (void)xl; (void)xu;
  if (zl == 0) { c.zl = 1; zl = 1; }
  if (zu == 0) { c.zu = -1; zu = -1; }
  if (zl > zu) return;
  int lo = TBD_NINF, hi = TBD_PINF;
  if (fin(zl, zu)) {
    long long m = max(llabs((long long)zl), llabs((long long)zu)) - 1;
    if (m < 0) m = 0;
    lo = clamp64(-m); hi = clamp64(m);
  }
  if (yl >= 0) { lo = max(lo, 0); hi = min(hi, yu); }
  if (yu <= 0) { hi = min(hi, 0); lo = max(lo, yl); }
  if (fin(yl, yu) && fin(zl, zu) && yl == yu && zl == zu) {
    int r = (int)((long long)yl % (long long)zl);
    lo = max(lo, r); hi = min(hi, r);
  }
  c.xl = lo; c.xu = hi;
}

// Operators whose arithmetic needs extended integers or 64 bits (ADD_G, MUL, TDIV, TMOD): out of
// line, so that the specialised loops stay small; candidates come back through one struct.
__device__ __noinline__ void wide_eval(int op, int xl, int xu, int yl, int yu, int zl, int zu, Cand* out) {
  Cand c;
  c.xl = TBD_NINF; c.xu = TBD_PINF; c.yl = TBD_NINF; c.yu = TBD_PINF; c.zl = TBD_NINF; c.zu = TBD_PINF;
  if (op == TB_OP_ADD) add_ext(xl, xu, yl, yu, zl, zu, c);
  else if (op == TB_OP_MUL) mul(xl, xu, yl, yu, zl, zu, c);
  else if (op == TB_OP_TDIV) tdiv(xl, xu, yl, yu, zl, zu, c);
  else tmod(xl, xu, yl, yu, zl, zu, c);
  *out = c;
}

// ---- class traits ----------------------------------------------------------------------------------------
// Which operands a class loads from the store; the others are constants carried in the propagator word
// (…_XK / …_ZK) or implied by the class (…_T: x = 1, …_F: x = 0).
__host__ __device__ constexpr bool cls_loads_x(int c) {
  return !(c == TBC_ADD_XK || c == TBC_EQ_T || c == TBC_EQ_F || c == TBC_LEQ_T || c == TBC_LEQ_F);
}
__host__ __device__ constexpr bool cls_loads_z(int c) { return !(c == TBC_ADD_ZK || c == TBC_EQ_ZK || c == TBC_LEQ_ZK); }
__host__ __device__ constexpr int cls_op(int c) {
  return (c == TBC_ADD_S || c == TBC_ADD_XK || c == TBC_ADD_ZK || c == TBC_ADD_G) ? TB_OP_ADD
       : c == TBC_MUL ? TB_OP_MUL : c == TBC_TDIV ? TB_OP_TDIV : c == TBC_TMOD ? TB_OP_TMOD
       : c == TBC_MIN ? TB_OP_MIN : c == TBC_MAX ? TB_OP_MAX
       : (c == TBC_EQ_S || c == TBC_EQ_T || c == TBC_EQ_F || c == TBC_EQ_ZK || c == TBC_EQ_G) ? TB_OP_EQ : TB_OP_LEQ;
}
// 32-bit arithmetic on the bounds is exact (every operand lives within +-2^28 at the root)
__host__ __device__ constexpr bool cls_small(int c) {
  return !(c == TBC_ADD_G || c == TBC_MUL || c == TBC_TDIV || c == TBC_TMOD || c == TBC_EQ_G || c == TBC_LEQ_G);
}

// New bounds n* (already met with the snapshot s) of one propagator of class CLS.
template <int CLS>
__device__ __forceinline__ void narrow(const Snap& s, Snap& n) {
  constexpr int op = cls_op(CLS);
  const int xl = s.xl, xu = s.xu, yl = s.yl, yu = s.yu, zl = s.zl, zu = s.zu;
  n = s;
  if (CLS == TBC_ADD_G || op == TB_OP_MUL || op == TB_OP_TDIV || op == TB_OP_TMOD) {
    Cand c;
    wide_eval(op, xl, xu, yl, yu, zl, zu, &c);
    n.xl = max(xl, c.xl); n.xu = min(xu, c.xu); n.yl = max(yl, c.yl); n.yu = min(yu, c.yu); n.zl = max(zl, c.zl); n.zu = min(zu, c.zu);
  } else if (op == TB_OP_ADD) {
    // max(a + b, c) / min(a + b, c) are single instructions on sm_100 (VIADDMNMX)
    n.xl = __viaddmax_s32(yl, zl, xl); n.xu = __viaddmin_s32(yu, zu, xu);
    n.yl = __viaddmax_s32(xl, -zu, yl); n.yu = __viaddmin_s32(xu, -zl, yu);
    n.zl = __viaddmax_s32(xl, -yu, zl); n.zu = __viaddmin_s32(xu, -yl, zu);
  } else if (op == TB_OP_MIN) {
    n.xl = max(xl, min(yl, zl)); n.xu = min(xu, min(yu, zu));
    n.yl = max(yl, xl); n.zl = max(zl, xl);
    if (yl > xu) n.zu = min(zu, xu);
    if (zl > xu) n.yu = min(yu, xu);
  } else if (op == TB_OP_MAX) {
    n.xl = max(xl, max(yl, zl)); n.xu = min(xu, max(yu, zu));
    n.yu = min(yu, xu); n.zu = min(zu, xu);
    if (yu < xl) n.zl = max(zl, xl);
    if (zu < xl) n.yl = max(yl, xl);
  } else if (op == TB_OP_LEQ) {
    // The rules of the three cases (x true / x false / x undecided) are applied side by side: a rule of
    // the "wrong" case only fires on stores the right case fails on too.
    const bool t = xl >= 1, f = xu <= 0;
    const int zl1 = cls_small(CLS) ? zl + 1 : succ(zl), yu1 = cls_small(CLS) ? yu - 1 : pred(yu);
    if (t) { n.yu = min(yu, zu); n.zl = max(zl, yl); }
    if (f) { n.yl = max(yl, zl1); n.zu = min(zu, yu1); }
    if (yu <= zl) n.xl = max(xl, 1);
    if (yl > zu) n.xu = min(xu, 0);
  } else {   // TB_OP_EQ
    const bool t = xl >= 1, f = xu <= 0;
    if (t) { n.yl = max(yl, zl); n.yu = min(yu, zu); n.zl = n.yl; n.zu = n.yu; }
    if (f) {
      if (yl == yu && (cls_small(CLS) || fin(yl, yu))) { if (zl == yl) n.zl = yl + 1; if (zu == yl) n.zu = yl - 1; }
      if (zl == zu && (cls_small(CLS) || fin(zl, zu))) { if (yl == zl) n.yl = zl + 1; if (yu == zl) n.yu = zl - 1; }
    }
    if (yu < zl || zu < yl) n.xu = min(xu, 0);
    else if (yl == yu && zl == zu && yl == zl) n.xl = max(xl, 1);
  }
}

// PIR::ask on a snapshot: is the propagator entailed?
template <int CLS>
__device__ __forceinline__ bool entailed(const Snap& s) {
  constexpr int op = cls_op(CLS);
  if (op == TB_OP_LEQ) return s.xl >= 1 ? s.yu <= s.zl : (s.xu <= 0 ? s.yl > s.zu : false);
  if (op == TB_OP_EQ)
    return s.xl >= 1 ? ((s.yl == s.yu) & (s.zl == s.zu) & (s.yl == s.zl)) : (s.xu <= 0 ? ((s.yu < s.zl) | (s.zu < s.yl)) : false);
  return (s.xl == s.xu) & (s.yl == s.yu) & (s.zl == s.zu);
}

// One evaluation = three phases, so that a thread can interleave the phases of several propagators.
// a, b, c are the three fields of the propagator word (slots, or a constant for …K classes).
template <int CLS, class Store>
__device__ __forceinline__ void load_snap(const Store& st, int a, int b, int c, Snap& s) {
  if (cls_loads_x(CLS)) st.ld(a, s.xl, s.xu);
  else if (CLS == TBC_ADD_XK) s.xl = s.xu = a;
  else s.xl = s.xu = (CLS == TBC_EQ_T || CLS == TBC_LEQ_T) ? 1 : 0;
  st.ld(b, s.yl, s.yu);
  if (cls_loads_z(CLS)) st.ld(c, s.zl, s.zu); else s.zl = s.zu = c;
}
template <int CLS>
__device__ __forceinline__ bool snap_changed(const Snap& s, const Snap& n) {
  bool changed = (n.yl != s.yl) | (n.yu != s.yu);
  if (cls_loads_x(CLS)) changed |= (n.xl != s.xl) | (n.xu != s.xu);
  if (cls_loads_z(CLS)) changed |= (n.zl != s.zl) | (n.zu != s.zu);
  return changed;
}
// A loaded interval is empty.
template <int CLS>
__device__ __forceinline__ bool snapshot_empty(const Snap& s) {
  bool e = s.yl > s.yu;
  if (cls_loads_x(CLS)) e |= s.xl > s.xu;
  if (cls_loads_z(CLS)) e |= s.zl > s.zu;
  return e;
}

// "This evaluation has something to do": `narrow` would move a bound of a loaded operand, or a loaded interval
// is empty. This is what every evaluation computes; the new bounds themselves are only computed when some lane
// of the warp answers yes. For the additions the six tests "candidate beyond current bound" are six 3-input
// adds whose maximum is positive iff something moves, about half the ALU work of min/max + compare; for the
// three-variable addition that maximum is also positive whenever an operand is empty (if none of the six
// candidates improves a bound then yl + zl <= xl <= yl + zu, ... which forces xl <= xu, yl <= yu, zl <= zu).
template <int CLS>
__device__ __forceinline__ bool has_work(const Snap& s) {
  if (CLS == TBC_ADD_S) {
    const int d1 = s.yl + s.zl - s.xl, d2 = s.xu - s.yu - s.zu;
    const int d3 = s.xl - s.zu - s.yl, d4 = s.yu - s.xu + s.zl;
    const int d5 = s.xl - s.yu - s.zl, d6 = s.zu - s.xu + s.yl;
    return max(__vimax3_s32(d1, d2, d3), __vimax3_s32(d4, d5, d6)) > 0;
  }
  if (CLS == TBC_ADD_XK) {      // x is the constant s.xl
    const int d3 = s.xl - s.zu - s.yl, d4 = s.yu - s.xl + s.zl;
    const int d5 = s.xl - s.yu - s.zl, d6 = s.zu - s.xl + s.yl;
    return (max(__vimax3_s32(d3, d4, d5), d6) > 0) | snapshot_empty<CLS>(s);
  }
  if (CLS == TBC_ADD_ZK) {      // z is the constant s.zl
    const int d1 = s.yl + s.zl - s.xl, d2 = s.xu - s.yu - s.zl;
    const int d3 = s.xl - s.zl - s.yl, d4 = s.yu - s.xu + s.zl;
    return (max(__vimax3_s32(d1, d2, d3), d4) > 0) | snapshot_empty<CLS>(s);
  }
  Snap n;
  narrow<CLS>(s, n);
  return snap_changed<CLS>(s, n) | snapshot_empty<CLS>(s);
}

// Publish the bounds that moved (lanes whose propagator changed nothing publish nothing).
template <int CLS, class Store>
__device__ __forceinline__ void publish(const Store& st, int a, int b, int c, const Snap& s, const Snap& n, unsigned& narrowed) {
  if (cls_loads_x(CLS)) {
    st.tell_lb(a, n.xl, s.xl, narrowed);
    st.tell_ub(a, n.xu, s.xu, narrowed);
  }
  st.tell_lb(b, n.yl, s.yl, narrowed);
  st.tell_ub(b, n.yu, s.yu, narrowed);
  if (cls_loads_z(CLS)) {
    st.tell_lb(c, n.zl, s.zl, narrowed);
    st.tell_ub(c, n.zu, s.zu, narrowed);
  }
}

// Failure detection: an empty interval gives every propagator that loads it "work" (has_work), and whoever has
// work re-reads the intervals it loads after publishing and reports the empty ones. So an interval emptied by
// this thread's own update is reported at once, and one emptied by two publishers together (one raises lb, the
// other lowers ub, each from a non-empty snapshot) at the latest by the next evaluation that loads it: no
// sweep can end without change while a loaded interval is empty.
template <int CLS, class Store>
__device__ __forceinline__ bool emptied(const Store& st, int a, int b, int c) {
  int l, u;
  st.ld(b, l, u);
  bool e = l > u;
  if (cls_loads_x(CLS)) { st.ld(a, l, u); e |= l > u; }
  if (cls_loads_z(CLS)) { st.ld(c, l, u); e |= l > u; }
  return e;
}

// Non-zero iff the propagator is not entailed on the snapshot (the fused `ask`).
template <int CLS>
__device__ __forceinline__ int not_entailed_bits(const Snap& s) {
  constexpr int op = cls_op(CLS);
  if (op == TB_OP_LEQ || op == TB_OP_EQ) return entailed<CLS>(s) ? 0 : 1;
  return (s.xl ^ s.xu) | (s.yl ^ s.yu) | (s.zl ^ s.zu);
}

}  // namespace tbd
