// tnf_device.cuh — device-side interval narrowing for one TNF propagator  x = y op z.
//
// B200-native replacement of PIR::deduce / PIR::ask (called at reference
// include/barebones_dive_and_solve.hpp:931,944,977; bodies live in lala-pc, not in the tree).
// One thread evaluates one propagator on values it has already loaded into registers and returns
// the candidate bounds; the caller publishes them with shared-memory atomicMax/atomicMin.
// All candidates are computed from the *loaded* snapshot (Jacobi style); the greatest fixpoint is
// schedule independent, so this agrees bit for bit with the sequential CPU oracle.
//
// Result flags: bit0 = a bound moved, bit1 = an interval is (or became) empty,
//               bit2 = propagator not entailed on the loaded snapshot (fused `ask`).
#pragma once
#include <stdint.h>
#include "../../include/turbo_b200.h"

#define TBD_NINF INT32_MIN
#define TBD_PINF INT32_MAX
#define TB_OP_NOP 8            // padding propagator: never changes anything, always entailed

#define F_CHANGED 1
#define F_FAILED 2
#define F_NOT_ENTAILED 4

namespace tbd {

struct Cand {          // candidate bounds for x, y, z (initialised to "no information")
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

// ---- x = y + z ---------------------------------------------------------------------------------
__device__ __forceinline__ void add(int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
  // Fast path: every bound in [-2^30, 2^30): plain 32-bit arithmetic cannot wrap and no infinity
  // is involved. (v + 2^30) has its sign bit clear exactly for those values.
  const unsigned B = 0x40000000u;
  unsigned t = ((unsigned)xl + B) | ((unsigned)xu + B) | ((unsigned)yl + B) | ((unsigned)yu + B) |
               ((unsigned)zl + B) | ((unsigned)zu + B);
  if ((int)t >= 0) {
    c.xl = yl + zl; c.xu = yu + zu;
    c.yl = xl - zu; c.yu = xu - zl;
    c.zl = xl - yu; c.zu = xu - yl;
  } else {
    long long exl = ext64(xl), exu = ext64(xu), eyl = ext64(yl), eyu = ext64(yu), ezl = ext64(zl), ezu = ext64(zu);
    c.xl = clamp64(eyl + ezl); c.xu = clamp64(eyu + ezu);
    c.yl = clamp64(exl - ezu); c.yu = clamp64(exu - ezl);
    c.zl = clamp64(exl - eyu); c.zu = clamp64(exu - eyl);
  }
}

// quotient hull for  f = p / d  (d without 0, everything finite), rounded inward
__device__ __forceinline__ void quot(int pl, int pu, int dl, int du, int& ql, int& qu) {
  long long a = pl, b = pu, c = dl, d = du;
  ql = clamp64(min4(cdiv64(a, c), cdiv64(a, d), cdiv64(b, c), cdiv64(b, d)));
  qu = clamp64(max4(fdiv64(a, c), fdiv64(a, d), fdiv64(b, c), fdiv64(b, d)));
}

// ---- x = y * z ---------------------------------------------------------------------------------
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

// ---- x = y tdiv z, z != 0 ----------------------------------------------------------------------
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

// ---- x = y tmod z, z != 0 ----------------------------------------------------------------------
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

// The rare operators (MUL / TDIV / TMOD): kept out of line so that the hot loop stays small and
// register-resident; candidates come back through one by-value struct.
__device__ __noinline__ void rare_eval(int op, int xl, int xu, int yl, int yu, int zl, int zu, Cand* out) {
  Cand c;
  c.xl = TBD_NINF; c.xu = TBD_PINF; c.yl = TBD_NINF; c.yu = TBD_PINF; c.zl = TBD_NINF; c.zu = TBD_PINF;
  if (op == TB_OP_MUL) mul(xl, xu, yl, yu, zl, zu, c);
  else if (op == TB_OP_TDIV) tdiv(xl, xu, yl, yu, zl, zu, c);
  else tmod(xl, xu, yl, yu, zl, zu, c);
  *out = c;
}

// Hot operators, fully inlined. On entry n* hold the current bounds; on exit the narrowed ones
// (new = current meet candidate). Returns whether the propagator is entailed on the snapshot.
// `op` must be one of ADD, LEQ, EQ, MIN, MAX, NOP.
__device__ __forceinline__ bool hot_eval(int op, int xl, int xu, int yl, int yu, int zl, int zu,
                                         int& nxl, int& nxu, int& nyl, int& nyu, int& nzl, int& nzu) {
  const bool ground = (xl == xu) & (yl == yu) & (zl == zu);
  if (op == TB_OP_ADD) {
    const unsigned B = 0x40000000u;
    unsigned t = ((unsigned)xl + B) | ((unsigned)xu + B) | ((unsigned)yl + B) | ((unsigned)yu + B) |
                 ((unsigned)zl + B) | ((unsigned)zu + B);
    if ((int)t >= 0) {          // all bounds in [-2^30, 2^30): 32-bit arithmetic is exact, no infinity involved
      nxl = max(xl, yl + zl); nxu = min(xu, yu + zu);
      nyl = max(yl, xl - zu); nyu = min(yu, xu - zl);
      nzl = max(zl, xl - yu); nzu = min(zu, xu - yl);
    } else {
      Cand c;
      add(xl, xu, yl, yu, zl, zu, c);
      nxl = max(xl, c.xl); nxu = min(xu, c.xu); nyl = max(yl, c.yl); nyu = min(yu, c.yu); nzl = max(zl, c.zl); nzu = min(zu, c.zu);
    }
    return ground;
  }
  if (op == TB_OP_LEQ) {
    if (xl >= 1) { nyu = min(yu, zu); nzl = max(zl, yl); return yu <= zl; }
    if (xu <= 0) { nyl = max(yl, succ(zl)); nzu = min(zu, pred(yu)); return yl > zu; }
    if (yu <= zl) nxl = 1; else if (yl > zu) nxu = 0;
    return false;
  }
  if (op == TB_OP_EQ) {
    if (xl >= 1) {
      nyl = max(yl, zl); nyu = min(yu, zu); nzl = nyl; nzu = nyu;
      return (yl == yu) & (zl == zu) & (yl == zl);
    }
    if (xu <= 0) {
      if (yl == yu && fin(yl, yu)) { if (zl == yl) nzl = yl + 1; if (zu == yl) nzu = yl - 1; }
      if (zl == zu && fin(zl, zu)) { if (yl == zl) nyl = zl + 1; if (yu == zl) nyu = zl - 1; }
      return (yu < zl) | (zu < yl);
    }
    if (yu < zl || zu < yl) nxu = 0;
    else if (yl == yu && zl == zu && yl == zl) nxl = 1;
    return false;
  }
  if (op == TB_OP_MIN) {
    nxl = max(xl, min(yl, zl)); nxu = min(xu, min(yu, zu));
    nyl = max(yl, xl); nzl = max(zl, xl);
    if (yl > xu) nzu = min(zu, xu);
    if (zl > xu) nyu = min(yu, xu);
    return ground;
  }
  if (op == TB_OP_MAX) {
    nxl = max(xl, max(yl, zl)); nxu = min(xu, max(yu, zu));
    nyu = min(yu, xu); nzu = min(zu, xu);
    if (yu < xl) nzl = max(zl, xl);
    if (zu < xl) nyl = max(yl, xl);
    return ground;
  }
  return true;   // TB_OP_NOP
}

__device__ __forceinline__ bool is_rare(int op) { return op >= TB_OP_MUL && op <= TB_OP_TMOD; }

}  // namespace tbd
