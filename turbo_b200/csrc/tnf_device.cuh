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
__device__ __noinline__ void mul(int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
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
__device__ __noinline__ void tdiv(int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
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
__device__ __noinline__ void tmod(int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
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

// Evaluate propagator `op` on the loaded snapshot. Returns F_NOT_ENTAILED if `ask` is false.
__device__ __forceinline__ int eval(int op, int xl, int xu, int yl, int yu, int zl, int zu, Cand& c) {
  c.xl = TBD_NINF; c.xu = TBD_PINF; c.yl = TBD_NINF; c.yu = TBD_PINF; c.zl = TBD_NINF; c.zu = TBD_PINF;
  bool entailed;
  switch (op) {
    case TB_OP_ADD:
      add(xl, xu, yl, yu, zl, zu, c);
      entailed = (xl == xu) & (yl == yu) & (zl == zu);
      break;
    case TB_OP_LEQ:
      if (xl >= 1) { c.yu = zu; c.zl = yl; entailed = yu <= zl; }
      else if (xu <= 0) { c.yl = succ(zl); c.zu = pred(yu); entailed = yl > zu; }
      else {
        if (yu <= zl) c.xl = 1; else if (yl > zu) c.xu = 0;
        entailed = false;
      }
      break;
    case TB_OP_EQ:
      if (xl >= 1) {
        c.yl = zl; c.yu = zu; c.zl = yl; c.zu = yu;
        entailed = (yl == yu) & (zl == zu) & (yl == zl);
      }
      else if (xu <= 0) {
        if (yl == yu && fin(yl, yu)) { if (zl == yl) c.zl = yl + 1; if (zu == yl) c.zu = yl - 1; }
        if (zl == zu && fin(zl, zu)) { if (yl == zl) c.yl = zl + 1; if (yu == zl) c.yu = zl - 1; }
        entailed = (yu < zl) | (zu < yl);
      }
      else {
        if (yu < zl || zu < yl) c.xu = 0;
        else if (yl == yu && zl == zu && yl == zl) c.xl = 1;
        entailed = false;
      }
      break;
    case TB_OP_MIN:
      c.xl = min(yl, zl); c.xu = min(yu, zu);
      c.yl = xl; c.zl = xl;
      if (yl > xu) c.zu = xu;
      if (zl > xu) c.yu = xu;
      entailed = (xl == xu) & (yl == yu) & (zl == zu);
      break;
    case TB_OP_MAX:
      c.xl = max(yl, zl); c.xu = max(yu, zu);
      c.yu = xu; c.zu = xu;
      if (yu < xl) c.zl = xl;
      if (zu < xl) c.yl = xl;
      entailed = (xl == xu) & (yl == yu) & (zl == zu);
      break;
    case TB_OP_MUL:
      mul(xl, xu, yl, yu, zl, zu, c);
      entailed = (xl == xu) & (yl == yu) & (zl == zu);
      break;
    case TB_OP_TDIV:
      tdiv(xl, xu, yl, yu, zl, zu, c);
      entailed = (xl == xu) & (yl == yu) & (zl == zu);
      break;
    case TB_OP_TMOD:
      tmod(xl, xu, yl, yu, zl, zu, c);
      entailed = (xl == xu) & (yl == yu) & (zl == zu);
      break;
    default:  // TB_OP_NOP
      entailed = true;
      break;
  }
  return entailed ? 0 : F_NOT_ENTAILED;
}

}  // namespace tbd
