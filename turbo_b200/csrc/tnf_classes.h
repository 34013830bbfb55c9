// tnf_classes.h — propagator classes of the device table (shared by the host layout pass and the kernels).
//
// The public operator set is tb_op (include/turbo_b200.h; the reference's lala::Sig inside PIR's
// bytecode, include/common_solving.hpp:738-771).  On the device every propagator belongs to one class:
// operator x "which operands are constants at the root" x "is 32-bit arithmetic exact".  The table is
// sorted by class and each class is padded to whole chunks of 32, so a warp never mixes classes.
//
// Device word (64 bits): three 21-bit fields  f0 | f1 << 21 | f2 << 42.
//   f0 = slot of x   (…_XK: the constant value of x, two's complement; …_T / …_F: unused)
//   f1 = slot of y
//   f2 = slot of z   (…_ZK: the constant value of z, two's complement)
#pragma once

enum {
  TBC_ADD_S = 0,   // x = y + z, every operand within +-2^28 at the root
  TBC_ADD_XK,      // k = y + z
  TBC_ADD_ZK,      // x = y + k
  TBC_ADD_G,       // x = y + z on extended integers (infinite or huge bounds)
  TBC_MUL, TBC_TDIV, TBC_TMOD,
  TBC_MIN, TBC_MAX,
  TBC_EQ_S,        // x = (y == z), y and z within +-2^28
  TBC_EQ_T,        // y == z          (x is the constant 1)
  TBC_EQ_F,        // y != z          (x is the constant 0)
  TBC_EQ_ZK,       // x = (y == k)
  TBC_EQ_G,        // x = (y == z) with infinite bounds around
  TBC_LEQ_S,       // x = (y <= z)
  TBC_LEQ_T,       // y <= z
  TBC_LEQ_F,       // y > z
  TBC_LEQ_ZK,      // x = (y <= k)
  TBC_LEQ_G,
  TBC_NUM
};

// Propagators per thread and chunk visit: a chunk is 32 * TBC_U propagators of one class (TBC_U "rows" of 32);
// lane l evaluates lane l of every row. The shipped build uses one row (two rows per visit measured 13 % slower:
// DESIGN.md); the Makefile passes the same value to every translation unit.
#ifndef TBC_U
#define TBC_U 1
#endif

#define TBC_FIELD_BITS 21
#define TBC_FIELD_MASK 0x1FFFFFu
#define TBC_MAX_VARS (1 << TBC_FIELD_BITS)
#define TBC_CONST_LIMIT (1 << (TBC_FIELD_BITS - 1))     // constants in [-2^20, 2^20) fit a field
#define TBC_SMALL_LIMIT (1 << 28)                       // "small": every bound within [-2^28, 2^28]: sums of three fit 32 bits
