// layout.h — host-side layout pass: tb_problem -> device propagator table + variable placement.
//
// The reference hands PIR's bytecode array to the device as it comes out of the ternariser
// (include/barebones_dive_and_solve.hpp:82,561).  Here the table is compiled for the B200's shared
// memory: propagators are classified (tnf_classes.h), sorted by class into chunks of 32 (one warp
// evaluates one chunk), packed to 8 bytes, and the variables are renumbered so that the 32 lanes of
// a chunk hit 32 different shared-memory banks whenever possible.  The fixpoint does not depend on
// any of this (monotone contracting operators: the greatest fixpoint is unique).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/turbo_b200.h"
#include "tnf_classes.h"

struct TnfLayout {
  int nvars = 0;
  int nslots = 0;                          // size of a store image in variables (>= nvars, padded)
  std::vector<int> slot_of;                // variable -> slot in the store image
  std::vector<uint64_t> words;             // nchunks * 32 device words, class-sorted
  int cls_begin[TBC_NUM + 1] = {0};        // first chunk of each class; cls_begin[TBC_NUM] = nchunks
  int cls_last[TBC_NUM] = {0};             // real propagators in the last chunk of the class (1..32)
  std::vector<uint8_t> referenced;         // per variable: appears in some propagator
  uint64_t loads_per_sweep = 0;            // 8-byte {lb, ub} loads one sweep over the table issues (real lanes)
  double wavefronts_per_load = 0.0;        // bank model: average shared-memory wavefronts per half-warp load
  bool identity = true;                    // slot_of[v] == v
  std::vector<int> watch_off, watch_list;  // slot -> chunks that load it (CSR over nslots; active-set fixpoint)
  std::vector<int> chunk_of_prop;          // propagator -> chunk of the device table
  double cluster_local_fraction = 0.0;     // STORE_CLUSTER: share of the operand loads of a sweep that stay in the evaluating CTA
  std::vector<uint64_t> watch_inline;      // per slot: first three watchers as 16-bit ids (0xFFFF = none), top 16 bits 0xFFFE = more in the list
};

struct TnfLayoutOptions {
  int nbanks = 0;        // 0: keep the caller's variable numbering; 16: the 8-byte banks of one SM
  int lanes_per_set = 16; // lanes whose loads are served together (a half-warp for 8-byte {lb, ub} pairs)
  int slot_align = 4;    // nslots is rounded up to a multiple of this (and of nbanks)
  // STORE_CLUSTER: the store is striped over `cluster` CTAs (slot s lives in CTA s mod cluster) of `cluster_warps` warps
  // each, and chunk ch is evaluated by the cluster's warp ch mod (cluster * cluster_warps). With cluster > 1 the pass
  // places variables that occur together in the same CTA (breadth-first order over the propagators, cut into `cluster`
  // parts) and gives every CTA the chunks whose operands mostly live in it: only the cut goes through DSMEM.
  int cluster = 0, cluster_warps = 0;
};

// Returns TB_OK or TB_ERR_UNSUPPORTED (too many variables for a 21-bit field) with *err set.
tb_status tb_build_layout(const tb_problem* pb, const TnfLayoutOptions& opt, TnfLayout* out, std::string* err);

// Class of one propagator given the root domains (exposed for tests / statistics).
int tb_classify(const tb_prop& p, const int32_t* lb, const int32_t* ub, bool* swap_yz);
const char* tb_class_name(int cls);
