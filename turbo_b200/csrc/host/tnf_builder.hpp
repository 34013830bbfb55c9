// tnf_builder.hpp — FlatZinc model -> ternary normal form (host side, C++17).
//
// Replaces lala-core's ternarize/normalize + the interpretation into PIR
// (reference include/common_solving.hpp:520-529; the rewriting itself is un-vendored).
// Every constraint becomes propagators  x = y op z  over int32 interval variables; integer
// literals are variables with singleton domains ("no constant in a TCN", common_solving.hpp:725-727)
// with 0, 1, 2 pre-created as variables 0, 1, 2 (ternarize(f, env, {0,1,2}), :521).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/turbo_b200.h"
#include "fzn_parser.hpp"

struct tb_model {
  std::unique_ptr<fzn::Model> src;           // null for synthetic / .tnf models
  std::vector<int32_t> lb, ub;
  std::vector<tb_prop> props;
  std::vector<std::vector<int32_t>> strat_vars;  // empty list = all store variables
  std::vector<std::pair<int, int>> strat_orders; // (var_order, val_order), parallel to strat_vars
  std::vector<tb_strategy> strategies;
  bool has_eps_strategy = false;
  std::vector<int32_t> var_of_model;         // FlatZinc variable -> TNF variable
  tb_problem problem{};
  int objective_kind = -1;                   // -1 satisfy, 0 minimise, 1 maximise
  int user_obj_var = -1;                     // TNF variable of the user's objective
  int obj_var = -1;                          // TNF variable that is minimised
  bool root_failed = false;
  int parsed_variables = 0, parsed_constraints = 0;
  std::string error;
  // ---- set by tb_model_simplify (tnf_simplify.cpp): lb/ub/props/strategies above then describe the REDUCED
  // network; the full one (what var_of_model indexes) is kept here for printing and re-checking ------------
  bool simplified = false;
  tb_simplify_stats simplify_stats{};
  std::vector<int32_t> full_lb, full_ub;     // root-propagated domains of the full network
  std::vector<tb_prop> full_props;
  std::vector<int32_t> red_of_full;          // full variable -> reduced variable, -1 = eliminated (takes full_lb, or its definition)
  std::vector<int32_t> rep_of_full;          // full variable -> representative of its equivalence class (full index)
  std::vector<tb_prop> defs;                 // x = y op z of functionally defined variables (full indices of representatives)

  void finalize();                           // (re)build `problem` from the vectors
  // Prepend the EPS strategy (-eps_var_order / -eps_value_order, common_solving.hpp:652-667).
  void push_eps_strategy(int var_order, int val_order);
};

// Throws std::runtime_error on unsupported constraints.
std::unique_ptr<tb_model> build_tnf(std::unique_ptr<fzn::Model> src);

// Store of the reduced network -> point of the full one (tnf_simplify.cpp).
void tb_model_expand_internal(const tb_model* m, const int32_t* lb, const int32_t* ub, std::vector<int32_t>& flb, std::vector<int32_t>& fub);

// Number of violated FlatZinc constraints / domains at the point value[v] (FlatZinc variable index).
int check_flatzinc(const fzn::Model& m, const std::vector<int64_t>& value, std::string* first_violation);
