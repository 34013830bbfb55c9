// fzn_parser.hpp — FlatZinc front-end (host side, C++17).
//
// Replaces lala-parsing's parse_flatzinc (called at reference include/common_solving.hpp:404-408;
// the parser itself is an un-vendored dependency).  Covers the FlatZinc subset of SURVEY.md
// Appendix B: parameter arrays, bounded / boolean / set-literal variable declarations, variable
// arrays mixing identifiers and literals, output annotations, nested predicate calls used as terms
// (test_data/bug1.fzn), solve items with int_search / bool_search / seq_search. A set variable `var set of lo..hi`
// is declared as one Boolean membership variable per value of its universe (what MiniZinc's nosets.mzn does at the
// model level; unsolved_bugs_data/valve6.fzn was flattened without it).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace fzn {

struct ParseError {
  std::string message;
  int line;
};

// A term after identifier resolution.
struct Expr {
  enum Kind { INT, BOOL, VAR, ARRAY, SET, CALL, STRING, IDENT, SETVAR } kind = INT;
  int64_t value = 0;                    // INT / BOOL literal
  int var = -1;                         // VAR: index into Model::vars
  std::vector<Expr> elems;              // ARRAY elements / CALL arguments / SETVAR: one membership variable per universe value (ascending)
  std::vector<std::pair<int64_t, int64_t>> ranges;  // SET: sorted disjoint ranges; SETVAR: its universe
  std::string name;                     // CALL / IDENT / STRING text
};

struct Var {
  std::string name;
  bool is_bool = false;
  bool has_lb = false, has_ub = false;
  int64_t lb = 0, ub = 0;
  std::vector<int64_t> holes;           // values inside [lb,ub] excluded by a set-literal domain
  bool output = false;
  bool introduced = false;
  int alias_var = -1;                   // `var int: x = y;`
  bool has_alias_const = false;
  int64_t alias_const = 0;
};

struct OutputArray {
  std::string name;
  std::vector<std::pair<int64_t, int64_t>> dims;
  std::vector<Expr> elems;              // VAR / INT / BOOL
  bool is_bool = false;
};

struct Constraint {
  std::string name;
  std::vector<Expr> args;
  int line = 0;
};

struct SearchAnn {                      // one int_search / bool_search
  std::vector<Expr> vars;
  std::string var_sel, val_sel;
};

struct Model {
  std::vector<Var> vars;
  std::map<std::string, int> var_index;
  std::map<std::string, Expr> names;    // parameters, parameter arrays and variable arrays by name
  std::vector<Constraint> constraints;
  std::vector<OutputArray> output_arrays;
  std::vector<std::string> output_order;   // names in declaration order ("v:<name>" or "a:<name>")
  enum { SATISFY, MINIMIZE, MAXIMIZE } solve = SATISFY;
  Expr objective;
  std::vector<SearchAnn> search;
};

// Throws ParseError.
std::unique_ptr<Model> parse(const char* text, size_t len);

}  // namespace fzn
