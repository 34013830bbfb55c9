/* tnf_oracle.h — CPU oracle for the dive-and-solve path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libturbo_b200.so, bin/turbo) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in lattice-land/lala-pc v1.2.8
 * (lala/pir.hpp: PIR::deduce/ask) and lala-core (interval.hpp, vstore.hpp, fixpoint.hpp,
 * split_strategy.hpp), pulled by CMake FetchContent (reference CMakeLists.txt:46-65) and absent
 * from /root/reference and from this machine.  The reference ships no fixpoint-level golden
 * vector.  What IS pinned: the 32 FlatZinc optima of benchmarks/test_list.csv (status + objective,
 * end to end) and brute-force soundness of every operator.  The control flow below follows the
 * reference files cited at each function.
 */
#ifndef TNF_ORACLE_H
#define TNF_ORACLE_H

#include "../include/turbo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One narrowing step of propagator p on (lb, ub); returns 1 when a bound moved.
 * *failed is set (never cleared) when some touched interval is or becomes empty.
 * Follows the contract of PIR::deduce(i) as used at barebones_dive_and_solve.hpp:931. */
int tbo_deduce(const tb_prop* p, int32_t* lb, int32_t* ub, int32_t* failed);

/* Entailment of propagator p on (lb, ub): PIR::ask(i), barebones_dive_and_solve.hpp:977. */
int tbo_ask(const tb_prop* p, const int32_t* lb, const int32_t* ub);

/* Gauss-Seidel fixpoint: in-order sweeps until a sweep changes nothing or the store fails
 * (GaussSeidelIteration, cpu_solving.hpp:20-26; stop-on-bot as barebones :932).
 * `order` (may be NULL) is a permutation of 0..nprops-1 used for schedule-independence tests.
 * Returns the number of sweeps. */
int64_t tbo_fixpoint(const tb_problem* pb, int32_t* lb, int32_t* ub, int32_t* failed,
                     const int32_t* order, uint64_t* num_deductions);

/* EPS dive (barebones_dive_and_solve.hpp:663-741). leaf_kind: 0 reached, 1 failed, 2 solution. */
int tbo_dive(const tb_problem* pb, uint64_t idx, int32_t depth,
             int32_t* lb_out, int32_t* ub_out, int32_t* remaining_depth, int32_t* leaf_kind);

/* Sequential dive-and-solve with the semantics of one barebones block that processes the
 * subproblems 0 .. 2^depth-1 in order (barebones_dive_and_solve.hpp:620-901, 903-1031).
 * nthreads > 1 runs EPS-parallel workers sharing the incumbent (the "all host cores" baseline). */
int tbo_solve(const tb_problem* pb, int32_t depth, uint64_t cutnodes, uint64_t timeout_ms,
              int32_t nthreads, volatile int32_t* stop_flag,
              int32_t* best_lb, int32_t* best_ub, int32_t* has_solution, int32_t* exhaustive,
              tb_stats* stats);

/* One GPU's shard of the subproblems (idx = k * world + rank) with an incumbent the
 * other shards may already have found (initial_bound, TB_POS_INF for none). Used by the world_size-2
 * gloo tests of the multi-rank host logic. */
int tbo_solve_shard(const tb_problem* pb, int32_t depth, uint64_t cutnodes, uint64_t timeout_ms,
                    int32_t nthreads, int32_t rank, int32_t world, int32_t initial_bound,
                    volatile int32_t* stop_flag,
                    int32_t* best_lb, int32_t* best_ub, int32_t* has_solution, int32_t* exhaustive,
                    tb_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
