// turbo_main.cpp — the `turbo` command-line driver: the reference's process-level drop-in surface
// (SURVEY.md §8b(1)).  Same flags (reference src/config.cpp:11-44,128-220), same stdout protocol
// (include/statistics.hpp:232-412, include/config.hpp:237-266, include/memory_gpu.hpp:113-122), so
// MiniZinc's solns2out and test_turbo.sh's greps (`objective=`, `solveTime=`) keep working.
// All solving goes through the C ABI (tb_*); there is no CPU solving path in this binary.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cinttypes>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/turbo_b200.h"

namespace {

struct Config {                      // Configuration<> (include/config.hpp:32-57)
  bool print_intermediate_solutions = false;
  size_t stop_after_n_solutions = 1;
  size_t stop_after_n_nodes = 0;     // 0 = no limit (printed as cutnodes=0, config.hpp:265)
  bool free_search = false, print_statistics = false, print_ast = false, only_global_memory = false;
  bool force_ternarize = false, disable_simplify = false, disable_network_analysis = false;
  int verbose = 0;
  size_t timeout_ms = 0, or_nodes = 0, subproblems_factor = 300, stack_kb = 0, wac1_threshold = 0, seed = 0;
  int subproblems_power = -1;
  std::string arch = "barebones", fixpoint = "auto", eps_var_order = "default", eps_value_order = "default";
  std::string problem_path, version, hardware;
  int gpus = 1;                      // extension: number of B200s to shard the subproblems over
  int threads_per_block = 0;         // extension: 0 = placement policy decides
};

void usage_and_exit(const std::string& prog) {
  std::cout << "usage: " << prog << " [-t 2000] [-a] [-n 10] [-i] [-f] [-s] [-v] [-p <i>] [-arch <barebones|gpu|hybrid>] [-or 48] [-sub 12] "
               "[-subfactor 300] [-stack 100] [-fp <ac1|wac1|ac1_active|wac1_active>] [-wac1_threshold 0] [-eps_var_order <input_order|first_fail|anti_first_fail|smallest|largest>] "
               "[-eps_value_order <min|max|split|reverse_split>] [-seed 0] [-cutnodes 0] [-disable_simplify] [-force_ternarize] [-globalmem] "
               "[-disable_network_analysis] [-version 1.0.0] [-hardware <desc>] [-gpus 1] [-tpb 0] fzninstance.fzn|instance.tnf" << std::endl;
  std::cout << "\t-t 2000: Run the solver with a timeout of 2000 milliseconds (-timeout overrides -t)." << std::endl;
  std::cout << "\t-a / -i: accepted; the dive-and-solve architecture prints the best solution at the end." << std::endl;
  std::cout << "\t-n 10: accepted (satisfaction problems stop at the first solution)." << std::endl;
  std::cout << "\t-f: accepted and ignored (free search)." << std::endl;
  std::cout << "\t-s: Print statistics during and after the search for solutions." << std::endl;
  std::cout << "\t-v: Print log messages; repeat for more." << std::endl;
  std::cout << "\t-p 48 / -or 48: number of thread blocks searching in parallel (0 = automatic)." << std::endl;
  std::cout << "\t-arch: barebones is the only GPU architecture; gpu and hybrid are mapped onto it. cpu is not available in this build." << std::endl;
  std::cout << "\t-fp <ac1|wac1|ac1_active|wac1_active>: fixpoint strategy (default: wac1_active where it pays, i.e. on tables of 128 chunks and more in shared memory, wac1 elsewhere); the _active kinds only re-evaluate propagators whose variables changed: same fixpoints, same search." << std::endl;
  std::cout << "\t-sub 12: create 2^12 subproblems (-1: at least subfactor * blocks * gpus)." << std::endl;
  std::cout << "\t-cutnodes 1000: stop a block after 1000 nodes (0 for no limit)." << std::endl;
  std::cout << "\t-globalmem: keep the variable store in global memory." << std::endl;
  std::cout << "\t-gpus 8: shard the subproblems over 8 GPUs, sharing the incumbent through peer memory." << std::endl;
  exit(EXIT_FAILURE);
}

// The reference's InputParser (src/config.cpp:47-126): flags are looked up anywhere in argv, unknown
// flags are ignored, the input file is the last token.
struct InputParser {
  std::string program;
  std::vector<std::string> tokens;
  size_t tokens_read = 0;
  InputParser(int argc, char** argv) : program(argv[0]) { for (int i = 1; i < argc; ++i) tokens.push_back(argv[i]); }
  const std::string& option(const std::string& o) const {
    static const std::string empty;
    auto it = std::find(tokens.begin(), tokens.end(), o);
    if (it != tokens.end() && ++it != tokens.end()) return *it;
    return empty;
  }
  bool exists(const std::string& o) const { return std::find(tokens.begin(), tokens.end(), o) != tokens.end(); }
  bool read_size(const std::string& o, size_t& r) { const std::string& v = option(o); if (!v.empty()) { sscanf(v.c_str(), "%zu", &r); tokens_read += 2; return true; } return false; }
  bool read_int(const std::string& o, int& r) { const std::string& v = option(o); if (!v.empty()) { r = atoi(v.c_str()); tokens_read += 2; return true; } return false; }
  bool read_bool(const std::string& o, bool& r) { r = exists(o); if (r) ++tokens_read; return r; }
  bool read_string(const std::string& o, std::string& r) { r = option(o); if (!r.empty()) { tokens_read += 2; return true; } return false; }
};

Config parse_args(int argc, char** argv) {
  Config c;
  InputParser in(argc, argv);
  if (in.exists("-or") && in.exists("-p")) {
    std::cerr << "The options -or and -p cannot be used at the same time" << std::endl;
    usage_and_exit(argv[0]);
  }
  in.read_int("-sub", c.subproblems_power);
  in.read_size("-subfactor", c.subproblems_factor);
  in.read_size("-p", c.or_nodes);
  in.read_size("-or", c.or_nodes);
  in.read_size("-t", c.timeout_ms);
  in.read_size("-timeout", c.timeout_ms);
  in.read_size("-stack", c.stack_kb);
  in.read_size("-n", c.stop_after_n_solutions);
  in.read_size("-cutnodes", c.stop_after_n_nodes);
  in.read_size("-seed", c.seed);
  in.read_bool("-i", c.print_intermediate_solutions);
  bool all = false;
  in.read_bool("-a", all);
  if (all) { c.stop_after_n_solutions = 0; c.print_intermediate_solutions = true; }
  in.read_bool("-f", c.free_search);
  c.verbose = (int)std::count(in.tokens.begin(), in.tokens.end(), std::string("-v"));
  in.tokens_read += (size_t)c.verbose;
  in.read_bool("-ast", c.print_ast);
  in.read_bool("-s", c.print_statistics);
  in.read_bool("-globalmem", c.only_global_memory);
  in.read_bool("-disable_simplify", c.disable_simplify);
  in.read_bool("-force_ternarize", c.force_ternarize);
  in.read_bool("-disable_network_analysis", c.disable_network_analysis);
  std::string arch;
  if (in.read_string("-arch", arch)) {
    if (arch != "cpu" && arch != "hybrid" && arch != "gpu" && arch != "barebones") {
      std::cerr << "Unknown architecture -arch " << arch << std::endl;
      exit(EXIT_FAILURE);
    }
    c.arch = arch;
  }
  std::string fp;
  if (in.read_string("-fp", fp)) {
    if (fp != "ac1" && fp != "wac1" && fp != "ac1_active" && fp != "wac1_active") { std::cerr << "Unknown fixpoint -fp " << fp << std::endl; exit(EXIT_FAILURE); }
    c.fixpoint = fp;
  }
  in.read_size("-wac1_threshold", c.wac1_threshold);
  std::string evar, eval;
  if (in.read_string("-eps_var_order", evar)) c.eps_var_order = evar;
  if (in.read_string("-eps_value_order", eval)) c.eps_value_order = eval;
  if (evar.empty() != eval.empty()) {
    printf("-eps_var_order and -eps_value_order must be specified together.\n");
    exit(EXIT_FAILURE);
  }
  in.read_string("-version", c.version);
  in.read_string("-hardware", c.hardware);
  in.read_int("-gpus", c.gpus);
  in.read_int("-tpb", c.threads_per_block);
  if (in.tokens.size() <= in.tokens_read) usage_and_exit(argv[0]);
  c.problem_path = in.tokens.back();
  return c;
}

void print_commandline(const Config& c, const char* prog) {      // config.hpp:168-207
  printf("%s -t %zu %s-n %zu %s%s%s%s", prog, c.timeout_ms, c.print_intermediate_solutions ? "-a " : "", c.stop_after_n_solutions,
         c.print_intermediate_solutions ? "-i " : "", c.free_search ? "-f " : "", c.print_statistics ? "-s " : "", c.print_ast ? "-ast " : "");
  for (int i = 0; i < c.verbose; ++i) printf("-v ");
  printf("-arch %s -or %zu -sub %d -subfactor %zu -stack %zu ", c.arch.c_str(), c.or_nodes, c.subproblems_power, c.subproblems_factor, c.stack_kb);
  if (c.only_global_memory) printf("-globalmem ");
  if (c.disable_simplify) printf("-disable_simplify ");
  if (c.force_ternarize) printf("-force_ternarize ");
  if (c.disable_network_analysis) printf("-disable_network_analysis ");
  if (c.fixpoint != "auto") printf("-fp %s ", c.fixpoint.c_str());
  if (c.fixpoint == "wac1" || c.fixpoint == "wac1_active") printf("-wac1_threshold %zu ", c.wac1_threshold);
  printf("-seed %zu -eps_var_order %s -eps_value_order %s ", c.seed, c.eps_var_order.c_str(), c.eps_value_order.c_str());
  if (!c.version.empty()) printf("-version %s ", c.version.c_str());
  if (!c.hardware.empty()) printf("-hardware '%s' ", c.hardware.c_str());
  printf("-cutnodes %zu -gpus %d %s", c.stop_after_n_nodes, c.gpus, c.problem_path.c_str());
}

struct Stat {
  bool on;
  void s(const char* k, const char* v) const { if (on) printf("%%%%%%mzn-stat: %s=\"%s\"\n", k, v); }
  void u(const char* k, uint64_t v) const { if (on) printf("%%%%%%mzn-stat: %s=%" PRIu64 "\n", k, v); }
  void i(const char* k, int64_t v) const { if (on) printf("%%%%%%mzn-stat: %s=%" PRId64 "\n", k, v); }
  void d(const char* k, double v) const { if (on) printf("%%%%%%mzn-stat: %s=%lf\n", k, v != v ? 0.0 : v); }
  void end() const { if (on) printf("%%%%%%mzn-stat-end\n"); }
  void mem(int verbose, const char* k, uint64_t bytes) const {       // statistics.hpp:304-319
    u(k, bytes);
    if (on && verbose) {
      if (bytes < 1000 * 1000) printf("%%   [%.2fKB]\n", bytes / 1e3);
      else if (bytes < 1000ull * 1000 * 1000) printf("%%   [%.2fMB]\n", bytes / 1e6);
      else printf("%%   [%.2fGB]\n", bytes / 1e9);
    }
  }
};

double to_sec(int64_t ns) { return (double)(ns / 1000 / 1000) / 1000.; }   // statistics.hpp:325-327 (ms resolution)

std::atomic<int> g_signal{0};
volatile int32_t g_stop = 0;
void (*g_prev_int)(int) = nullptr;
void (*g_prev_term)(int) = nullptr;
void on_signal(int sig) {                                   // common_solving.hpp:60-71
  std::signal(SIGINT, on_signal);
  std::signal(SIGTERM, on_signal);
  g_signal = 1;
  g_stop = 1;
  if (sig == SIGINT && g_prev_int && g_prev_int != SIG_DFL && g_prev_int != SIG_IGN) g_prev_int(sig);
  if (sig == SIGTERM && g_prev_term && g_prev_term != SIG_DFL && g_prev_term != SIG_IGN) g_prev_term(sig);
}

int order_of(const std::string& s, bool var) {
  if (var) {
    if (s == "input_order" || s == "random") return TB_VAR_INPUT_ORDER;
    if (s == "first_fail") return TB_VAR_FIRST_FAIL;
    if (s == "anti_first_fail") return TB_VAR_ANTI_FIRST_FAIL;
    if (s == "smallest") return TB_VAR_SMALLEST;
    if (s == "largest") return TB_VAR_LARGEST;
    return -1;
  }
  if (s == "min") return TB_VAL_MIN;
  if (s == "max") return TB_VAL_MAX;
  if (s == "split") return TB_VAL_SPLIT;
  if (s == "reverse_split") return TB_VAL_REVERSE_SPLIT;
  return -1;
}

bool ends_with(const std::string& s, const char* suf) {
  size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

void print_config_stats(const Config& c, const tb_stats& st) {       // config.hpp:237-266
  printf("%%%%%%mzn-stat: problem_path=\"%s\"\n", c.problem_path.c_str());
  printf("%%%%%%mzn-stat: solver=\"Turbo\"\n");
  printf("%%%%%%mzn-stat: version=\"%s\"\n", c.version.empty() ? "1.3.0-b200" : c.version.c_str());
  printf("%%%%%%mzn-stat: hardware=\"%s\"\n", c.hardware.empty() ? "unspecified" : c.hardware.c_str());
  printf("%%%%%%mzn-stat: arch=\"%s\"\n", "barebones");
  static const char* kFp[] = {"ac1", "wac1", "ac1_active", "wac1_active"};
  printf("%%%%%%mzn-stat: fixpoint=\"%s\"\n", kFp[st.fixpoint_in_effect & 3]);
  printf("%%%%%%mzn-stat: subproblems_factor=%zu\n", c.subproblems_factor);
  if (st.fixpoint_in_effect & 1) printf("%%%%%%mzn-stat: wac1_threshold=%zu\n", c.wac1_threshold);
  printf("%%%%%%mzn-stat: seed=%zu\n", c.seed);
  printf("%%%%%%mzn-stat: eps_var_order=\"%s\"\n", c.eps_var_order.c_str());
  printf("%%%%%%mzn-stat: eps_value_order=\"%s\"\n", c.eps_value_order.c_str());
  printf("%%%%%%mzn-stat: free_search=\"%s\"\n", c.free_search ? "yes" : "no");
  printf("%%%%%%mzn-stat: or_nodes=%zu\n", c.or_nodes);
  printf("%%%%%%mzn-stat: timeout_ms=%zu\n", c.timeout_ms);
  printf("%%%%%%mzn-stat: threads_per_block=%d\n", st.threads_per_block);
  printf("%%%%%%mzn-stat: stack_size=%zu\n", c.stack_kb * 1000);
  {
    tb_device_info di;                                               // config.hpp:258-260
    if (tb_get_device_info(0, &di) == TB_OK) printf("%%%%%%mzn-stat: cuda_version=%d\n", di.cuda_runtime_version);
  }
  printf("%%%%%%mzn-stat: cuda_architecture=%d\n", 1000);
  printf("%%%%%%mzn-stat: num_gpus=%d\n", c.gpus);
  printf("%%%%%%mzn-stat: cutnodes=%zu\n", c.stop_after_n_nodes);
}

void print_solver_stats(const Stat& S, const tb_stats& st, int verbose, size_t variables, size_t constraints, int64_t init_ns, int64_t overall_ns) {
  // statistics.hpp:338-371; per-block timers are sums over blocks divided by num_blocks (:330-332)
  (void)verbose;
  const int nb = std::max(1, st.num_blocks);
  S.i("num_blocks", st.num_blocks);
  S.u("nodes", st.nodes);
  S.u("failures", st.fails);
  S.u("variables", variables);
  S.u("propagators", constraints);
  S.i("peakDepth", st.depth_max);
  S.d("initTime", to_sec(init_ns));
  S.d("solveTime", to_sec(overall_ns));
  S.u("num_solutions", st.solutions);
  S.u("eps_num_subproblems", st.eps_num_subproblems);
  S.u("eps_solved_subproblems", st.eps_solved_subproblems);
  S.u("eps_skipped_subproblems", st.eps_skipped_subproblems);
  if (st.eps_stolen_subproblems) S.u("eps_stolen_subproblems", st.eps_stolen_subproblems);
  if (st.eps_split_subproblems) { S.u("eps_split_subproblems", st.eps_split_subproblems); S.u("eps_split_parts_solved", st.eps_split_parts_solved); }
  S.u("num_blocks_done", st.num_blocks_done);
  S.u("fixpoint_iterations", st.fixpoint_iterations);
  S.u("num_deductions", st.num_deductions);
  S.d("cumulative_time_block_sec", to_sec(st.cumulative_time_block_ns));
  S.d("deductions_per_block_second", (double)(st.num_deductions / (uint64_t)nb) / to_sec(st.cumulative_time_block_ns));
  S.d("solve_time", to_sec(overall_ns / nb));
  S.d("search_time", to_sec(st.timers_ns[TB_TIMER_SEARCH] / nb));
  S.d("fixpoint_time", to_sec(st.timers_ns[TB_TIMER_FIXPOINT] / nb));
  S.d("transfer_cpu2gpu_time", to_sec(st.timers_ns[TB_TIMER_TRANSFER_CPU2GPU] / nb));
  S.d("transfer_gpu2cpu_time", to_sec(st.timers_ns[TB_TIMER_TRANSFER_GPU2CPU] / nb));
  S.d("select_fp_functions_time", to_sec(st.timers_ns[TB_TIMER_SELECT_FP_FUNCTIONS] / nb));
  S.d("wait_cpu_time", to_sec(st.timers_ns[TB_TIMER_WAIT_CPU] / nb));
  S.d("dive_time", to_sec(st.timers_ns[TB_TIMER_DIVE] / nb));
  S.d("best_obj_time", to_sec(st.timers_ns[TB_TIMER_LATEST_BEST_OBJ_FOUND]));
  S.d("first_block_idle_time", to_sec(st.timers_ns[TB_TIMER_FIRST_BLOCK_IDLE]));
  // extensions (not in the reference): device-side throughput of the hot path
  S.d("kernel_time_sec", st.kernel_ms / 1e3);
  S.u("bounds_narrowed", st.bounds_narrowed);
  if (st.kernel_ms > 0) {
    S.d("propagations_per_second", (double)st.num_deductions / (st.kernel_ms / 1e3));
    S.d("nodes_per_second", (double)st.nodes / (st.kernel_ms / 1e3));
  }
}

void print_final_separator(const tb_stats& st) {                      // statistics.hpp:394-412
  if (st.solutions > 0) { if (st.exhaustive) printf("==========\n"); }
  else if (st.exhaustive) printf("=====UNSATISFIABLE=====\n");
  else printf("=====UNKNOWN=====\n");
}

const char* mem_name(int k) {
  return k == TB_MEM_GLOBAL ? "global" : k == TB_MEM_STORE_SHARED ? "store_shared" : k == TB_MEM_TCN_SHARED ? "tcn_shared" : "store_cluster";
}

}  // namespace

int main(int argc, char** argv) {
  const auto start = std::chrono::steady_clock::now();
  auto since_ns = [&]() { return (int64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - start).count(); };
  Config config = parse_args(argc, argv);
  const Stat S{config.print_statistics};
  if (config.print_statistics) {
    printf("%%%%%%mzn-stat: command_line=\"");
    print_commandline(config, argv[0]);
    printf("\"\n");
  }
  if (config.arch == "cpu") {
    std::cerr << "-arch cpu: this build solves on B200 GPUs only and has no CPU fallback (the CPU restatement lives in oracle/ as test infrastructure)." << std::endl;
    return EXIT_FAILURE;
  }
  if (config.arch != "barebones" && config.verbose)
    printf("%% WARNING: -arch %s is served by the barebones dive-and-solve architecture.\n", config.arch.c_str());

  // ---- preprocessing (CP::preprocess, common_solving.hpp:605-637) ----------------------------------------
  tb_model* model = nullptr;
  tb_status rc;
  if (ends_with(config.problem_path, ".fzn")) rc = tb_model_load_fzn(&model, config.problem_path.c_str(), config.disable_simplify ? 1u : 0u);
  else if (ends_with(config.problem_path, ".tnf")) rc = tb_model_load_tnf(&model, config.problem_path.c_str());
  else if (ends_with(config.problem_path, ".xml")) { std::cerr << "XCSP3 input is not supported by this build." << std::endl; return EXIT_FAILURE; }
  else { printf("ERROR: Unknown input format for the file %s [supported extension: .fzn and .tnf].\n", config.problem_path.c_str()); return EXIT_FAILURE; }
  if (rc != TB_OK) {
    std::cerr << "Could not parse input file." << std::endl;
    std::cerr << tb_last_error() << std::endl;
    return EXIT_FAILURE;
  }
  if (config.verbose) printf("%% Input file parsed\n");
  if (config.eps_var_order != "default") {
    int vo = order_of(config.eps_var_order, true), va = order_of(config.eps_value_order, false);
    if (vo < 0) { printf("Unrecognized option `-eps_var_order %s`\n", config.eps_var_order.c_str()); return EXIT_FAILURE; }
    if (va < 0) { printf("Unrecognized option `-eps_value_order %s`\n", config.eps_value_order.c_str()); return EXIT_FAILURE; }
    tb_model_push_eps_strategy(model, vo, va);
  }
  const tb_problem* pb = tb_model_problem(model);
  S.u("parsed_variables", (uint64_t)tb_model_num_parsed_variables(model));
  S.u("parsed_constraints", (uint64_t)tb_model_num_parsed_constraints(model));
  S.s("abstract_domain", "pir_itv32_z");
  S.s("entailed_prop_removal", "deactivated");
  S.u("tcn_variables", (uint64_t)pb->nvars);
  S.u("tcn_constraints", (uint64_t)pb->nprops);
  // TNF simplifier (preprocess_tcn, common_solving.hpp:538-565); its root fixpoint runs on the GPU
  if (!config.disable_simplify && !tb_model_root_failed(model)) {
    if (tb_device_count() <= 0) { std::cerr << "No CUDA device found: turbo (B200 build) has no CPU fallback." << std::endl; return EXIT_FAILURE; }
    int32_t dev = 0;
    tb_simplify_stats ss;
    rc = tb_model_simplify(model, tb_fixpoint_on_device, &dev, &ss);
    if (rc != TB_OK) { std::cerr << "tb_model_simplify failed: " << tb_last_error() << std::endl; return EXIT_FAILURE; }
    pb = tb_model_problem(model);
    S.i("preprocessing_iterations", ss.iterations);
    S.i("preprocessing_icse_eliminated_constraints", ss.eliminated_icse);
    S.i("preprocessing_algsimp_eliminated_constraints", ss.eliminated_equalities);
    S.i("preprocessing_algsimp_eliminated_eq_constraints", ss.merged_variables);
    S.i("preprocessing_entailment_eliminated_constraints", ss.eliminated_entailed);
    S.i("preprocessing_functional_eliminated_constraints", ss.eliminated_functional);
    S.i("preprocessing_eliminated_variables", ss.eliminated_variables + ss.merged_variables);
    S.u("preprocessed_tcn_variables", (uint64_t)pb->nvars);
    S.u("preprocessed_tcn_constraints", (uint64_t)pb->nprops);
    if (config.verbose) printf("%% Formula simplified.\n");
  }
  const int64_t init_ns = since_ns();
  S.d("preprocessing_time", to_sec(init_ns));
  S.end();

  tb_stats total;
  memset(&total, 0, sizeof(total));
  total.exhaustive = 1;
  total.fixpoint_in_effect = config.fixpoint == "ac1" ? TB_FP_AC1 : config.fixpoint == "ac1_active" ? TB_FP_AC1_ACTIVE : config.fixpoint == "wac1_active" ? TB_FP_WAC1_ACTIVE : TB_FP_WAC1;
  const int okind = tb_model_objective_kind(model);
  if (tb_model_root_failed(model)) {                      // barebones :474-478
    print_final_separator(total);
    if (config.print_statistics) { print_config_stats(config, total); print_solver_stats(S, total, config.verbose, pb->nvars, pb->nprops, init_ns, since_ns()); S.end(); }
    return 0;
  }

  // ---- configure the GPUs (configure_gpu_barebones, barebones :527-606) -----------------------------------
  const int ndev = tb_device_count();
  if (ndev <= 0) { std::cerr << "No CUDA device found: turbo (B200 build) has no CPU fallback." << std::endl; return EXIT_FAILURE; }
  const int G = std::max(1, std::min(config.gpus, ndev));
  if (config.gpus > ndev && config.verbose) printf("%% WARNING: -gpus %d is more than the %d visible devices.\n", config.gpus, ndev);
  config.gpus = G;
  std::vector<tb_solver*> solvers((size_t)G, nullptr);
  if (config.stack_kb) {                                             // -stack (barebones :588-593)
    for (int g = 0; g < G; ++g)
      if (tb_set_stack_limit(g, (uint64_t)config.stack_kb * 1000) != TB_OK) { std::cerr << "-stack: " << tb_last_error() << std::endl; return EXIT_FAILURE; }
  }
  S.i("subproblems_power", config.subproblems_power);
  {
    // one host thread per GPU: context creation, uploads and attribute queries of the G devices overlap
    std::vector<tb_status> crc((size_t)G, TB_OK);
    std::vector<std::string> cerr_msg((size_t)G);
    auto make = [&](int g) {
      tb_options o;
      memset(&o, 0, sizeof(o));
      // default: the active-set kind (same fixpoints and search tree, 1.7 x the nodes/s on the large networks); the engine
      // itself runs the plain sweeps where tracking does not pay (small tables, store outside shared memory)
      o.fixpoint = config.fixpoint == "ac1" ? TB_FP_AC1 : config.fixpoint == "ac1_active" ? TB_FP_AC1_ACTIVE : config.fixpoint == "wac1" ? TB_FP_WAC1 : TB_FP_WAC1_ACTIVE;
      o.wac1_threshold = (int32_t)config.wac1_threshold;
      o.subproblems_power = config.subproblems_power;
      o.subproblems_factor = (int32_t)config.subproblems_factor;
      o.or_blocks = (int32_t)config.or_nodes;
      o.threads_per_block = config.threads_per_block;
      o.mem_kind = config.only_global_memory ? TB_MEM_GLOBAL : TB_MEM_AUTO;
      o.verbose = config.verbose;
      o.gpu_rank = g; o.gpu_world = G; o.device = g;
      o.cutnodes = config.stop_after_n_nodes;
      o.seed = config.seed;
      if (config.timeout_ms) {
        int64_t left = (int64_t)config.timeout_ms - since_ns() / 1000000;
        o.timeout_ms = (uint64_t)std::max<int64_t>(1, left);
      }
      crc[(size_t)g] = tb_create(&solvers[(size_t)g], pb, &o);
      if (crc[(size_t)g] != TB_OK) cerr_msg[(size_t)g] = tb_last_error();
    };
    std::vector<std::thread> th;
    for (int g = 1; g < G; ++g) th.emplace_back(make, g);
    make(0);
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g)
      if (crc[(size_t)g] != TB_OK) { std::cerr << "tb_create failed: " << cerr_msg[(size_t)g] << std::endl; return EXIT_FAILURE; }
  }
  if (G > 1) {
    rc = tb_link_peers(solvers.data(), G);
    if (rc != TB_OK) { std::cerr << "tb_link_peers failed: " << tb_last_error() << std::endl; return EXIT_FAILURE; }
  }
  tb_stats cfg;
  tb_get_config(solvers[0], &cfg);
  {
    // barebones :579-581 prints the device heap it raised and the global memory size; here every byte is allocated by
    // the host before the launch, so heap_memory is what the solvers hold
    tb_device_info di;
    uint64_t held = 0;
    for (int g = 0; g < G; ++g) { tb_stats c; tb_get_config(solvers[(size_t)g], &c); held += c.device_bytes; }
    if (tb_get_device_info(0, &di) == TB_OK) {
      S.mem(config.verbose, "heap_memory", held);
      S.mem(config.verbose, "total_global_mem_bytes", di.total_global_mem_bytes);
      S.mem(config.verbose, "stack_memory", di.stack_limit_bytes * (uint64_t)cfg.num_blocks * (uint64_t)cfg.threads_per_block);
    }
  }
  S.mem(config.verbose, "mem_per_block", cfg.store_bytes * 2 + 320000);
  S.i("num_blocks", (int64_t)cfg.num_blocks * G);
  S.s("memory_configuration", mem_name(cfg.mem_kind));
  S.mem(config.verbose, "shared_mem", cfg.shared_bytes);
  S.mem(config.verbose, "store_mem", cfg.store_bytes);
  S.mem(config.verbose, "propagator_mem", cfg.prop_bytes);
  S.end();
  fflush(stdout);

  // ---- solve ----------------------------------------------------------------------------------------------------
  g_prev_int = std::signal(SIGINT, on_signal);
  g_prev_term = std::signal(SIGTERM, on_signal);
  const size_t nv = (size_t)std::max(1, pb->nvars);
  std::vector<std::vector<int32_t>> blb((size_t)G, std::vector<int32_t>(nv)), bub((size_t)G, std::vector<int32_t>(nv));
  std::vector<int32_t> has((size_t)G, 0), exh((size_t)G, 0);
  std::vector<tb_stats> sts((size_t)G);
  std::vector<tb_status> rcs((size_t)G, TB_OK);
  std::vector<std::string> cerr_solve((size_t)G);
  if (config.verbose) printf("%% GPU kernel started, starting solving...\n");
  const int64_t kernel_start_ns = since_ns();
  if (config.verbose) printf("%% start-up: preprocessing (parse, ternarise, simplify on the GPU) %.3f s, engine creation %.3f s\n", to_sec(init_ns), to_sec(kernel_start_ns - init_ns));
  // -i / -a: improving solutions are printed as they are found (the reference's barebones architecture cannot,
  // barebones :465-467; its `gpu` architecture does it with a consumer thread, gpu_dive_and_solve.hpp:100-132)
  std::atomic<bool> search_over{false};
  std::mutex print_mutex;
  bool printed_any = false;
  int32_t last_printed = 0;                 // objective (of the minimised TNF variable) of the latest printed solution
  std::vector<int32_t> plb(nv), pub(nv);
  auto print_if_improving = [&](const int32_t* lb, const int32_t* ub, int32_t objective) {
    std::lock_guard<std::mutex> lock(print_mutex);
    if (printed_any && (pb->obj_var < 0 || objective >= last_printed)) return false;
    int bad = tb_model_check_solution(model, lb, ub);
    if (bad < 0) bad = tb_model_check_tnf(model, lb);
    if (bad != 0) { std::cerr << "% ERROR: an intermediate solution violates " << bad << " constraint(s)" << std::endl; return false; }
    size_t n = tb_model_format_solution(model, lb, ub, nullptr, 0);
    std::string text(n + 1, '\0');
    tb_model_format_solution(model, lb, ub, &text[0], n + 1);
    fputs(text.c_str(), stdout);
    printf("----------\n");
    fflush(stdout);
    printed_any = true; last_printed = objective;
    return true;
  };
  std::thread consumer;
  if (config.print_intermediate_solutions) {
    for (int g = 0; g < G; ++g)
      if (tb_stream_solutions(solvers[(size_t)g], 16) != TB_OK) { std::cerr << "tb_stream_solutions: " << tb_last_error() << std::endl; return EXIT_FAILURE; }
    consumer = std::thread([&]() {
      for (;;) {
        const bool last_round = search_over.load();
        bool got = false;
        for (int g = 0; g < G; ++g) {
          int32_t obj = 0;
          while (tb_poll_solution(solvers[(size_t)g], plb.data(), pub.data(), &obj, nullptr) == 1) { got = true; print_if_improving(plb.data(), pub.data(), obj); }
        }
        if (last_round) break;
        if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(2));
      }
    });
  }
  if (config.timeout_ms) {
    // what is left of -t now that the engines exist (their creation overlapped on G host threads but is not free)
    const int64_t left = (int64_t)config.timeout_ms - since_ns() / 1000000;
    for (int g = 0; g < G; ++g) tb_set_timeout(solvers[(size_t)g], (uint64_t)std::max<int64_t>(1, left));
  }
  {
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g)
      th.emplace_back([&, g]() {
        rcs[(size_t)g] = tb_solve(solvers[(size_t)g], &g_stop, blb[(size_t)g].data(), bub[(size_t)g].data(), &has[(size_t)g], &exh[(size_t)g], &sts[(size_t)g]);
        if (rcs[(size_t)g] != TB_OK) cerr_solve[(size_t)g] = tb_last_error();      // (the error text is per thread)
      });
    for (auto& t : th) t.join();
  }
  search_over.store(true);
  if (consumer.joinable()) consumer.join();
  int exit_code = 0;
  for (int g = 0; g < G; ++g)
    if (rcs[(size_t)g] != TB_OK) { std::cerr << "tb_solve failed on GPU " << g << ": " << cerr_solve[(size_t)g] << std::endl; exit_code = EXIT_FAILURE; }

  // ---- reduce over GPUs (reduce_blocks, barebones :1033-1067): the same pack / reduce pair a one-process-per-GPU
  // launcher gathers with (tb_result_pack + NCCL gather + tb_result_reduce) ------------------------------------------
  int best = -1;
  {
    const size_t rsz = tb_result_size(solvers[0]);
    std::vector<char> packed(rsz * (size_t)G);
    for (int g = 0; g < G; ++g)
      if (tb_result_pack(solvers[(size_t)g], packed.data() + rsz * (size_t)g, rsz) != TB_OK) { std::cerr << "tb_result_pack: " << tb_last_error() << std::endl; return EXIT_FAILURE; }
    int32_t any = 0, exh_all = 0;
    if (tb_result_reduce(packed.data(), G, rsz, nullptr, nullptr, &any, &exh_all, &total, &best) != TB_OK) { std::cerr << "tb_result_reduce: " << tb_last_error() << std::endl; return EXIT_FAILURE; }
  }
  total.eps_num_subproblems = sts[0].eps_num_subproblems;
  total.threads_per_block = sts[0].threads_per_block;
  if (g_signal || (config.timeout_ms && since_ns() / 1000000 >= (int64_t)config.timeout_ms)) total.exhaustive = 0;
  if (best >= 0) {
    // time-to-optimum includes the time before the kernel started (barebones :501)
    total.timers_ns[TB_TIMER_LATEST_BEST_OBJ_FOUND] = sts[(size_t)best].timers_ns[TB_TIMER_LATEST_BEST_OBJ_FOUND] + kernel_start_ns;
    if (total.timers_ns[TB_TIMER_FIRST_BLOCK_IDLE] != 0) total.timers_ns[TB_TIMER_FIRST_BLOCK_IDLE] += kernel_start_ns;
    const int32_t* lb = blb[(size_t)best].data();
    const int32_t* ub = bub[(size_t)best].data();
    // every reported solution is re-checked against the FlatZinc constraints (north_star)
    int bad = tb_model_check_solution(model, lb, ub);
    if (bad < 0) bad = tb_model_check_tnf(model, lb);
    if (bad != 0) {
      std::cerr << "% ERROR: the solution violates " << bad << " constraint(s): " << tb_last_error() << std::endl;
      exit_code = 2;
    }
    // (with -i / -a the best solution has usually been printed already, when it was found)
    const bool already = config.print_intermediate_solutions && printed_any && (pb->obj_var < 0 || lb[(size_t)pb->obj_var] >= last_printed);
    if (!already) {
      size_t n = tb_model_format_solution(model, lb, ub, nullptr, 0);
      std::string text(n + 1, '\0');
      tb_model_format_solution(model, lb, ub, &text[0], n + 1);
      fputs(text.c_str(), stdout);
      printf("----------\n");
    }
  }
  print_final_separator(total);
  if (config.print_statistics) {
    print_config_stats(config, total);
    print_solver_stats(S, total, config.verbose, (size_t)pb->nvars, (size_t)pb->nprops, init_ns, since_ns());
    if (okind >= 0 && best >= 0) {
      const int uv = tb_model_user_objective_var(model);
      // lb for minimisation, ub of the original variable for maximisation (statistics.hpp:378-388)
      printf("%%%%%%mzn-stat: objective=%d\n", okind == 0 ? blb[(size_t)best][(size_t)uv] : bub[(size_t)best][(size_t)uv]);
    }
    S.end();
  }
  for (tb_solver* s : solvers) tb_destroy(s);
  tb_model_destroy(model);
  return exit_code;
}
