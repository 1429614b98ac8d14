// Host-side lowering of a decoding factor graph to the kernel schedules of libtqec_cuda.so (internal header).
// The algorithms are documented where they were first written down -- tensorqec.jl_b200/schedule.py (frontier
// recurrence, step tables), sweep.py (in-place patch sweep), wide.py (global-memory passes) -- and the Python versions
// stay in the repository as the test oracle of this file: tests/test_lower_cpp.py compares every emitted table.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "tqec.h"

namespace tqec {
namespace lower {

struct Factor {
  std::vector<int> vars;
  std::vector<double> table;  // 2^rank entries, index a = sum_j a_j << j (first variable fastest)
};

struct Check {
  std::vector<int> vars;
  int kind;   // 0 = "syn" (clamped by syndrome bit `index`), 1 = "obs" (open output axis `index`)
  int index;
};

struct Step {
  int factor = 0;
  std::vector<int> vars;
  int w_in = 0, w_out = 0;
  std::vector<int> opened;                    // check ids
  std::vector<std::pair<int, int>> closed;    // (slot in the full index, syndrome bit), ascending slot
  std::vector<int> perm;                      // perm[b] = slot in the full index of output bit b
  std::vector<int64_t> M;                     // (2^r) masks in the full index space
  std::vector<int64_t> a0;                    // (2^n_open) representative candidate per opened pattern, -1 = infeasible
  std::vector<int64_t> ker;                   // kernel candidates, ascending
  std::vector<double> table;                  // (2^r) values in the semiring's domain
  bool quad = false;
};

struct Schedule {
  int semiring = 0, n_vars = 0, n_checks = 0, n_obs = 0;
  std::vector<Step> steps;
  std::vector<int> obs_slot;
  std::vector<int> order;
  std::vector<Factor> factors;                // merged
  std::vector<Check> checks;
  int w_max = 0;
  double cost = 0.0;
  int log2_scale = 0;
  std::vector<int32_t> hdr, ints;
  std::vector<double> tables;
};

struct SweepPlan {
  int semiring = 0, n_vars = 0, n_checks = 0, n_obs = 0;
  int W = 0, sg = 0, head_steps = 0, n_ss = 0, bp_words = 0, conflicts = 0;
  std::vector<int> head_bits;
  std::vector<double> head_state;             // (2^nh, 2^W)
  std::vector<uint64_t> head_cfg;             // (2^nh, 2^W, ncw) for max-plus, (2^nh, 1, ncw) zeros otherwise
  std::vector<int> out_index;
  std::vector<int32_t> rec, tb;
  std::vector<uint32_t> lanetab;
  std::vector<double> tvals;
};

struct WidePlan {
  int semiring = 0, n_vars = 0, n_checks = 0, n_obs = 0;
  int w_cap = 0, w_peak = 0, t_max = 0, n_pass = 0, n_steps = 0, log2_scale = 0;   // w_cap: between passes (HBM); w_peak: any step
  std::vector<int> obs_pos, order;
  double cost = 0.0, bytes_per_shot = 0.0;
  std::vector<int32_t> pass_hdr, step_hdr, ints;
  std::vector<double> tables;
  // butterfly encoding of the passes k_wide_bf can run (tqec_lower_wide.cpp:bf_encode_pass): bf_off[pass] = offset of the
  // pass's block in bf_ints or -1; the product of the normalised-away factor entries is bf_mant * 2^bf_log2
  std::vector<int32_t> bf_off, bf_ints;
  std::vector<double> bf_vals;
  double bf_mant = 1.0;
  int bf_log2 = 0;
};

struct Problem {
  int semiring = 0, n_vars = 0, n_checks = 0, n_obs = 0;
  std::vector<Factor> factors;
  std::vector<Check> checks;
  std::vector<int> order;                     // empty = chosen here
  bool has_order = false;
};

// All functions throw std::runtime_error with a message on invalid input (the ABI layer turns it into an error code).
std::vector<Factor> merge_overlapping(const std::vector<Factor> &factors, int n_vars, const std::vector<Check> &checks,
                                      bool allow_negative = false);
std::vector<int> map_order(const std::vector<Factor> &original, const std::vector<Factor> &merged, const std::vector<int> &order);
std::vector<int> choose_order(const std::vector<Factor> &factors, const std::vector<Check> &checks, int max_starts = 24);
std::pair<int, double> evaluate_order(const std::vector<Factor> &factors, const std::vector<Check> &checks, const std::vector<int> &order);
// fuse: -1 = default (max-plus: pair factors into quad steps unless TQEC_NO_FUSE is set), 0 / 1 = off / on
Schedule lower_schedule(const std::vector<Factor> &factors, const std::vector<Check> &checks, int semiring, int n_vars,
                        int n_checks, int n_obs, const std::vector<int> *order, int max_width, int fuse, bool stable);
// -> false when the plan does not fit the in-place form (the caller keeps the general kernels)
bool lower_sweep(const Schedule &sch, int max_head_bits, SweepPlan &out);
WidePlan lower_wide(const std::vector<Factor> &factors, const std::vector<Check> &checks, int semiring, int n_vars, int n_checks,
                    int n_obs, const std::vector<int> *order, int t_max, int low_bits, double max_drop_bits = 0.0);

}  // namespace lower
}  // namespace tqec
