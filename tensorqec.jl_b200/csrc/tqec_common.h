// Internal definitions shared by the translation units of libtqec_cuda.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "tqec.h"

namespace tqec {

void set_error(const char *fmt, ...);

// NVTX range covering a scope: the library's entry points show up as named ranges in Nsight Systems / Compute timelines
// (plan compile, plan create, decode, fused Monte-Carlo pipeline).  Header-only NVTX v3: no link dependency; without a
// profiler attached a range costs a few nanoseconds.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

#define TQEC_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      tqec::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return TQEC_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define TQEC_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      tqec::set_error(__VA_ARGS__);    \
      return TQEC_ERR_INVALID;         \
    }                                  \
  } while (0)

#define TQEC_H_FAST 14  // internal (reserved slot of the ABI header): 0 or 1 + offset of the step's fast record in ints
#define TQEC_H_AM 15    // internal (reserved slot): offset in ints of the traceback (assignment, mask) pairs

static inline int words_for(int nbits) { return nbits <= 0 ? 1 : (nbits + 63) / 64; }

// Device view of a compiled schedule (passed to kernels by value).
struct PlanDev {
  const int32_t *hdr;      // n_steps * TQEC_HDR_INTS
  const int32_t *ints;
  const double *tables;
  const int32_t *bp_off;   // n_steps + 1: word offset of each step's back-pointer block inside a team's scratch
  const int32_t *obs_slot; // n_obs
  int32_t n_steps, n_vars, n_checks, n_obs;
  int32_t w_max;           // widest state in bits
  int32_t sg_log2;         // log2(shots per team)
  int32_t nsw, ncw;        // syndrome / configuration words per shot
  int32_t bp_words;        // back-pointer words per team
  int32_t n_ints, n_tables; // pool sizes (for the shared-memory copies of warp teams)
  int32_t sub_minor;       // 1: state entry sigma of shot s at (sigma << sg) | s ("shot in lane"), else (s << w) | sigma
  int32_t defer;           // 1: warp teams run 32 forward sweeps, then trace 32 shots in parallel (one lane each)
  int32_t off_states, off_ints, off_tables, off_words;  // warp-team shared-memory layout (bytes from the dynamic array)
};

// Device view of the in-place patch sweep of a plan (tqec_sweep.cu; tables described in tensorqec.jl_b200/sweep.py).
struct SweepDev {
  const int32_t *rec, *tb;
  const uint32_t *lanetab;
  const double *tvals, *head_state;
  const uint64_t *head_cfg;
  int32_t head_bits[16];
  int32_t out_index[16];            // state index of output entry i (max-plus: entry 0 only)
  int32_t n_ss, W, sg, nh, nsw, ncw, bp_words, n_tvals, n_obs;
  int32_t sync_mode;                // CTA barrier between teams: 0 none, 1 per group of 32 shots, 2 per pass
  int32_t head_tma;                 // 1: head rows are stored pre-swizzled and fetched by cp.async.bulk (TMA)
  int32_t grp;                      // shots per deferred-traceback group (2^sg .. 32): the back-pointer ring of a team holds this many
  int32_t off_states, off_rec, off_lanetab, off_tvals, off_words, words_bytes;   // shared-memory layout (bytes)
  int32_t off_tb;                   // >= 0: the traceback records have a shared-memory copy at this offset (max-plus plans)
};

// Device view of the global-memory lowering of a plan (tqec_wide.cu; tables described in tensorqec.jl_b200/wide.py).
struct WideDev {
  const int32_t *pass_hdr, *step_hdr, *ints;
  const double *tables;
  int32_t n_pass, n_steps, w_cap, t_max, n_obs, nsw;
  int32_t obs_pos[16];
  const int32_t *bf_ints;   // butterfly encoding (k_wide_bf); bf_off is kept on the host (one launch per pass)
  const double *bf_vals;
  double out_mant;          // marginals are multiplied by this on the way out (1.0 unless butterfly passes are in use)
};

}  // namespace tqec

struct tqec_plan {
  tqec::PlanDev dev;
  int device;
  int semiring;
  int log2_scale;        // sum-product: marginals are multiplied by 2^log2_scale on the way out
  int team_threads;
  int warp_teams;        // 1: k_frontier_warp (a team is a warp, tables in shared memory); 0: k_frontier_cta
  int teams_per_cta;
  int layout;            // warp-team index layout: 0 standard, 1 wide (64 / 128 entries per thread), 2 shot-minor
  int shots_per_team;
  int smem_bytes;
  int grid_max;          // persistent grid size (teams resident on the whole GPU)
  int teams_per_sm;
  int sm_count;
  double candidates_per_shot;
  int64_t launches;
  // in-place patch sweep (optional)
  tqec::SweepDev sw;
  int has_sweep, sw_teams, sw_smem, sw_maxt, sw_ext;   // sw_ext: the plan uses fresh-pin shapes (extended instantiation of k_sweep)
  // fully tabulated plan (n_checks <= 16): outputs of every syndrome, filled once by the plan's own kernels
  int has_table;
  uint64_t *d_tab_corr;
  double *d_tab_out;
  int32_t *d_tab_arg;
  void *d_sw[8];
  uint32_t *d_sw_bp;
  // global-memory executor (optional): two state arrays of wd_batch << w_cap doubles, allocated at the first decode
  tqec::WideDev wd;
  int has_wide, wd_smem, wd_grid;
  int wd_bf_smem, wd_bf_grid;      // k_wide_bf launch configuration (0: no butterfly passes)
  int32_t *wd_bf_off;              // host, per pass: offset of its block in bf_ints or -1 (NULL: no butterfly passes)
  void *d_wd[6];
  double *d_wd_state[2];
  unsigned long long *d_wd_max;   // dynamic rescaling: per-shot maxima of the last two passes
  int32_t *d_wd_exp;              // ... and per-shot accumulated exponents
  int wd_dynamic;
  int64_t wd_batch;
  int wd_full;            // the state arrays already take all the memory the executor may use
  void *d_hdr, *d_ints, *d_tables, *d_bp_off, *d_obs_slot;
  uint32_t *d_bp;        // back-pointer scratch: grid_max * bp_words
  void *d_mc;            // scratch of the fused Monte-Carlo pipeline (tqec_mc_run), reused across calls
  size_t mc_cap;
  // host staging for the host-pointer entry points
  void *d_io[4];
  size_t io_cap[4];
  cudaStream_t stream;
  // host-pointer entry points: copy-in / copy-out streams and per-chunk events of the three-stage pipeline
  cudaStream_t s_in, s_out;
  cudaEvent_t ev_in[2], ev_cmp[2], ev_out[2];
  void *h_pin[3];        // pinned host staging of the byte-per-bit entry points (two slots each)
  size_t h_pin_cap[3];
};

struct tqec_gf2 {
  int device;
  int rows, cols;
  int rw, cw;            // words per packed output (rows) / input (cols)
  uint64_t *d_rows;      // rows * cw
  void *d_io[4];
  size_t io_cap[4];
  cudaStream_t stream;
  int64_t launches;
};

namespace tqec {
int ensure_cap(void **ptr, size_t *cap, size_t bytes);
int launch_decode(tqec_plan *plan, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_out,
                  int32_t *d_argmax, cudaStream_t stream);
int launch_sweep(tqec_plan *plan, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_out, int32_t *d_argmax,
                 cudaStream_t stream);
int sweep_create(tqec_plan *p, const tqec_plan_desc *d, const cudaDeviceProp &prop);
void sweep_destroy(tqec_plan *p);
int wide_create(tqec_plan *p, const tqec_plan_desc *d, const cudaDeviceProp &prop);
void wide_destroy(tqec_plan *p);
int comm_allreduce_dev(tqec_comm *c, unsigned long long *d_counts, cudaStream_t stream);
int launch_wide(tqec_plan *plan, const uint64_t *d_synd, int64_t B, double *d_out, int32_t *d_argmax, cudaStream_t stream,
                int32_t *d_log2 = nullptr);
int launch_gf2_apply(tqec_gf2 *m, const uint64_t *d_in, int64_t B, uint64_t *d_out, cudaStream_t stream);
int launch_sample(int model, int n_sites, const double *d_p, uint64_t seed, int64_t shot_offset, int64_t B,
                  uint64_t *d_err, int words, cudaStream_t stream);
int launch_flags(tqec_gf2 *L, const int32_t *d_row_class, const uint64_t *d_e1, const uint64_t *d_e2, int64_t B,
                 uint8_t *d_flags, unsigned long long *d_counts, cudaStream_t stream);
}  // namespace tqec
