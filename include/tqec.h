/* libtqec_cuda.so -- C ABI of the B200-native TNMAP / TNMMAP decoding hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns 0 (TQEC_OK) or a negative error code,
 * and records a message retrievable with tqec_last_error() (thread-local).  The caller owns all host buffers; the
 * library owns device memory behind opaque handles.  A handle is bound to ONE device and is not re-entrant
 * (the reference's compiled decoders are not thread-safe either: src/decoding/tndecoder.jl:52); different handles
 * are independent.  `*_dev` variants take DEVICE pointers and enqueue on the given cudaStream_t (passed as void*)
 * without synchronising; the plain variants take HOST pointers and return when the results are in host memory.
 *
 * Bit-packed layout (all GF(2) data): SHOT-MAJOR, ceil(nbits/64) uint64 words per shot, bit k of word w is bit
 * 64*w + k -- the `compresscol` layout of the reference's own packed product (src/codes/mod2.jl:58-71).
 *
 * Reference interfaces replaced (paths relative to the reference repo, TensorQEC.jl v2.2.1):
 *   tqec_plan_compile       compile(decoder, problem)  (factor graph in, plan out)     src/decoding/interfaces.jl:67-79, 129-142
 *   tqec_plan_create        (the same from an already lowered schedule)
 *                           compile(::TNMAP, ::GeneralDecodingProblem)                 src/decoding/tndecoder.jl:46-50
 *                           compile(::TNMMAP, ::IndependentDepolarizingDecodingProblem) src/decoding/tndecoder.jl:97-146
 *                           compile(::TNMMAP, ::DetectorErrorModel)                     src/decoding/tndecoder.jl:186-219
 *                           (the contraction order found by OMEinsum's optimize_code becomes the step order)
 *   tqec_decode_map         decode(::CompiledTNMAP, ::SimpleSyndrome)                   src/decoding/tndecoder.jl:53-57
 *   tqec_decode_marginal    update_syndrome! + ct.code(ct.tensors...) + findmax         src/decoding/tndecoder.jl:148-165, 240-253
 *   tqec_coset_rep          error_pattern + _mixed_integer_programming_for_one_solution src/decoding/tndecoder.jl:167-174, 255-271;
 *                                                                                       src/decoding/ipdecoder.jl:150-169
 *   tqec_sample_errors      random_error_pattern                                        src/decoding/error_model.jl:69-71, 97-117;
 *                                                                                       src/decoding/dem.jl:162-164
 *   tqec_gf2_apply          syndrome_extraction  (model: bitmul!)                       src/decoding/error_model.jl:131-146; src/codes/mod2.jl:44-57
 *   tqec_logical_flags      check_logical_error                                         src/decoding/error_model.jl:161-163, 179-181
 *   tqec_mc_run             multi_round_qec (inner loop + counters)                     src/decoding/threshold.jl:1-19
 */
#ifndef TQEC_H
#define TQEC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQEC_OK 0
#define TQEC_ERR_INVALID (-1)     /* bad argument / malformed schedule */
#define TQEC_ERR_CUDA (-2)        /* CUDA runtime error (message has the CUDA string) */
#define TQEC_ERR_UNSUPPORTED (-3) /* valid request the library cannot serve (e.g. frontier too wide) */
#define TQEC_ERR_NOMEM (-4)

#define TQEC_SEMIRING_MAXPLUS 0 /* TNMAP : (max, +) on log-weights, with traceback */
#define TQEC_SEMIRING_SUMPROD 1 /* TNMMAP: (+, *) on weights, open observable axes  */

#define TQEC_HDR_INTS 16
/* step header fields (int32 each); entries 14 and 15 are reserved (the library overwrites them on its device copy) */
enum {
  TQEC_H_R = 0,       /* number of variables of the absorbed factor                                  */
  TQEC_H_WIN = 1,     /* state width (bits) before the step                                          */
  TQEC_H_NOPEN = 2,   /* checks opened by the step: they take slots w_in .. w_in+n_open-1            */
  TQEC_H_NCLOSE = 3,  /* checks closed by the step                                                   */
  TQEC_H_WOUT = 4,    /* state width after the step = w_in + n_open - n_close                        */
  TQEC_H_NK = 5,      /* candidates per output element (power of two)                                */
  TQEC_H_KB = 6,      /* log2(nk) = back-pointer bits per output element                             */
  TQEC_H_OFF_T = 7,   /* tables[off + pat*nk + k]  : factor value of candidate k for opened pattern  */
  TQEC_H_OFF_ML = 8,  /* ints[off + pat]           : in-state mask of the coset representative       */
  TQEC_H_OFF_MK = 9,  /* ints[off + k]             : in-state mask of kernel candidate k             */
  TQEC_H_OFF_A0 = 10, /* ints[off + pat]           : variable assignment of the coset representative */
  TQEC_H_OFF_KER = 11,/* ints[off + k]             : variable assignment of kernel candidate k       */
  TQEC_H_OFF_VARS = 12,/* ints[off + j]            : id of the factor's j-th variable (output bit)    */
  TQEC_H_OFF_CLOSE = 13/* ints[off + 2c], [off+2c+1]: (full slot, syndrome bit) of closed check c; then, at
                          ints[off + 2*n_close + b], b < w_out: perm[b] = full slot of output bit b.  The full index of a
                          step has slots 0..w_in-1 = the input state bits and w_in.. = the opened checks; the output
                          index is a bit permutation of the full slots that survive (closed slots are not in the image) */
};

typedef struct tqec_plan tqec_plan; /* a compiled schedule resident on one device      */
typedef struct tqec_comm tqec_comm; /* one rank of a job: NCCL communicator + stream   */
typedef struct tqec_gf2 tqec_gf2;   /* a bit-packed GF(2) matrix resident on one device */

/* Optional second lowering of a plan (either semiring): the in-place patch sweep (tensorqec.jl_b200/sweep.py, executed by
 * k_sweep).  Same recurrence, same results bit for bit; plans that carry it decode through k_sweep unless the
 * environment variable TQEC_NO_SWEEP is set.  All tables are host pointers copied by tqec_plan_create. */
typedef struct {
  int32_t W;               /* slot bits of the state index                                        */
  int32_t sg;              /* log2(shots per team pass); W + sg = 10                               */
  int32_t n_ss;            /* super-steps                                                          */
  int32_t n_head_bits;     /* syndrome bits the tabulated head depends on (<= 16)                  */
  int32_t bp_words;        /* back-pointer words per lane per pass                                 */
  int32_t n_tvals;
  const int32_t *rec;      /* n_ss * 32: forward records                                           */
  const int32_t *tb;       /* n_ss * 64: traceback records                                         */
  const uint32_t *lanetab; /* n_ss * 32: lane byte address | lane shot bits << 16                  */
  const double *tvals;     /* pooled layer tables                                                  */
  const int32_t *head_bits;/* n_head_bits syndrome bit indices                                     */
  const double *head_state;/* 2^n_head_bits * 2^W state values after the head                      */
  const uint64_t *head_cfg;/* max-plus: 2^n_head_bits * 2^W * ceil(n_vars/64) partial configurations   */
  const int32_t *out_index;/* 2^n_obs state indices of the output entries (max-plus: one)              */
} tqec_sweep_desc;

/* Optional lowering of a sum-product plan for the GLOBAL-MEMORY executor (tensorqec.jl_b200/wide.py, executed by
 * k_wide_pass): plans whose frontier does not fit on chip (14..31 bits; the state, 2^w FP64 entries per shot, lives in
 * HBM).  Steps are grouped into passes; a pass loads a tile of 2^t_in entries per value of its spectator bits, runs
 * its steps in shared memory in tile-local coordinates and stores 2^t_out entries.  A descriptor that carries `wide`
 * needs no `hdr` / `ints` / `tables` (n_steps = 0).  All tables are host pointers copied by tqec_plan_create. */
#define TQEC_WIDE_PASS_INTS 16
#define TQEC_WIDE_STEP_INTS 16
enum { /* pass header */
  TQEC_WP_WIN = 0, TQEC_WP_WOUT = 1,         /* global state width before / after the pass                           */
  TQEC_WP_TIN = 2, TQEC_WP_TOUT = 3,         /* tile width before / after                                            */
  TQEC_WP_NSTEPS = 4, TQEC_WP_STEP0 = 5,     /* the pass's steps in step_hdr                                         */
  TQEC_WP_TINMASK = 6, TQEC_WP_TOUTMASK = 7, /* positions of the tile bits in the global index before / after        */
  TQEC_WP_OFF_INTS = 8, TQEC_WP_N_INTS = 9,  /* the pass's block of `ints` (step offsets are relative to it)         */
  TQEC_WP_OFF_TAB = 10, TQEC_WP_N_TAB = 11   /* the pass's block of `tables`                                         */
};
enum { /* local step header: tile-local widths; tables as in the TQEC_H_* step header */
  TQEC_WL_WIN = 0, TQEC_WL_NOPEN = 1, TQEC_WL_NCLOSE = 2, TQEC_WL_WOUT = 3, TQEC_WL_NK = 4,
  TQEC_WL_OFF_T = 5,     /* tables[off + pat*nk + k]                                                                 */
  TQEC_WL_OFF_ML = 6,    /* ints[off + pat]                                                                          */
  TQEC_WL_OFF_MK = 7,    /* ints[off + k]                                                                            */
  TQEC_WL_OFF_CLOSE = 8, /* ints[off + 2c], [off + 2c + 1] = (full slot, syndrome bit) of closed check c             */
  TQEC_WL_KEEPMASK = 9   /* full slots that survive; output bit b is the b-th set bit (monotone permutation)         */
};
typedef struct {
  int32_t n_pass, n_steps;
  int32_t w_cap;           /* widest global state over the plan (bits, <= 31)                     */
  int32_t t_max;           /* widest tile (bits, <= 13)                                           */
  const int32_t *pass_hdr; /* n_pass * TQEC_WIDE_PASS_INTS                                        */
  const int32_t *step_hdr; /* n_steps * TQEC_WIDE_STEP_INTS                                       */
  const int32_t *ints;
  int64_t n_ints;
  const double *tables;
  int64_t n_tables;
  const int32_t *obs_pos;  /* n_obs: position of observable i in the final index                  */
  /* Optional second encoding of the passes made of rank-1 factors only (detector error models), executed by k_wide_bf
   * as register butterflies (tqec_lower_wide.cpp:bf_encode_pass describes the blocks).  bf_off == NULL: none. */
  const int32_t *bf_off;   /* n_pass: offset of the pass's block in bf_ints, or -1 (k_wide_pass runs the pass)   */
  const int32_t *bf_ints;
  int64_t n_bf_ints;
  const double *bf_vals;   /* ratios r = t1 / t0: per group TQEC_BF_G unit steps, then its dependent steps */
  int64_t n_bf_vals;
  double bf_mant;          /* product of the t0 normalised away = bf_mant * 2^bf_log2; applied to the output when   */
  int32_t bf_log2;         /* the butterfly passes are in use                                                     */
} tqec_wide_desc;
#ifndef TQEC_BF_SWZ_WIDE
#define TQEC_BF_SWZ_WIDE 0  /* 1: the k_wide_bf tile swizzle folds index bits 8..10 as well as 5..7 into bits 1..3 (measured
                               7 % slower at d = 5 x 5: benchmarks/ab_swz.sh) */
#endif
#define TQEC_BF_G 5            /* dimensions of a butterfly group: a thread holds 2^5 state entries in registers */
#define TQEC_BF_GROUP_INTS 16  /* group record: n_free, dep0, n_dep, close0, n_closes, basis[5], zero mask, val0, -, order */

typedef struct {
  int32_t semiring;        /* TQEC_SEMIRING_*                                                     */
  int32_t n_vars;          /* error variables = bits of a decoded configuration                   */
  int32_t n_checks;        /* syndrome bits per shot                                              */
  int32_t n_obs;           /* open observable axes (sum-product only); output has 2^n_obs entries */
  int32_t n_steps;
  int32_t w_max;           /* widest state (bits) over all steps                                  */
  const int32_t *hdr;      /* n_steps * TQEC_HDR_INTS                                             */
  const int32_t *ints;     /* pooled integer tables                                               */
  int64_t n_ints;
  const double *tables;    /* pooled FP64 tables: log-weights (max-plus) or weights (sum-product) */
  int64_t n_tables;
  const int32_t *obs_slot; /* n_obs: slot of observable i in the final state                      */
  int32_t device;          /* CUDA device ordinal                                                 */
  const tqec_sweep_desc *sweep; /* optional (NULL): in-place patch sweep of the same plan          */
  int32_t table_bits;      /* plans with n_checks <= table_bits are fully tabulated at creation (0 = default 16,
                              -1 = never, at most 26; table = 2^n_checks x (configuration words + outputs)) */
  const tqec_wide_desc *wide; /* optional (NULL): global-memory lowering; takes precedence over hdr / sweep */
  int32_t flags;           /* TQEC_PLAN_*                                                                    */
  int32_t log2_scale;      /* sum-product: the factor tables were pre-multiplied by powers of two against under- and
                              overflow and their exponents sum to -log2_scale; the library multiplies every marginal
                              it returns (host and *_dev entry points alike) by 2^log2_scale */
} tqec_plan_desc;

#define TQEC_PLAN_DYNAMIC_RESCALE 1 /* global-memory executor: per-shot dynamic rescaling.  After every pass the largest
                                      state entry of each shot is known; when it has fallen below 2^-300 the next pass
                                      multiplies the shot's state by a power of two (exact) and an int32 exponent per shot
                                      absorbs it, so neither the marginals' ratio nor the argmax is lost to underflow
                                      however unlikely the syndrome.  The static scaling (log2_scale) stays in place. */

const char *tqec_last_error(void);
int tqec_version(void);
int tqec_device_count(int32_t *out);

/* ---- schedule --------------------------------------------------------------------------------------------- */
int tqec_plan_create(const tqec_plan_desc *desc, tqec_plan **out);
int tqec_plan_destroy(tqec_plan *plan);
/* launch geometry and per-shot cost the library derived for the plan */
enum {
  TQEC_Q_TEAM_THREADS = 0, TQEC_Q_SHOTS_PER_TEAM = 1, TQEC_Q_SMEM_BYTES = 2, TQEC_Q_GRID = 3,
  TQEC_Q_TEAMS_PER_SM = 4, TQEC_Q_BP_BYTES_PER_TEAM = 5, TQEC_Q_CANDIDATES_PER_SHOT = 6, TQEC_Q_SM_COUNT = 7,
  TQEC_Q_LAUNCHES = 8, /* kernels launched through this plan so far */
  TQEC_Q_SWEEP = 9,    /* 1 if the plan decodes through the in-place patch sweep (k_sweep) */
  TQEC_Q_TABLE = 10,   /* 1 if the plan is fully tabulated (n_checks <= 16): decode is a table look-up filled once, at
                          plan creation, by the plan's own kernels; TQEC_NO_TABLE=1 in the environment disables it */
  TQEC_Q_WIDE = 11,    /* 1 if the plan decodes through the global-memory executor (k_wide_pass) */
  TQEC_Q_WIDE_BATCH = 12 /* shots whose states are resident in HBM at a time (global-memory executor) */
};
int tqec_plan_query(const tqec_plan *plan, int32_t what, int64_t *out);

/* ---- compile: factor graph -> plan, entirely inside the library -------------------------------------------- */
/* The decoder's factor graph exactly as the reference builds its tensor network: prior factors over binary error
 * variables (`single_qubit_tensor`, src/decoding/general_decoding.jl:5; `[1-p, p]`, src/decoding/tndecoder.jl:202-205;
 * any `SimpleTensorNetwork` of a GeneralDecodingProblem) and parity rows (checks / detectors clamped by a syndrome bit,
 * logical rows left open: src/decoding/tndecoder.jl:33-50, 97-146, 186-219).  tqec_plan_compile lowers it (frontier
 * schedule, in-place patch sweep with its tabulated head, or global-memory passes -- whichever the frontier width
 * calls for) and creates the plan: what `compile(decoder, problem)` does (src/decoding/interfaces.jl:67-79, 129-142).
 * A host binding therefore needs no lowering code of its own: it lists factors and rows and, optionally, passes the
 * leaf order of the contraction tree its optimiser chose (OMEinsum's `optimize_code`) as `order`. */
typedef struct {
  int32_t semiring;            /* TQEC_SEMIRING_*                                                             */
  int32_t n_vars, n_checks, n_obs;
  int32_t n_factors;
  const int32_t *factor_ptr;   /* n_factors + 1: factor f has variables factor_vars[factor_ptr[f] .. factor_ptr[f+1]) */
  const int32_t *factor_vars;  /* 0-based variable ids                                                        */
  const double *factor_tables; /* concatenated; 2^rank entries per factor, first variable fastest (column-major,
                                  the layout of a Julia array)                                                */
  int32_t n_rows;              /* parity rows: the n_checks clamped ones and the n_obs open ones, in any order  */
  const int32_t *row_ptr;      /* n_rows + 1                                                                  */
  const int32_t *row_vars;
  const int32_t *row_kind;     /* 0: clamped by syndrome bit row_index; 1: open output axis row_index          */
  const int32_t *row_index;
  const int32_t *order;        /* NULL, or n_factors entries: absorption order of the prior factors            */
  int32_t head_bits;           /* syndrome bits the tabulated head of the sweep may depend on (0 = default 14;
                                  at most 16)                                                                  */
  int32_t table_bits;          /* as in tqec_plan_desc                                                         */
  int32_t device;
  int32_t flags;               /* TQEC_COMPILE_*                                                               */
  int32_t wide_t_max;          /* tile bits of the global-memory executor (0 = default 12)                     */
} tqec_problem_desc;
#define TQEC_COMPILE_NO_SWEEP 1   /* never use the in-place patch sweep                                        */
#define TQEC_COMPILE_NO_FUSE 2    /* general max-plus kernels: one factor per step                             */
#define TQEC_COMPILE_FORCE_WIDE 4 /* sum-product: global-memory executor even if the frontier fits on chip     */
#define TQEC_COMPILE_DYNAMIC_RESCALE 8 /* sum-product: global-memory executor with per-shot dynamic rescaling
                                     (TQEC_PLAN_DYNAMIC_RESCALE); implies FORCE_WIDE                           */

typedef struct tqec_lowered tqec_lowered; /* host-side result of the lowering (no device memory)               */
int tqec_lower(const tqec_problem_desc *prob, tqec_lowered **out);
int tqec_lowered_destroy(tqec_lowered *lw);
/* Tables of a lowering, for inspection and for tests (the Python lowering is kept as the oracle of the C++ one):
 * *data points into the handle (valid until it is destroyed), *count = number of elements. */
enum {
  TQEC_LW_META = 0,        /* int32[20]: kind (0 schedule, 1 schedule + sweep, 2 wide), n_steps, w_max, log2_scale,
                              sweep {W, sg, n_ss, n_head_bits, bp_words, head_steps, conflicts}, wide {n_pass, n_steps,
                              w_cap, t_max}, table_bits, n_obs, n_checks, n_vars, semiring                     */
  TQEC_LW_COST = 1,        /* double[2]: candidate evaluations per shot, HBM bytes per shot (wide)              */
  TQEC_LW_ORDER = 2,       /* int32: absorption order of the merged factors                                    */
  TQEC_LW_HDR = 3, TQEC_LW_INTS = 4, TQEC_LW_TABLES = 5 /* double */, TQEC_LW_OBS_SLOT = 6,
  TQEC_LW_SW_REC = 7, TQEC_LW_SW_TB = 8, TQEC_LW_SW_LANETAB = 9, TQEC_LW_SW_TVALS = 10 /* double */,
  TQEC_LW_SW_HEAD_BITS = 11, TQEC_LW_SW_HEAD_STATE = 12 /* double */, TQEC_LW_SW_HEAD_CFG = 13 /* uint64 */,
  TQEC_LW_SW_OUT_INDEX = 14,
  TQEC_LW_WD_PASS_HDR = 15, TQEC_LW_WD_STEP_HDR = 16, TQEC_LW_WD_INTS = 17, TQEC_LW_WD_TABLES = 18 /* double */,
  TQEC_LW_WD_OBS_POS = 19,
  TQEC_LW_WD_BF_OFF = 20, TQEC_LW_WD_BF_INTS = 21, TQEC_LW_WD_BF_VALS = 22 /* double */,
  TQEC_LW_WD_BF_SCALE = 23 /* double[2]: bf_mant, bf_log2 */
};
int tqec_lowered_get(const tqec_lowered *lw, int32_t what, const void **data, int64_t *count);
int tqec_plan_from_lowered(const tqec_lowered *lw, int32_t device, tqec_plan **out);
/* Plan serialisation: a lowered plan as one file (every table tqec_plan_from_lowered reads, behind a magic word and the
 * library's table-format version), so that a host can lower once -- 0.7 s at d = 9 with the 14-bit head, 1.4 s for the
 * d = 5 x 5 circuit-level DEM -- and create plans from the file afterwards (on any device / rank).  The reference has no
 * counterpart (a CompiledTNMAP is rebuilt by compile(), src/decoding/tndecoder.jl:46-50).  A file of another table
 * format, a truncated or a corrupt one is refused with TQEC_ERR_INVALID.                                           */
int tqec_lowered_save(const tqec_lowered *lw, const char *path);
int tqec_lowered_load(const char *path, tqec_lowered **out);
/* = tqec_lower + tqec_plan_from_lowered + tqec_lowered_destroy */
int tqec_plan_compile(const tqec_problem_desc *prob, tqec_plan **out);

/* ---- decoding --------------------------------------------------------------------------------------------- */
/* TNMAP: synd = B * ceil(n_checks/64) words; corr_out = B * ceil(n_vars/64) words (bit v = variable v of the most
 * probable configuration); logp_out (may be NULL) = B log-weights of that configuration (-inf: infeasible). */
int tqec_decode_map(tqec_plan *plan, const uint64_t *synd, int64_t n_shots, uint64_t *corr_out, double *logp_out);
int tqec_decode_map_dev(tqec_plan *plan, const uint64_t *d_synd, int64_t n_shots, uint64_t *d_corr, double *d_logp,
                        void *stream);
/* TNMMAP: mar_out = B * 2^n_obs weights (the static scaling of the plan's tables already undone), entry index =
 * sum_i obs_i << i (observable 0 fastest = the reference's column-major `mar`, tndecoder.jl:134); argmax_out (may be
 * NULL) = first maximal entry per shot (findmax). */
int tqec_decode_marginal(tqec_plan *plan, const uint64_t *synd, int64_t n_shots, double *mar_out,
                         int32_t *argmax_out);
int tqec_decode_marginal_dev(tqec_plan *plan, const uint64_t *d_synd, int64_t n_shots, double *d_mar,
                             int32_t *d_argmax, void *stream);

/* TNMMAP with the exponents kept apart: true marginal = mar_out[b][i] * 2^log2_out[b].  For plans compiled with dynamic
 * rescaling the entries of mar_out stay in the FP64 range even when the syndrome's probability does not. */
int tqec_decode_marginal_log2(tqec_plan *plan, const uint64_t *synd, int64_t n_shots, double *mar_out, int32_t *log2_out,
                              int32_t *argmax_out);
/* The same two calls with ONE BYTE PER BIT in host memory, the layout of the reference's own containers (Mod2 wraps
 * Bool, src/codes/mod2.jl:20-41; a batch `Matrix{Mod2}` holds one shot per column = n_bits contiguous bytes per shot):
 * synd_bits = B * n_checks bytes, corr_bits = B * n_vars bytes.  The bytes cross PCIe as they are and are packed /
 * unpacked on the device next to the decode kernel (host-side packing of 1e7 shots costs more than decoding them). */
int tqec_decode_map_bytes(tqec_plan *plan, const uint8_t *synd_bits, int64_t n_shots, uint8_t *corr_bits, double *logp_out);
/* TNMAP with the syndrome kept as two arrays -- a CSS code's sx (n_a bits per shot) and sz (n_b bits per shot), the layout of
 * the reference's CSSSyndrome (src/decoding/interfaces.jl): bit k of a shot is synd_a[shot][k] for k < n_a, else
 * synd_b[shot][k - n_a]; n_a + n_b = n_checks.  Saves the host-side concatenation of the two arrays. */
int tqec_decode_map_bytes2(tqec_plan *plan, const uint8_t *synd_a, int32_t n_a, const uint8_t *synd_b, int32_t n_b,
                           int64_t n_shots, uint8_t *corr_bits, double *logp_out);
int tqec_decode_marginal_bytes(tqec_plan *plan, const uint8_t *synd_bits, int64_t n_shots, double *mar_out,
                               int32_t *argmax_out);

/* ---- GF(2) ------------------------------------------------------------------------------------------------ */
/* rows x cols matrix, each row packed into ceil(cols/64) words. */
int tqec_gf2_create(int32_t rows, int32_t cols, const uint64_t *packed_rows, int32_t device, tqec_gf2 **out);
int tqec_gf2_destroy(tqec_gf2 *m);
/* out[shot] = M * in[shot] over GF(2): in = B * ceil(cols/64) words, out = B * ceil(rows/64) words. */
int tqec_gf2_apply(tqec_gf2 *m, const uint64_t *in, int64_t n_shots, uint64_t *out);
int tqec_gf2_apply_dev(tqec_gf2 *m, const uint64_t *d_in, int64_t n_shots, uint64_t *d_out, void *stream);
/* Logical check of e1 against e2 (both B * ceil(cols/64) words; e2 may be NULL = zero): row i of L has class
 * row_class[i] in {0, 1} (0: X-type logical flip, rows of lz acting on the x block; 1: Z-type, rows of lx acting on
 * the z block).  flags_out (may be NULL) = B bytes, bit 0 / bit 1 = some class-0 / class-1 row has odd parity on
 * e1 xor e2.  counts (may be NULL) += {#shots with bit 0, #shots with bit 1, #shots with any, #shots}. */
int tqec_logical_flags(tqec_gf2 *L, const int32_t *row_class, const uint64_t *e1, const uint64_t *e2,
                       int64_t n_shots, uint8_t *flags_out, int64_t counts[4]);
/* TNMMAP error pattern: e = R * synd (any solution of H e = synd), then moved into sector[shot] by the rows of FIX:
 * undetectable patterns (H f = 0) whose sector flips d_j = L f_j are in reduced echelon form (the lowest set bit of d_j
 * is its pivot, no other row has it) -- for a CSS code the conjugate logicals, in general a basis of the reachable
 * flips (joint flips of several observables included).  R: n_vars x n_checks; L: n_obs x n_vars (n_obs <= 16);
 * FIX: at most n_obs rows x n_vars.  ok_out (may be NULL): B bytes, 1 iff e ends in the requested sector. */
int tqec_coset_rep(tqec_gf2 *R, tqec_gf2 *L, tqec_gf2 *FIX, const uint64_t *synd, const int32_t *sector,
                   int64_t n_shots, uint64_t *err_out, uint8_t *ok_out);

/* ---- lookup-table decoder (a batched baseline sharing the GF(2) front / back end) ----------- */
/* The reference's TableDecoder (src/decoding/truthtable.jl:138-166): syndrome -> error pattern, decode = look-up.  Keys:
 * n_entries * ceil(n_checks/64) words, strictly increasing (most significant word = last word); values: n_entries *
 * ceil(n_vars/64) words.  found_out (may be NULL): 1 if the syndrome is in the table (else the pattern is zero). */
typedef struct tqec_table tqec_table;
int tqec_table_create(int64_t n_entries, int32_t n_checks, int32_t n_vars, const uint64_t *keys_sorted,
                      const uint64_t *values, int32_t device, tqec_table **out);
int tqec_table_destroy(tqec_table *table);
int tqec_table_decode(tqec_table *table, const uint64_t *synd, int64_t n_shots, uint64_t *corr_out, uint8_t *found_out);

/* ---- belief propagation + OSD --------------------------------------------------------------- */
/* The reference's BPDecoder (src/decoding/bposd.jl): tanh-rule sum-product on log-likelihood ratios, flooding schedule,
 * messages clamped to [-10, 10], at most max_iter iterations, order-0 OSD when it does not converge (osd != 0).  The
 * Tanner graph in CSR form: check s touches bits s_adj[s_ptr[s] .. s_ptr[s+1]); p[q] = flip probability of bit q.
 * flags_out (may be NULL): bit 0 = BP converged, bit 1 = the pattern comes from OSD; 0 = neither (zero pattern). */
typedef struct tqec_bp tqec_bp;
int tqec_bp_create(int32_t n_bits, int32_t n_checks, const int32_t *s_ptr, const int32_t *s_adj, const double *p,
                   int32_t max_iter, int32_t osd, int32_t device, tqec_bp **out);
int tqec_bp_destroy(tqec_bp *bp);
int tqec_bp_decode(tqec_bp *bp, const uint64_t *synd, int64_t n_shots, uint64_t *corr_out, uint8_t *flags_out);

/* ---- sampling --------------------------------------------------------------------------------------------- */
#define TQEC_MODEL_FLIP 0  /* IndependentFlipError: bit i flips iff u < p0[i]            (error_model.jl:69-71)  */
#define TQEC_MODEL_DEPOL 1 /* IndependentDepolarizingError on n qubits: one u per qubit, Y tested first:
                              u < py -> Y; u < px+py -> X; u < px+py+pz -> Z (error_model.jl:97-117); output has
                              2n bits: x errors in bits 0..n-1, z errors in bits n..2n-1 (reduce2general order)    */
/* u = Philox4x32-10(key = seed, counter = (shot, site)) -> 53-bit uniform; shot = shot_offset + local index, so
 * the sampled errors do not depend on how shots are split over calls / GPUs. p0,p1,p2 = (p) or (px,py,pz). */
int tqec_sample_errors(int32_t model, int32_t n_sites, const double *p0, const double *p1, const double *p2,
                       uint64_t seed, int64_t shot_offset, int64_t n_shots, uint64_t *err_out, int32_t device);

/* ---- fused Monte-Carlo pipeline (sample -> syndrome -> decode -> logical check) ---------------------------- */
typedef struct {
  tqec_plan *plan;         /* max-plus plan over n_vars variables                     */
  tqec_gf2 *H;             /* n_checks x n_vars                                        */
  tqec_gf2 *L;             /* logical rows x n_vars                                    */
  const int32_t *row_class;
  int32_t model;           /* TQEC_MODEL_*                                             */
  int32_t n_sites;         /* qubits (DEPOL: n_vars = 2 n_sites) or bits (FLIP)        */
  const double *p0, *p1, *p2;
  int64_t chunk;           /* shots per internal batch (0 = library default)           */
  tqec_comm *comm;         /* optional (NULL): all-reduce the counters over the job's ranks before returning */
} tqec_mc_desc;
/* counts += {logical X-type failures, logical Z-type failures, any failure, shots} over shots
 * shot_offset .. shot_offset + n_shots - 1 -- of THIS rank, or, with `comm`, summed over all ranks (one ncclAllReduce of
 * the four device counters on the pipeline's stream, inside elapsed_ms).  elapsed_ms (may be NULL) = device time of
 * the whole pipeline. */
int tqec_mc_run(const tqec_mc_desc *mc, uint64_t seed, int64_t shot_offset, int64_t n_shots, int64_t counts[4],
                float *elapsed_ms);

/* ---- multi-GPU: shots shard over ranks (one process per GPU), the counters are the only thing exchanged ---------- */
/* Replaces the job farm of src/multiprocessing.jl:41-52.  Rank 0 makes the id (128 bytes = ncclUniqueId) and hands it to
 * the other ranks by whatever means the host has (Distributed.jl, MPI, a file); every rank then calls tqec_comm_init.
 * NCCL is loaded at run time (libnccl.so.2); without it these calls return TQEC_ERR_UNSUPPORTED. */
int tqec_comm_unique_id(void *id_out_128_bytes);
int tqec_comm_init(int32_t nranks, int32_t rank, const void *unique_id_128_bytes, int32_t device, tqec_comm **out);
int tqec_comm_destroy(tqec_comm *comm);
/* counts[0..3] <- sum over ranks (in place; every rank calls it with its own counters) */
int tqec_comm_allreduce_counts(tqec_comm *comm, int64_t counts[4]);

/* ---- measurement helper -------------------------------------------------------------------------------------- */
/* FP64 CUDA-core peak of the device, measured with register-resident chains: DADD instructions/s (T/s), DFMA
 * TFLOP/s (2 flops each), and the max-plus candidate rate counted as 2 ops (add + compare) in Tops/s.  Roofline
 * denominator of the decode kernels (MEASURED_PEAKS.json has no FP64 entry). */
int tqec_fp64_peak(int32_t device, double *dadd_tops, double *dfma_tflops, double *maxplus_tops);
/* FP64 tensor-core rate (mma.sync m8n8k4 f64 = DMMA.8x8x4, 512 flops per warp instruction), TFLOP/s: measured beside the
 * DFMA rate to decide whether any sum-product step should run as a dense product (none does). */
int tqec_dmma_peak(int32_t device, double *dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* TQEC_H */
