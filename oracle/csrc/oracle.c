/* C restatements of the decoding hot path, used ONLY as timed CPU baselines and as a fast checker
 * (test infrastructure; see oracle/__init__.py).  Plain C + OpenMP over shots; one shot at a time per thread,
 * no cross-shot batching.
 *
 * (1) oracle_dense_*    : what the reference does per decode call -- pairwise contraction of the DENSE tensor
 *     network (unity vectors, dense parity tensors, priors; src/decoding/tndecoder.jl:33-57, 97-165) along a binary
 *     tree fixed at compile time, all intermediates cached, then (max-plus) a root-to-leaves traceback.  The tree,
 *     the evidence slicing and the index tables come from oracle/cref.py (greedy tree of oracle/dense.py).  It is an
 *     optimistic stand-in for OMEinsum/TensorInference: no dynamic dispatch, no allocation, no permutedims, and the
 *     traceback touches one output element per node instead of TensorInference's two extra einsums per node.
 * (2) oracle_frontier_* : the frontier recurrence executed by the CUDA kernels, from the same lowered tables
 *     (tensorqec.jl_b200/schedule.py) -- the algorithm-for-algorithm CPU port.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_leaves, n_steps, n_checks, n_vars;
  const int64_t *node_size;  /* n_leaves + n_steps */
  /* leaves: data offset for evidence value 0 / 1 (equal when the leaf carries no evidence label), evidence bit or -1 */
  const int64_t *leaf_off0, *leaf_off1;
  const int32_t *leaf_ev;
  const double *leaf_data;   /* values already in the semiring's domain (log for max-plus) */
  /* steps */
  const int32_t *st_a, *st_b;
  const int64_t *st_no, *st_nk; /* number of output elements / contracted combinations */
  const int64_t *st_tab;        /* offset of this step's tables in `tabs`: la_o[no], lb_o[no], la_k[nk], lb_k[nk] */
  const int32_t *tabs;
  const int32_t *var_leaf;      /* n_vars: unity leaf of each variable (its selected index is the variable's value) */
} dense_plan;

static int bit_of(const uint64_t *w, int b) { return (int)((w[b >> 6] >> (b & 63)) & 1ull); }

/* one shot: returns log-weight (max-plus) or fills `mar` (sum-product, root elements) */
static double dense_one(const dense_plan *P, int maxplus, const uint64_t *syn, double *work, const double **node,
                        int64_t *sel, uint64_t *cfg, double *mar) {
  const int nl = P->n_leaves;
  for (int i = 0; i < nl; ++i) {
    int ev = P->leaf_ev[i];
    int64_t off = (ev >= 0 && bit_of(syn, ev)) ? P->leaf_off1[i] : P->leaf_off0[i];
    node[i] = P->leaf_data + off;
  }
  double *wp = work;
  for (int s = 0; s < P->n_steps; ++s) {
    const double *A = node[P->st_a[s]], *B = node[P->st_b[s]];
    const int64_t no = P->st_no[s], nk = P->st_nk[s];
    const int32_t *la_o = P->tabs + P->st_tab[s], *lb_o = la_o + no, *la_k = lb_o + no, *lb_k = la_k + nk;
    double *O = wp;
    if (maxplus) {
      for (int64_t o = 0; o < no; ++o) {
        const double *a = A + la_o[o], *b = B + lb_o[o];
        double best = a[la_k[0]] + b[lb_k[0]];
        for (int64_t k = 1; k < nk; ++k) {
          double v = a[la_k[k]] + b[lb_k[k]];
          if (v > best) best = v;
        }
        O[o] = best;
      }
    } else {
      for (int64_t o = 0; o < no; ++o) {
        const double *a = A + la_o[o], *b = B + lb_o[o];
        double acc = 0.0;
        for (int64_t k = 0; k < nk; ++k) acc += a[la_k[k]] * b[lb_k[k]];
        O[o] = acc;
      }
    }
    node[nl + s] = O;
    wp += no;
  }
  const int root = nl + P->n_steps - 1;
  if (!maxplus) {
    if (mar) memcpy(mar, node[root], sizeof(double) * (size_t)P->node_size[root]);
    return 0.0;
  }
  const double logp = node[root][0];
  if (cfg) {
    sel[root] = 0;
    for (int s = P->n_steps - 1; s >= 0; --s) {
      const double *A = node[P->st_a[s]], *B = node[P->st_b[s]];
      const int64_t no = P->st_no[s], nk = P->st_nk[s];
      const int32_t *la_o = P->tabs + P->st_tab[s], *lb_o = la_o + no, *la_k = lb_o + no, *lb_k = la_k + nk;
      const int64_t o = sel[nl + s];
      const double *a = A + la_o[o], *b = B + lb_o[o];
      int64_t bk = 0;
      double best = a[la_k[0]] + b[lb_k[0]];
      for (int64_t k = 1; k < nk; ++k) {
        double v = a[la_k[k]] + b[lb_k[k]];
        if (v > best) { best = v; bk = k; }
      }
      sel[P->st_a[s]] = la_o[o] + la_k[bk];
      sel[P->st_b[s]] = lb_o[o] + lb_k[bk];
    }
    const int cw = (P->n_vars + 63) / 64 > 0 ? (P->n_vars + 63) / 64 : 1;
    memset(cfg, 0, sizeof(uint64_t) * (size_t)cw);
    for (int v = 0; v < P->n_vars; ++v)
      if (sel[P->var_leaf[v]] & 1) cfg[v >> 6] |= 1ull << (v & 63);
  }
  return logp;
}

/* syn: B x sw words; cfg_out: B x cw words (may be NULL); out: B log-weights (max-plus) or B x root_size marginals */
int oracle_dense_run(const dense_plan *P, int maxplus, const uint64_t *syn, int64_t B, uint64_t *cfg_out, double *out,
                     int n_threads) {
  const int sw = (P->n_checks + 63) / 64 > 0 ? (P->n_checks + 63) / 64 : 1;
  const int cw = (P->n_vars + 63) / 64 > 0 ? (P->n_vars + 63) / 64 : 1;
  const int nn = P->n_leaves + P->n_steps;
  int64_t total = 0;
  for (int s = 0; s < P->n_steps; ++s) total += P->st_no[s];
  const int64_t root_size = P->node_size[nn - 1];
  int fail = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    double *work = (double *)malloc(sizeof(double) * (size_t)(total > 0 ? total : 1));
    const double **node = (const double **)malloc(sizeof(double *) * (size_t)nn);
    int64_t *sel = (int64_t *)malloc(sizeof(int64_t) * (size_t)nn);
    if (!work || !node || !sel) {
#pragma omp atomic write
      fail = 1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int64_t b = 0; b < B; ++b) {
        if (maxplus) {
          double lp = dense_one(P, 1, syn + b * sw, work, node, sel, cfg_out ? cfg_out + b * cw : NULL, NULL);
          if (out) out[b] = lp;
        } else {
          dense_one(P, 0, syn + b * sw, work, node, sel, NULL, out + b * root_size);
        }
      }
    }
    free(work); free((void *)node); free(sel);
  }
  return fail;
}

/* ---------------------------------------------------------------------------------------------------------------- */
enum { H_R = 0, H_WIN, H_NOPEN, H_NCLOSE, H_WOUT, H_NK, H_KB, H_OFF_T, H_OFF_ML, H_OFF_MK, H_OFF_A0, H_OFF_KER,
       H_OFF_VARS, H_OFF_CLOSE, HDR_INTS = 16 };

typedef struct {
  int32_t semiring, n_vars, n_checks, n_obs, n_steps, w_max;
  const int32_t *hdr, *ints;
  const double *tables;
  const int32_t *obs_slot;
} frontier_plan;

/* output index -> full index: bit b goes to slot perm[b] (perm follows the closed list); the closed slots take the
 * shot's syndrome bits.  The scatter of every output index is tabulated once per plan (it does not depend on the shot);
 * the closed-bit values are computed once per step and shot. */
static int scatter_of(int tau, int w_out, const int32_t *perm) {
  int full = 0;
  for (int b = 0; b < w_out; ++b) full |= ((tau >> b) & 1) << perm[b];
  return full;
}

static int closed_bits(int n_close, const int32_t *CL, const uint64_t *syn) {
  int v = 0;
  for (int c = 0; c < n_close; ++c) v |= bit_of(syn, CL[2 * c + 1]) << CL[2 * c];
  return v;
}

static double frontier_one(const frontier_plan *P, const int32_t *const *scat, const uint64_t *syn, double *S0, double *S1,
                           uint16_t *bp, uint64_t *cfg, double *mar) {
  const int maxplus = P->semiring == 0;
  double *Sin = S0, *Sout = S1;
  Sin[0] = maxplus ? 0.0 : 1.0;
  const size_t stride = (size_t)1 << P->w_max;
  for (int t = 0; t < P->n_steps; ++t) {
    const int32_t *h = P->hdr + t * HDR_INTS;
    const int w_in = h[H_WIN], w_out = h[H_WOUT], nk = h[H_NK], n_close = h[H_NCLOSE];
    const double *T = P->tables + h[H_OFF_T];
    const int32_t *ML = P->ints + h[H_OFF_ML], *MK = P->ints + h[H_OFF_MK], *CL = P->ints + h[H_OFF_CLOSE];
    const int inmask = (1 << w_in) - 1;
    const int cbv = closed_bits(n_close, CL, syn);
    const int32_t *sc = scat[t];
    uint16_t *bpt = bp ? bp + (size_t)t * stride : NULL;
    for (int tau = 0; tau < (1 << w_out); ++tau) {
      const int full = sc[tau] | cbv;
      const int pat = full >> w_in;
      const int low = (full & inmask) ^ ML[pat];
      const double *tb = T + pat * nk;
      if (maxplus) {
        double best = Sin[low ^ MK[0]] + tb[0];
        int bk = 0;
        for (int k = 1; k < nk; ++k) {
          const double v = Sin[low ^ MK[k]] + tb[k];
          if (v > best) { best = v; bk = k; }
        }
        Sout[tau] = best;
        if (bpt) bpt[tau] = (uint16_t)bk;
      } else {
        double acc = Sin[low ^ MK[0]] * tb[0];
        for (int k = 1; k < nk; ++k) acc += Sin[low ^ MK[k]] * tb[k];
        Sout[tau] = acc;
      }
    }
    double *tmp = Sin; Sin = Sout; Sout = tmp;
  }
  if (!maxplus) {
    const int NO = 1 << P->n_obs;
    for (int idx = 0; idx < NO; ++idx) {
      int src = 0;
      for (int o = 0; o < P->n_obs; ++o) src |= ((idx >> o) & 1) << P->obs_slot[o];
      mar[idx] = Sin[src];
    }
    return 0.0;
  }
  const double logp = Sin[0];
  if (cfg) {
    const int cw = (P->n_vars + 63) / 64 > 0 ? (P->n_vars + 63) / 64 : 1;
    memset(cfg, 0, sizeof(uint64_t) * (size_t)cw);
    int tau = 0;
    for (int t = P->n_steps - 1; t >= 0; --t) {
      const int32_t *h = P->hdr + t * HDR_INTS;
      const int w_in = h[H_WIN], r = h[H_R];
      const int k = h[H_KB] ? bp[(size_t)t * stride + tau] : 0;
      const int full = scat[t][tau] | closed_bits(h[H_NCLOSE], P->ints + h[H_OFF_CLOSE], syn);
      const int pat = full >> w_in;
      const int a = P->ints[h[H_OFF_A0] + pat] ^ P->ints[h[H_OFF_KER] + k];
      for (int j = 0; j < r; ++j)
        if ((a >> j) & 1) {
          const int v = P->ints[h[H_OFF_VARS] + j];
          cfg[v >> 6] |= 1ull << (v & 63);
        }
      tau = (full & ((1 << w_in) - 1)) ^ P->ints[h[H_OFF_ML] + pat] ^ P->ints[h[H_OFF_MK] + k];
    }
  }
  return logp;
}

int oracle_frontier_run(const frontier_plan *P, const uint64_t *syn, int64_t B, uint64_t *cfg_out, double *out,
                        int n_threads) {
  const int sw = (P->n_checks + 63) / 64 > 0 ? (P->n_checks + 63) / 64 : 1;
  const int cw = (P->n_vars + 63) / 64 > 0 ? (P->n_vars + 63) / 64 : 1;
  const size_t stride = (size_t)1 << P->w_max;
  const int NO = 1 << P->n_obs;
  int fail = 0;
  /* per-plan scatter tables (shared, read-only) */
  int32_t **scat = (int32_t **)calloc((size_t)P->n_steps, sizeof(int32_t *));
  if (!scat) return 1;
  for (int t = 0; t < P->n_steps; ++t) {
    const int32_t *h = P->hdr + t * HDR_INTS;
    const int w_out = h[H_WOUT];
    const int32_t *perm = P->ints + h[H_OFF_CLOSE] + 2 * h[H_NCLOSE];
    scat[t] = (int32_t *)malloc(sizeof(int32_t) * ((size_t)1 << w_out));
    if (!scat[t]) { fail = 1; break; }
    for (int tau = 0; tau < (1 << w_out); ++tau) scat[t][tau] = scatter_of(tau, w_out, perm);
  }
  if (fail) { for (int t = 0; t < P->n_steps; ++t) free(scat[t]); free(scat); return 1; }
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    double *S0 = (double *)malloc(sizeof(double) * stride), *S1 = (double *)malloc(sizeof(double) * stride);
    uint16_t *bp = (P->semiring == 0 && cfg_out) ? (uint16_t *)malloc(sizeof(uint16_t) * stride * (size_t)P->n_steps) : NULL;
    if (!S0 || !S1 || (P->semiring == 0 && cfg_out && !bp)) {
#pragma omp atomic write
      fail = 1;
    } else {
#pragma omp for schedule(static)
      for (int64_t b = 0; b < B; ++b) {
        if (P->semiring == 0) {
          double lp = frontier_one(P, (const int32_t *const *)scat, syn + b * sw, S0, S1, bp, cfg_out ? cfg_out + b * cw : NULL, NULL);
          if (out) out[b] = lp;
        } else {
          frontier_one(P, (const int32_t *const *)scat, syn + b * sw, S0, S1, NULL, NULL, out + (size_t)b * NO);
        }
      }
    }
    free(S0); free(S1); free(bp);
  }
  for (int t = 0; t < P->n_steps; ++t) free(scat[t]);
  free(scat);
  return fail;
}

/* (3) oracle_frontier_wide_run: the same recurrence for plans whose state has up to 2^31 entries (circuit-level detector
 *     error models): one shot at a time, OpenMP over the OUTPUT entries of a step, no per-step scatter tables -- the
 *     schedule must be lowered with the stable layout (monotone perm: schedule.lower(..., stable=True)), so that the
 *     scatter is one bit deposit.  Sum-product only.  Used to make the d = 5 x 5 rounds golden marginals
 *     (tests/golden/make_dem_d5_golden.py) with an absorption order DIFFERENT from the one the CUDA path uses. */
#include <immintrin.h>
int oracle_frontier_wide_run(const frontier_plan *P, const uint64_t *syn, int64_t B, double *out, int n_threads) {
  if (P->semiring == 0) return 2;
  const int sw = (P->n_checks + 63) / 64 > 0 ? (P->n_checks + 63) / 64 : 1;
  const size_t stride = (size_t)1 << P->w_max;
  const int NO = 1 << P->n_obs;
  double *S0 = (double *)malloc(sizeof(double) * stride), *S1 = (double *)malloc(sizeof(double) * stride);
  if (!S0 || !S1) { free(S0); free(S1); return 1; }
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  int bad = 0;
  for (int64_t b = 0; b < B && !bad; ++b) {
    const uint64_t *sy = syn + b * sw;
    double *Sin = S0, *Sout = S1;
    Sin[0] = 1.0;
    for (int t = 0; t < P->n_steps; ++t) {
      const int32_t *h = P->hdr + t * HDR_INTS;
      const int w_in = h[H_WIN], w_out = h[H_WOUT], nk = h[H_NK], n_close = h[H_NCLOSE];
      const double *T = P->tables + h[H_OFF_T];
      const int32_t *ML = P->ints + h[H_OFF_ML], *MK = P->ints + h[H_OFF_MK], *CL = P->ints + h[H_OFF_CLOSE];
      const int32_t *perm = CL + 2 * n_close;
      uint32_t keep = 0;
      for (int q = 0; q < w_out; ++q) {
        if (q > 0 && perm[q] <= perm[q - 1]) bad = 1;            /* not the stable layout */
        keep |= 1u << perm[q];
      }
      if (bad) break;
      const uint32_t inmask = (uint32_t)(((uint64_t)1 << w_in) - 1);
      const uint32_t cbv = (uint32_t)closed_bits(n_close, CL, sy);
      const int64_t n = (int64_t)1 << w_out;
#pragma omp parallel for schedule(static)
      for (int64_t tau = 0; tau < n; ++tau) {
        const uint32_t full = _pdep_u32((uint32_t)tau, keep) | cbv;
        const uint32_t pat = full >> w_in;
        const uint32_t low = (full & inmask) ^ (uint32_t)ML[pat];
        const double *tb = T + (size_t)pat * nk;
        double acc = Sin[low ^ (uint32_t)MK[0]] * tb[0];
        for (int k = 1; k < nk; ++k) acc += Sin[low ^ (uint32_t)MK[k]] * tb[k];
        Sout[tau] = acc;
      }
      double *tmp = Sin; Sin = Sout; Sout = tmp;
    }
    for (int idx = 0; idx < NO; ++idx) {
      int src = 0;
      for (int o = 0; o < P->n_obs; ++o) src |= ((idx >> o) & 1) << P->obs_slot[o];
      out[(size_t)b * NO + idx] = Sin[src];
    }
  }
  free(S0); free(S1);
  return bad ? 3 : 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
