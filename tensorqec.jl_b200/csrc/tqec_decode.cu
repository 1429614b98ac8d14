// Frontier-recurrence decoding kernels (TNMAP max-plus with traceback, TNMMAP sum-product) for sm_100a.
//
// One TEAM (= one CTA of `team_threads` threads) owns 2^sg_log2 shots at a time and keeps their whole state
// tensor S[sub][sigma] (FP64, <= 2^(w_max+sg_log2) entries, ping-pong) in shared memory; nothing but the
// bit-packed syndromes (in), the corrections / marginals (out) and 1..kb back-pointer bits per state entry
// (L2-resident scratch, max-plus only) ever touches global memory.  The grid is persistent: teams stride over
// shot groups.  See tensorqec.jl_b200/schedule.py for the recurrence and the table layout, DESIGN.md for the
// roofline argument (FP64-pipe / issue bound, not HBM bound).
#include <cmath>
#include <cstring>

#include "tqec_common.h"

namespace tqec {

__device__ __forceinline__ int insert_bit(int x, int slot, int bit) {
  return ((x >> slot) << (slot + 1)) | (bit << slot) | (x & ((1 << slot) - 1));
}

__device__ __forceinline__ int rebuild_full(int tau, int sub, int n_close, const int32_t *__restrict__ CL,
                                            const uint64_t *__restrict__ sh_syn, int nsw) {
  int full = tau;
  for (int c = 0; c < n_close; ++c) {
    const int slot = __ldg(CL + 2 * c), bit = __ldg(CL + 2 * c + 1);
    const int sb = (int)((sh_syn[sub * nsw + (bit >> 6)] >> (bit & 63)) & 1ull);
    full = insert_bit(full, slot, sb);
  }
  return full;
}

template <int SEMI, int NK>
__device__ __forceinline__ void run_step(const PlanDev &P, const int32_t *__restrict__ h, const double *__restrict__ Sin,
                                         double *__restrict__ Sout, const uint64_t *__restrict__ sh_syn,
                                         uint32_t *__restrict__ bpt, int T, int tid) {
  const int w_in = h[TQEC_H_WIN], n_close = h[TQEC_H_NCLOSE], w_out = h[TQEC_H_WOUT];
  const int nk = NK > 0 ? NK : h[TQEC_H_NK];
  const int kb = h[TQEC_H_KB];
  const double *__restrict__ Tt = P.tables + h[TQEC_H_OFF_T];
  const int32_t *__restrict__ ML = P.ints + h[TQEC_H_OFF_ML];
  const int32_t *__restrict__ MK = P.ints + h[TQEC_H_OFF_MK];
  const int32_t *__restrict__ CL = P.ints + h[TQEC_H_OFF_CLOSE];
  const int n_tot = 1 << (w_out + P.sg_log2);
  const int inmask = (1 << w_in) - 1, outmask = (1 << w_out) - 1;
  const int per_word = kb ? 32 / kb : 1;
  int mk[NK > 0 ? NK : 1];
  if (NK > 0) {
#pragma unroll
    for (int k = 0; k < NK; ++k) mk[k] = __ldg(MK + k);
  }
  uint32_t word = 0;
  int jw = 0, wi = 0;
  for (int e = tid; e < n_tot; e += T) {
    const int tau = e & outmask, sub = e >> w_out;
    const int full = rebuild_full(tau, sub, n_close, CL, sh_syn, P.nsw);
    const int pat = full >> w_in;
    const int low = (full & inmask) ^ __ldg(ML + pat);
    const double *__restrict__ tb = Tt + pat * nk;
    const double *__restrict__ Sb = Sin + (sub << w_in);
    double best;
    int bk = 0;
    if (NK > 0) {
      double v[NK > 0 ? NK : 1];
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const double s = Sb[low ^ mk[k]], tv = __ldg(tb + k);
        v[k] = SEMI == TQEC_SEMIRING_MAXPLUS ? s + tv : s * tv;
      }
      best = v[0];
#pragma unroll
      for (int k = 1; k < NK; ++k) {
        if (SEMI == TQEC_SEMIRING_MAXPLUS) {
          if (v[k] > best) { best = v[k]; bk = k; }
        } else {
          best += v[k];
        }
      }
    } else {
      const double s0 = Sb[low ^ __ldg(MK)], t0 = __ldg(tb);
      best = SEMI == TQEC_SEMIRING_MAXPLUS ? s0 + t0 : s0 * t0;
      for (int k = 1; k < nk; ++k) {
        const double s = Sb[low ^ __ldg(MK + k)], tv = __ldg(tb + k);
        if (SEMI == TQEC_SEMIRING_MAXPLUS) {
          const double v = s + tv;
          if (v > best) { best = v; bk = k; }
        } else {
          best += s * tv;
        }
      }
    }
    Sout[e] = best;
    if (SEMI == TQEC_SEMIRING_MAXPLUS && kb) {
      word |= (uint32_t)bk << (kb * jw);
      if (++jw == per_word) {
        bpt[wi * T + tid] = word;
        word = 0; jw = 0; ++wi;
      }
    }
  }
  if (SEMI == TQEC_SEMIRING_MAXPLUS && kb && jw) bpt[wi * T + tid] = word;
}

template <int SEMI>
__global__ void k_frontier(const PlanDev P, const uint64_t *__restrict__ synd, const int64_t B,
                           uint64_t *__restrict__ corr, double *__restrict__ out, int32_t *__restrict__ argmax_out,
                           uint32_t *__restrict__ bp_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x, tid = threadIdx.x;
  const int SG = 1 << P.sg_log2;
  const int NS = 1 << (P.w_max + P.sg_log2);
  double *S0 = reinterpret_cast<double *>(smem_raw);
  double *S1 = S0 + NS;
  uint64_t *sh_syn = reinterpret_cast<uint64_t *>(S1 + NS);
  uint64_t *sh_cfg = sh_syn + SG * P.nsw;
  uint32_t *bp = bp_all + (size_t)blockIdx.x * P.bp_words;
  const int64_t n_groups = (B + SG - 1) >> P.sg_log2;
  const double ONE = SEMI == TQEC_SEMIRING_MAXPLUS ? 0.0 : 1.0;

  for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int64_t shot0 = g << P.sg_log2;
    for (int i = tid; i < SG * P.nsw; i += T) {
      const int64_t s = shot0 + i / P.nsw;
      sh_syn[i] = s < B ? synd[shot0 * P.nsw + i] : 0ull;
    }
    if (SEMI == TQEC_SEMIRING_MAXPLUS)
      for (int i = tid; i < SG * P.ncw; i += T) sh_cfg[i] = 0ull;
    for (int i = tid; i < SG; i += T) S0[i] = ONE;
    __syncthreads();

    double *Sin = S0, *Sout = S1;
    for (int t = 0; t < P.n_steps; ++t) {
      const int32_t *h = P.hdr + t * TQEC_HDR_INTS;
      uint32_t *bpt = bp + P.bp_off[t];
      switch (h[TQEC_H_NK]) {
        case 1: run_step<SEMI, 1>(P, h, Sin, Sout, sh_syn, bpt, T, tid); break;
        case 2: run_step<SEMI, 2>(P, h, Sin, Sout, sh_syn, bpt, T, tid); break;
        case 4: run_step<SEMI, 4>(P, h, Sin, Sout, sh_syn, bpt, T, tid); break;
        default: run_step<SEMI, 0>(P, h, Sin, Sout, sh_syn, bpt, T, tid); break;
      }
      __syncthreads();
      double *tmp = Sin; Sin = Sout; Sout = tmp;
    }

    if (SEMI == TQEC_SEMIRING_MAXPLUS) {
      // traceback: one thread per shot walks the back-pointers from the scalar root to the first step
      for (int sub = tid; sub < SG; sub += T) {
        int tau = 0;
        uint64_t *cfg = sh_cfg + sub * P.ncw;
        for (int t = P.n_steps - 1; t >= 0; --t) {
          const int32_t *h = P.hdr + t * TQEC_HDR_INTS;
          const int w_in = h[TQEC_H_WIN], w_out = h[TQEC_H_WOUT], kb = h[TQEC_H_KB], r = h[TQEC_H_R];
          const int32_t *CL = P.ints + h[TQEC_H_OFF_CLOSE];
          int k = 0;
          if (kb) {
            const int e = (sub << w_out) | tau;
            const int j = e / T, lane = e - j * T, per_word = 32 / kb;
            const uint32_t wv = __ldcg(bp + P.bp_off[t] + (j / per_word) * T + lane);
            k = (wv >> (kb * (j % per_word))) & ((1u << kb) - 1u);
          }
          const int full = rebuild_full(tau, sub, h[TQEC_H_NCLOSE], CL, sh_syn, P.nsw);
          const int pat = full >> w_in;
          const int a = __ldg(P.ints + h[TQEC_H_OFF_A0] + pat) ^ __ldg(P.ints + h[TQEC_H_OFF_KER] + k);
          const int32_t *V = P.ints + h[TQEC_H_OFF_VARS];
          for (int j = 0; j < r; ++j)
            if ((a >> j) & 1) {
              const int v = __ldg(V + j);
              cfg[v >> 6] |= 1ull << (v & 63);
            }
          tau = (full & ((1 << w_in) - 1)) ^ __ldg(P.ints + h[TQEC_H_OFF_ML] + pat) ^ __ldg(P.ints + h[TQEC_H_OFF_MK] + k);
        }
        if (shot0 + sub < B && out) out[shot0 + sub] = Sin[sub];
      }
      __syncthreads();
      for (int i = tid; i < SG * P.ncw; i += T)
        if (shot0 + i / P.ncw < B) corr[shot0 * P.ncw + i] = sh_cfg[i];
    } else {
      const int NO = 1 << P.n_obs;
      for (int i = tid; i < SG * NO; i += T) {
        const int sub = i >> P.n_obs, idx = i & (NO - 1);
        int src = 0;
        for (int o = 0; o < P.n_obs; ++o) src |= ((idx >> o) & 1) << __ldg(P.obs_slot + o);
        if (shot0 + sub < B) out[(shot0 + sub) * NO + idx] = Sin[(sub << P.n_obs) | src];
      }
      if (argmax_out) {
        for (int sub = tid; sub < SG; sub += T) {
          if (shot0 + sub >= B) continue;
          double best = -1.0;
          int bi = 0;
          for (int idx = 0; idx < NO; ++idx) {
            int src = 0;
            for (int o = 0; o < P.n_obs; ++o) src |= ((idx >> o) & 1) << __ldg(P.obs_slot + o);
            const double v = Sin[(sub << P.n_obs) | src];
            if (v > best) { best = v; bi = idx; }
          }
          argmax_out[shot0 + sub] = bi;
        }
      }
    }
    __syncthreads();
  }
}

int launch_decode(tqec_plan *plan, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_out,
                  int32_t *d_argmax, cudaStream_t stream) {
  if (B <= 0) return TQEC_OK;
  const int64_t groups = (B + plan->shots_per_team - 1) / plan->shots_per_team;
  const int grid = (int)(groups < plan->grid_max ? groups : plan->grid_max);
  if (plan->semiring == TQEC_SEMIRING_MAXPLUS)
    k_frontier<TQEC_SEMIRING_MAXPLUS><<<grid, plan->team_threads, plan->smem_bytes, stream>>>(
        plan->dev, d_synd, B, d_corr, d_out, nullptr, plan->d_bp);
  else
    k_frontier<TQEC_SEMIRING_SUMPROD><<<grid, plan->team_threads, plan->smem_bytes, stream>>>(
        plan->dev, d_synd, B, nullptr, d_out, d_argmax, plan->d_bp);
  TQEC_CUDA(cudaGetLastError());
  plan->launches += 1;
  return TQEC_OK;
}

}  // namespace tqec

using namespace tqec;

// ---------------------------------------------------------------------------------------------------------------
static int validate_desc(const tqec_plan_desc *d) {
  TQEC_REQUIRE(d != nullptr, "tqec_plan_create: desc is NULL");
  TQEC_REQUIRE(d->semiring == TQEC_SEMIRING_MAXPLUS || d->semiring == TQEC_SEMIRING_SUMPROD,
               "tqec_plan_create: unknown semiring %d", d->semiring);
  TQEC_REQUIRE(d->n_steps > 0 && d->hdr && d->ints && d->tables, "tqec_plan_create: empty schedule");
  TQEC_REQUIRE(d->n_vars >= 0 && d->n_checks >= 0 && d->n_obs >= 0 && d->n_obs <= 16,
               "tqec_plan_create: bad sizes (n_vars=%d n_checks=%d n_obs=%d)", d->n_vars, d->n_checks, d->n_obs);
  TQEC_REQUIRE(d->semiring == TQEC_SEMIRING_SUMPROD || d->n_obs == 0, "tqec_plan_create: max-plus plans have no open axes");
  TQEC_REQUIRE(d->n_obs == 0 || d->obs_slot, "tqec_plan_create: obs_slot is NULL");
  int w = 0, wmax = 0;
  for (int t = 0; t < d->n_steps; ++t) {
    const int32_t *h = d->hdr + t * TQEC_HDR_INTS;
    const int r = h[TQEC_H_R], w_in = h[TQEC_H_WIN], n_open = h[TQEC_H_NOPEN], n_close = h[TQEC_H_NCLOSE];
    const int w_out = h[TQEC_H_WOUT], nk = h[TQEC_H_NK], kb = h[TQEC_H_KB];
    TQEC_REQUIRE(w_in == w, "step %d: w_in=%d does not continue the previous width %d", t, w_in, w);
    TQEC_REQUIRE(r >= 0 && r <= 10 && n_open >= 0 && n_close >= 0 && w_out == w_in + n_open - n_close && w_out >= 0,
                 "step %d: inconsistent widths (r=%d w_in=%d open=%d close=%d w_out=%d)", t, r, w_in, n_open, n_close, w_out);
    TQEC_REQUIRE(nk >= 1 && (nk & (nk - 1)) == 0 && (1 << kb) == nk && nk <= (1 << r), "step %d: bad candidate count %d", t, nk);
    const int64_t np = (int64_t)1 << n_open;
    TQEC_REQUIRE(h[TQEC_H_OFF_T] >= 0 && h[TQEC_H_OFF_T] + np * nk <= d->n_tables, "step %d: table offset out of range", t);
    const int32_t offs[6] = {h[TQEC_H_OFF_ML], h[TQEC_H_OFF_MK], h[TQEC_H_OFF_A0], h[TQEC_H_OFF_KER], h[TQEC_H_OFF_VARS], h[TQEC_H_OFF_CLOSE]};
    const int64_t lens[6] = {np, nk, np, nk, r, 2 * (int64_t)n_close};
    for (int i = 0; i < 6; ++i)
      TQEC_REQUIRE(offs[i] >= 0 && offs[i] + lens[i] <= d->n_ints, "step %d: int table %d out of range", t, i);
    for (int p = 0; p < np; ++p)
      TQEC_REQUIRE((d->ints[offs[0] + p] >> w_in) == 0, "step %d: representative mask leaves the in-state", t);
    for (int k = 0; k < nk; ++k)
      TQEC_REQUIRE((d->ints[offs[1] + k] >> w_in) == 0, "step %d: kernel mask leaves the in-state", t);
    for (int j = 0; j < r; ++j)
      TQEC_REQUIRE(d->ints[offs[4] + j] >= 0 && d->ints[offs[4] + j] < d->n_vars, "step %d: variable id out of range", t);
    int prev = -1;
    for (int c = 0; c < n_close; ++c) {
      const int slot = d->ints[offs[5] + 2 * c], bit = d->ints[offs[5] + 2 * c + 1];
      TQEC_REQUIRE(slot > prev && slot < w_in + n_open, "step %d: closed slots must ascend inside the full index", t);
      TQEC_REQUIRE(bit >= 0 && bit < d->n_checks, "step %d: syndrome bit %d out of range", t, bit);
      prev = slot;
    }
    w = w_out;
    wmax = wmax > w_in ? wmax : w_in;
    wmax = wmax > w_out ? wmax : w_out;
  }
  TQEC_REQUIRE(w == d->n_obs, "final state has %d bits but n_obs=%d", w, d->n_obs);
  TQEC_REQUIRE(wmax == d->w_max, "w_max=%d does not match the schedule (%d)", d->w_max, wmax);
  for (int o = 0; o < d->n_obs; ++o)
    TQEC_REQUIRE(d->obs_slot[o] >= 0 && d->obs_slot[o] < d->n_obs, "obs_slot[%d] out of range", o);
  return TQEC_OK;
}

template <typename T>
static int upload(void **dst, const T *src, size_t n) {
  TQEC_CUDA(cudaMalloc(dst, (n ? n : 1) * sizeof(T)));
  if (n) TQEC_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return TQEC_OK;
}

extern "C" int tqec_plan_create(const tqec_plan_desc *d, tqec_plan **out) {
  TQEC_REQUIRE(out != nullptr, "tqec_plan_create: out is NULL");
  *out = nullptr;
  int rc = validate_desc(d);
  if (rc) return rc;
  int ndev = 0;
  TQEC_CUDA(cudaGetDeviceCount(&ndev));
  TQEC_REQUIRE(d->device >= 0 && d->device < ndev, "tqec_plan_create: device %d not present (%d visible)", d->device, ndev);
  TQEC_CUDA(cudaSetDevice(d->device));
  cudaDeviceProp prop;
  TQEC_CUDA(cudaGetDeviceProperties(&prop, d->device));

  tqec_plan *p = new tqec_plan();
  std::memset(p, 0, sizeof(*p));
  p->device = d->device;
  p->semiring = d->semiring;
  p->sm_count = prop.multiProcessorCount;

  // launch geometry: ~1024 state entries per team; narrow plans pack several shots per team
  const int target_bits = 10;
  int sg = d->w_max < target_bits ? target_bits - d->w_max : 0;
  if (sg > 6) sg = 6;
  const int tot_bits = d->w_max + sg;
  int T = 32;
  if (tot_bits > 10) T = 1 << (tot_bits - 5 > 8 ? 8 : tot_bits - 5);
  const int nsw = words_for(d->n_checks), ncw = words_for(d->n_vars);
  const size_t smem = 2 * ((size_t)1 << tot_bits) * sizeof(double) + ((size_t)1 << sg) * (nsw + ncw) * sizeof(uint64_t);
  if (smem > (size_t)prop.sharedMemPerBlockOptin) {
    delete p;
    set_error("schedule needs %zu B of shared memory per team (w_max=%d) > %zu available", smem, d->w_max,
              (size_t)prop.sharedMemPerBlockOptin);
    return TQEC_ERR_UNSUPPORTED;
  }
  p->team_threads = T;
  p->shots_per_team = 1 << sg;
  p->smem_bytes = (int)smem;

  const void *kern = d->semiring == TQEC_SEMIRING_MAXPLUS ? (const void *)k_frontier<TQEC_SEMIRING_MAXPLUS>
                                                          : (const void *)k_frontier<TQEC_SEMIRING_SUMPROD>;
  TQEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  TQEC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
  if (per_sm < 1) per_sm = 1;
  p->teams_per_sm = per_sm;
  p->grid_max = per_sm * p->sm_count;

  // back-pointer layout per team
  std::vector<int32_t> bp_off(d->n_steps + 1, 0);
  double cand = 0.0;
  for (int t = 0; t < d->n_steps; ++t) {
    const int32_t *h = d->hdr + t * TQEC_HDR_INTS;
    const int kb = h[TQEC_H_KB];
    int words = 0;
    if (d->semiring == TQEC_SEMIRING_MAXPLUS && kb) {
      const int n_tot = 1 << (h[TQEC_H_WOUT] + sg);
      const int J = (n_tot + T - 1) / T, per_word = 32 / kb;
      words = ((J + per_word - 1) / per_word) * T;
    }
    bp_off[t + 1] = bp_off[t] + words;
    cand += std::ldexp(1.0, h[TQEC_H_WOUT]) * h[TQEC_H_NK];
  }
  p->candidates_per_shot = cand;

  PlanDev &D = p->dev;
  D.n_steps = d->n_steps; D.n_vars = d->n_vars; D.n_checks = d->n_checks; D.n_obs = d->n_obs;
  D.w_max = d->w_max; D.sg_log2 = sg; D.nsw = nsw; D.ncw = ncw; D.bp_words = bp_off[d->n_steps];
  rc = upload(&p->d_hdr, d->hdr, (size_t)d->n_steps * TQEC_HDR_INTS);
  if (!rc) rc = upload(&p->d_ints, d->ints, (size_t)d->n_ints);
  if (!rc) rc = upload(&p->d_tables, d->tables, (size_t)d->n_tables);
  if (!rc) rc = upload(&p->d_bp_off, bp_off.data(), bp_off.size());
  if (!rc) rc = upload(&p->d_obs_slot, d->obs_slot, (size_t)d->n_obs);
  if (!rc) {
    const size_t bytes = (size_t)p->grid_max * (D.bp_words ? D.bp_words : 1) * sizeof(uint32_t);
    cudaError_t e = cudaMalloc((void **)&p->d_bp, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu B back-pointer scratch): %s", bytes, cudaGetErrorString(e)); rc = TQEC_ERR_NOMEM; }
  }
  if (!rc) {
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); rc = TQEC_ERR_CUDA; }
  }
  if (rc) { tqec_plan_destroy(p); return rc; }
  D.hdr = (const int32_t *)p->d_hdr; D.ints = (const int32_t *)p->d_ints; D.tables = (const double *)p->d_tables;
  D.bp_off = (const int32_t *)p->d_bp_off; D.obs_slot = (const int32_t *)p->d_obs_slot;
  *out = p;
  return TQEC_OK;
}

extern "C" int tqec_plan_destroy(tqec_plan *p) {
  if (!p) return TQEC_OK;
  cudaSetDevice(p->device);
  cudaFree(p->d_hdr); cudaFree(p->d_ints); cudaFree(p->d_tables); cudaFree(p->d_bp_off); cudaFree(p->d_obs_slot);
  cudaFree(p->d_bp);
  for (int i = 0; i < 4; ++i) cudaFree(p->d_io[i]);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return TQEC_OK;
}

extern "C" int tqec_plan_query(const tqec_plan *p, int32_t what, int64_t *out) {
  TQEC_REQUIRE(p && out, "tqec_plan_query: NULL argument");
  switch (what) {
    case TQEC_Q_TEAM_THREADS: *out = p->team_threads; break;
    case TQEC_Q_SHOTS_PER_TEAM: *out = p->shots_per_team; break;
    case TQEC_Q_SMEM_BYTES: *out = p->smem_bytes; break;
    case TQEC_Q_GRID: *out = p->grid_max; break;
    case TQEC_Q_TEAMS_PER_SM: *out = p->teams_per_sm; break;
    case TQEC_Q_BP_BYTES_PER_TEAM: *out = (int64_t)p->dev.bp_words * 4; break;
    case TQEC_Q_CANDIDATES_PER_SHOT: *out = (int64_t)p->candidates_per_shot; break;
    case TQEC_Q_SM_COUNT: *out = p->sm_count; break;
    case TQEC_Q_LAUNCHES: *out = p->launches; break;
    default: set_error("tqec_plan_query: unknown item %d", what); return TQEC_ERR_INVALID;
  }
  return TQEC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" int tqec_decode_map_dev(tqec_plan *p, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_logp,
                                   void *stream) {
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_MAXPLUS, "tqec_decode_map: plan is not a max-plus (TNMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (d_synd && d_corr)), "tqec_decode_map: NULL buffer");
  TQEC_CUDA(cudaSetDevice(p->device));
  return launch_decode(p, d_synd, B, d_corr, d_logp, nullptr, (cudaStream_t)stream);
}

extern "C" int tqec_decode_marginal_dev(tqec_plan *p, const uint64_t *d_synd, int64_t B, double *d_mar,
                                        int32_t *d_argmax, void *stream) {
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_SUMPROD, "tqec_decode_marginal: plan is not a sum-product (TNMMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (d_synd && d_mar)), "tqec_decode_marginal: NULL buffer");
  TQEC_CUDA(cudaSetDevice(p->device));
  return launch_decode(p, d_synd, B, nullptr, d_mar, d_argmax, (cudaStream_t)stream);
}

extern "C" int tqec_decode_map(tqec_plan *p, const uint64_t *synd, int64_t B, uint64_t *corr_out, double *logp_out) {
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_MAXPLUS, "tqec_decode_map: plan is not a max-plus (TNMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (synd && corr_out)), "tqec_decode_map: NULL buffer");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(p->device));
  const size_t sb = (size_t)B * p->dev.nsw * 8, cb = (size_t)B * p->dev.ncw * 8, lb = (size_t)B * 8;
  int rc;
  if ((rc = ensure_cap(&p->d_io[0], &p->io_cap[0], sb))) return rc;
  if ((rc = ensure_cap(&p->d_io[1], &p->io_cap[1], cb))) return rc;
  if ((rc = ensure_cap(&p->d_io[2], &p->io_cap[2], lb))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(p->d_io[0], synd, sb, cudaMemcpyHostToDevice, p->stream));
  if ((rc = launch_decode(p, (const uint64_t *)p->d_io[0], B, (uint64_t *)p->d_io[1], (double *)p->d_io[2], nullptr, p->stream))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(corr_out, p->d_io[1], cb, cudaMemcpyDeviceToHost, p->stream));
  if (logp_out) TQEC_CUDA(cudaMemcpyAsync(logp_out, p->d_io[2], lb, cudaMemcpyDeviceToHost, p->stream));
  TQEC_CUDA(cudaStreamSynchronize(p->stream));
  return TQEC_OK;
}

extern "C" int tqec_decode_marginal(tqec_plan *p, const uint64_t *synd, int64_t B, double *mar_out, int32_t *argmax_out) {
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_SUMPROD, "tqec_decode_marginal: plan is not a sum-product (TNMMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (synd && mar_out)), "tqec_decode_marginal: NULL buffer");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(p->device));
  const size_t sb = (size_t)B * p->dev.nsw * 8, mb = ((size_t)B << p->dev.n_obs) * 8, ab = (size_t)B * 4;
  int rc;
  if ((rc = ensure_cap(&p->d_io[0], &p->io_cap[0], sb))) return rc;
  if ((rc = ensure_cap(&p->d_io[1], &p->io_cap[1], mb))) return rc;
  if ((rc = ensure_cap(&p->d_io[2], &p->io_cap[2], ab))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(p->d_io[0], synd, sb, cudaMemcpyHostToDevice, p->stream));
  if ((rc = launch_decode(p, (const uint64_t *)p->d_io[0], B, nullptr, (double *)p->d_io[1], (int32_t *)p->d_io[2], p->stream))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(mar_out, p->d_io[1], mb, cudaMemcpyDeviceToHost, p->stream));
  if (argmax_out) TQEC_CUDA(cudaMemcpyAsync(argmax_out, p->d_io[2], ab, cudaMemcpyDeviceToHost, p->stream));
  TQEC_CUDA(cudaStreamSynchronize(p->stream));
  return TQEC_OK;
}
