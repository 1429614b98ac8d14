// Frontier-recurrence decoding kernels (TNMAP max-plus with traceback, TNMMAP sum-product) for sm_100a.
//
// A TEAM owns 2^sg_log2 shots at a time and keeps their whole state tensor S[sub][sigma] (FP64, <= 2^(w_max+sg_log2)
// entries, ping-pong) in shared memory; nothing but the bit-packed syndromes (in), the corrections / marginals (out)
// and 1..kb back-pointer bits per state entry (L2-resident scratch, max-plus only) ever touches global memory.
// Two launch shapes share one body:
//   warp teams (WT = true) : a team is ONE WARP; a CTA holds as many independent warp-teams as shared memory allows
//                            plus one copy of the schedule tables, so every table read is an LDS and steps are
//                            separated by __syncwarp only.  Used whenever the state fits 1024 entries per team.
//   CTA teams  (WT = false): a team is a whole CTA of 32..256 threads, tables stay in global memory (L1/L2), steps are
//                            separated by __syncthreads.  Used for wide frontiers (w_max > 10) or oversized tables.
// The grid is persistent: teams stride over shot groups.  See tensorqec.jl_b200/schedule.py for the recurrence and the
// table layout, DESIGN.md for the roofline argument (FP64-pipe / issue bound, not HBM bound).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <future>
#include <system_error>
#include <thread>

#include "tqec_common.h"

namespace tqec {

template <bool SM> __device__ __forceinline__ int ldi(const int32_t *p) { return SM ? *p : __ldg(p); }
template <bool SM> __device__ __forceinline__ double ldd(const double *p) { return SM ? *p : __ldg(p); }
template <bool WT> __device__ __forceinline__ void team_sync() {
  if (WT) __syncwarp(); else __syncthreads();
}

// schedule tables as seen by a team (shared-memory copies for warp teams, global memory otherwise)
struct Tabs {
  const int32_t *hdr, *ints, *bp_off, *obs_slot;
  const double *tables;
};

// full index of an output index: bit b goes to full slot perm[b] (perm[w_out] follows the closed list); the shot's
// closed-bit values (precomputed once per shot and step in `cb`, already at their slots) are OR-ed in by the caller.
template <bool SM>
__device__ __forceinline__ int scatter_bits(int tau, int w_out, const int32_t *__restrict__ perm) {
  int full = 0;
  for (int b = 0; b < w_out; ++b) full |= ((tau >> b) & 1) << ldi<SM>(perm + b);
  return full;
}

// ---- generic step: any geometry, per-element index arithmetic ------------------------------------------------------
// LY = index layout of the plan: 0 standard (<= 32 entries per thread), 1 wide (64 / 128 entries per thread), 2 shot-minor
template <int SEMI, int NK, bool SM, int LY>
__device__ __forceinline__ void run_step(const PlanDev &P, const Tabs &X, const int32_t *__restrict__ h,
                                         const double *__restrict__ Sin, double *__restrict__ Sout,
                                         const int32_t *__restrict__ cbt, uint32_t *__restrict__ gtab,
                                         uint32_t *__restrict__ bpt, int T, int tid) {
  const int w_in = ldi<SM>(h + TQEC_H_WIN), n_close = ldi<SM>(h + TQEC_H_NCLOSE), w_out = ldi<SM>(h + TQEC_H_WOUT);
  const int nk = NK > 0 ? NK : ldi<SM>(h + TQEC_H_NK);
  const int kb = ldi<SM>(h + TQEC_H_KB);
  const double *__restrict__ Tt = X.tables + ldi<SM>(h + TQEC_H_OFF_T);
  const int32_t *__restrict__ ML = X.ints + ldi<SM>(h + TQEC_H_OFF_ML);
  const int32_t *__restrict__ MK = X.ints + ldi<SM>(h + TQEC_H_OFF_MK);
  const int32_t *__restrict__ CL = X.ints + ldi<SM>(h + TQEC_H_OFF_CLOSE);
  const int n_tot = 1 << (w_out + P.sg_log2);
  const int inmask = (1 << w_in) - 1, outmask = (1 << w_out) - 1;
  const int per_word = kb ? 32 / kb : 1;
  uint32_t word = 0;
  int jw = 0, wi = 0;
  // the scattered image of the five low output bits is tabulated once per step; higher bits are scattered one by one
  const int32_t *__restrict__ perm = CL + 2 * n_close;
  const int wl = w_out < 5 ? w_out : 5;
  if (tid < (1 << wl)) gtab[tid] = (uint32_t)scatter_bits<SM>(tid, wl, perm);
  if (T == 32) __syncwarp(); else __syncthreads();
  for (int e = tid; e < n_tot; e += T) {
    const int tau = LY == 2 ? e >> P.sg_log2 : e & outmask;
    const int sub = LY == 2 ? e & ((1 << P.sg_log2) - 1) : e >> w_out;
    int full = (int)gtab[tau & 31] | cbt[sub];
    for (int b = 5; b < w_out; ++b) full |= ((tau >> b) & 1) << ldi<SM>(perm + b);
    const int pat = full >> w_in;
    const int low = (full & inmask) ^ ldi<SM>(ML + pat);
    const double *__restrict__ tb = Tt + pat * nk;
    // state entry sigma of shot sub lives at (sub << w) | sigma, or at (sigma << sg) | sub in the shot-minor layout
    const int sh = LY == 2 ? P.sg_log2 : 0;
    const double *__restrict__ Sb = Sin + (LY == 2 ? sub : (sub << w_in));
    const double s0 = Sb[(low ^ ldi<SM>(MK)) << sh], t0 = ldd<SM>(tb);
    double best = SEMI == TQEC_SEMIRING_MAXPLUS ? s0 + t0 : s0 * t0;
    int bk = 0;
#pragma unroll
    for (int k = 1; k < (NK > 0 ? NK : nk); ++k) {
      const double s = Sb[(low ^ ldi<SM>(MK + k)) << sh], tv = ldd<SM>(tb + k);
      if (SEMI == TQEC_SEMIRING_MAXPLUS) {
        const double v = s + tv;
        if (v > best) { best = v; bk = k; }      // strict: the smallest candidate wins exact ties
      } else {
        best += s * tv;
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

// ---- fast step ("state in lane", T = 32) --------------------------------------------------------------------------------
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v));
}

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// One block of U consecutive elements of a thread (j = j0 .. j0+U-1, j0 a multiple of U): U*NK gathers (all issued
// before the first use), U stores, U*kb back-pointer bits.  `gj8` is the j0-dependent byte offset (one broadcast load
// per block); the other elements differ from it by the constants dl[u] (the deposited low bits of j).  Addresses are
// ABSOLUTE shared-memory addresses: the state's base is a multiple of its size, so base, lane part, candidate mask and
// j part combine with XOR only and an address costs one LOP3.
template <int SEMI, int NK, int U, int CJ>
__device__ __forceinline__ uint32_t fast_block(uint32_t so, const uint32_t (&ck8)[NK], const double (&tv)[NK],
                                               const uint32_t (&dl)[8], uint32_t gj8) {
  // CJ > 0 (nk = 2 only): candidate 1 of element u reads exactly what candidate 0 of element u ^ CJ reads (the kernel
  // candidate flips only low j bits), so each input is loaded once and used by both outputs of the pair.
  constexpr int NL = CJ > 0 ? 1 : NK;
  double v[U][NK];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int k = 0; k < NL; ++k) v[u][k] = lds_f64(u == 0 ? (ck8[k] ^ gj8) : xor3(ck8[k], gj8, dl[u]));
  if (CJ > 0) {
#pragma unroll
    for (int u = 0; u < U; ++u) v[u][1 % NK] = v[(u ^ CJ) % U][0];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u][1 % NK] = SEMI == TQEC_SEMIRING_MAXPLUS ? v[u][1 % NK] + tv[1 % NK] : v[u][1 % NK] * tv[1 % NK];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u][0] = SEMI == TQEC_SEMIRING_MAXPLUS ? v[u][0] + tv[0] : v[u][0] * tv[0];
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < NK; ++k) v[u][k] = SEMI == TQEC_SEMIRING_MAXPLUS ? v[u][k] + tv[k] : v[u][k] * tv[k];
  }
  uint32_t m = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    double best;
    if (SEMI == TQEC_SEMIRING_MAXPLUS) {
      // the smallest candidate index wins exact ties: the right operand of every comparison must be strictly greater
      if (NK == 1) {
        best = v[u][0];
      } else if (NK == 2) {
        const bool p = v[u][1] > v[u][0];
        best = p ? v[u][1] : v[u][0];
        if (p) m |= 1u << u;
      } else {
        const bool p01 = v[u][1] > v[u][0], p23 = v[u][3 % NK] > v[u][2 % NK];
        const double b01 = p01 ? v[u][1] : v[u][0], b23 = p23 ? v[u][3 % NK] : v[u][2 % NK];
        const bool pf = b23 > b01;
        best = pf ? b23 : b01;
        const uint32_t bk = pf ? (2u | (uint32_t)p23) : (uint32_t)p01;
        m |= bk << (2 * u);
      }
    } else {
      best = v[u][0];
#pragma unroll
      for (int k = 1; k < NK; ++k) best += v[u][k];
    }
    sts_f64(so + (u << 8), best);
  }
  return m;
}

template <int SEMI, int NK, int U, int CJ>
__device__ __forceinline__ void fast_blocks(int nblk, const double *__restrict__ rec, uint32_t pl8, uint32_t so,
                                            const uint32_t (&dl)[8], const volatile uint32_t *__restrict__ gtab,
                                            uint32_t *__restrict__ bpt, int tid) {
  constexpr int KB = NK == 1 ? 0 : (NK == 2 ? 1 : 2);
  constexpr int RB = NK + (NK + 1) / 2;                 // doubles per block record: NK factor values + packed masks
  uint32_t word = 0;
  int pos = 0, wi = 0;
  for (int b = 0; b < nblk; ++b, rec += RB) {
    uint32_t ck8[NK];
    double tv[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) tv[k] = rec[k];
#pragma unroll
    for (int k = 0; k < NK; k += 2) {
      const int2 c2 = *reinterpret_cast<const int2 *>(rec + NK + k / 2);
      ck8[k] = (uint32_t)c2.x ^ pl8;
      if (k + 1 < NK) ck8[k + 1] = (uint32_t)c2.y ^ pl8;
    }
    const uint32_t m = fast_block<SEMI, NK, U, CJ>(so + ((b * U) << 8), ck8, tv, dl, gtab[b * U]);
    if (SEMI == TQEC_SEMIRING_MAXPLUS && KB) {
      word |= m << pos;
      pos += U * KB;
      if (pos == 32) { bpt[wi * 32 + tid] = word; word = 0; pos = 0; ++wi; }
    }
  }
  if (SEMI == TQEC_SEMIRING_MAXPLUS && KB && pos) bpt[wi * 32 + tid] = word;
}

// The element index e = tid + 32*j splits into lane bits and j bits, and so does every quantity derived from it: the
// re-inserted full index is deposit(tid) | deposit(j << 5) | closed-bit values, the opened pattern and the shot sub-index
// depend on j only.  The host (build_device_tables) precomputes, per step, the deposits of the three low j bits and one
// record per block of U elements {factor values, XOR constants per candidate}; the device adds the lane part and the
// state's base address (a multiple of the state size, so XOR is enough) and keeps the j-dependent part in a 32-entry
// per-team table (one broadcast LDS per block).  A candidate then costs LOP3 + LDS.64 + DADD (+ DSETP + 2 FSEL beyond
// the first).  `sin_abs` / `sout_abs` are absolute shared addresses of the ping-pong states.
template <int SEMI, int NK, bool SM, int LY>
__device__ __forceinline__ void fast_step(const PlanDev &P, const Tabs &X, const int32_t *__restrict__ h,
                                          const int32_t *__restrict__ frec, uint32_t sin_abs, uint32_t sout_abs,
                                          uint32_t *__restrict__ gtab, const int32_t *__restrict__ cbt,
                                          uint32_t *__restrict__ bpt, int tid) {
  const int4 q0 = *reinterpret_cast<const int4 *>(h);            // r, w_in, n_open, n_close
  const int w_in = q0.y, w_out = ldi<SM>(h + TQEC_H_WOUT);
  const int4 f0 = *reinterpret_cast<const int4 *>(frec);         // dl1, dl2, dl4, nblk
  const int4 f1 = *reinterpret_cast<const int4 *>(frec + 4);     // U, off_blk, cj, -
  const int4 f2 = *reinterpret_cast<const int4 *>(frec + 8);     // full slots of output bits 0..3
  const int4 f3 = *reinterpret_cast<const int4 *>(frec + 12);    // full slot of output bit 4, then bits 5..7
  const int4 f4 = *reinterpret_cast<const int4 *>(frec + 16);    // full slots of output bits 8..11 (10, 11: wide states)
  const int lane = tid;
  const int inmask = (1 << w_in) - 1;
  // lane part: output bits 0..4 are lane bits
  const int pl = ((lane & 1) << f2.x) | (((lane >> 1) & 1) << f2.y) | (((lane >> 2) & 1) << f2.z) |
                 (((lane >> 3) & 1) << f2.w) | (((lane >> 4) & 1) << f3.x);
  // j part: output bits 5.. are j bits 0.. (bits beyond w_out select the shot); lane l fills entries l, l+32, ..
  const int nj = w_out - 5;                                      // state bits among the j bits
  const int J = 1 << (nj + P.sg_log2);
  {
    // common case (<= 32 entries per thread): one table entry per lane, straight-line
    const int j = lane;
    int gfull = 0;
    if (nj > 0) gfull |= (j & 1) << f3.y;
    if (nj > 1) gfull |= ((j >> 1) & 1) << f3.z;
    if (nj > 2) gfull |= ((j >> 2) & 1) << f3.w;
    if (nj > 3) gfull |= ((j >> 3) & 1) << f4.x;
    if (nj > 4) gfull |= ((j >> 4) & 1) << f4.y;
    const int gsub = (j >> nj) & ((1 << P.sg_log2) - 1);
    gtab[j] = (uint32_t)(((gfull | cbt[gsub]) & inmask) << 3);
  }
  if (LY == 1)
  for (int j = lane + 32; j < J; j += 32) {                        // wide states: 64 / 128 entries per thread
    int gfull = ((j & 1) << f3.y) | (((j >> 1) & 1) << f3.z) | (((j >> 2) & 1) << f3.w) | (((j >> 3) & 1) << f4.x) |
                (((j >> 4) & 1) << f4.y);
    if (nj > 5) gfull |= ((j >> 5) & 1) << f4.z;
    if (nj > 6) gfull |= ((j >> 6) & 1) << f4.w;
    const int gsub = (j >> nj) & ((1 << P.sg_log2) - 1);
    gtab[j] = (uint32_t)(((gfull | cbt[gsub]) & inmask) << 3);
  }
  const uint32_t pl8 = (uint32_t)(pl << 3) ^ sin_abs;
  uint32_t dl[8];
  dl[0] = 0; dl[1] = (uint32_t)f0.x; dl[2] = (uint32_t)f0.y; dl[4] = (uint32_t)f0.z;
  dl[3] = dl[1] | dl[2]; dl[5] = dl[1] | dl[4]; dl[6] = dl[2] | dl[4]; dl[7] = dl[3] | dl[4];
  const uint32_t so = sout_abs + (tid << 3);
  const double *__restrict__ rec = X.tables + f1.y;
  __syncwarp();
  if (NK == 2 && f1.z == 1) {
    if (f1.x == 8) fast_blocks<SEMI, NK, 8, 1>(f0.w, rec, pl8, so, dl, gtab, bpt, tid);
    else fast_blocks<SEMI, NK, 4, 1>(f0.w, rec, pl8, so, dl, gtab, bpt, tid);
  } else if (NK == 2 && f1.z == 3) {
    if (f1.x == 8) fast_blocks<SEMI, NK, 8, 3>(f0.w, rec, pl8, so, dl, gtab, bpt, tid);
    else fast_blocks<SEMI, NK, 4, 3>(f0.w, rec, pl8, so, dl, gtab, bpt, tid);
  } else if (f1.x == 8) fast_blocks<SEMI, NK, (NK == 4 ? 4 : 8), 0>(f0.w, rec, pl8, so, dl, gtab, bpt, tid);
  else if (f1.x == 4) fast_blocks<SEMI, NK, 4, 0>(f0.w, rec, pl8, so, dl, gtab, bpt, tid);
  else fast_blocks<SEMI, NK, 1, 0>(f0.w, rec, pl8, so, dl, gtab, bpt, tid);
}

// ---- quad step: two factors absorbed at once (4 candidates per output, 2 opened checks), 16-output blocks -----------------
// The planner (schedule.py, "quad") arranges that the candidates leaving the closed checks alone form a 4-element group
// that realises the 4 opened patterns p and whose generators flip exactly the checks at output bits 5 and 6 (u): output
// (u, p) then reads the inputs L[u ^ (p & PM)][k], k = 0..3, where L is a 4 x 4 patch of the input state (4 values of the
// two u-checks x the 4 kernel candidates).  A thread's block of 16 outputs (u, p) therefore needs 16 loads -- one per
// output instead of four -- and every loaded value feeds 4 outputs from registers.  PM says which generator flips a
// surviving check (3: both; 1 / 2: only the first / second -- the other u bit is then an unrelated check).
template <int SEMI, int PM>
__device__ __forceinline__ void quad_step(const Tabs &X, const int32_t *__restrict__ h, const int32_t *__restrict__ qrec,
                                          uint32_t sin_abs, uint32_t sout_abs, const int32_t *__restrict__ cbt,
                                          uint32_t *__restrict__ bpt, int lane) {
  const int w_in = h[TQEC_H_WIN];
  const int4 r0 = *reinterpret_cast<const int4 *>(qrec);        // byte masks of u bit 0, u bit 1, kernel candidates 1, 2
  const int4 r1 = *reinterpret_cast<const int4 *>(qrec + 4);    // off_T, blocks per thread (shots per pass), pm, -
  const int4 f2 = *reinterpret_cast<const int4 *>(qrec + 8);    // full slots of output bits 0..3
  const int f3 = qrec[12];                                      // full slot of output bit 4
  const int pl = ((lane & 1) << f2.x) | (((lane >> 1) & 1) << f2.y) | (((lane >> 2) & 1) << f2.z) |
                 (((lane >> 3) & 1) << f2.w) | (((lane >> 4) & 1) << f3);
  const double *__restrict__ T = X.tables + r1.x;               // T[p * 4 + k]
  double tv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) tv[i] = T[i];
  const uint32_t inmask = (1u << w_in) - 1u;
  const uint32_t km[4] = {0u, (uint32_t)r0.z, (uint32_t)r0.w, (uint32_t)(r0.z ^ r0.w)};
#pragma unroll 1
  for (int sub = 0; sub < r1.y; ++sub) {
    const uint32_t base = sin_abs ^ (uint32_t)(pl << 3) ^ ((((uint32_t)cbt[sub] & inmask) | ((uint32_t)sub << w_in)) << 3);
    const uint32_t so = sout_abs + (lane << 3) + ((uint32_t)sub << 12);   // bits 8..11 (u, p) of the offset are free
    uint32_t word = 0;
    // one row of the patch per iteration (NOT unrolled: the body must stay small enough for the instruction cache):
    // the 4 inputs L[v][0..3] feed the 4 outputs (u = v ^ (p & PM), p), p = 0..3.  Back-pointers of a quad step are
    // filed under (v, p) -- 2 bits at position 2 * (v + 4 p) of the shot's word -- so that every shift is static.
#pragma unroll 1
    for (int v = 0; v < 4; ++v) {
      const uint32_t row = base ^ ((v & 1) ? (uint32_t)r0.x : 0u) ^ ((v & 2) ? (uint32_t)r0.y : 0u);
      const uint32_t so_v = so ^ ((uint32_t)v << 8);
      double L[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) L[k] = lds_f64(row ^ km[k]);
      uint32_t rb = 0;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        double c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = SEMI == TQEC_SEMIRING_MAXPLUS ? L[k] + tv[p * 4 + k] : L[k] * tv[p * 4 + k];
        double best;
        if (SEMI == TQEC_SEMIRING_MAXPLUS) {
          // the smallest candidate index wins exact ties
          const bool p01 = c[1] > c[0], p23 = c[3] > c[2];
          const double b01 = p01 ? c[1] : c[0], b23 = p23 ? c[3] : c[2];
          const bool pf = b23 > b01;
          best = pf ? b23 : b01;
          const uint32_t bk = pf ? (2u | (uint32_t)p23) : (uint32_t)p01;
          rb |= bk << (8 * p);
        } else {
          best = (c[0] + c[1]) + (c[2] + c[3]);
        }
        sts_f64((so_v ^ (uint32_t)((p & PM) << 8)) + (p << 10), best);       // u = v ^ (p & PM)
      }
      word |= rb << (2 * v);
    }
    if (SEMI == TQEC_SEMIRING_MAXPLUS) bpt[sub * 32 + lane] = word;
  }
}

// ---- fast step, shot-minor layout ("shot in lane": plans with w_max <= 5 pack 32 shots per team, lane = shot) ---------------
// The state of shot `lane` sits at ((sigma << 5) | lane) * 8, so every lane walks ALL output indices tau of its own shot:
// the index arithmetic is uniform across the warp (host-precomputed per step: UK[tau][k] = source index << 8 and the
// factor value TV[tau][k]), the only per-lane term is the shot's closed-bit value, bank conflicts cannot occur, and a
// candidate costs LOP3 + LDS.64 + DADD plus a share of two broadcast table loads.
template <int SEMI, int NK>
__device__ __forceinline__ void fast_step_b(const int32_t *__restrict__ brec, const int32_t *__restrict__ ints,
                                            const double *__restrict__ tables, uint32_t sin_abs, uint32_t sout_abs,
                                            int cbv, int w_in, uint32_t *__restrict__ bpt, int lane) {
  constexpr int KB = NK == 1 ? 0 : (NK == 2 ? 1 : 2);
  const int4 b0 = *reinterpret_cast<const int4 *>(brec);          // n_tau, off_uk, off_tv, -
  const int n_tau = b0.x;
  const int32_t *__restrict__ UK = ints + b0.y;
  const double *__restrict__ TV = tables + b0.z;
  const uint32_t lc = sin_abs ^ (uint32_t)(lane << 3) ^ (uint32_t)((cbv & ((1 << w_in) - 1)) << 8);
  const uint32_t so = sout_abs + (lane << 3);
  uint32_t word = 0;
  int pos = 0, wi = 0;
#pragma unroll 4
  for (int tau = 0; tau < n_tau; ++tau) {
    double v[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) v[k] = lds_f64(lc ^ (uint32_t)UK[tau * NK + k]);
#pragma unroll
    for (int k = 0; k < NK; ++k) v[k] = SEMI == TQEC_SEMIRING_MAXPLUS ? v[k] + TV[tau * NK + k] : v[k] * TV[tau * NK + k];
    double best;
    uint32_t bk = 0;
    if (SEMI == TQEC_SEMIRING_MAXPLUS) {
      if (NK == 1) {
        best = v[0];
      } else if (NK == 2) {
        const bool p = v[1] > v[0];
        best = p ? v[1] : v[0];
        bk = p;
      } else {
        const bool p01 = v[1] > v[0], p23 = v[3 % NK] > v[2 % NK];
        const double b01 = p01 ? v[1] : v[0], b23 = p23 ? v[3 % NK] : v[2 % NK];
        const bool pf = b23 > b01;
        best = pf ? b23 : b01;
        bk = pf ? (2u | (uint32_t)p23) : (uint32_t)p01;
      }
    } else {
      best = v[0];
#pragma unroll
      for (int k = 1; k < NK; ++k) best += v[k];
    }
    sts_f64(so + (tau << 8), best);
    if (SEMI == TQEC_SEMIRING_MAXPLUS && KB) {
      word |= bk << pos;
      pos += KB;
      if (pos == 32) { bpt[wi * 32 + lane] = word; word = 0; pos = 0; ++wi; }
    }
  }
  if (SEMI == TQEC_SEMIRING_MAXPLUS && KB && pos) bpt[wi * 32 + lane] = word;
}

// ---- one traceback step (shared by both traceback shapes): returns the previous state index and the factor's bits ------
template <bool SM>
__device__ __forceinline__ int trace_step(const Tabs &X, const int32_t *__restrict__ h, int tau, int cbv, int k, int &a_out) {
  const int4 q0 = *reinterpret_cast<const int4 *>(h);            // r, w_in, n_open, n_close
  const int w_in = q0.y;
  const int32_t *CL = X.ints + ldi<SM>(h + TQEC_H_OFF_CLOSE);
  const int full = scatter_bits<SM>(tau, ldi<SM>(h + TQEC_H_WOUT), CL + 2 * q0.w) | cbv;
  const int pat = full >> w_in;
  // (assignment, in-state mask) of candidate k for this opened pattern, precomputed by the host
  const int2 am = *reinterpret_cast<const int2 *>(X.ints + ldi<SM>(h + TQEC_H_AM) + 2 * (pat * ldi<SM>(h + TQEC_H_NK) + k));
  a_out = am.x;
  return (full & ((1 << w_in) - 1)) ^ am.y;
}

// ---- the team body ---------------------------------------------------------------------------------------------------
// forward sweep over all steps for the SG shots whose syndrome words sit in sh_syn; returns the final state
template <int SEMI, bool WT, int LY>
__device__ __forceinline__ double *forward_pass(const PlanDev &P, const Tabs &X, double *S0, const uint64_t *sh_syn,
                                                uint32_t *gtab, int32_t *cb, uint32_t *__restrict__ bp, int T, int tid) {
  constexpr bool SM = WT;
  const int SG = 1 << P.sg_log2;
  const int NS = 1 << (P.w_max + P.sg_log2);
  double *S1 = S0 + NS;
  const double ONE = SEMI == TQEC_SEMIRING_MAXPLUS ? 0.0 : 1.0;
  for (int i = tid; i < SG; i += T) S0[i] = ONE;
  // closed-bit values of every (step, shot): sum_c syndrome_bit(c) << slot(c)
  for (int i = tid; i < (P.n_steps << P.sg_log2); i += T) {
    const int t = i >> P.sg_log2, sub = i & (SG - 1);
    const int32_t *h = X.hdr + t * TQEC_HDR_INTS;
    const int n_close = ldi<SM>(h + TQEC_H_NCLOSE);
    const int32_t *CL = X.ints + ldi<SM>(h + TQEC_H_OFF_CLOSE);
    int v = 0;
    for (int c = 0; c < n_close; ++c) {
      const int bit = ldi<SM>(CL + 2 * c + 1);
      v |= (int)((sh_syn[sub * P.nsw + (bit >> 6)] >> (bit & 63)) & 1ull) << ldi<SM>(CL + 2 * c);
    }
    cb[i] = v;
  }
  team_sync<WT>();

  double *Sin = S0, *Sout = S1;
  uint32_t sin_abs = (uint32_t)__cvta_generic_to_shared(S0), sout_abs = (uint32_t)__cvta_generic_to_shared(S1);
  for (int t = 0; t < P.n_steps; ++t) {
    const int32_t *h = X.hdr + t * TQEC_HDR_INTS;
    uint32_t *bpt = bp + ldi<SM>(X.bp_off + t);
    const int nk = ldi<SM>(h + TQEC_H_NK);
    const int fo = WT ? ldi<SM>(h + TQEC_H_FAST) : 0;
    const int32_t *cbt = cb + (t << P.sg_log2);
    if (WT && fo > 0 && LY == 2) {
      const int32_t *brec = X.ints + (fo - 1);
      const int w_in = ldi<SM>(h + TQEC_H_WIN);
      if (nk == 2) fast_step_b<SEMI, 2>(brec, X.ints, X.tables, sin_abs, sout_abs, cbt[tid], w_in, bpt, tid);
      else if (nk == 1) fast_step_b<SEMI, 1>(brec, X.ints, X.tables, sin_abs, sout_abs, cbt[tid], w_in, bpt, tid);
      else fast_step_b<SEMI, 4>(brec, X.ints, X.tables, sin_abs, sout_abs, cbt[tid], w_in, bpt, tid);
    } else if (WT && LY == 0 && fo < 0) {
      const int32_t *qrec = X.ints + (-fo - 1);
      const int pm = qrec[6];
      if (pm == 3) quad_step<SEMI, 3>(X, h, qrec, sin_abs, sout_abs, cbt, bpt, tid);
      else if (pm == 1) quad_step<SEMI, 1>(X, h, qrec, sin_abs, sout_abs, cbt, bpt, tid);
      else quad_step<SEMI, 2>(X, h, qrec, sin_abs, sout_abs, cbt, bpt, tid);
    } else if (WT && fo > 0 && LY != 2) {
      const int32_t *frec = X.ints + (fo - 1);
      if (nk == 2) fast_step<SEMI, 2, SM, LY>(P, X, h, frec, sin_abs, sout_abs, gtab, cbt, bpt, tid);
      else if (nk == 1) fast_step<SEMI, 1, SM, LY>(P, X, h, frec, sin_abs, sout_abs, gtab, cbt, bpt, tid);
      else fast_step<SEMI, 4, SM, LY>(P, X, h, frec, sin_abs, sout_abs, gtab, cbt, bpt, tid);
    } else {
      switch (nk) {
        case 1: run_step<SEMI, 1, SM, LY>(P, X, h, Sin, Sout, cbt, gtab, bpt, T, tid); break;
        case 2: run_step<SEMI, 2, SM, LY>(P, X, h, Sin, Sout, cbt, gtab, bpt, T, tid); break;
        case 4: run_step<SEMI, 4, SM, LY>(P, X, h, Sin, Sout, cbt, gtab, bpt, T, tid); break;
        default: run_step<SEMI, 0, SM, LY>(P, X, h, Sin, Sout, cbt, gtab, bpt, T, tid); break;
      }
    }
    team_sync<WT>();
    double *tmp = Sin; Sin = Sout; Sout = tmp;
    const uint32_t to = sin_abs; sin_abs = sout_abs; sout_abs = to;
  }
  return Sin;
}

// back-pointers of a quad step (standard layout, w_out = 9): element tau = lane | u << 5 | p << 7 of shot `sub` is
// filed under (v = u ^ (p & pm), p) in word sub * 32 + lane
__device__ __forceinline__ void bp_locate_quad(int tau, int sub, int pm, int &word, int &sh) {
  const int u = (tau >> 5) & 3, p = (tau >> 7) & 3;
  word = sub * 32 + (tau & 31);
  sh = 2 * ((u ^ (p & pm)) + 4 * p);
}

// back-pointer of output element e of a step: word index / shift inside the team's per-step block
__device__ __forceinline__ void bp_locate(int e, int LT, int T, int kb, int &word, int &sh) {
  const int j = e >> LT, ln = e & (T - 1);
  int wi;
  if (kb == 1) { wi = j >> 5; sh = j & 31; }
  else if (kb == 2) { wi = j >> 4; sh = (j & 15) << 1; }
  else { const int pw = 32 / kb; wi = j / pw; sh = (j - wi * pw) * kb; }
  word = wi * T + ln;
}

template <int SEMI, bool WT, int LY>
__device__ __forceinline__ void team_run(const PlanDev &P, const Tabs &X, unsigned char *smem, int s0_off, uint64_t *sh_syn,
                                         uint64_t *sh_cfg, uint32_t *gtab, int32_t *cb, uint32_t *__restrict__ bp, int T,
                                         int LT, int tid, int64_t g_first, int64_t g_stride,
                                         const uint64_t *__restrict__ synd, int64_t B, uint64_t *__restrict__ corr,
                                         double *__restrict__ out, int32_t *__restrict__ argmax_out) {
  constexpr bool SM = WT;
  const int SG = 1 << P.sg_log2;
  double *S0 = reinterpret_cast<double *>(smem + s0_off);
  // Deferred traceback (max-plus warp teams): the team runs the forward sweeps of 32 consecutive shots one pass after
  // the other (each pass packs SG shots), keeping every pass's back-pointers in its own slice of the scratch; then lane
  // q walks the back-pointers of shot q, so the serial traceback costs one warp-instruction stream per 32 shots, not
  // per shot.  Lane q keeps its shot's syndrome and configuration words in registers (nsw, ncw <= 4).
  const bool defer = SEMI == TQEC_SEMIRING_MAXPLUS && WT && P.defer;
  const int QS = defer ? 32 : SG;                       // shots per group
  const int NF = defer ? (32 >> P.sg_log2) : 1;         // forward passes per group
  const int64_t n_groups = (B + QS - 1) / QS;

  for (int64_t g = g_first; g < n_groups; g += g_stride) {
    const int64_t group0 = g * QS, myshot = group0 + tid;
    uint64_t syn[4] = {0ull, 0ull, 0ull, 0ull};
    if (defer && myshot < B)
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (w < P.nsw) syn[w] = synd[myshot * P.nsw + w];

    for (int f = 0; f < NF; ++f) {
      const int64_t shot0 = group0 + ((int64_t)f << P.sg_log2);
      if (shot0 >= B) break;
      if (defer) {
        if ((tid >> P.sg_log2) == f) {
          const int sub = tid & (SG - 1);
#pragma unroll
          for (int w = 0; w < 4; ++w)
            if (w < P.nsw) sh_syn[sub * P.nsw + w] = syn[w];
        }
      } else {
        for (int i = tid; i < SG * P.nsw; i += T) {
          const int64_t s = shot0 + i / P.nsw;
          sh_syn[i] = s < B ? synd[shot0 * P.nsw + i] : 0ull;
        }
        if (SEMI == TQEC_SEMIRING_MAXPLUS)
          for (int i = tid; i < SG * P.ncw; i += T) sh_cfg[i] = 0ull;
      }
      team_sync<WT>();
      const double *Sin = forward_pass<SEMI, WT, LY>(P, X, S0, sh_syn, gtab, cb, bp + (size_t)f * P.bp_words, T, tid);

      if (defer) {
        if (out && tid < SG && shot0 + tid < B) out[shot0 + tid] = Sin[tid];
      } else if (SEMI == TQEC_SEMIRING_MAXPLUS) {
        // one thread per shot walks the back-pointers from the scalar root to the first step
        for (int sub = tid; sub < SG; sub += T) {
          int tau = 0;
          uint64_t *cfg = sh_cfg + sub * P.ncw;
          for (int t = P.n_steps - 1; t >= 0; --t) {
            const int32_t *h = X.hdr + t * TQEC_HDR_INTS;
            const int kb = ldi<SM>(h + TQEC_H_KB), r = ldi<SM>(h + TQEC_H_R);
            int k = 0;
            if (kb) {
              int word, sh;
              const int fq = WT ? ldi<SM>(h + TQEC_H_FAST) : 0;
              if (LY == 0 && fq < 0) bp_locate_quad(tau, sub, ldi<SM>(X.ints + (-fq - 1) + 6), word, sh);
              else bp_locate(LY == 2 ? ((tau << P.sg_log2) | sub) : ((sub << ldi<SM>(h + TQEC_H_WOUT)) | tau), LT, T, kb, word, sh);
              k = (__ldcg(bp + ldi<SM>(X.bp_off + t) + word) >> sh) & ((1u << kb) - 1u);
            }
            int a;
            tau = trace_step<SM>(X, h, tau, cb[(t << P.sg_log2) + sub], k, a);
            const int32_t *V = X.ints + ldi<SM>(h + TQEC_H_OFF_VARS);
            for (int j = 0; j < r; ++j)
              if ((a >> j) & 1) {
                const int v = ldi<SM>(V + j);
                cfg[v >> 6] |= 1ull << (v & 63);
              }
          }
          if (shot0 + sub < B && out) out[shot0 + sub] = Sin[sub];
        }
        team_sync<WT>();
        for (int i = tid; i < SG * P.ncw; i += T)
          if (shot0 + i / P.ncw < B) corr[shot0 * P.ncw + i] = sh_cfg[i];
      } else {
        const int NO = 1 << P.n_obs;
        for (int i = tid; i < SG * NO; i += T) {
          const int sub = i >> P.n_obs, idx = i & (NO - 1);
          int src = 0;
          for (int o = 0; o < P.n_obs; ++o) src |= ((idx >> o) & 1) << ldi<SM>(X.obs_slot + o);
          if (shot0 + sub < B) out[(shot0 + sub) * NO + idx] = Sin[LY == 2 ? ((src << P.sg_log2) | sub) : ((sub << P.n_obs) | src)];
        }
        if (argmax_out) {
          for (int sub = tid; sub < SG; sub += T) {
            if (shot0 + sub >= B) continue;
            double best = -1.0;
            int bi = 0;
            for (int idx = 0; idx < NO; ++idx) {             // first maximal entry (findmax)
              int src = 0;
              for (int o = 0; o < P.n_obs; ++o) src |= ((idx >> o) & 1) << ldi<SM>(X.obs_slot + o);
              const double v = Sin[LY == 2 ? ((src << P.sg_log2) | sub) : ((sub << P.n_obs) | src)];
              if (v > best) { best = v; bi = idx; }
            }
            argmax_out[shot0 + sub] = bi;
          }
        }
      }
      team_sync<WT>();
    }

    if (defer) {
      // lane q: traceback of shot q of the group
      const int f = tid >> P.sg_log2, sub = tid & (SG - 1);
      const uint32_t *bpq = bp + (size_t)f * P.bp_words;
      uint64_t cfg[4] = {0ull, 0ull, 0ull, 0ull};
      int tau = 0;
      for (int t = P.n_steps - 1; t >= 0; --t) {
        const int32_t *h = X.hdr + t * TQEC_HDR_INTS;
        const int4 q0 = *reinterpret_cast<const int4 *>(h);          // r, w_in, n_open, n_close
        const int4 q1 = *reinterpret_cast<const int4 *>(h + 4);      // w_out, nk, kb, off_T
        const int4 q3 = *reinterpret_cast<const int4 *>(h + 12);     // off_vars, off_close, fast, am
        const int kb = q1.z;
        int k = 0;
        if (kb) {
          int word, sh;
          if (LY == 0 && q3.z < 0) bp_locate_quad(tau, sub, ldi<SM>(X.ints + (-q3.z - 1) + 6), word, sh);
          else bp_locate(LY == 2 ? ((tau << P.sg_log2) | sub) : ((sub << q1.x) | tau), 5, 32, kb, word, sh);
          k = (__ldcg(bpq + ldi<SM>(X.bp_off + t) + word) >> sh) & ((1u << kb) - 1u);
        }
        const int32_t *CL = X.ints + q3.y;
        int full = scatter_bits<SM>(tau, q1.x, CL + 2 * q0.w);
        for (int c = 0; c < q0.w; ++c) {
          const int slot = ldi<SM>(CL + 2 * c), bit = ldi<SM>(CL + 2 * c + 1), w = bit >> 6;
          const uint64_t sw = w == 0 ? syn[0] : (w == 1 ? syn[1] : (w == 2 ? syn[2] : syn[3]));
          full |= (int)((sw >> (bit & 63)) & 1ull) << slot;
        }
        const int pat = full >> q0.y;
        const int2 am = *reinterpret_cast<const int2 *>(X.ints + q3.w + 2 * (pat * q1.y + k));
        const int32_t *V = X.ints + q3.x;
        for (int j = 0; j < q0.x; ++j) {
          const int v = ldi<SM>(V + j);
          const uint64_t bitv = (uint64_t)((am.x >> j) & 1) << (v & 63);
          const int w = v >> 6;
          cfg[0] |= w == 0 ? bitv : 0ull; cfg[1] |= w == 1 ? bitv : 0ull;
          cfg[2] |= w == 2 ? bitv : 0ull; cfg[3] |= w == 3 ? bitv : 0ull;
        }
        tau = (full & ((1 << q0.y) - 1)) ^ am.y;
      }
      if (myshot < B)
#pragma unroll
        for (int w = 0; w < 4; ++w)
          if (w < P.ncw) corr[myshot * P.ncw + w] = cfg[w];
      __syncwarp();
    }
  }
}

// entries of the per-team j table: one per element a thread owns in a step (32 threads per warp team), at least 32
__host__ __device__ inline int gtab_entries(int w_max, int sg) {
  const int tot = w_max + sg;
  return tot > 10 && tot <= 12 ? 1 << (tot - 5) : 32;
}
// per-team words: syndrome words, configuration words, the j table, closed-bit values per (step, shot)
__host__ __device__ inline size_t team_words_bytes(int w_max, int sg, int nsw, int ncw, int n_steps) {
  const size_t b = ((size_t)1 << sg) * (size_t)(nsw + ncw) * sizeof(uint64_t) + gtab_entries(w_max, sg) * sizeof(uint32_t) +
                   ((size_t)n_steps << sg) * sizeof(int32_t);
  return (b + 15) & ~(size_t)15;
}
// per-team shared memory of a CTA team: state ping-pong + words
__host__ __device__ inline size_t team_smem_bytes(int w_max, int sg, int nsw, int ncw, int n_steps) {
  return 2 * ((size_t)1 << (w_max + sg)) * sizeof(double) + team_words_bytes(w_max, sg, nsw, ncw, n_steps);
}

// CTA teams: blockDim.x threads form one team, tables in global memory
template <int SEMI>
__global__ void k_frontier_cta(const PlanDev P, const uint64_t *__restrict__ synd, const int64_t B,
                               uint64_t *__restrict__ corr, double *__restrict__ out, int32_t *__restrict__ argmax_out,
                               uint32_t *__restrict__ bp_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x, tid = threadIdx.x;
  const int LT = 31 - __clz(T);
  const int SG = 1 << P.sg_log2;
  uint64_t *sh_syn = reinterpret_cast<uint64_t *>(smem_raw + ((size_t)16 << (P.w_max + P.sg_log2)));
  uint64_t *sh_cfg = sh_syn + SG * P.nsw;
  uint32_t *gtab = reinterpret_cast<uint32_t *>(sh_cfg + SG * P.ncw);
  int32_t *cb = reinterpret_cast<int32_t *>(gtab + gtab_entries(P.w_max, P.sg_log2));
  Tabs X{P.hdr, P.ints, P.bp_off, P.obs_slot, P.tables};
  team_run<SEMI, false, 0>(P, X, smem_raw, 0, sh_syn, sh_cfg, gtab, cb, bp_all + (size_t)blockIdx.x * P.bp_words, T, LT, tid, blockIdx.x,
                        gridDim.x, synd, B, corr, out, argmax_out);
}

// warp teams: every warp of the CTA is an independent team; the schedule tables are staged in shared memory once
template <int SEMI, int LY>
__global__ void k_frontier_warp(const PlanDev P, const uint64_t *__restrict__ synd, const int64_t B,
                                uint64_t *__restrict__ corr, double *__restrict__ out, int32_t *__restrict__ argmax_out,
                                uint32_t *__restrict__ bp_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NW = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SG = 1 << P.sg_log2;
  // layout (offsets chosen by the host, tqec_plan_create): the NW ping-pong states start at an ABSOLUTE shared address
  // that is a multiple of the state size, so a state base can be XOR-folded into element offsets; the gap in front of
  // them (the driver reserves the first KiB of the window) and the tail hold the table copies and the per-team words.
  const size_t state_bytes = (size_t)16 << (P.w_max + P.sg_log2);
  if (((uint32_t)__cvta_generic_to_shared(smem_raw) + P.off_states) & (uint32_t)(state_bytes / 2 - 1)) __trap();
  const size_t words_bytes = team_words_bytes(P.w_max, P.sg_log2, P.nsw, P.ncw, P.n_steps);
  unsigned char *words0 = smem_raw + P.off_words;
  double *sm_tables = reinterpret_cast<double *>(smem_raw + P.off_tables);
  int32_t *sm_hdr = reinterpret_cast<int32_t *>(smem_raw + P.off_ints);   // 16-byte aligned; 64 B per step
  int32_t *sm_ints = sm_hdr + P.n_steps * TQEC_HDR_INTS;                  // stays 16-byte aligned (int4 / int2 loads)
  int32_t *sm_bpoff = sm_ints + P.n_ints;
  int32_t *sm_obs = sm_bpoff + P.n_steps + 1;
  for (int i = threadIdx.x; i < P.n_tables; i += blockDim.x) sm_tables[i] = P.tables[i];
  for (int i = threadIdx.x; i < P.n_steps * TQEC_HDR_INTS; i += blockDim.x) sm_hdr[i] = P.hdr[i];
  for (int i = threadIdx.x; i <= P.n_steps; i += blockDim.x) sm_bpoff[i] = P.bp_off[i];
  for (int i = threadIdx.x; i < P.n_obs; i += blockDim.x) sm_obs[i] = P.obs_slot[i];
  for (int i = threadIdx.x; i < P.n_ints; i += blockDim.x) sm_ints[i] = P.ints[i];
  __syncthreads();
  uint64_t *sh_syn = reinterpret_cast<uint64_t *>(words0 + words_bytes * warp);
  uint64_t *sh_cfg = sh_syn + SG * P.nsw;
  uint32_t *gtab = reinterpret_cast<uint32_t *>(sh_cfg + SG * P.ncw);
  int32_t *cb = reinterpret_cast<int32_t *>(gtab + gtab_entries(P.w_max, P.sg_log2));
  Tabs X{sm_hdr, sm_ints, sm_bpoff, sm_obs, sm_tables};
  const int64_t team = (int64_t)blockIdx.x * NW + warp;
  team_run<SEMI, true, LY>(P, X, smem_raw, (int)(P.off_states + state_bytes * warp), sh_syn, sh_cfg, gtab, cb, bp_all + (size_t)team * P.bp_words * (P.defer ? (32 >> P.sg_log2) : 1), 32, 5, lane, team,
                       (int64_t)gridDim.x * NW, synd, B, corr, out, argmax_out);
}

// ---- fully tabulated plans -------------------------------------------------------------------------------------------
// A plan with at most TQEC_TABLE_BITS syndrome bits is decoded ONCE for every syndrome at compile time (by the kernels
// above, so the table holds exactly their outputs) and `decode` becomes a gather: HBM-bound, 8 B in and 8..40 B out per
// shot.  This is the tabulated head of the sweep taken to its end (and the reference's own TableDecoder,
// src/decoding/truthtable.jl, built from the MAP decoder instead of from error enumeration).
#define TQEC_TABLE_BITS 16
__global__ void k_lookup(const uint64_t *__restrict__ synd, int64_t B, int nsw, uint64_t mask,
                         const uint64_t *__restrict__ tab_corr, int ncw, const double *__restrict__ tab_out, int n_out,
                         const int32_t *__restrict__ tab_arg, uint64_t *__restrict__ corr, double *__restrict__ out,
                         int32_t *__restrict__ argmax_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t s = synd[i * nsw] & mask;
    if (corr)
      for (int w = 0; w < ncw; ++w) corr[i * ncw + w] = __ldg(tab_corr + s * ncw + w);
    if (out)
      for (int k = 0; k < n_out; ++k) out[i * n_out + k] = __ldg(tab_out + s * n_out + k);
    if (argmax_out) argmax_out[i] = __ldg(tab_arg + s);
  }
}

// undo the static power-of-two scaling of a sum-product plan's tables (exact)
__global__ void k_scale_out(double *__restrict__ out, int64_t n, int e) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ldexp(out[i], e);
}

static int launch_decode_raw(tqec_plan *plan, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_out,
                             int32_t *d_argmax, cudaStream_t stream);

int launch_decode(tqec_plan *plan, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_out,
                  int32_t *d_argmax, cudaStream_t stream) {
  int rc = launch_decode_raw(plan, d_synd, B, d_corr, d_out, d_argmax, stream);
  if (rc || B <= 0) return rc;
  // tabulated plans hold already rescaled values (the table was filled through this function)
  if (plan->semiring == TQEC_SEMIRING_SUMPROD && plan->log2_scale != 0 && !plan->has_table && d_out) {
    const int64_t n = B << plan->dev.n_obs;
    const int64_t want = (n + 255) / 256;
    const int grid = (int)(want < (int64_t)plan->sm_count * 8 ? want : (int64_t)plan->sm_count * 8);
    k_scale_out<<<grid, 256, 0, stream>>>(d_out, n, plan->log2_scale);
    TQEC_CUDA(cudaGetLastError());
  }
  return TQEC_OK;
}

static int launch_decode_raw(tqec_plan *plan, const uint64_t *d_synd, int64_t B, uint64_t *d_corr, double *d_out,
                             int32_t *d_argmax, cudaStream_t stream) {
  if (B <= 0) return TQEC_OK;
  if (plan->has_table) {
    const bool mp = plan->semiring == TQEC_SEMIRING_MAXPLUS;
    const int64_t want = (B + 255) / 256;
    const int grid = (int)(want < (int64_t)plan->sm_count * 16 ? want : (int64_t)plan->sm_count * 16);
    k_lookup<<<grid, 256, 0, stream>>>(d_synd, B, plan->dev.nsw, ((uint64_t)1 << plan->dev.n_checks) - 1, plan->d_tab_corr,
                                       plan->dev.ncw, plan->d_tab_out, mp ? 1 : (1 << plan->dev.n_obs), plan->d_tab_arg,
                                       mp ? d_corr : nullptr, d_out, mp ? nullptr : d_argmax);
    TQEC_CUDA(cudaGetLastError());
    plan->launches += 1;
    return TQEC_OK;
  }
  if (plan->has_wide) return launch_wide(plan, d_synd, B, d_out, d_argmax, stream);
  if (plan->has_sweep) return launch_sweep(plan, d_synd, B, d_corr, d_out, d_argmax, stream);
  const int64_t per_group = plan->dev.defer ? 32 : plan->shots_per_team;
  const int64_t groups = (B + per_group - 1) / per_group;
  const bool mp = plan->semiring == TQEC_SEMIRING_MAXPLUS;
  if (plan->warp_teams) {
    const int64_t ctas = (groups + plan->teams_per_cta - 1) / plan->teams_per_cta;
    const int grid = (int)(ctas < plan->grid_max ? ctas : plan->grid_max);
    const int threads = 32 * plan->teams_per_cta;
    const int ly = plan->layout;
#define TQEC_LAUNCH_WARP(SEMI, LY, CORR, ARG) \
  k_frontier_warp<SEMI, LY><<<grid, threads, plan->smem_bytes, stream>>>(plan->dev, d_synd, B, CORR, d_out, ARG, plan->d_bp)
    if (mp) {
      if (ly == 2) TQEC_LAUNCH_WARP(TQEC_SEMIRING_MAXPLUS, 2, d_corr, nullptr);
      else if (ly == 1) TQEC_LAUNCH_WARP(TQEC_SEMIRING_MAXPLUS, 1, d_corr, nullptr);
      else TQEC_LAUNCH_WARP(TQEC_SEMIRING_MAXPLUS, 0, d_corr, nullptr);
    } else {
      if (ly == 2) TQEC_LAUNCH_WARP(TQEC_SEMIRING_SUMPROD, 2, nullptr, d_argmax);
      else if (ly == 1) TQEC_LAUNCH_WARP(TQEC_SEMIRING_SUMPROD, 1, nullptr, d_argmax);
      else TQEC_LAUNCH_WARP(TQEC_SEMIRING_SUMPROD, 0, nullptr, d_argmax);
    }
#undef TQEC_LAUNCH_WARP
  } else {
    const int grid = (int)(groups < plan->grid_max ? groups : plan->grid_max);
    if (mp) k_frontier_cta<TQEC_SEMIRING_MAXPLUS><<<grid, plan->team_threads, plan->smem_bytes, stream>>>(plan->dev, d_synd, B, d_corr, d_out, nullptr, plan->d_bp);
    else k_frontier_cta<TQEC_SEMIRING_SUMPROD><<<grid, plan->team_threads, plan->smem_bytes, stream>>>(plan->dev, d_synd, B, nullptr, d_out, d_argmax, plan->d_bp);
  }
  TQEC_CUDA(cudaGetLastError());
  plan->launches += 1;
  return TQEC_OK;
}

}  // namespace tqec

using namespace tqec;

// Device-side tables derived from the ABI schedule (host, once per plan):
//   hdr[t][TQEC_H_AM]   -> ints: (assignment, in-state mask) pairs of candidate k for opened pattern p at 2*(p*nk + k),
//                          i.e. A0[p]^KER[k] and ML[p]^MK[k] -- one 8-byte load per traceback step
//   hdr[t][TQEC_H_FAST] -> 0, or 1 + offset (16-byte aligned) of the fast-step record {dl1, dl2, dl4, nblk, U, off_blk}:
//                          dl* = byte offsets of the deposited three low j bits; one record per block of U elements in
//                          `tables` at off_blk: NK factor values, then the NK XOR constants ((ML^MK | sub << w_in) << 3)
//                          packed two per double.  A step is fast iff the element index splits into lane bits and j
//                          bits (see fast_step): 32-thread teams, <= 32 elements per thread, every closed slot below
//                          w_in, the opened pattern determined by j alone, nk in {1, 2, 4}.
static void build_device_tables(const tqec_plan_desc *d, int sg, bool want_fast, bool sub_minor, std::vector<int32_t> &hdr,
                                std::vector<int32_t> &ints, std::vector<double> &tables) {
  hdr.assign(d->hdr, d->hdr + (size_t)d->n_steps * TQEC_HDR_INTS);
  ints.assign(d->ints, d->ints + d->n_ints);
  tables.assign(d->tables, d->tables + d->n_tables);
  const int LT = 5;
  for (int t = 0; t < d->n_steps; ++t) {
    int32_t *h = hdr.data() + (size_t)t * TQEC_HDR_INTS;
    const int w_in = h[TQEC_H_WIN], n_open = h[TQEC_H_NOPEN], n_close = h[TQEC_H_NCLOSE], w_out = h[TQEC_H_WOUT];
    const int nk = h[TQEC_H_NK], np = 1 << n_open;
    const int32_t *perm = d->ints + h[TQEC_H_OFF_CLOSE] + 2 * n_close;
    const int32_t *ML = d->ints + h[TQEC_H_OFF_ML], *MK = d->ints + h[TQEC_H_OFF_MK];
    const int32_t *A0 = d->ints + h[TQEC_H_OFF_A0], *KER = d->ints + h[TQEC_H_OFF_KER];
    if (ints.size() & 1) ints.push_back(0);
    h[TQEC_H_AM] = (int32_t)ints.size();
    for (int pp = 0; pp < np; ++pp)
      for (int k = 0; k < nk; ++k) { ints.push_back(A0[pp] ^ KER[k]); ints.push_back(ML[pp] ^ MK[k]); }
    h[TQEC_H_FAST] = 0;
    if (sub_minor) {
      // shot-minor fast step (fast_step_b): uniform per-tau tables; needs nk in {1,2,4} and an opened pattern that does
      // not depend on the shot (no closed slot among the opened slots)
      bool okb = want_fast && (nk == 1 || nk == 2 || nk == 4) && w_out <= 5;
      for (int c = 0; c < n_close && okb; ++c) if (d->ints[h[TQEC_H_OFF_CLOSE] + 2 * c] >= w_in) okb = false;
      if (!okb) continue;
      const int inmask_b = (1 << w_in) - 1;
      const double *Tb = d->tables + h[TQEC_H_OFF_T];
      while (ints.size() & 3) ints.push_back(0);
      const size_t rec_at = ints.size();
      ints.push_back(1 << w_out); ints.push_back(0); ints.push_back((int32_t)tables.size()); ints.push_back(0);
      ints[rec_at + 1] = (int32_t)ints.size();
      for (int tau = 0; tau < (1 << w_out); ++tau) {
        int full = 0;
        for (int b = 0; b < w_out; ++b) full |= ((tau >> b) & 1) << perm[b];
        const int pat = full >> w_in;
        for (int k = 0; k < nk; ++k) {
          ints.push_back((int32_t)((uint32_t)(((full & inmask_b) ^ ML[pat] ^ MK[k])) << 8));
          tables.push_back(Tb[pat * nk + k]);
        }
      }
      h[TQEC_H_FAST] = (int32_t)rec_at + 1;
      continue;
    }
    // quad step (quad_step): 16-output blocks (u = output bits 5, 6; p = the two opened bits on top), w_out = 9
    if (want_fast && nk == 4 && n_open == 2 && w_out == 9 && sg <= 1 && std::getenv("TQEC_NO_QUAD") == nullptr) {
      bool okq = perm[7] == w_in && perm[8] == w_in + 1 && MK[3] == (MK[1] ^ MK[2]);
      for (int b = 0; b < 7 && okq; ++b) if (perm[b] >= w_in) okq = false;
      int pm = 0;
      for (int cand = 3; cand >= 1 && okq && pm == 0; --cand) {
        bool m = true;
        for (int pp = 0; pp < 4 && m; ++pp) {
          const int want = (((pp & 1) && (cand & 1)) ? (1 << perm[5]) : 0) ^ (((pp & 2) && (cand & 2)) ? (1 << perm[6]) : 0);
          if (ML[pp] != want) m = false;
        }
        if (m) pm = cand;
      }
      if (okq && pm) {
        while (ints.size() & 3) ints.push_back(0);
        h[TQEC_H_FAST] = -((int32_t)ints.size() + 1);
        ints.push_back((1 << perm[5]) << 3); ints.push_back((1 << perm[6]) << 3);
        ints.push_back(MK[1] << 3); ints.push_back(MK[2] << 3);
        ints.push_back(h[TQEC_H_OFF_T]); ints.push_back(1 << sg); ints.push_back(pm); ints.push_back(0);
        for (int b = 0; b < 5; ++b) ints.push_back(perm[b]);
        ints.push_back(0); ints.push_back(0); ints.push_back(0);
        continue;
      }
    }
    const int lgJ = w_out + sg - LT, lg_jj = w_out - LT - n_open;
    bool ok = want_fast && lgJ >= 0 && lgJ <= 7 && lg_jj >= 0 && (nk == 1 || nk == 2 || nk == 4);
    // lane bits must land inside the input state; the opened slots must be the top output bits, in order
    for (int b = 0; b < 5 && ok; ++b) if (perm[b] >= w_in) ok = false;
    for (int i = 0; i < n_open && ok; ++i) if (perm[w_out - n_open + i] != w_in + i) ok = false;
    if (!ok) continue;
    const int inmask = (1 << w_in) - 1;
    const int J = 1 << lgJ, njj = 1 << lg_jj;
    const int umax = nk == 4 ? 4 : 8;
    const int U = njj >= umax ? umax : (njj >= 4 ? 4 : 1);
    auto slot8 = [&](int b) { return b < w_out ? (int32_t)(((1 << perm[b]) & inmask) << 3) : 0; };
    // pairing (nk = 2): the kernel candidate's mask must be exactly the image of the lowest one or two j bits
    int cj = 0;
    if (nk == 2 && U >= 4) {
      const int m1 = MK[1];
      if (w_out > 5 && m1 == ((1 << perm[5]) & inmask) && m1 != 0) cj = 1;
      else if (w_out > 6 && m1 == (((1 << perm[5]) | (1 << perm[6])) & inmask) && perm[5] < w_in && perm[6] < w_in) cj = 3;
    }
    while (ints.size() & 3) ints.push_back(0);
    h[TQEC_H_FAST] = (int32_t)ints.size() + 1;
    ints.push_back(slot8(5)); ints.push_back(slot8(6)); ints.push_back(slot8(7));
    ints.push_back(J / U);
    ints.push_back(U == umax ? 8 : U);       // 8 selects the widest block of this candidate count
    ints.push_back((int32_t)tables.size());
    ints.push_back(cj); ints.push_back(0);
    for (int b = 0; b < 12; ++b) ints.push_back(b < w_out ? perm[b] : 0);
    const double *Tt = d->tables + h[TQEC_H_OFF_T];
    for (int b = 0; b < J / U; ++b) {
      const int grp = (b * U) >> lg_jj, pat = grp & (np - 1), sub = grp >> n_open;
      for (int k = 0; k < nk; ++k) tables.push_back(Tt[pat * nk + k]);
      for (int k = 0; k < nk; k += 2) {
        int32_t c2[2] = {0, 0};
        for (int q = 0; q < 2 && k + q < nk; ++q) c2[q] = (int32_t)((((uint32_t)(ML[pat] ^ MK[k + q]) | ((uint32_t)sub << w_in)) << 3));
        double packed;
        std::memcpy(&packed, c2, sizeof(packed));
        tables.push_back(packed);
      }
    }
  }
}

static const void *warp_kernel(int semiring, int layout) {
  if (semiring == TQEC_SEMIRING_MAXPLUS)
    return layout == 2 ? (const void *)k_frontier_warp<TQEC_SEMIRING_MAXPLUS, 2>
                       : (layout == 1 ? (const void *)k_frontier_warp<TQEC_SEMIRING_MAXPLUS, 1> : (const void *)k_frontier_warp<TQEC_SEMIRING_MAXPLUS, 0>);
  return layout == 2 ? (const void *)k_frontier_warp<TQEC_SEMIRING_SUMPROD, 2>
                     : (layout == 1 ? (const void *)k_frontier_warp<TQEC_SEMIRING_SUMPROD, 1> : (const void *)k_frontier_warp<TQEC_SEMIRING_SUMPROD, 0>);
}

// ---------------------------------------------------------------------------------------------------------------
static int validate_desc(const tqec_plan_desc *d) {
  TQEC_REQUIRE(d != nullptr, "tqec_plan_create: desc is NULL");
  TQEC_REQUIRE(d->semiring == TQEC_SEMIRING_MAXPLUS || d->semiring == TQEC_SEMIRING_SUMPROD,
               "tqec_plan_create: unknown semiring %d", d->semiring);
  TQEC_REQUIRE(d->n_vars >= 0 && d->n_checks >= 0 && d->n_obs >= 0 && d->n_obs <= 16,
               "tqec_plan_create: bad sizes (n_vars=%d n_checks=%d n_obs=%d)", d->n_vars, d->n_checks, d->n_obs);
  TQEC_REQUIRE(d->semiring == TQEC_SEMIRING_SUMPROD || d->n_obs == 0, "tqec_plan_create: max-plus plans have no open axes");
  if (d->wide) return TQEC_OK;                                   // validated by wide_create
  TQEC_REQUIRE(d->n_steps > 0 && d->hdr && d->ints && d->tables, "tqec_plan_create: empty schedule");
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
    uint32_t used = 0;
    for (int c = 0; c < n_close; ++c) {
      const int slot = d->ints[offs[5] + 2 * c], bit = d->ints[offs[5] + 2 * c + 1];
      TQEC_REQUIRE(slot >= 0 && slot < w_in + n_open && !((used >> slot) & 1u), "step %d: bad closed slot %d", t, slot);
      TQEC_REQUIRE(bit >= 0 && bit < d->n_checks, "step %d: syndrome bit %d out of range", t, bit);
      used |= 1u << slot;
    }
    TQEC_REQUIRE(offs[5] + 2 * (int64_t)n_close + w_out <= d->n_ints, "step %d: output permutation out of range", t);
    for (int b = 0; b < w_out; ++b) {
      const int slot = d->ints[offs[5] + 2 * n_close + b];
      TQEC_REQUIRE(slot >= 0 && slot < w_in + n_open && !((used >> slot) & 1u), "step %d: output bit %d maps to a bad or repeated slot %d", t, b, slot);
      used |= 1u << slot;
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

// Plans with few syndrome bits: decode every syndrome once with the kernels of this plan; decode() is a table look-up
// from here on.  table_bits: 0 = default (TQEC_TABLE_BITS), < 0 = never, at most 26.
static int tabulate_plan(tqec_plan *p, const tqec_plan_desc *d) {
  if (d->table_bits < 0) return TQEC_OK;
  const int ncw = words_for(d->n_vars);
  int rc = TQEC_OK;
  const int table_bits = d->table_bits == 0 ? TQEC_TABLE_BITS : (d->table_bits > 26 ? 26 : d->table_bits);
  if (d->n_checks >= 1 && d->n_checks <= table_bits && std::getenv("TQEC_NO_TABLE") == nullptr) {
    const int64_t N = (int64_t)1 << d->n_checks;
    const bool mp = d->semiring == TQEC_SEMIRING_MAXPLUS;
    const int n_out = mp ? 1 : (1 << d->n_obs);
    // cap the table by bytes as well: (configuration words + outputs + argmax) per syndrome, at most 2 GiB
    if ((double)N * (ncw * 8.0 + n_out * 8.0 + 4.0) > 2147483648.0) return TQEC_OK;
    std::vector<uint64_t> all((size_t)N);
    for (int64_t s = 0; s < N; ++s) all[(size_t)s] = (uint64_t)s;
    uint64_t *d_all = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_all, (size_t)N * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_tab_corr, (size_t)N * ncw * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_tab_out, (size_t)N * n_out * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->d_tab_arg, (size_t)N * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_all, all.data(), (size_t)N * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(p->d_tab_corr, 0, (size_t)N * ncw * 8);
    if (e != cudaSuccess) { set_error("tabulating the plan: %s", cudaGetErrorString(e)); cudaFree(d_all); return TQEC_ERR_NOMEM; }
    rc = launch_decode(p, d_all, N, mp ? p->d_tab_corr : nullptr, p->d_tab_out, mp ? nullptr : p->d_tab_arg, p->stream);
    if (!rc && cudaStreamSynchronize(p->stream) != cudaSuccess) { set_error("tabulating the plan: %s", cudaGetErrorString(cudaGetLastError())); rc = TQEC_ERR_CUDA; }
    cudaFree(d_all);
    if (rc) return rc;
    p->has_table = 1;
    p->launches = 0;
  }
  return TQEC_OK;
}

extern "C" int tqec_plan_create(const tqec_plan_desc *d, tqec_plan **out) {
  tqec::NvtxRange nvtx_range("tqec_plan_create");
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
  p->log2_scale = d->semiring == TQEC_SEMIRING_SUMPROD ? d->log2_scale : 0;
  p->sm_count = prop.multiProcessorCount;
  if (d->wide) {
    // global-memory executor: none of the on-chip machinery below applies
    p->dev.n_vars = d->n_vars; p->dev.n_checks = d->n_checks; p->dev.n_obs = d->n_obs;
    p->dev.nsw = words_for(d->n_checks); p->dev.ncw = words_for(d->n_vars);
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); delete p; return TQEC_ERR_CUDA; }
    rc = wide_create(p, d, prop);
    if (rc) { tqec_plan_destroy(p); return rc; }
    rc = tabulate_plan(p, d);
    if (rc) { tqec_plan_destroy(p); return rc; }
    *out = p;
    return TQEC_OK;
  }

  // launch geometry: ~1024 state entries per team; narrow plans pack several shots per team
  int target_bits = 10;
  if (const char *e = std::getenv("TQEC_TARGET_BITS")) { const int v = std::atoi(e); if (v >= 5 && v <= 10) target_bits = v; }
  int sg = d->w_max < target_bits ? target_bits - d->w_max : 0;
  if (sg > 6) sg = 6;
  // narrow plans (w_max <= 5): 32 shots per team in the shot-minor layout (lane = shot), see fast_step_b
  const bool narrow = d->w_max <= 5 && std::getenv("TQEC_NO_SUB_MINOR") == nullptr && std::getenv("TQEC_NO_WARP_TEAMS") == nullptr;
  if (narrow) sg = 5;
  const int tot_bits = d->w_max + sg;
  int T = 32;
  // up to 2^12 state entries a team is one warp (64 / 128 entries per thread and step); beyond that a 256-thread CTA
  if (tot_bits > 12 || (tot_bits > 10 && std::getenv("TQEC_NO_WIDE_WARP") != nullptr)) T = 1 << (tot_bits - 5 > 8 ? 8 : tot_bits - 5);
  const int layout = narrow ? 2 : (tot_bits > 10 ? 1 : 0);   // index layout of the warp-team kernel (see run_step)
  p->layout = layout;
  const int nsw = words_for(d->n_checks), ncw = words_for(d->n_vars);
  const bool want_warp = T == 32 && std::getenv("TQEC_NO_WARP_TEAMS") == nullptr;

  // device tables: the ABI pools plus, per step, the traceback (assignment, mask) pairs and the fast-step records
  std::vector<int32_t> hdr, ints;
  std::vector<double> tables;
  build_device_tables(d, sg, want_warp && std::getenv("TQEC_NO_FAST") == nullptr, narrow && want_warp, hdr, ints, tables);

  const size_t per_team = team_smem_bytes(d->w_max, sg, nsw, ncw, d->n_steps);
  const size_t budget = (size_t)prop.sharedMemPerBlockOptin;
  // warp-team layout: [front gap: whatever fits of the three blobs] [states, size-aligned absolute address] [rest]
  const size_t state_bytes = (size_t)16 << tot_bits, align = state_bytes / 2;
  const size_t ints_bytes = ((hdr.size() + d->n_steps + 1 + d->n_obs + ints.size()) * 4 + 15) & ~(size_t)15;
  const size_t tables_bytes = (tables.size() * 8 + 15) & ~(size_t)15;
  const size_t words_team = team_words_bytes(d->w_max, sg, nsw, ncw, d->n_steps);
  int reserved = 1024;
  cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, d->device);
  const size_t gap = (align - (size_t)reserved % align) % align;
  int nw = 0;
  size_t off_states = gap, off_ints = 0, off_tables = 0, off_words = 0, smem_warp = 0;
  if (want_warp) {
    // teams per CTA: bounded by the register file (one CTA per SM: teams x 32 threads x registers per thread)
    int cap = 32;
    {
      cudaFuncAttributes fa;
      const void *wk = warp_kernel(d->semiring, layout);
      if (cudaFuncGetAttributes(&fa, wk) == cudaSuccess && fa.numRegs > 0) {
        const int regs = (fa.numRegs + 7) & ~7;
        int by_regs = prop.regsPerBlock / (regs * 32);
        if (fa.maxThreadsPerBlock / 32 < by_regs) by_regs = fa.maxThreadsPerBlock / 32;
        if (by_regs < cap) cap = by_regs;
      }
      if (cap < 1) cap = 1;
    }
    if (const char *e = std::getenv("TQEC_TEAMS_PER_CTA")) { const int v = std::atoi(e); if (v >= 1 && v < cap) cap = v; }
    for (int cand = cap; cand >= 1 && nw == 0; --cand) {
      size_t front = 0, tail = gap + state_bytes * cand;
      const size_t sizes[3] = {ints_bytes, tables_bytes, words_team * cand};
      size_t offs[3];
      for (int i = 0; i < 3; ++i) {
        if (front + sizes[i] <= gap) { offs[i] = front; front += sizes[i]; }
        else { offs[i] = tail; tail += sizes[i]; }
      }
      if (tail <= budget) { nw = cand; off_ints = offs[0]; off_tables = offs[1]; off_words = offs[2]; smem_warp = tail; }
    }
  }
  p->warp_teams = nw >= 1;
  p->teams_per_cta = nw >= 1 ? nw : 1;
  const size_t smem = p->warp_teams ? smem_warp : per_team;
  if (smem > budget) {
    delete p;
    set_error("schedule needs %zu B of shared memory per team (w_max=%d) > %zu available", smem, d->w_max, budget);
    return TQEC_ERR_UNSUPPORTED;
  }
  p->team_threads = T;
  p->shots_per_team = 1 << sg;
  p->smem_bytes = (int)smem;

  const void *kern;
  if (p->warp_teams) kern = warp_kernel(d->semiring, layout);
  else kern = d->semiring == TQEC_SEMIRING_MAXPLUS ? (const void *)k_frontier_cta<TQEC_SEMIRING_MAXPLUS>
                                                   : (const void *)k_frontier_cta<TQEC_SEMIRING_SUMPROD>;
  TQEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  TQEC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, p->warp_teams ? 32 * nw : T, smem));
  if (per_sm < 1) per_sm = 1;
  p->teams_per_sm = per_sm * p->teams_per_cta;
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
  D.n_ints = (int32_t)ints.size(); D.n_tables = (int32_t)tables.size();
  D.sub_minor = (narrow && p->warp_teams) ? 1 : 0;
  D.defer = (p->warp_teams && d->semiring == TQEC_SEMIRING_MAXPLUS && nsw <= 4 && ncw <= 4 && sg <= 5 &&
             std::getenv("TQEC_NO_DEFER") == nullptr) ? 1 : 0;
  D.off_states = (int32_t)off_states; D.off_ints = (int32_t)off_ints; D.off_tables = (int32_t)off_tables; D.off_words = (int32_t)off_words;
  rc = upload(&p->d_hdr, hdr.data(), hdr.size());
  if (!rc) rc = upload(&p->d_ints, ints.data(), ints.size());
  if (!rc) rc = upload(&p->d_tables, tables.data(), tables.size());
  if (!rc) rc = upload(&p->d_bp_off, bp_off.data(), bp_off.size());
  if (!rc) rc = upload(&p->d_obs_slot, d->obs_slot, (size_t)d->n_obs);
  if (!rc) {
    const size_t passes = D.defer ? (size_t)(32 >> sg) : 1;
    const size_t bytes = (size_t)p->grid_max * p->teams_per_cta * passes * (D.bp_words ? D.bp_words : 1) * sizeof(uint32_t);
    cudaError_t e = cudaMalloc((void **)&p->d_bp, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu B back-pointer scratch): %s", bytes, cudaGetErrorString(e)); rc = TQEC_ERR_NOMEM; }
  }
  if (!rc) {
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); rc = TQEC_ERR_CUDA; }
  }
  if (!rc) rc = sweep_create(p, d, prop);
  if (rc) { tqec_plan_destroy(p); return rc; }
  D.hdr = (const int32_t *)p->d_hdr; D.ints = (const int32_t *)p->d_ints; D.tables = (const double *)p->d_tables;
  D.bp_off = (const int32_t *)p->d_bp_off; D.obs_slot = (const int32_t *)p->d_obs_slot;
  rc = tabulate_plan(p, d);
  if (rc) { tqec_plan_destroy(p); return rc; }
  *out = p;
  return TQEC_OK;
}

extern "C" int tqec_plan_destroy(tqec_plan *p) {
  if (!p) return TQEC_OK;
  cudaSetDevice(p->device);
  cudaFree(p->d_hdr); cudaFree(p->d_ints); cudaFree(p->d_tables); cudaFree(p->d_bp_off); cudaFree(p->d_obs_slot);
  cudaFree(p->d_bp);
  cudaFree(p->d_tab_corr); cudaFree(p->d_tab_out); cudaFree(p->d_tab_arg);
  sweep_destroy(p);
  wide_destroy(p);
  cudaFree(p->d_mc);
  for (int i = 0; i < 4; ++i) cudaFree(p->d_io[i]);
  for (int i = 0; i < 3; ++i) if (p->h_pin[i]) cudaFreeHost(p->h_pin[i]);
  if (p->stream) cudaStreamDestroy(p->stream);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  for (int i = 0; i < 2; ++i) {
    if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
    if (p->ev_cmp[i]) cudaEventDestroy(p->ev_cmp[i]);
    if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
  }
  delete p;
  return TQEC_OK;
}

extern "C" int tqec_plan_query(const tqec_plan *p, int32_t what, int64_t *out) {
  TQEC_REQUIRE(p && out, "tqec_plan_query: NULL argument");
  switch (what) {
    case TQEC_Q_TEAM_THREADS: *out = p->team_threads; break;
    case TQEC_Q_SHOTS_PER_TEAM: *out = p->has_sweep ? (1 << p->sw.sg) : p->shots_per_team; break;
    case TQEC_Q_SMEM_BYTES: *out = p->has_sweep ? p->sw_smem : p->smem_bytes; break;
    case TQEC_Q_GRID: *out = p->has_sweep ? p->sm_count : p->grid_max; break;
    case TQEC_Q_TEAMS_PER_SM: *out = p->has_sweep ? p->sw_teams : p->teams_per_sm; break;
    case TQEC_Q_BP_BYTES_PER_TEAM: *out = p->has_sweep ? (int64_t)p->sw.bp_words * 128 : (int64_t)p->dev.bp_words * 4; break;
    case TQEC_Q_CANDIDATES_PER_SHOT: *out = (int64_t)p->candidates_per_shot; break;
    case TQEC_Q_SM_COUNT: *out = p->sm_count; break;
    case TQEC_Q_LAUNCHES: *out = p->launches; break;
    case TQEC_Q_SWEEP: *out = p->has_sweep; break;
    case TQEC_Q_TABLE: *out = p->has_table; break;
    case TQEC_Q_WIDE: *out = p->has_wide; break;
    case TQEC_Q_WIDE_BATCH: *out = p->wd_batch; break;
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

// chunk size of the host pipelines: about `want` shots, rounded down to whole rounds of k_sweep when the plan runs on it
static int64_t pipeline_chunk(const tqec_plan *p, int64_t want) {
  if (const char *e = std::getenv("TQEC_PIPE_CHUNK_LOG2")) { const int v = std::atoi(e); if (v >= 10 && v <= 24) want = (int64_t)1 << v; }
  if (!p->has_sweep) return want;
  const int64_t per_round = (int64_t)p->sm_count * p->sw_teams * p->sw.grp;
  return per_round > 0 && want > per_round ? (want / per_round) * per_round : want;
}

static int ensure_pipeline(tqec_plan *p) {
  if (p->s_in) return TQEC_OK;
  TQEC_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  TQEC_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    TQEC_CUDA(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
    TQEC_CUDA(cudaEventCreateWithFlags(&p->ev_cmp[i], cudaEventDisableTiming));
    TQEC_CUDA(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
  }
  return TQEC_OK;
}

extern "C" int tqec_decode_map(tqec_plan *p, const uint64_t *synd, int64_t B, uint64_t *corr_out, double *logp_out) {
  tqec::NvtxRange nvtx_range("tqec_decode_map");
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_MAXPLUS, "tqec_decode_map: plan is not a max-plus (TNMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (synd && corr_out)), "tqec_decode_map: NULL buffer");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(p->device));
  const size_t sb = (size_t)B * p->dev.nsw * 8, cb = (size_t)B * p->dev.ncw * 8, lb = (size_t)B * 8;
  int rc;
  if ((rc = ensure_cap(&p->d_io[0], &p->io_cap[0], sb))) return rc;
  if ((rc = ensure_cap(&p->d_io[1], &p->io_cap[1], cb))) return rc;
  if ((rc = ensure_cap(&p->d_io[2], &p->io_cap[2], lb))) return rc;
  // Chunked three-stage pipeline: H2D of chunk c+1 and D2H of chunk c-1 overlap the decode of chunk c (three streams,
  // one event pair per chunk; the device buffers hold the whole batch, so chunks never alias).
  if ((rc = ensure_pipeline(p))) return rc;
  const int64_t CH = pipeline_chunk(p, std::min<int64_t>((int64_t)1 << 20, std::max<int64_t>((int64_t)1 << 18, B / 4)));   // >= 4 chunks in flight when the batch allows
  const int nsw = p->dev.nsw, ncw = p->dev.ncw;
  const uint64_t *d_syn = (const uint64_t *)p->d_io[0];
  uint64_t *d_cor = (uint64_t *)p->d_io[1];
  double *d_lp = (double *)p->d_io[2];
  int slot = 0;
  for (int64_t o = 0; o < B; o += CH, slot ^= 1) {
    const int64_t n = B - o < CH ? B - o : CH;
    TQEC_CUDA(cudaMemcpyAsync((void *)(d_syn + o * nsw), synd + o * nsw, (size_t)n * nsw * 8, cudaMemcpyHostToDevice, p->s_in));
    TQEC_CUDA(cudaEventRecord(p->ev_in[slot], p->s_in));
    TQEC_CUDA(cudaStreamWaitEvent(p->stream, p->ev_in[slot], 0));
    if ((rc = launch_decode(p, d_syn + o * nsw, n, d_cor + o * ncw, d_lp + o, nullptr, p->stream))) return rc;
    TQEC_CUDA(cudaEventRecord(p->ev_cmp[slot], p->stream));
    TQEC_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_cmp[slot], 0));
    TQEC_CUDA(cudaMemcpyAsync(corr_out + o * ncw, d_cor + o * ncw, (size_t)n * ncw * 8, cudaMemcpyDeviceToHost, p->s_out));
    if (logp_out) TQEC_CUDA(cudaMemcpyAsync(logp_out + o, d_lp + o, (size_t)n * 8, cudaMemcpyDeviceToHost, p->s_out));
    // (an event slot is re-recorded two chunks later; a stream wait binds to the record that preceded it)
  }
  TQEC_CUDA(cudaStreamSynchronize(p->s_out));
  TQEC_CUDA(cudaStreamSynchronize(p->stream));
  return TQEC_OK;
}

extern "C" int tqec_decode_marginal(tqec_plan *p, const uint64_t *synd, int64_t B, double *mar_out, int32_t *argmax_out) {
  tqec::NvtxRange nvtx_range("tqec_decode_marginal");
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_SUMPROD, "tqec_decode_marginal: plan is not a sum-product (TNMMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (synd && mar_out)), "tqec_decode_marginal: NULL buffer");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(p->device));
  const size_t sb = (size_t)B * p->dev.nsw * 8, mb = ((size_t)B << p->dev.n_obs) * 8, ab = (size_t)B * 4;
  int rc;
  if ((rc = ensure_cap(&p->d_io[0], &p->io_cap[0], sb))) return rc;
  if ((rc = ensure_cap(&p->d_io[1], &p->io_cap[1], mb))) return rc;
  if ((rc = ensure_cap(&p->d_io[2], &p->io_cap[2], ab))) return rc;
  // same chunked three-stage pipeline as tqec_decode_map
  if ((rc = ensure_pipeline(p))) return rc;
  const int64_t CH = pipeline_chunk(p, std::min<int64_t>((int64_t)1 << 20, std::max<int64_t>((int64_t)1 << 18, B / 4)));   // >= 4 chunks in flight when the batch allows
  const int nsw = p->dev.nsw;
  const int64_t NO = (int64_t)1 << p->dev.n_obs;
  const uint64_t *d_syn = (const uint64_t *)p->d_io[0];
  double *d_mar = (double *)p->d_io[1];
  int32_t *d_arg = (int32_t *)p->d_io[2];
  int slot = 0;
  for (int64_t o = 0; o < B; o += CH, slot ^= 1) {
    const int64_t n = B - o < CH ? B - o : CH;
    TQEC_CUDA(cudaMemcpyAsync((void *)(d_syn + o * nsw), synd + o * nsw, (size_t)n * nsw * 8, cudaMemcpyHostToDevice, p->s_in));
    TQEC_CUDA(cudaEventRecord(p->ev_in[slot], p->s_in));
    TQEC_CUDA(cudaStreamWaitEvent(p->stream, p->ev_in[slot], 0));
    if ((rc = launch_decode(p, d_syn + o * nsw, n, nullptr, d_mar + o * NO, d_arg + o, p->stream))) return rc;
    TQEC_CUDA(cudaEventRecord(p->ev_cmp[slot], p->stream));
    TQEC_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_cmp[slot], 0));
    TQEC_CUDA(cudaMemcpyAsync(mar_out + o * NO, d_mar + o * NO, (size_t)n * NO * 8, cudaMemcpyDeviceToHost, p->s_out));
    if (argmax_out) TQEC_CUDA(cudaMemcpyAsync(argmax_out + o, d_arg + o, (size_t)n * 4, cudaMemcpyDeviceToHost, p->s_out));
  }
  TQEC_CUDA(cudaStreamSynchronize(p->s_out));
  TQEC_CUDA(cudaStreamSynchronize(p->stream));
  return TQEC_OK;
}

// ---- byte-per-bit entry points: Vector{Mod2} / numpy uint8 in and out, packing done on the device ------------------------
// The reference's containers hold one byte per bit (Mod2 wraps Bool: src/codes/mod2.jl:20-41; a batch is a Matrix{Mod2}
// with one shot per column = `n_bits` contiguous bytes per shot).  Packing 1e7 x 80 bits on the host costs more than the
// decode; here the bytes go over PCIe as they are and two small kernels convert them next to the decode kernel.
namespace tqec {
__global__ void k_pack_bits(const uint8_t *__restrict__ bytes, int64_t B, int n_bits, int words, uint64_t *__restrict__ out) {
  const int64_t n = B * words;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t shot = i / words;
    const int w = (int)(i - shot * words);
    const uint8_t *src = bytes + shot * n_bits + 64 * w;
    const int m = n_bits - 64 * w < 64 ? n_bits - 64 * w : 64;
    uint64_t v = 0;
    for (int k = 0; k < m; ++k) v |= (uint64_t)(src[k] & 1u) << k;
    out[i] = v;
  }
}
// the same for syndromes kept as two arrays (a CSS code: sx with n_a bits and sz with n_b bits per shot): the chunk arrives as
// the block of sx rows followed by the block of sz rows, bit k of a shot is read from the array it belongs to
__global__ void k_pack_bits2(const uint8_t *__restrict__ bytes, int64_t B, int n_a, int n_b, int words, uint64_t *__restrict__ out) {
  const int64_t n = B * words;
  const uint8_t *A = bytes, *Bs = bytes + B * n_a;
  const int n_bits = n_a + n_b;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t shot = i / words;
    const int w = (int)(i - shot * words);
    const int m = n_bits - 64 * w < 64 ? n_bits - 64 * w : 64;
    uint64_t v = 0;
    for (int k = 0; k < m; ++k) {
      const int b = 64 * w + k;
      const uint8_t x = b < n_a ? A[shot * n_a + b] : Bs[shot * n_b + (b - n_a)];
      v |= (uint64_t)(x & 1u) << k;
    }
    out[i] = v;
  }
}
__global__ void k_unpack_bits(const uint64_t *__restrict__ wordsv, int64_t B, int n_bits, int words, uint8_t *__restrict__ out) {
  const int64_t n = B * (int64_t)n_bits;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t shot = i / n_bits;
    const int b = (int)(i - shot * n_bits);
    out[i] = (uint8_t)((wordsv[shot * words + (b >> 6)] >> (b & 63)) & 1ull);
  }
}
static int grid_for(int64_t n, int sm) {
  const int64_t want = (n + 255) / 256;
  return (int)(want < (int64_t)sm * 16 ? (want > 0 ? want : 1) : (int64_t)sm * 16);
}
}  // namespace tqec

// host-side copy between the caller's (pageable) arrays and the pinned staging buffers, split over a few threads: one
// core moves ~10 GB/s, the decode of a chunk needs ~18 GB/s of bytes at d = 9
static void par_memcpy(void *dst, const void *src, size_t n) {
  const size_t MIN_PART = (size_t)4 << 20;
  unsigned hw = std::thread::hardware_concurrency();
  size_t parts = n / MIN_PART;
  static const size_t max_parts = [] { const char *e = std::getenv("TQEC_COPY_THREADS"); const int v = e ? std::atoi(e) : 0; return (size_t)(v >= 1 && v <= 64 ? v : 8); }();
  if (parts > max_parts) parts = max_parts;
  if (hw && parts > hw) parts = hw;
  if (parts <= 1) { std::memcpy(dst, src, n); return; }
  std::vector<std::thread> th;
  const size_t step = ((n / parts) + 63) & ~(size_t)63;
  for (size_t i = 1; i < parts; ++i) {
    const size_t o = i * step, m = o >= n ? 0 : (i + 1 == parts ? n - o : (o + step > n ? n - o : step));
    if (!m) continue;
    try {
      th.emplace_back([=] { std::memcpy((char *)dst + o, (const char *)src + o, m); });
    } catch (const std::system_error &) {                        // no thread to be had: this part is copied here
      std::memcpy((char *)dst + o, (const char *)src + o, m);
    }
  }
  std::memcpy(dst, src, step < n ? step : n);
  for (auto &t : th) t.join();
}

static int ensure_pinned(void **slot, size_t *cap, size_t bytes) {
  if (*cap >= bytes) return TQEC_OK;
  if (*slot) cudaFreeHost(*slot);
  *slot = nullptr; *cap = 0;
  cudaError_t e = cudaMallocHost(slot, bytes);
  if (e != cudaSuccess) { tqec::set_error("cudaMallocHost(%zu B staging): %s", bytes, cudaGetErrorString(e)); return TQEC_ERR_NOMEM; }
  *cap = bytes;
  return TQEC_OK;
}

// shared body: bytes in -> pack -> decode -> (unpack) -> out.  Three-stage pipeline over chunks of 2^19 shots with two
// slots: the caller's arrays are pageable, so every chunk goes through pinned staging buffers (host copy by par_memcpy,
// then a true asynchronous transfer); H2D of chunk c + 1, the kernels of chunk c, D2H of chunk c - 1 and the host copies
// overlap.
static int decode_bytes(tqec_plan *p, const uint8_t *synd_bits, int64_t B, uint8_t *corr_bits, double *out, int32_t *argmax_out,
                        const uint8_t *synd_b = nullptr, int n_a = 0) {
  tqec::NvtxRange nvtx_range("tqec_decode_bytes");
  const bool mp = p->semiring == TQEC_SEMIRING_MAXPLUS;
  const int nc = p->dev.n_checks, nv = p->dev.n_vars, nsw = p->dev.nsw, ncw = p->dev.ncw;
  const int64_t NO = mp ? 1 : ((int64_t)1 << p->dev.n_obs);
  const int64_t CH = pipeline_chunk(p, (int64_t)1 << 19);
  const int64_t nb = B < CH ? B : CH;
  const size_t nc1 = (size_t)(nc ? nc : 1), nv1 = (size_t)(nv ? nv : 1);
  int rc;
  if ((rc = ensure_pipeline(p))) return rc;
  // device staging, two slots each: [0] syndrome bytes, [1] packed syndromes + packed corrections, [2] outputs,
  // [3] correction bytes (max-plus) or argmax (sum-product)
  const size_t sz0 = (size_t)nb * nc1, sz1 = (size_t)nb * (nsw + ncw) * 8, sz2 = (size_t)nb * NO * 8;
  const size_t sz3 = mp ? (((size_t)nb * nv1 + 15) & ~(size_t)15) : (size_t)nb * 4;
  if ((rc = ensure_cap(&p->d_io[0], &p->io_cap[0], 2 * sz0))) return rc;
  if ((rc = ensure_cap(&p->d_io[1], &p->io_cap[1], 2 * sz1))) return rc;
  if ((rc = ensure_cap(&p->d_io[2], &p->io_cap[2], 2 * sz2))) return rc;
  if ((rc = ensure_cap(&p->d_io[3], &p->io_cap[3], 2 * sz3))) return rc;
  // pinned host staging, two slots: [0] syndrome bytes in, [1] outputs (doubles), [2] correction bytes / argmax out
  if ((rc = ensure_pinned(&p->h_pin[0], &p->h_pin_cap[0], 2 * sz0))) return rc;
  if ((rc = ensure_pinned(&p->h_pin[1], &p->h_pin_cap[1], 2 * sz2))) return rc;
  if ((rc = ensure_pinned(&p->h_pin[2], &p->h_pin_cap[2], 2 * sz3))) return rc;
  const int64_t n_chunks = (B + nb - 1) / nb;
  bool ev_out_used[2] = {false, false};
  std::future<int> out_task;                                     // copy-out of chunk c - 1 (at most one in flight)
  struct Joiner { std::future<int> &f; ~Joiner() { if (f.valid()) f.wait(); } } joiner{out_task};   // early returns wait for it
  for (int64_t c = 0; c <= n_chunks; ++c) {
    const int slot = (int)(c & 1);
    if (c < n_chunks) {
      const int64_t o = c * nb, n = B - o < nb ? B - o : nb;
      uint8_t *d_sb = (uint8_t *)p->d_io[0] + slot * sz0;
      uint64_t *d_syn = (uint64_t *)((char *)p->d_io[1] + slot * sz1), *d_cor = d_syn + (size_t)nb * nsw;
      double *d_out = (double *)((char *)p->d_io[2] + slot * sz2);
      uint8_t *d_x = (uint8_t *)p->d_io[3] + slot * sz3;
      uint8_t *h_in = (uint8_t *)p->h_pin[0] + slot * sz0;
      if (c >= 2) TQEC_CUDA(cudaEventSynchronize(p->ev_in[slot]));      // the staging buffer's previous transfer has left
      if (synd_b) {                                            // two source arrays: block of sx rows, then block of sz rows
        par_memcpy(h_in, synd_bits + (size_t)o * n_a, (size_t)n * n_a);
        par_memcpy(h_in + (size_t)n * n_a, synd_b + (size_t)o * (nc - n_a), (size_t)n * (nc - n_a));
      } else {
        par_memcpy(h_in, synd_bits + (size_t)o * nc, (size_t)n * nc);
      }
      if (c >= 2) TQEC_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_cmp[slot], 0));   // chunk c - 2 has consumed the device slot
      TQEC_CUDA(cudaMemcpyAsync(d_sb, h_in, (size_t)n * nc, cudaMemcpyHostToDevice, p->s_in));
      TQEC_CUDA(cudaEventRecord(p->ev_in[slot], p->s_in));
      TQEC_CUDA(cudaStreamWaitEvent(p->stream, p->ev_in[slot], 0));
      if (ev_out_used[slot]) TQEC_CUDA(cudaStreamWaitEvent(p->stream, p->ev_out[slot], 0));   // chunk c - 2's results have left
      if (synd_b) k_pack_bits2<<<grid_for(n * nsw, p->sm_count), 256, 0, p->stream>>>(d_sb, n, n_a, nc - n_a, nsw, d_syn);
      else k_pack_bits<<<grid_for(n * nsw, p->sm_count), 256, 0, p->stream>>>(d_sb, n, nc, nsw, d_syn);
      TQEC_CUDA(cudaGetLastError());
      if ((rc = launch_decode(p, d_syn, n, mp ? d_cor : nullptr, d_out, mp ? nullptr : (int32_t *)d_x, p->stream))) return rc;
      if (mp) {
        k_unpack_bits<<<grid_for(n * nv, p->sm_count), 256, 0, p->stream>>>(d_cor, n, nv, ncw, d_x);
        TQEC_CUDA(cudaGetLastError());
      }
      TQEC_CUDA(cudaEventRecord(p->ev_cmp[slot], p->stream));
      TQEC_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_cmp[slot], 0));
      // chunk c - 2's results must have left the staging slot before chunk c's transfers are enqueued into it
      if (out_task.valid() && out_task.get() != 0) { tqec::set_error("decode_bytes: waiting for a result transfer failed"); return TQEC_ERR_CUDA; }
      if (mp) {
        TQEC_CUDA(cudaMemcpyAsync((uint8_t *)p->h_pin[2] + slot * sz3, d_x, (size_t)n * nv, cudaMemcpyDeviceToHost, p->s_out));
        if (out) TQEC_CUDA(cudaMemcpyAsync((char *)p->h_pin[1] + slot * sz2, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, p->s_out));
      } else {
        TQEC_CUDA(cudaMemcpyAsync((char *)p->h_pin[1] + slot * sz2, d_out, (size_t)n * NO * 8, cudaMemcpyDeviceToHost, p->s_out));
        if (argmax_out) TQEC_CUDA(cudaMemcpyAsync((uint8_t *)p->h_pin[2] + slot * sz3, d_x, (size_t)n * 4, cudaMemcpyDeviceToHost, p->s_out));
      }
      TQEC_CUDA(cudaEventRecord(p->ev_out[slot], p->s_out));
      ev_out_used[slot] = true;
      p->launches += mp ? 2 : 1;
    }
    if (c >= 1) {
      // results of the previous chunk: pinned staging -> caller's arrays, on a helper thread so that the copy overlaps the
      // next chunk's input copy on this one (joined before the staging slot's next transfer is enqueued, i.e. next pass)
      const int ps = (int)((c - 1) & 1);
      const int64_t o = (c - 1) * nb, n = B - o < nb ? B - o : nb;
      const int dev = p->device;
      if (out_task.valid() && out_task.get() != 0) { tqec::set_error("decode_bytes: waiting for a result transfer failed"); return TQEC_ERR_CUDA; }
      auto copy_out = [=]() -> int {
        if (cudaSetDevice(dev) != cudaSuccess || cudaEventSynchronize(p->ev_out[ps]) != cudaSuccess) return 1;
        if (mp) {
          par_memcpy(corr_bits + (size_t)o * nv, (uint8_t *)p->h_pin[2] + ps * sz3, (size_t)n * nv);
          if (out) std::memcpy(out + o, (char *)p->h_pin[1] + ps * sz2, (size_t)n * 8);
        } else {
          par_memcpy(out + o * NO, (char *)p->h_pin[1] + ps * sz2, (size_t)n * NO * 8);
          if (argmax_out) std::memcpy(argmax_out + o, (uint8_t *)p->h_pin[2] + ps * sz3, (size_t)n * 4);
        }
        return 0;
      };
      try {
        out_task = std::async(std::launch::async, copy_out);
      } catch (const std::system_error &) {                      // no thread to be had: copy on this one
        if (copy_out() != 0) { tqec::set_error("decode_bytes: waiting for a result transfer failed"); return TQEC_ERR_CUDA; }
      }
    }
  }
  if (out_task.valid() && out_task.get() != 0) { tqec::set_error("decode_bytes: waiting for a result transfer failed"); return TQEC_ERR_CUDA; }
  TQEC_CUDA(cudaStreamSynchronize(p->stream));
  return TQEC_OK;
}

extern "C" int tqec_decode_map_bytes(tqec_plan *p, const uint8_t *synd_bits, int64_t B, uint8_t *corr_bits, double *logp_out) {
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_MAXPLUS, "tqec_decode_map_bytes: plan is not a max-plus (TNMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (synd_bits && corr_bits)), "tqec_decode_map_bytes: NULL buffer");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(p->device));
  return decode_bytes(p, synd_bits, B, corr_bits, logp_out, nullptr);
}

extern "C" int tqec_decode_map_bytes2(tqec_plan *p, const uint8_t *synd_a, int32_t n_a, const uint8_t *synd_b, int32_t n_b, int64_t B,
                                      uint8_t *corr_bits, double *logp_out) {
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_MAXPLUS, "tqec_decode_map_bytes2: plan is not a max-plus (TNMAP) plan");
  TQEC_REQUIRE(n_a > 0 && n_b > 0 && n_a + n_b == p->dev.n_checks, "tqec_decode_map_bytes2: %d + %d syndrome bits, the plan has %d checks",
               n_a, n_b, p->dev.n_checks);
  TQEC_REQUIRE(B >= 0 && (B == 0 || (synd_a && synd_b && corr_bits)), "tqec_decode_map_bytes2: NULL buffer");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(p->device));
  return decode_bytes(p, synd_a, B, corr_bits, logp_out, nullptr, synd_b, n_a);
}

extern "C" int tqec_decode_marginal_bytes(tqec_plan *p, const uint8_t *synd_bits, int64_t B, double *mar_out, int32_t *argmax_out) {
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_SUMPROD, "tqec_decode_marginal_bytes: plan is not a sum-product (TNMMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (synd_bits && mar_out)), "tqec_decode_marginal_bytes: NULL buffer");
  if (B == 0) return TQEC_OK;
  TQEC_CUDA(cudaSetDevice(p->device));
  return decode_bytes(p, synd_bits, B, nullptr, mar_out, argmax_out);
}

// TNMMAP with the exponents kept apart (see include/tqec.h): the marginals are returned as mantissas, the per-shot
// power of two (dynamic exponent + the plan's static log2_scale) in log2_out.
extern "C" int tqec_decode_marginal_log2(tqec_plan *p, const uint64_t *synd, int64_t B, double *mar_out, int32_t *log2_out,
                                         int32_t *argmax_out) {
  TQEC_REQUIRE(p && p->semiring == TQEC_SEMIRING_SUMPROD, "tqec_decode_marginal_log2: plan is not a sum-product (TNMMAP) plan");
  TQEC_REQUIRE(B >= 0 && (B == 0 || (synd && mar_out && log2_out)), "tqec_decode_marginal_log2: NULL buffer");
  if (B == 0) return TQEC_OK;
  if (!(p->has_wide && p->wd_dynamic) || p->has_table) {
    const int rc = tqec_decode_marginal(p, synd, B, mar_out, argmax_out);
    for (int64_t b = 0; b < B; ++b) log2_out[b] = 0;
    return rc;
  }
  TQEC_CUDA(cudaSetDevice(p->device));
  const int64_t NO = (int64_t)1 << p->dev.n_obs;
  int rc;
  if ((rc = ensure_cap(&p->d_io[0], &p->io_cap[0], (size_t)B * p->dev.nsw * 8))) return rc;
  if ((rc = ensure_cap(&p->d_io[1], &p->io_cap[1], (size_t)B * NO * 8))) return rc;
  if ((rc = ensure_cap(&p->d_io[2], &p->io_cap[2], (size_t)B * 4))) return rc;
  if ((rc = ensure_cap(&p->d_io[3], &p->io_cap[3], (size_t)B * 4))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(p->d_io[0], synd, (size_t)B * p->dev.nsw * 8, cudaMemcpyHostToDevice, p->stream));
  if ((rc = launch_wide(p, (const uint64_t *)p->d_io[0], B, (double *)p->d_io[1], (int32_t *)p->d_io[2], p->stream, (int32_t *)p->d_io[3]))) return rc;
  TQEC_CUDA(cudaMemcpyAsync(mar_out, p->d_io[1], (size_t)B * NO * 8, cudaMemcpyDeviceToHost, p->stream));
  TQEC_CUDA(cudaMemcpyAsync(log2_out, p->d_io[3], (size_t)B * 4, cudaMemcpyDeviceToHost, p->stream));
  if (argmax_out) TQEC_CUDA(cudaMemcpyAsync(argmax_out, p->d_io[2], (size_t)B * 4, cudaMemcpyDeviceToHost, p->stream));
  TQEC_CUDA(cudaStreamSynchronize(p->stream));
  for (int64_t b = 0; b < B; ++b) log2_out[b] += p->log2_scale;
  return TQEC_OK;
}
